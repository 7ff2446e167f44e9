// K5 — SinkhornDistance: squared-L2 cost matrix, log-domain Sinkhorn with the reference's
// batch-mean early stop, transport plan and cost; forward and exact unrolled backward.
//
// Replaces utils/sinkhorn_distance.py:27-86 of the reference, which materialises a
// [B,P1,P2,D] broadcast temp for the cost matrix, launches ~14 kernels per iteration and does
// a device->host sync per iteration for `err.item() < 0.1`.  Here one CTA per batch element
// keeps C, u, v in shared memory for all iterations.  The early exit depends on the MEAN of
// err over the batch, so every CTA runs all max_iter iterations, records (u_t, v_t, err_t),
// and the finalize kernel picks the first t whose batch-mean err is below the threshold —
// bit-identical control flow to the reference's break, with no host sync and no grid barrier.
// Algorithmic bytes: 4*B*D*(P1+P2) in + 8*B*P1*P2 out (C and pi).
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int SD_THREADS = 256;
constexpr int SD_WARPS = SD_THREADS / 32;
constexpr int DC = 32;  // feature chunk staged through shared memory for the cost matrix

// `spill`: the P1 x P2 matrices do not fit one CTA's shared memory (node sets of ~250-320 rows, config 3): C lives in
// its global output array and the backward's dC accumulator in its global output array instead (both are L2-resident:
// <= 0.5 MB), everything else is unchanged.
__host__ __device__ inline size_t sd_fwd_smem_floats(int P1, int P2, bool spill) {
    return (spill ? 0 : (size_t)P1 * P2) + P1 + P2 + (size_t)(P1 + P2) * (DC + 1) + 32;
}
__host__ __device__ inline size_t sd_bwd_smem_floats(int P1, int P2, bool spill) {
    return (spill ? 0 : (size_t)2 * P1 * P2) + 4 * (size_t)(P1 + P2) + 2 * (size_t)(P1 > P2 ? P1 : P2) + 32;
}

__global__ void __launch_bounds__(SD_THREADS)
sd_iterate_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ Cout,
                  float* __restrict__ hist_u, float* __restrict__ hist_v, float* __restrict__ err,
                  int P1, int P2, int D, float eps, int max_iter, int spill) {
    extern __shared__ __align__(16) float smem[];
    float* C = spill ? Cout + (size_t)blockIdx.x * P1 * P2 : smem;       // [P1][P2]
    float* u = spill ? smem : smem + (size_t)P1 * P2;                    // [P1]
    float* v = u + P1;                     // [P2]
    float* xs = v + P2;                    // [P1][DC+1]
    float* ys = xs + (size_t)P1 * (DC + 1);// [P2][DC+1]
    float* scratch = ys + (size_t)P2 * (DC + 1);

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    x += (size_t)b * P1 * D;
    y += (size_t)b * P2 * D;
    Cout += (size_t)b * P1 * P2;
    hist_u += (size_t)b * max_iter * P1;
    hist_v += (size_t)b * max_iter * P2;
    err += (size_t)b * max_iter;

    // ---- cost matrix C_ij = sum_d (x_id - y_jd)^2   [sinkhorn_distance.py:81-86] ----
    const int npairs = P1 * P2;
    for (int p = tid; p < npairs; p += SD_THREADS) C[p] = 0.f;
    for (int d0 = 0; d0 < D; d0 += DC) {
        const int dc = min(DC, D - d0);
        __syncthreads();
        for (int e = tid; e < P1 * DC; e += SD_THREADS) {
            const int i = e / DC, d = e - i * DC;
            xs[i * (DC + 1) + d] = (d < dc) ? x[(size_t)i * D + d0 + d] : 0.f;
        }
        for (int e = tid; e < P2 * DC; e += SD_THREADS) {
            const int j = e / DC, d = e - j * DC;
            ys[j * (DC + 1) + d] = (d < dc) ? y[(size_t)j * D + d0 + d] : 0.f;
        }
        __syncthreads();
        for (int p = tid; p < npairs; p += SD_THREADS) {
            const int i = p / P2, j = p - i * P2;
            const float* xi = xs + i * (DC + 1);
            const float* yj = ys + j * (DC + 1);
            float acc = C[p];
#pragma unroll 8
            for (int d = 0; d < DC; ++d) {
                const float df = xi[d] - yj[d];
                acc += df * df;
            }
            C[p] = acc;
        }
    }
    __syncthreads();
    if (!spill)
        for (int p = tid; p < npairs; p += SD_THREADS) Cout[p] = C[p];
    for (int i = tid; i < P1; i += SD_THREADS) u[i] = 0.f;
    for (int j = tid; j < P2; j += SD_THREADS) v[j] = 0.f;
    __syncthreads();

    const float logmu = logf((float)(1.0 / (double)P1) + 1e-8f);
    const float lognu = logf((float)(1.0 / (double)P2) + 1e-8f);

    for (int t = 0; t < max_iter; ++t) {
        // u <- eps*(log(mu+1e-8) - LSE_j((-C+u+v)/eps)) + u      [sinkhorn_distance.py:54]
        float errloc = 0.f;
        for (int i = warp; i < P1; i += SD_WARPS) {
            const float ui = u[i];
            float m = -INFINITY;
            for (int j = lane; j < P2; j += 32) m = fmaxf(m, (-C[i * P2 + j] + ui + v[j]) / eps);
            m = ge::warp_max(m);
            float acc = 0.f;
            for (int j = lane; j < P2; j += 32) acc += __expf((-C[i * P2 + j] + ui + v[j]) / eps - m);
            acc = ge::warp_sum(acc);
            const float un = eps * (logmu - (m + logf(acc))) + ui;
            if (lane == 0) {
                errloc += fabsf(un - ui);
                hist_u[(size_t)t * P1 + i] = un;
            }
            __syncwarp();
            if (lane == 0) u[i] = un;
        }
        errloc = ge::block_sum(errloc, scratch);   // includes the barrier that publishes u
        if (tid == 0) err[t] = errloc;
        // v <- eps*(log(nu+1e-8) - LSE_i((-C+u+v)/eps)) + v  with the updated u   [:55]
        for (int j = tid; j < P2; j += SD_THREADS) {
            const float vj = v[j];
            float m = -INFINITY;
            for (int i = 0; i < P1; ++i) m = fmaxf(m, (-C[i * P2 + j] + u[i] + vj) / eps);
            float acc = 0.f;
            for (int i = 0; i < P1; ++i) acc += __expf((-C[i * P2 + j] + u[i] + vj) / eps - m);
            const float vn = eps * (lognu - (m + logf(acc))) + vj;
            v[j] = vn;
            hist_v[(size_t)t * P2 + j] = vn;
        }
        __syncthreads();
    }
}

// Number of executed iterations, exactly as the reference's loop would break
// (sinkhorn_distance.py:56-60): err = mean over batch of sum_i|u-u_prev|; stop once
// float(err) < thresh compared in double.
__device__ __forceinline__ int sd_executed_iters(const float* err, int B, int max_iter, double thresh) {
    for (int t = 0; t < max_iter; ++t) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += err[(size_t)b * max_iter + t];
        const float mean = s / (float)B;
        if ((double)mean < thresh) return t + 1;
    }
    return max_iter;
}

__global__ void __launch_bounds__(SD_THREADS)
sd_finalize_kernel(const float* __restrict__ C, const float* __restrict__ hist_u,
                   const float* __restrict__ hist_v, const float* __restrict__ err,
                   float* __restrict__ pi, float* __restrict__ cost, int* __restrict__ nits_out,
                   int B, int P1, int P2, float eps, int max_iter, double thresh) {
    __shared__ float scratch[32];
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        s_n = sd_executed_iters(err, B, max_iter, thresh);
        if (b == 0) nits_out[0] = s_n;
    }
    __syncthreads();
    const int n = s_n;
    C += (size_t)b * P1 * P2;
    pi += (size_t)b * P1 * P2;
    const float* u = (n > 0) ? hist_u + ((size_t)b * max_iter + (n - 1)) * P1 : nullptr;
    const float* v = (n > 0) ? hist_v + ((size_t)b * max_iter + (n - 1)) * P2 : nullptr;
    float acc = 0.f;
    for (int p = tid; p < P1 * P2; p += SD_THREADS) {
        const int i = p / P2, j = p - i * P2;
        const float c = C[p];
        const float ui = u ? u[i] : 0.f, vj = v ? v[j] : 0.f;
        const float pv = expf((-c + ui + vj) / eps);      // [sinkhorn_distance.py:64]
        pi[p] = pv;
        acc += pv * c;                                     // [:66]
    }
    acc = ge::block_sum(acc, scratch);
    if (tid == 0) cost[b] = acc;
}

// Reverse sweep over the executed iterations; writes dLoss/dC (given dLoss/dcost_b).
__global__ void __launch_bounds__(SD_THREADS)
sd_bwd_kernel(const float* __restrict__ Cg, const float* __restrict__ hist_u,
              const float* __restrict__ hist_v, const int* __restrict__ nits,
              const float* __restrict__ gcost, float* __restrict__ dC,
              int P1, int P2, float eps, int max_iter, int spill) {
    extern __shared__ __align__(16) float smem[];
    const int PM = P1 > P2 ? P1 : P2;
    const float* C = spill ? Cg + (size_t)blockIdx.x * P1 * P2 : smem;             // [P1][P2]
    float* E = spill ? dC + (size_t)blockIdx.x * P1 * P2 : smem + (size_t)P1 * P2;  // [P1][P2]  dC accumulator
    float* ut = spill ? smem : smem + (size_t)2 * P1 * P2;                          // [P1] u_t
    float* up = ut + P1;                    // [P1] u_{t-1}
    float* vt = up + P1;                    // [P2] v_t
    float* vp = vt + P2;                    // [P2] v_{t-1}
    float* gu = vp + P2;                    // [P1]
    float* gv = gu + P1;                    // [P2]
    float* lse = gv + P2;                   // [PM]
    float* tmp = lse + PM;                  // [PM]
    (void)tmp;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = nits[0];
    const float g = gcost[b];
    Cg += (size_t)b * P1 * P2;
    dC += (size_t)b * P1 * P2;
    hist_u += (size_t)b * max_iter * P1;
    hist_v += (size_t)b * max_iter * P2;
    const float inv_eps = 1.f / eps;

    if (!spill)
        for (int p = tid; p < P1 * P2; p += SD_THREADS) smem[p] = Cg[p];
    for (int i = tid; i < P1; i += SD_THREADS) ut[i] = (n > 0) ? hist_u[(size_t)(n - 1) * P1 + i] : 0.f;
    for (int j = tid; j < P2; j += SD_THREADS) vt[j] = (n > 0) ? hist_v[(size_t)(n - 1) * P2 + j] : 0.f;
    __syncthreads();
    // cost = sum pi*C, pi = exp(L), L = (-C+u+v)/eps:
    //   dC = g*pi*(1 - C/eps);  gu_i = g/eps * sum_j pi*C;  gv_j = g/eps * sum_i pi*C
    for (int i = warp; i < P1; i += SD_WARPS) {
        float acc = 0.f;
        for (int j = lane; j < P2; j += 32) {
            const float c = C[i * P2 + j];
            const float pv = expf((-c + ut[i] + vt[j]) / eps);
            const float dl = g * pv * c * inv_eps;
            E[i * P2 + j] = g * pv - dl;
            acc += dl;
        }
        acc = ge::warp_sum(acc);
        if (lane == 0) gu[i] = acc;
    }
    __syncthreads();
    for (int j = tid; j < P2; j += SD_THREADS) {
        float acc = 0.f;
        for (int i = 0; i < P1; ++i) {
            const float c = C[i * P2 + j];
            acc += g * expf((-c + ut[i] + vt[j]) / eps) * c * inv_eps;
        }
        gv[j] = acc;
    }
    __syncthreads();

    for (int t = n; t >= 1; --t) {
        // potentials before this iteration
        for (int i = tid; i < P1; i += SD_THREADS) up[i] = (t > 1) ? hist_u[(size_t)(t - 2) * P1 + i] : 0.f;
        for (int j = tid; j < P2; j += SD_THREADS) vp[j] = (t > 1) ? hist_v[(size_t)(t - 2) * P2 + j] : 0.f;
        __syncthreads();
        // ---- adjoint of v_t = eps*(lognu - LSE_i Lv) + v_{t-1},  Lv = (-C + u_t + v_{t-1})/eps
        for (int j = tid; j < P2; j += SD_THREADS) {
            float m = -INFINITY;
            for (int i = 0; i < P1; ++i) m = fmaxf(m, (-C[i * P2 + j] + ut[i] + vp[j]) / eps);
            float acc = 0.f;
            for (int i = 0; i < P1; ++i) acc += __expf((-C[i * P2 + j] + ut[i] + vp[j]) / eps - m);
            lse[j] = m + logf(acc);
        }
        __syncthreads();
        for (int i = warp; i < P1; i += SD_WARPS) {
            float acc = 0.f;
            for (int j = lane; j < P2; j += 32) {
                const float w = __expf((-C[i * P2 + j] + ut[i] + vp[j]) / eps - lse[j]) * gv[j];
                E[i * P2 + j] += w;     // dv_j/dC_ij = +W
                acc += w;               // dv_j/du_i  = -W
            }
            acc = ge::warp_sum(acc);
            if (lane == 0) gu[i] -= acc;
        }
        __syncthreads();
        // ---- adjoint of u_t = eps*(logmu - LSE_j Lu) + u_{t-1},  Lu = (-C + u_{t-1} + v_{t-1})/eps
        for (int i = warp; i < P1; i += SD_WARPS) {
            float m = -INFINITY;
            for (int j = lane; j < P2; j += 32) m = fmaxf(m, (-C[i * P2 + j] + up[i] + vp[j]) / eps);
            m = ge::warp_max(m);
            float acc = 0.f;
            for (int j = lane; j < P2; j += 32) acc += __expf((-C[i * P2 + j] + up[i] + vp[j]) / eps - m);
            acc = ge::warp_sum(acc);
            if (lane == 0) lse[i] = m + logf(acc);
        }
        __syncthreads();
        for (int j = tid; j < P2; j += SD_THREADS) {
            float acc = 0.f;
            for (int i = 0; i < P1; ++i) {
                const float w = __expf((-C[i * P2 + j] + up[i] + vp[j]) / eps - lse[i]) * gu[i];
                E[i * P2 + j] += w;     // du_i/dC_ij = +W
                acc += w;               // du_i/dv_j  = -W   (du_i/du_{t-1} = 0, dv_j/dv_{t-1} = 0)
            }
            gv[j] = -acc;
        }
        __syncthreads();
        for (int i = tid; i < P1; i += SD_THREADS) { gu[i] = 0.f; ut[i] = up[i]; }
        for (int j = tid; j < P2; j += SD_THREADS) vt[j] = vp[j];
        __syncthreads();
    }
    if (!spill)
        for (int p = tid; p < P1 * P2; p += SD_THREADS) dC[p] = E[p];
}

// dx_id = 2*sum_j dC_ij (x_id - y_jd),  dy_jd = -2*sum_i dC_ij (x_id - y_jd)
__global__ void __launch_bounds__(256)
sd_bwd_xy_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dC,
                 float* __restrict__ dx, float* __restrict__ dy, int P1, int P2, int D, int spill) {
    extern __shared__ __align__(16) float gs[];  // [P1][P2]
    const int b = blockIdx.y;
    x += (size_t)b * P1 * D;
    y += (size_t)b * P2 * D;
    dC += (size_t)b * P1 * P2;
    dx += (size_t)b * P1 * D;
    dy += (size_t)b * P2 * D;
    const float* g = dC;
    if (!spill) {
        for (int p = threadIdx.x; p < P1 * P2; p += blockDim.x) gs[p] = dC[p];
        g = gs;
    }
    __syncthreads();
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    for (int i = 0; i < P1; ++i) {
        const float xi = x[(size_t)i * D + d];
        float acc = 0.f;
        for (int j = 0; j < P2; ++j) acc = fmaf(g[i * P2 + j], xi - __ldg(y + (size_t)j * D + d), acc);
        dx[(size_t)i * D + d] = 2.f * acc;
    }
    for (int j = 0; j < P2; ++j) {
        const float yj = y[(size_t)j * D + d];
        float acc = 0.f;
        for (int i = 0; i < P1; ++i) acc = fmaf(g[i * P2 + j], __ldg(x + (size_t)i * D + d) - yj, acc);
        dy[(size_t)j * D + d] = -2.f * acc;
    }
}

constexpr size_t kCap = 220 * 1024;

}  // namespace

extern "C" int ge_sinkhorn_distance_fwd(const float* x, const float* y, float* C, float* pi, float* cost,
                                        float* hist_u, float* hist_v, float* err, int* nits,
                                        int B, int P1, int P2, int D, float eps, int max_iter,
                                        double thresh, ge_stream_t stream) {
    GE_REQUIRE(x && y && C && pi && cost && hist_u && hist_v && err && nits, GE_ERR_ARG,
               "ge_sinkhorn_distance_fwd: null pointer");
    GE_REQUIRE(B > 0 && P1 > 0 && P2 > 0 && D > 0 && max_iter >= 0 && eps > 0.f, GE_ERR_ARG,
               "ge_sinkhorn_distance_fwd: bad dimension");
    const int spill = sd_fwd_smem_floats(P1, P2, false) * sizeof(float) > kCap;
    const size_t smem = sd_fwd_smem_floats(P1, P2, spill) * sizeof(float);
    GE_REQUIRE(smem <= kCap, GE_ERR_CAPACITY,
               "ge_sinkhorn_distance_fwd: P1=%d + P2=%d rows do not fit one CTA's shared memory", P1, P2);
    cudaStream_t st = (cudaStream_t)stream;
    { static size_t ge_max_smem__ = 0; if ((size_t)(smem) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(sd_iterate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)), "ge_sinkhorn_distance_fwd(attr)"); ge_max_smem__ = (size_t)(smem); } }
    sd_iterate_kernel<<<B, SD_THREADS, smem, st>>>(x, y, C, hist_u, hist_v, err, P1, P2, D, eps, max_iter, spill);
    GE_CHECK_LAUNCH("ge_sinkhorn_distance_fwd(iterate)");
    sd_finalize_kernel<<<B, SD_THREADS, 0, st>>>(C, hist_u, hist_v, err, pi, cost, nits, B, P1, P2, eps, max_iter, thresh);
    GE_CHECK_LAUNCH("ge_sinkhorn_distance_fwd(finalize)");
    return GE_OK;
}

extern "C" int ge_sinkhorn_distance_bwd(const float* x, const float* y, const float* C, const float* hist_u,
                                        const float* hist_v, const int* nits, const float* gcost,
                                        float* dC, float* dx, float* dy,
                                        int B, int P1, int P2, int D, float eps, int max_iter,
                                        ge_stream_t stream) {
    GE_REQUIRE(x && y && C && hist_u && hist_v && nits && gcost && dC && ((dx && dy) || (!dx && !dy)), GE_ERR_ARG,
               "ge_sinkhorn_distance_bwd: null pointer");
    GE_REQUIRE(B > 0 && P1 > 0 && P2 > 0 && D > 0 && max_iter >= 0 && eps > 0.f, GE_ERR_ARG,
               "ge_sinkhorn_distance_bwd: bad dimension");
    const int spill = sd_bwd_smem_floats(P1, P2, false) * sizeof(float) > kCap;
    const size_t smem = sd_bwd_smem_floats(P1, P2, spill) * sizeof(float);
    GE_REQUIRE(smem <= kCap, GE_ERR_CAPACITY,
               "ge_sinkhorn_distance_bwd: P1=%d + P2=%d rows do not fit one CTA's shared memory", P1, P2);
    cudaStream_t st = (cudaStream_t)stream;
    { static size_t ge_max_smem__ = 0; if ((size_t)(smem) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(sd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)), "ge_sinkhorn_distance_bwd(attr)"); ge_max_smem__ = (size_t)(smem); } }
    sd_bwd_kernel<<<B, SD_THREADS, smem, st>>>(C, hist_u, hist_v, nits, gcost, dC, P1, P2, eps, max_iter, spill);
    GE_CHECK_LAUNCH("ge_sinkhorn_distance_bwd(sweep)");
    if (dx == nullptr) return GE_OK;      // the caller turns dC into dx, dy itself (two GEMMs for large node sets)
    const int spill2 = (size_t)P1 * P2 * sizeof(float) > kCap;
    const size_t smem2 = spill2 ? 0 : (size_t)P1 * P2 * sizeof(float);
    { static size_t ge_max_smem__ = 0; if ((size_t)(smem2) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(sd_bwd_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem2)), "ge_sinkhorn_distance_bwd(attr2)"); ge_max_smem__ = (size_t)(smem2); } }
    dim3 grid(ge::cdiv(D, 256), B);
    sd_bwd_xy_kernel<<<grid, 256, smem2, st>>>(x, y, dC, dx, dy, P1, P2, D, spill2);
    GE_CHECK_LAUNCH("ge_sinkhorn_distance_bwd(xy)");
    return GE_OK;
}
