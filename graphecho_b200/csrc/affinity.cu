// K3 — pairwise affinity  M_ij = sum_k w2_k * relu(A_ik + B_jk) + b2
//
// Replaces the reference's broadcast-concat + Linear(512,512)+ReLU+Linear(512,1)
// (/root/reference/models/affinity_layer.py:52-73) with its separable form: the first
// Linear acts on [P_s x_i ; P_t y_j], so W1 [x;y] + b1 = A_i + B_j with
//   A = X P_s^T W1[:, :256]^T           [N1, H]
//   B = Y P_t^T W1[:, 256:]^T + b1      [N2, H]
// (two small dense projections, done by the caller with cuBLAS), and only the
// relu-coupled pairwise reduction over k needs a custom kernel.  The [N1,N2,512] tensor
// the reference materialises never exists here.
//
// Work per problem: 3*H*N1*N2 fp32 ALU ops (add, max, fma) — CUDA-core bound, not HBM
// bound: algorithmic bytes are 4*H*(N1+N2) + 4*N1*N2.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int TI = 32;       // output tile rows per CTA
constexpr int TJ = 32;       // output tile cols per CTA
constexpr int FWD_THREADS = 256;
constexpr int FWD_WARPS = FWD_THREADS / 32;
constexpr int RED_LD = 40;   // padded row stride of the cross-warp reduction tile

// Each CTA: 32x32 outputs, full K in shared memory.  The 8 warps split K (split-K inside
// the CTA, so a single 250x250 problem still yields 64 CTAs x 8 warps), every lane owns an
// 8x4 register tile with rows li+4r / cols lj+8c so that the 128-bit shared loads are
// bank-conflict free with a row stride of H+4 floats.
__global__ void __launch_bounds__(FWD_THREADS)
affinity_pairwise_fwd_kernel(const float* __restrict__ A, const float* __restrict__ B,
                             const float* __restrict__ w2, const float* __restrict__ b2,
                             float* __restrict__ M, int N1, int N2, int H) {
    extern __shared__ __align__(16) float smem[];
    const int ld = H + 4;
    float* As = smem;                 // [TI][ld]
    float* Bs = As + TI * ld;         // [TJ][ld]
    float* ws = Bs + TJ * ld;         // [H]

    const int b = blockIdx.z;
    A += (size_t)b * N1 * H;
    B += (size_t)b * N2 * H;
    M += (size_t)b * N1 * N2;
    const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
    const int tid = threadIdx.x;
    const int h4 = H >> 2;

    for (int e = tid; e < TI * h4; e += FWD_THREADS) {
        const int r = e / h4, k4 = e - r * h4;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (i0 + r < N1) va = __ldg(reinterpret_cast<const float4*>(A + (size_t)(i0 + r) * H) + k4);
        if (j0 + r < N2) vb = __ldg(reinterpret_cast<const float4*>(B + (size_t)(j0 + r) * H) + k4);
        *reinterpret_cast<float4*>(As + r * ld + 4 * k4) = va;
        *reinterpret_cast<float4*>(Bs + r * ld + 4 * k4) = vb;
    }
    for (int k = tid; k < H; k += FWD_THREADS) ws[k] = __ldg(w2 + k);
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int li = lane >> 3, lj = lane & 7;
    const int kslice = H / FWD_WARPS;   // host guarantees H % 32 == 0
    const int kbeg = warp * kslice, kend = kbeg + kslice;

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int k = kbeg; k < kend; k += 4) {
        float4 a4[8], b4[4];
#pragma unroll
        for (int r = 0; r < 8; ++r) a4[r] = *reinterpret_cast<const float4*>(As + (li + 4 * r) * ld + k);
#pragma unroll
        for (int c = 0; c < 4; ++c) b4[c] = *reinterpret_cast<const float4*>(Bs + (lj + 8 * c) * ld + k);
        const float4 w4 = *reinterpret_cast<const float4*>(ws + k);
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                acc[r][c] = fmaf(fmaxf(a4[r].x + b4[c].x, 0.f), w4.x, acc[r][c]);
                acc[r][c] = fmaf(fmaxf(a4[r].y + b4[c].y, 0.f), w4.y, acc[r][c]);
                acc[r][c] = fmaf(fmaxf(a4[r].z + b4[c].z, 0.f), w4.z, acc[r][c]);
                acc[r][c] = fmaf(fmaxf(a4[r].w + b4[c].w, 0.f), w4.w, acc[r][c]);
            }
    }
    __syncthreads();
    float* red = smem;  // reuse: [FWD_WARPS][TI * RED_LD]
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            red[warp * (TI * RED_LD) + (li + 4 * r) * RED_LD + lj + 8 * c] = acc[r][c];
    __syncthreads();
    const float bias = __ldg(b2);
    for (int o = tid; o < TI * TJ; o += FWD_THREADS) {
        const int i = o >> 5, j = o & 31;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < FWD_WARPS; ++w) s += red[w * (TI * RED_LD) + i * RED_LD + j];
        if (i0 + i < N1 && j0 + j < N2) M[(size_t)(i0 + i) * N2 + j0 + j] = s + bias;
    }
}

// Backward sweep.  One thread per k, RT rows of P per CTA, loop over all rows of Q.
//   TRANS=false: P=A (rows i), Q=B, g(r,q) = dM[(i0+r)*N2 + q]   -> dA, partial dw2
//   TRANS=true : P=B (rows j), Q=A, g(r,q) = dM[q*N2 + (j0+r)]   -> dB
constexpr int RT = 4;

template <bool TRANS>
__global__ void __launch_bounds__(512)
affinity_pairwise_bwd_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                             const float* __restrict__ w2, const float* __restrict__ dM,
                             float* __restrict__ dP, float* __restrict__ dw2_part,
                             int NP, int NQ, int H, int N1, int N2) {
    extern __shared__ __align__(16) float g[];  // [RT][NQ]
    const int b = blockIdx.y;
    P += (size_t)b * NP * H;
    Q += (size_t)b * NQ * H;
    dM += (size_t)b * N1 * N2;
    dP += (size_t)b * NP * H;
    const int r0 = blockIdx.x * RT;
    for (int e = threadIdx.x; e < RT * NQ; e += blockDim.x) {
        const int r = e / NQ, q = e - r * NQ;
        float v = 0.f;
        if (r0 + r < NP) v = TRANS ? dM[(size_t)q * N2 + (r0 + r)] : dM[(size_t)(r0 + r) * N2 + q];
        g[e] = v;
    }
    __syncthreads();
    float accW = 0.f;
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        float p[RT], accP[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            p[r] = (r0 + r < NP) ? __ldg(P + (size_t)(r0 + r) * H + k) : 0.f;
            accP[r] = 0.f;
        }
        int q = 0;
        for (; q + 4 <= NQ; q += 4) {
            float qv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) qv[u] = __ldg(Q + (size_t)(q + u) * H + k);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float s = p[r] + qv[u];
                    const float gv = g[r * NQ + q + u];
                    accP[r] += (s > 0.f) ? gv : 0.f;
                    if (!TRANS) accW = fmaf(gv, fmaxf(s, 0.f), accW);
                }
        }
        for (; q < NQ; ++q) {
            const float qv = __ldg(Q + (size_t)q * H + k);
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const float s = p[r] + qv;
                const float gv = g[r * NQ + q];
                accP[r] += (s > 0.f) ? gv : 0.f;
                if (!TRANS) accW = fmaf(gv, fmaxf(s, 0.f), accW);
            }
        }
        const float wk = __ldg(w2 + k);
#pragma unroll
        for (int r = 0; r < RT; ++r)
            if (r0 + r < NP) dP[(size_t)(r0 + r) * H + k] = wk * accP[r];
        if (!TRANS) {
            dw2_part[((size_t)b * gridDim.x + blockIdx.x) * H + k] = accW;
            accW = 0.f;
        }
    }
}

// dw2[k] = sum over (batch, row tiles) of the partials; db2 = sum(dM).  One CTA.
__global__ void __launch_bounds__(512)
affinity_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int H,
                             const float* __restrict__ dM, long long ndm,
                             float* __restrict__ dw2, float* __restrict__ db2) {
    __shared__ float scratch[32];
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += part[(size_t)p * H + k];
        dw2[k] = s;
    }
    float s = 0.f;
    for (long long e = threadIdx.x; e < ndm; e += blockDim.x) s += dM[e];
    s = ge::block_sum(s, scratch);
    if (threadIdx.x == 0) db2[0] = s;
}

}  // namespace

extern "C" int ge_affinity_pairwise_fwd(const float* A, const float* B, const float* w2, const float* b2,
                                        float* M, int batch, int N1, int N2, int H, ge_stream_t stream) {
    GE_REQUIRE(A && B && w2 && b2 && M, GE_ERR_ARG, "ge_affinity_pairwise_fwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && H > 0, GE_ERR_ARG, "ge_affinity_pairwise_fwd: non-positive dimension");
    GE_REQUIRE(H % 32 == 0 && H <= 640, GE_ERR_SHAPE,
               "ge_affinity_pairwise_fwd: hidden width H=%d must be a multiple of 32 and <= 640", H);
    const size_t smem = ((size_t)(TI + TJ) * (H + 4) + H) * sizeof(float);
    const size_t red = (size_t)FWD_WARPS * TI * RED_LD * sizeof(float);
    const size_t bytes = smem > red ? smem : red;
    { static size_t ge_max_smem__ = 0; if ((size_t)(bytes) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(affinity_pairwise_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)), "ge_affinity_pairwise_fwd(attr)"); ge_max_smem__ = (size_t)(bytes); } }
    dim3 grid(ge::cdiv(N2, TJ), ge::cdiv(N1, TI), batch);
    affinity_pairwise_fwd_kernel<<<grid, FWD_THREADS, bytes, (cudaStream_t)stream>>>(A, B, w2, b2, M, N1, N2, H);
    GE_CHECK_LAUNCH("ge_affinity_pairwise_fwd");
    return GE_OK;
}

extern "C" size_t ge_affinity_pairwise_bwd_workspace_bytes(int batch, int N1, int N2, int H) {
    (void)N2;
    if (batch <= 0 || N1 <= 0 || H <= 0) return 0;
    return (size_t)batch * ge::cdiv(N1, RT) * H * sizeof(float);
}

extern "C" int ge_affinity_pairwise_bwd(const float* A, const float* B, const float* w2, const float* dM,
                                        float* dA, float* dB, float* dw2, float* db2,
                                        void* workspace, size_t workspace_bytes,
                                        int batch, int N1, int N2, int H, ge_stream_t stream) {
    GE_REQUIRE(A && B && w2 && dM && dA && dB && dw2 && db2 && workspace, GE_ERR_ARG,
               "ge_affinity_pairwise_bwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && H > 0, GE_ERR_ARG, "ge_affinity_pairwise_bwd: non-positive dimension");
    GE_REQUIRE(workspace_bytes >= ge_affinity_pairwise_bwd_workspace_bytes(batch, N1, N2, H), GE_ERR_ARG,
               "ge_affinity_pairwise_bwd: workspace too small");
    GE_REQUIRE((size_t)RT * (N1 > N2 ? N1 : N2) * sizeof(float) <= 200 * 1024, GE_ERR_CAPACITY,
               "ge_affinity_pairwise_bwd: N=%d too large for the shared dM tile", N1 > N2 ? N1 : N2);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = static_cast<float*>(workspace);
    const int threads = H >= 512 ? 512 : ((H + 31) / 32) * 32;
    {
        const size_t smem = (size_t)RT * N2 * sizeof(float);
        { static size_t ge_max_smem__ = 0; if ((size_t)(smem) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(affinity_pairwise_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)), "ge_affinity_pairwise_bwd(attr)"); ge_max_smem__ = (size_t)(smem); } }
        dim3 grid(ge::cdiv(N1, RT), batch);
        affinity_pairwise_bwd_kernel<false><<<grid, threads, smem, st>>>(A, B, w2, dM, dA, part, N1, N2, H, N1, N2);
        GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(dA)");
    }
    {
        const size_t smem = (size_t)RT * N1 * sizeof(float);
        { static size_t ge_max_smem__ = 0; if ((size_t)(smem) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(affinity_pairwise_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)), "ge_affinity_pairwise_bwd(attr)"); ge_max_smem__ = (size_t)(smem); } }
        dim3 grid(ge::cdiv(N2, RT), batch);
        affinity_pairwise_bwd_kernel<true><<<grid, threads, smem, st>>>(B, A, w2, dM, dB, nullptr, N2, N1, H, N1, N2);
        GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(dB)");
    }
    affinity_bwd_finalize_kernel<<<1, 512, 0, st>>>(part, batch * ge::cdiv(N1, RT), H, dM,
                                                    (long long)batch * N1 * N2, dw2, db2);
    GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(finalize)");
    return GE_OK;
}
