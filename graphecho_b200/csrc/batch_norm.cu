// Fused BatchNorm2d (+ residual add) (+ ReLU) on NHWC feature maps, training and inference, forward and
// backward — the glue between the backbone's tensor-core convolutions.
//
// The reference runs nn.BatchNorm2d, the residual `out += identity` and nn.ReLU as separate full-tensor
// passes (/root/reference/models/fpnseg.py:192-212 Bottleneck.forward, :251-255 stem, :27-142 VGG16
// blocks): 8 tensor passes forward and 8 backward per residual block output.  Here:
//   forward : partial statistics (1 read) -> finalize (+ running-stat update) -> apply (1-2 reads, 1 write)
//   backward: partial sums of dy*relu' and dy*relu'*xhat (3 reads) -> finalize (dgamma, dbeta) ->
//             apply (3 reads, 1-2 writes)
// Statistics are reduced in two deterministic stages (per-CTA partials in a workspace, then one small
// CTA), never with atomics.  The statistics kernels move 8 consecutive channels per thread (one 128-bit access in
// bf16), the apply kernels 4 (half the registers, twice the occupancy).
// HBM-bound; algorithmic bytes per element (bf16): forward 6 (+2 with a residual), backward 14 (+2).
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using bf16 = __nv_bfloat16;
using namespace ge;

constexpr int BN_THREADS = 256;

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) { load8<float>(p, f); }

// ---- stage 1 (forward): per-CTA partial sums of (x - shift) and (x - shift)^2 --------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_partial_stats_kernel(const T* __restrict__ x, const float* __restrict__ shift_src, float* __restrict__ part,
                        long long P, int C, long long pix_per_cta) {
    extern __shared__ __align__(16) float sm[];       // [nPL][C] x 2
    const int c8 = C >> 3, nPL = BN_THREADS / c8;
    float* ssum = sm;
    float* ssq = sm + (size_t)nPL * C;
    const int tid = threadIdx.x, co = tid % c8, pl = tid / c8;
    const long long p0 = (long long)blockIdx.x * pix_per_cta;
    const long long p1 = min(P, p0 + pix_per_cta);
    float shift[8], a[8], b[8];
    load8f(shift_src + co * 8, shift);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = 0.f; b[u] = 0.f; }
    if (pl < nPL) {
        const T* xb = x + co * 8;
#pragma unroll 4
        for (long long p = p0 + pl; p < p1; p += nPL) {
            float v[8];
            load8<T>(xb + p * C, v);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float d = v[u] - shift[u];
                a[u] += d;
                b[u] = fmaf(d, d, b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { ssum[pl * C + co * 8 + u] = a[u]; ssq[pl * C + co * 8 + u] = b[u]; }
    }
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * 2 * C;
    for (int c = tid; c < C; c += BN_THREADS) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += ssum[q * C + c]; sb += ssq[q * C + c]; }
        out[c] = sa;
        out[C + c] = sb;
    }
}

// Sum the per-CTA partials of 32 channels with FL part-lanes (CTA = 32*FL threads, grid = ceil(C/32)).  The loop is
// latency-bound (one L2 round trip per row), so every thread keeps 8 independent rows (16 loads) in flight and
// FL = 32 lanes share the rows: <= 3 trips at 592 partials.
constexpr int FL = 32;

__device__ __forceinline__ void reduce_parts(const float* __restrict__ part, int nparts, int C, int c, int lane_p,
                                             float (*sh)[2][33], float& sa, float& sb) {
    float a = 0.f, b = 0.f;
    if (c < C) {
        int q = lane_p;
        for (; q + 7 * FL < nparts; q += 8 * FL) {
            float va[8], vb[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                va[u] = part[(size_t)(q + FL * u) * 2 * C + c];
                vb[u] = part[(size_t)(q + FL * u) * 2 * C + C + c];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { a += va[u]; b += vb[u]; }
        }
        for (; q < nparts; q += FL) {
            a += part[(size_t)q * 2 * C + c];
            b += part[(size_t)q * 2 * C + C + c];
        }
    }
    sh[lane_p][0][threadIdx.x & 31] = a;
    sh[lane_p][1][threadIdx.x & 31] = b;
    __syncthreads();
    sa = 0.f; sb = 0.f;
    if (lane_p == 0) {
#pragma unroll
        for (int q = 0; q < FL; ++q) { sa += sh[q][0][threadIdx.x & 31]; sb += sh[q][1][threadIdx.x & 31]; }
    }
}

// ---- stage 2 (forward): batch mean / rstd, running statistics (nn.BatchNorm2d semantics) --------
__global__ void __launch_bounds__(32 * FL)
bn_finalize_stats_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ shift_src,
                         float* __restrict__ save_mean, float* __restrict__ save_rstd,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         long long P, int C, float eps, float momentum) {
    __shared__ float sh[FL][2][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane8 = threadIdx.x >> 5;
    float sa, sb;
    reduce_parts(part, nparts, C, c, lane8, sh, sa, sb);
    if (lane8 != 0 || c >= C) return;
    const float inv = 1.f / (float)P;
    const float md = sa * inv;
    const float var = fmaxf(sb * inv - md * md, 0.f);          // biased, used to normalise
    const float mean = md + shift_src[c];
    save_mean[c] = mean;
    save_rstd[c] = 1.f / sqrtf(var + eps);
    if (running_mean != nullptr) {
        const float unbiased = P > 1 ? var * ((float)P / (float)(P - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// ---- stage 3 (forward): out = act( x*sc + sh (+ residual) ) --------------------------------------
// MODE 0: training (sc/sh from save_mean/save_rstd); MODE 1: inference (from running stats, `rstd` = var).
// A thread owns FOUR channels (one 64-bit bf16 access) of FPIX4 consecutive pixels: the per-channel constants are
// loaded once and FPIX4 independent loads are in flight before the first use (the first version moved 8 channels per
// thread at 93 registers -> 21 % occupancy in the round-1 ncu capture).
constexpr int FPIX4 = 4;

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
bn_apply_fwd4_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ mean,
                     const float* __restrict__ rstd_or_var, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ out, long long P, int C, float eps, int relu) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    const int cc = (int)(t % c4) * 4;
    const long long p0 = (t / c4) * FPIX4;
    if (p0 >= P) return;
    const float4 m4 = *reinterpret_cast<const float4*>(mean + cc), r4 = *reinterpret_cast<const float4*>(rstd_or_var + cc);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc), b4 = *reinterpret_cast<const float4*>(beta + cc);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    float sc[4], sh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float rs = (MODE == 0) ? r[u] : 1.f / sqrtf(r[u] + eps);
        sc[u] = rs * g[u];
        sh[u] = b[u] - m[u] * sc[u];
    }
    Vec4<T> v[FPIX4], rv[FPIX4];
#pragma unroll
    for (int q = 0; q < FPIX4; ++q)
        if (p0 + q < P) {
            v[q].load(x + (p0 + q) * C + cc);
            if (res != nullptr) rv[q].load(res + (p0 + q) * C + cc);
        }
#pragma unroll
    for (int q = 0; q < FPIX4; ++q) {
        if (p0 + q >= P) break;
        float f[4], rf[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
        v[q].get(f);
        if (res != nullptr) rv[q].get(rf);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            o[u] = fmaf(f[u], sc[u], sh[u]);
            if (res != nullptr) o[u] += rf[u];
            if (relu) o[u] = fmaxf(o[u], 0.f);
        }
        Vec4<T> w;
        w.set(o);
        w.store(out + (p0 + q) * C + cc);
    }
}

// ---- backward stage 1: partial sums of dyr = dy * relu'(out) and dyr * xhat ----------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_partial_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ out, const T* __restrict__ x,
                      const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ part,
                      long long P, int C, long long pix_per_cta, int relu) {
    extern __shared__ __align__(16) float sm[];
    const int c8 = C >> 3, nPL = BN_THREADS / c8;
    float* s1 = sm;
    float* s2 = sm + (size_t)nPL * C;
    const int tid = threadIdx.x, co = tid % c8, pl = tid / c8, cc = co * 8;
    const long long p0 = (long long)blockIdx.x * pix_per_cta;
    const long long p1 = min(P, p0 + pix_per_cta);
    float m[8], r[8], a[8], b[8];
    load8f(mean + cc, m); load8f(rstd + cc, r);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = 0.f; b[u] = 0.f; }
    if (pl < nPL) {
#pragma unroll 2
        for (long long p = p0 + pl; p < p1; p += nPL) {
            float g[8], xv[8];
            load8<T>(dy + p * C + cc, g);
            load8<T>(x + p * C + cc, xv);
            if (relu) {
                float ov[8];
                load8<T>(out + p * C + cc, ov);
#pragma unroll
                for (int u = 0; u < 8; ++u) g[u] = (ov[u] > 0.f) ? g[u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] += g[u];
                b[u] = fmaf(g[u], (xv[u] - m[u]) * r[u], b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1[pl * C + cc + u] = a[u]; s2[pl * C + cc + u] = b[u]; }
    }
    __syncthreads();
    float* o = part + (size_t)blockIdx.x * 2 * C;
    for (int c = tid; c < C; c += BN_THREADS) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += s1[q * C + c]; sb += s2[q * C + c]; }
        o[c] = sa;
        o[C + c] = sb;
    }
}

// ---- backward stage 2: dbeta = S1, dgamma = S2 ----------------------------------------------------
__global__ void __launch_bounds__(32 * FL)
bn_finalize_bwd_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, int C) {
    __shared__ float sh[FL][2][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane8 = threadIdx.x >> 5;
    float sa, sb;
    reduce_parts(part, nparts, C, c, lane8, sh, sa, sb);
    if (lane8 != 0 || c >= C) return;
    dbeta[c] = sa;
    dgamma[c] = sb;
}

// ---- backward stage 3: dx = rstd*gamma*(dyr - S1/P - xhat*S2/P);  dres = dyr ------------------------
// FOUR channels per thread (one 64-bit bf16 access) of BPIX4 consecutive pixels -- about half the live state of the
// first, 8-channel version (111 registers, 20 % occupancy, 44 % of DRAM peak in the round-1 ncu capture), so twice
// as many loads are in flight per SM.
constexpr int BPIX4 = 4;

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_bwd4_kernel(const T* __restrict__ dy, const T* __restrict__ out, const T* __restrict__ x,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ dgamma,
                     const float* __restrict__ dbeta, T* __restrict__ dx, T* __restrict__ dres,
                     long long P, int C, float invP, int relu) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    const int cc = (int)(t % c4) * 4;
    const long long p0 = (t / c4) * BPIX4;
    if (p0 >= P) return;
    const float4 m4 = *reinterpret_cast<const float4*>(mean + cc), r4 = *reinterpret_cast<const float4*>(rstd + cc);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc);
    const float4 b4 = *reinterpret_cast<const float4*>(dbeta + cc), q4 = *reinterpret_cast<const float4*>(dgamma + cc);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    const float s1[4] = {b4.x, b4.y, b4.z, b4.w}, s2[4] = {q4.x, q4.y, q4.z, q4.w};
    float k0[4], k1[4], k2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {            // dx = k0*dyr - k1 - (x-m)*k2
        k0[u] = r[u] * g[u];
        k1[u] = k0[u] * s1[u] * invP;
        k2[u] = k0[u] * s2[u] * invP * r[u];
    }
    Vec4<T> gv[BPIX4], xv[BPIX4], ov[BPIX4];
#pragma unroll
    for (int q = 0; q < BPIX4; ++q)
        if (p0 + q < P) {
            gv[q].load(dy + (p0 + q) * C + cc);
            xv[q].load(x + (p0 + q) * C + cc);
            if (relu) ov[q].load(out + (p0 + q) * C + cc);
        }
#pragma unroll
    for (int q = 0; q < BPIX4; ++q) {
        if (p0 + q >= P) break;
        float gf[4], xf[4], of[4] = {1.f, 1.f, 1.f, 1.f}, o[4];
        gv[q].get(gf); xv[q].get(xf);
        if (relu) ov[q].get(of);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float gq = (relu && !(of[u] > 0.f)) ? 0.f : gf[u];
            gf[u] = gq;
            o[u] = k0[u] * gq - k1[u] - (xf[u] - m[u]) * k2[u];
        }
        Vec4<T> w;
        if (dres != nullptr) { w.set(gf); w.store(dres + (p0 + q) * C + cc); }
        w.set(o);
        w.store(dx + (p0 + q) * C + cc);
    }
}

int bn_chunks(long long P, int C) {
    const int nPL = BN_THREADS / (C / 8);
    long long want = P / ((long long)nPL * 8);           // >= 8 pixels per pixel-lane
    const long long cap = (long long)ge::sm_count() * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

bool bn_shape_ok(int C) { return C % 8 == 0 && C / 8 <= BN_THREADS && BN_THREADS % (C / 8) == 0; }

size_t bn_smem(int C) { return (size_t)2 * (BN_THREADS / (C / 8)) * C * sizeof(float); }

}  // namespace

extern "C" size_t ge_bn_workspace_bytes(long long P, int C) {
    if (P <= 0 || C <= 0 || !bn_shape_ok(C)) return 0;
    return (size_t)bn_chunks(P, C) * 2 * C * sizeof(float);
}

// x, residual (or NULL), out: [P,C] NHWC-flattened (P = N*H*W) in `dtype`; gamma, beta, running_*,
// save_mean, save_rstd: fp32 [C].  running_* may be NULL (track_running_stats=False).
extern "C" int ge_bn_fwd_train(const void* x, const void* residual, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps,
                               void* out, float* save_mean, float* save_rstd, void* workspace, size_t workspace_bytes,
                               int dtype, long long P, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(x && gamma && beta && out && save_mean && save_rstd && workspace, GE_ERR_ARG, "ge_bn_fwd_train: null pointer");
    GE_REQUIRE(P > 0 && C > 0, GE_ERR_ARG, "ge_bn_fwd_train: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_fwd_train: unsupported channel count C=%d (C%%8==0, C/8 | 256)", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_fwd_train: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = bn_chunks(P, C);
    const long long ppc = ge::cdivll(P, chunks);
    const size_t smem = bn_smem(C);
    float* part = static_cast<float*>(workspace);
    // shift = running mean when tracked (close to the batch mean), else beta-free zero shift via gamma-less trick
    const float* shift = running_mean != nullptr ? running_mean : save_mean;
    if (running_mean == nullptr) GE_CUDA(cudaMemsetAsync(save_mean, 0, (size_t)C * sizeof(float), st), "ge_bn_fwd_train(memset)");
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_fwd_train(attr)"); c0 = smem; }
        bn_partial_stats_kernel<float><<<chunks, BN_THREADS, smem, st>>>((const float*)x, shift, part, P, C, ppc);
    } else if (dtype == GE_DTYPE_BF16) {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_fwd_train(attr)"); c1 = smem; }
        bn_partial_stats_kernel<bf16><<<chunks, BN_THREADS, smem, st>>>((const bf16*)x, shift, part, P, C, ppc);
    } else { ge_set_error("ge_bn_fwd_train: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_bn_fwd_train(stats)");
    bn_finalize_stats_kernel<<<ge::cdiv(C, 32), 32 * FL, 0, st>>>(part, chunks, shift, save_mean, save_rstd,
                                                               running_mean, running_var, P, C, eps, momentum);
    GE_CHECK_LAUNCH("ge_bn_fwd_train(finalize)");
    const unsigned blocks4 = (unsigned)ge::cdivll(ge::cdivll(P, FPIX4) * (C / 4), 256);
    if (dtype == GE_DTYPE_F32)
        bn_apply_fwd4_kernel<float, 0><<<blocks4, 256, 0, st>>>((const float*)x, (const float*)residual, save_mean, save_rstd,
                                                                 gamma, beta, (float*)out, P, C, eps, relu);
    else
        bn_apply_fwd4_kernel<bf16, 0><<<blocks4, 256, 0, st>>>((const bf16*)x, (const bf16*)residual, save_mean, save_rstd,
                                                                gamma, beta, (bf16*)out, P, C, eps, relu);
    GE_CHECK_LAUNCH("ge_bn_fwd_train(apply)");
    return GE_OK;
}

extern "C" int ge_bn_fwd_eval(const void* x, const void* residual, const float* gamma, const float* beta,
                              const float* running_mean, const float* running_var, float eps, void* out,
                              int dtype, long long P, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(x && gamma && beta && running_mean && running_var && out, GE_ERR_ARG, "ge_bn_fwd_eval: null pointer");
    GE_REQUIRE(P > 0 && C > 0, GE_ERR_ARG, "ge_bn_fwd_eval: bad dimension");
    GE_REQUIRE(C % 8 == 0, GE_ERR_SHAPE, "ge_bn_fwd_eval: C=%d must be a multiple of 8", C);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks4 = (unsigned)ge::cdivll(ge::cdivll(P, FPIX4) * (C / 4), 256);
    if (dtype == GE_DTYPE_F32)
        bn_apply_fwd4_kernel<float, 1><<<blocks4, 256, 0, st>>>((const float*)x, (const float*)residual, running_mean, running_var,
                                                                 gamma, beta, (float*)out, P, C, eps, relu);
    else if (dtype == GE_DTYPE_BF16)
        bn_apply_fwd4_kernel<bf16, 1><<<blocks4, 256, 0, st>>>((const bf16*)x, (const bf16*)residual, running_mean, running_var,
                                                                gamma, beta, (bf16*)out, P, C, eps, relu);
    else { ge_set_error("ge_bn_fwd_eval: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_bn_fwd_eval");
    return GE_OK;
}

// Training-mode backward: mean / rstd = save_mean / save_rstd of the forward.  dres (or NULL) receives the
// gradient of the residual input.  dgamma, dbeta fp32 [C] (overwritten).
extern "C" int ge_bn_bwd(const void* dy, const void* out, const void* x, const float* gamma,
                         const float* mean, const float* rstd_or_var, float eps, void* dx, void* dres,
                         float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                         int dtype, long long P, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(dy && x && gamma && mean && rstd_or_var && dx && dgamma && dbeta && workspace, GE_ERR_ARG, "ge_bn_bwd: null pointer");
    GE_REQUIRE(!relu || out, GE_ERR_ARG, "ge_bn_bwd: the forward output is needed for the ReLU mask");
    GE_REQUIRE(P > 0 && C > 0, GE_ERR_ARG, "ge_bn_bwd: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_bwd: unsupported channel count C=%d", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_bwd: workspace too small");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_bwd: unsupported dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = bn_chunks(P, C);
    const long long ppc = ge::cdivll(P, chunks);
    const size_t smem = bn_smem(C);
    float* part = static_cast<float*>(workspace);
    static size_t c0 = 0, c1 = 0;
    const float* rstd = rstd_or_var;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(bn_partial_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_bwd(attr)"); c0 = smem; }
    } else {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(bn_partial_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_bwd(attr)"); c1 = smem; }
    }
    if (dtype == GE_DTYPE_F32)
        bn_partial_bwd_kernel<float><<<chunks, BN_THREADS, smem, st>>>((const float*)dy, (const float*)out, (const float*)x,
                                                                       mean, rstd, part, P, C, ppc, relu);
    else
        bn_partial_bwd_kernel<bf16><<<chunks, BN_THREADS, smem, st>>>((const bf16*)dy, (const bf16*)out, (const bf16*)x,
                                                                      mean, rstd, part, P, C, ppc, relu);
    GE_CHECK_LAUNCH("ge_bn_bwd(partial)");
    bn_finalize_bwd_kernel<<<ge::cdiv(C, 32), 32 * FL, 0, st>>>(part, chunks, dgamma, dbeta, C);
    GE_CHECK_LAUNCH("ge_bn_bwd(finalize)");
    const float invP = 1.f / (float)P;
    const long long total4 = ge::cdivll(P, BPIX4) * (C / 4);
    const unsigned blocks4 = (unsigned)ge::cdivll(total4, 256);
    if (dtype == GE_DTYPE_F32)
        bn_apply_bwd4_kernel<float><<<blocks4, 256, 0, st>>>((const float*)dy, (const float*)out, (const float*)x, mean, rstd,
            gamma, dgamma, dbeta, (float*)dx, (float*)dres, P, C, invP, relu);
    else
        bn_apply_bwd4_kernel<bf16><<<blocks4, 256, 0, st>>>((const bf16*)dy, (const bf16*)out, (const bf16*)x, mean, rstd,
            gamma, dgamma, dbeta, (bf16*)dx, (bf16*)dres, P, C, invP, relu);
    GE_CHECK_LAUNCH("ge_bn_bwd(apply)");
    return GE_OK;
}
