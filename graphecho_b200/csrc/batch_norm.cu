// Fused BatchNorm2d (+ residual add) (+ ReLU) on NHWC feature maps, training and inference, forward and
// backward — the glue between the backbone's tensor-core convolutions.
//
// The reference runs nn.BatchNorm2d, the residual `out += identity` and nn.ReLU as separate full-tensor
// passes (/root/reference/models/fpnseg.py:192-212 Bottleneck.forward, :251-255 stem, :27-142 VGG16
// blocks): 8 tensor passes forward and 8 backward per residual block output.  Here:
//   forward : partial statistics (1 read) -> finalize (+ running-stat update) -> apply (1-2 reads, 1 write)
//   backward: partial sums of dy*relu' and dy*relu'*xhat (2 reads) -> finalize (dgamma, dbeta) ->
//             apply (2 reads, 1-2 writes)
// Statistics are reduced in two deterministic stages (per-CTA partials in a workspace, then one small
// CTA), never with atomics.  The statistics kernels move 8 consecutive channels per thread (one 128-bit access in
// bf16), the apply kernels 4 (half the registers, twice the occupancy).
// HBM-bound.  Compulsory bytes per element (bf16): forward 4 (+2 with a residual), backward 8 (+2 with a residual
// gradient); the kernels move forward 6 (+2) and backward 10.25 (+2): the ReLU mask is a 1-bit-per-element side
// output of the forward, so the backward never re-reads `out`.
//
// SEGMENTS (per-domain statistics): the reference trainer runs the network on the source batch and on the target
// batch in two separate train-mode calls (train_cardiac_uda.py:225, 234), so every BatchNorm normalises each domain
// with its own batch statistics and updates its running statistics twice per step.  Here both domains travel as ONE
// [source | target] batch through the convolutions; `P_split` (pixels of the first segment; 0 = one segment) makes
// the statistics, the running-stat updates (source first, then target) and the backward sums per segment.
#include "common.cuh"
#include "batch_norm_coop.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using bf16 = __nv_bfloat16;
using namespace ge;

constexpr int BN_THREADS = 256;

// How the pixel range [0,P) is cut into per-CTA chunks that never straddle the segment boundary: CTAs [0,chunks0)
// cover segment 0 = [0,P0), CTAs [chunks0,chunks) cover segment 1 = [P0,P).  One segment: P0 = P, chunks0 = chunks.
struct SegPlan {
    long long P, P0;
    int chunks, chunks0;
    long long ppc0, ppc1;
    __host__ __device__ __forceinline__ void range(int cta, long long& p0, long long& p1) const {
        if (cta < chunks0) {
            p0 = (long long)cta * ppc0;
            p1 = p0 + ppc0 < P0 ? p0 + ppc0 : P0;
        } else {
            p0 = P0 + (long long)(cta - chunks0) * ppc1;
            p1 = p0 + ppc1 < P ? p0 + ppc1 : P;
        }
    }
    __host__ __device__ __forceinline__ int nseg() const { return P0 < P ? 2 : 1; }
};

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) { load8<float>(p, f); }

// ---- stage 1 (forward): per-CTA partial sums of (x - shift) and (x - shift)^2 --------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_partial_stats_kernel(const T* __restrict__ x, const float* __restrict__ shift_src, float* __restrict__ part,
                        SegPlan sp, int C) {
    extern __shared__ __align__(16) float sm[];       // [nPL][C] x 2
    const int c8 = C >> 3, nPL = BN_THREADS / c8;
    float* ssum = sm;
    float* ssq = sm + (size_t)nPL * C;
    const int tid = threadIdx.x, co = tid % c8, pl = tid / c8;
    long long p0, p1;
    sp.range(blockIdx.x, p0, p1);
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

// Sum the per-CTA partials of FCH channels with FLN part-lanes (CTA = FCH*FLN threads, grid = ceil(C/FCH)).  The
// reduction is latency-bound (L2 round trips), so it is spread over many small CTAs (C/8 of them: 32 for C = 256) and
// every thread issues all of its (<= 5 x 2) loads before the first add.  Every thread of the CTA must call; the result
// is valid in the threads with lane_p == 0.
constexpr int FCH = 8, FLN = 64, FROWS = 5;      // up to FLN*FROWS = 320 partial rows per segment without a second trip

__device__ __forceinline__ void reduce_parts(const float* __restrict__ part, int nparts, int C, int c, int lane_p,
                                             float (*sh)[2][FCH], float& sa, float& sb) {
    float a = 0.f, b = 0.f;
    if (c < C) {
        for (int q0 = lane_p; q0 < nparts; q0 += FLN * FROWS) {
            float va[FROWS], vb[FROWS];
#pragma unroll
            for (int u = 0; u < FROWS; ++u) {
                const int q = q0 + FLN * u;
                va[u] = q < nparts ? part[(size_t)q * 2 * C + c] : 0.f;
                vb[u] = q < nparts ? part[(size_t)q * 2 * C + C + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < FROWS; ++u) { a += va[u]; b += vb[u]; }
        }
    }
    __syncthreads();                 // a previous call's readers are done with sh
    sh[lane_p][0][threadIdx.x % FCH] = a;
    sh[lane_p][1][threadIdx.x % FCH] = b;
    __syncthreads();
    sa = 0.f; sb = 0.f;
    if (lane_p == 0) {
#pragma unroll 8
        for (int q = 0; q < FLN; ++q) { sa += sh[q][0][threadIdx.x % FCH]; sb += sh[q][1][threadIdx.x % FCH]; }
    }
}

// ---- stage 2 (forward): batch mean / rstd per segment, running statistics (nn.BatchNorm2d semantics, one update
// per segment in segment order), num_batches_tracked += segments ------------------------------------------------
__global__ void __launch_bounds__(FCH * FLN)
bn_finalize_stats_kernel(const float* __restrict__ part, SegPlan sp, const float* __restrict__ shift_src,
                         float* __restrict__ save_mean, float* __restrict__ save_rstd,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         long long* __restrict__ num_batches_tracked, int C, float eps, float momentum,
                         int count_scale) {
    __shared__ float sh[FLN][2][FCH];
    const int c = blockIdx.x * FCH + (threadIdx.x % FCH), lane8 = threadIdx.x / FCH;
    const bool owner = lane8 == 0 && c < C;
    const int nseg = sp.nseg();
    const float shift = c < C ? shift_src[c] : 0.f;          // read before the running mean is updated below
    float rm = 0.f, rv = 0.f;
    if (owner && running_mean != nullptr) { rm = running_mean[c]; rv = running_var[c]; }
    for (int s = 0; s < nseg; ++s) {
        const int first = s == 0 ? 0 : sp.chunks0, n = s == 0 ? sp.chunks0 : sp.chunks - sp.chunks0;
        const long long Ps = s == 0 ? sp.P0 : sp.P - sp.P0;
        float sa, sb;
        reduce_parts(part + (size_t)first * 2 * C, n, C, c, lane8, sh, sa, sb);
        if (owner) {
            const float inv = 1.f / (float)Ps;
            const float md = sa * inv;
            const float var = fmaxf(sb * inv - md * md, 0.f);          // biased, used to normalise
            const float mean = md + shift;
            save_mean[s * C + c] = mean;
            save_rstd[s * C + c] = 1.f / sqrtf(var + eps);
            // count_scale > 1: the rows are AVERAGES over that many equal-sized ranks (SyncBatchNorm): the statistics are
            // those of Ps * count_scale pixels
            const long long Pt = Ps * count_scale;
            const float unbiased = Pt > 1 ? var * ((float)Pt / (float)(Pt - 1)) : var;
            rm = (1.f - momentum) * rm + momentum * mean;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        }
    }
    if (owner && running_mean != nullptr) { running_mean[c] = rm; running_var[c] = rv; }
    if (num_batches_tracked != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += nseg;
}

// ---- stage 3 (forward): out = act( x*sc + sh (+ residual) ) --------------------------------------
// MODE 0: training (sc/sh from save_mean/save_rstd of the pixel's segment); MODE 1: inference (running stats,
// `rstd` = var).  A thread owns FOUR channels (one 64-bit bf16 access) of FPIX4 consecutive pixels: the per-channel
// constants are loaded once and FPIX4 independent loads are in flight before the first use.
// ReLU mask: thread t (this exact mapping is shared with bn_apply_bwd4_kernel and bn_partial_bwd_kernel) writes one
// uint16 = bit (q*4 + u) set when out[pixel p0+q, channel cc+u] > 0.
constexpr int FPIX4 = 4;

template <typename T, int MODE>
__global__ void __launch_bounds__(256, 3)
bn_apply_fwd4_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ mean,
                     const float* __restrict__ rstd_or_var, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ out, unsigned short* __restrict__ mask,
                     long long P, long long P0, int C, float eps, int relu) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    const int cc = (int)(t % c4) * 4;
    const long long p0 = (t / c4) * FPIX4;
    if (p0 >= P) return;
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc), b4 = *reinterpret_cast<const float4*>(beta + cc);
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    float sc[4], sh[4];
    auto constants = [&](int seg) {
        const float4 m4 = *reinterpret_cast<const float4*>(mean + seg * C + cc);
        const float4 r4 = *reinterpret_cast<const float4*>(rstd_or_var + seg * C + cc);
        const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float rs = (MODE == 0) ? r[u] : 1.f / sqrtf(r[u] + eps);
            sc[u] = rs * g[u];
            sh[u] = b[u] - m[u] * sc[u];
        }
    };
    int seg = (MODE == 0 && p0 >= P0) ? 1 : 0;
    constants(seg);
    Vec4<T> v[FPIX4], rv[FPIX4];
#pragma unroll
    for (int q = 0; q < FPIX4; ++q)
        if (p0 + q < P) {
            v[q].load(x + (p0 + q) * C + cc);
            if (res != nullptr) rv[q].load(res + (p0 + q) * C + cc);
        }
    unsigned bits = 0;
#pragma unroll
    for (int q = 0; q < FPIX4; ++q) {
        if (p0 + q >= P) break;
        if (MODE == 0 && seg == 0 && p0 + q >= P0) { seg = 1; constants(1); }     // thread straddles the boundary (rare)
        float f[4], rf[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
        v[q].get(f);
        if (res != nullptr) rv[q].get(rf);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            o[u] = fmaf(f[u], sc[u], sh[u]);
            if (res != nullptr) o[u] += rf[u];
            if (relu) {
                if (o[u] > 0.f) bits |= 1u << (q * 4 + u);
                o[u] = fmaxf(o[u], 0.f);
            }
        }
        Vec4<T> w;
        w.set(o);
        w.store(out + (p0 + q) * C + cc);
    }
    if (mask != nullptr) mask[t] = (unsigned short)bits;
}

// ---- backward stage 1: partial sums of dyr = dy * relu'(out) and dyr * xhat ----------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_partial_bwd_kernel(const T* __restrict__ dy, const unsigned short* __restrict__ mask, const T* __restrict__ x,
                      const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ part,
                      SegPlan sp, int C) {
    extern __shared__ __align__(16) float sm[];
    const int c8 = C >> 3, nPL = BN_THREADS / c8, c4 = C >> 2;
    float* s1 = sm;
    float* s2 = sm + (size_t)nPL * C;
    const int tid = threadIdx.x, co = tid % c8, pl = tid / c8, cc = co * 8;
    long long p0, p1;
    sp.range(blockIdx.x, p0, p1);
    const int seg = blockIdx.x < sp.chunks0 ? 0 : 1;
    float m[8], r[8], a[8], b[8];
    load8f(mean + seg * C + cc, m); load8f(rstd + seg * C + cc, r);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = 0.f; b[u] = 0.f; }
    if (pl < nPL) {
#pragma unroll 2
        for (long long p = p0 + pl; p < p1; p += nPL) {
            float g[8], xv[8];
            load8<T>(dy + p * C + cc, g);
            load8<T>(x + p * C + cc, xv);
            if (mask != nullptr) {
                // the two uint16 words of the apply threads that own channels cc..cc+3 and cc+4..cc+7 of pixel p
                const unsigned w = *reinterpret_cast<const unsigned*>(mask + (p >> 2) * c4 + (cc >> 2));
                const unsigned sft = (unsigned)(p & 3) * 4;
                const unsigned lo = (w >> sft) & 0xFu, hi = (w >> (16 + sft)) & 0xFu;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    g[u] = (lo >> u) & 1u ? g[u] : 0.f;
                    g[4 + u] = (hi >> u) & 1u ? g[4 + u] : 0.f;
                }
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

// ---- backward stage 2: per segment S1 = sum dyr, S2 = sum dyr*xhat -> seg_sums [2][nseg][C]; dbeta = sum_s S1,
// dgamma = sum_s S2 ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FCH * FLN)
bn_finalize_bwd_kernel(const float* __restrict__ part, SegPlan sp, float* __restrict__ seg_sums,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
    __shared__ float sh[FLN][2][FCH];
    const int c = blockIdx.x * FCH + (threadIdx.x % FCH), lane8 = threadIdx.x / FCH;
    const bool owner = lane8 == 0 && c < C;
    const int nseg = sp.nseg();
    float ta = 0.f, tb = 0.f;
    for (int s = 0; s < nseg; ++s) {
        const int first = s == 0 ? 0 : sp.chunks0, n = s == 0 ? sp.chunks0 : sp.chunks - sp.chunks0;
        float sa, sb;
        reduce_parts(part + (size_t)first * 2 * C, n, C, c, lane8, sh, sa, sb);
        if (owner) {
            seg_sums[s * C + c] = sa;
            seg_sums[(nseg + s) * C + c] = sb;
            ta += sa; tb += sb;
        }
    }
    if (owner) { dbeta[c] = ta; dgamma[c] = tb; }
}

// ---- backward stage 3: dx = rstd*gamma*(dyr - S1/P - xhat*S2/P) with the segment's S1, S2, P;  dres = dyr -------
// FOUR channels per thread (one 64-bit bf16 access) of BPIX4 consecutive pixels; same thread <-> element mapping as
// bn_apply_fwd4_kernel (the ReLU mask word of thread t is mask[t]).
constexpr int BPIX4 = 4;
static_assert(BPIX4 == FPIX4, "the ReLU mask layout is shared by the forward and backward apply kernels");

template <typename T>
__global__ void __launch_bounds__(256, 3)
bn_apply_bwd4_kernel(const T* __restrict__ dy, const unsigned short* __restrict__ mask, const T* __restrict__ x,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ seg_sums,
                     T* __restrict__ dx, T* __restrict__ dres, long long P, long long P0, int C) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = C >> 2;
    const int cc = (int)(t % c4) * 4;
    const long long p0 = (t / c4) * BPIX4;
    if (p0 >= P) return;
    const int nseg = P0 < P ? 2 : 1;
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc);
    const float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float m[4], k0[4], k1[4], k2[4];
    auto constants = [&](int seg) {          // dx = k0*dyr - k1 - (x-m)*k2
        const float4 m4 = *reinterpret_cast<const float4*>(mean + seg * C + cc);
        const float4 r4 = *reinterpret_cast<const float4*>(rstd + seg * C + cc);
        const float4 b4 = *reinterpret_cast<const float4*>(seg_sums + seg * C + cc);
        const float4 q4 = *reinterpret_cast<const float4*>(seg_sums + (nseg + seg) * C + cc);
        const float r[4] = {r4.x, r4.y, r4.z, r4.w}, s1[4] = {b4.x, b4.y, b4.z, b4.w}, s2[4] = {q4.x, q4.y, q4.z, q4.w};
        const float invP = 1.f / (float)(seg == 0 ? P0 : P - P0);
        m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            k0[u] = r[u] * g[u];
            k1[u] = k0[u] * s1[u] * invP;
            k2[u] = k0[u] * s2[u] * invP * r[u];
        }
    };
    int seg = p0 >= P0 ? 1 : 0;
    constants(seg);
    Vec4<T> gv[BPIX4], xv[BPIX4];
#pragma unroll
    for (int q = 0; q < BPIX4; ++q)
        if (p0 + q < P) {
            gv[q].load(dy + (p0 + q) * C + cc);
            xv[q].load(x + (p0 + q) * C + cc);
        }
    const unsigned bits = mask != nullptr ? (unsigned)mask[t] : 0xFFFFu;
#pragma unroll
    for (int q = 0; q < BPIX4; ++q) {
        if (p0 + q >= P) break;
        if (seg == 0 && p0 + q >= P0) { seg = 1; constants(1); }
        float gf[4], xf[4], o[4];
        gv[q].get(gf); xv[q].get(xf);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float gq = (bits >> (q * 4 + u)) & 1u ? gf[u] : 0.f;
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
    const long long cap = (long long)ge::sm_count() * 4;       // 592 partial rows: two trips of the finalize lanes
    if (want > cap) want = cap;
    if (want < 2) want = 2;                              // room for one chunk per segment
    return (int)want;
}

SegPlan bn_plan(long long P, long long P_split, int C) {
    SegPlan sp;
    sp.P = P;
    sp.P0 = (P_split > 0 && P_split < P) ? P_split : P;
    sp.chunks = bn_chunks(P, C);
    if (sp.P0 < P) {
        long long c0 = (sp.chunks * sp.P0 + P / 2) / P;
        if (c0 < 1) c0 = 1;
        if (c0 > sp.chunks - 1) c0 = sp.chunks - 1;
        sp.chunks0 = (int)c0;
        sp.ppc0 = ge::cdivll(sp.P0, sp.chunks0);
        sp.ppc1 = ge::cdivll(P - sp.P0, sp.chunks - sp.chunks0);
    } else {
        sp.chunks0 = sp.chunks;
        sp.ppc0 = ge::cdivll(P, sp.chunks);
        sp.ppc1 = 1;
    }
    return sp;
}

bool bn_shape_ok(int C) { return C % 8 == 0 && C / 8 <= BN_THREADS && BN_THREADS % (C / 8) == 0; }

size_t bn_smem(int C) { return (size_t)2 * (BN_THREADS / (C / 8)) * C * sizeof(float); }

// 0 / 1 = three-kernel streaming path (default), 2 = single-launch cooperative kernels where the map fits on chip.
// The cooperative path is NOT the default: alone it is no faster (15-17 us vs 14-16 us at 4-6 MB, 29 vs 31 us at 26 MB:
// two grid barriers + an in-kernel finalize cost what two launches cost), and inside the training step it is slower
// (ge_bn_fwd_train 4.4 ms vs 2.6 ms per step): a cooperative grid needs every SM at once, so it cannot overlap the
// graph module's side-stream kernels or the tail of the previous convolution.  Kept for maps-on-chip experiments.
int g_bn_path = 0;

bool coop_supported() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess) v = 0;
        cached = v ? 1 : 0;
    }
    return cached == 1;
}

template <typename K>
int coop_launch(K kernel, int grid, size_t smem, void** args, cudaStream_t st, size_t* attr_cache, const char* name) {
    if (smem > *attr_cache) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        *attr_cache = smem;
    }
    GE_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(ge_bn_coop::CT), args, smem, st), name);
    ge_count_launches(1);
    return GE_OK;
}

// workspace rows: enough for the streaming path's chunks and for one partial per SM of the cooperative path
size_t bn_ws_rows(long long P, int C) {
    const size_t a = (size_t)bn_chunks(P, C), b = (size_t)ge::sm_count();
    return a > b ? a : b;
}

// rows -> per-segment sums in the row layout ([s][2][C]) the forward finalize reads (SyncBatchNorm exchange buffer)
__global__ void __launch_bounds__(FCH * FLN)
bn_rows_to_segments_kernel(const float* __restrict__ part, SegPlan sp, float* __restrict__ sums, int C) {
    __shared__ float sh[FLN][2][FCH];
    const int c = blockIdx.x * FCH + (threadIdx.x % FCH), lane8 = threadIdx.x / FCH;
    const bool owner = lane8 == 0 && c < C;
    const int nseg = sp.nseg();
    for (int s = 0; s < nseg; ++s) {
        const int first = s == 0 ? 0 : sp.chunks0, n = s == 0 ? sp.chunks0 : sp.chunks - sp.chunks0;
        float sa, sb;
        reduce_parts(part + (size_t)first * 2 * C, n, C, c, lane8, sh, sa, sb);
        if (owner) {
            sums[(size_t)s * 2 * C + c] = sa;
            sums[(size_t)s * 2 * C + C + c] = sb;
        }
    }
}

int bn_bwd_reduce_stage(const void* dy, const unsigned short* mk, const void* x, const float* mean, const float* rstd,
                        float* part, float* seg_sums, float* dgamma, float* dbeta, int dtype, long long P, long long P_split,
                        int C, cudaStream_t st) {
    const SegPlan sp = bn_plan(P, P_split, C);
    const size_t smem = bn_smem(C);
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(bn_partial_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_bwd(attr)"); c0 = smem; }
        bn_partial_bwd_kernel<float><<<sp.chunks, BN_THREADS, smem, st>>>((const float*)dy, mk, (const float*)x,
                                                                          mean, rstd, part, sp, C);
    } else {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(bn_partial_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_bwd(attr)"); c1 = smem; }
        bn_partial_bwd_kernel<bf16><<<sp.chunks, BN_THREADS, smem, st>>>((const bf16*)dy, mk, (const bf16*)x,
                                                                         mean, rstd, part, sp, C);
    }
    GE_CHECK_LAUNCH("ge_bn_bwd(partial)");
    bn_finalize_bwd_kernel<<<ge::cdiv(C, FCH), FCH * FLN, 0, st>>>(part, sp, seg_sums, dgamma, dbeta, C);
    GE_CHECK_LAUNCH("ge_bn_bwd(finalize)");
    return GE_OK;
}

int bn_bwd_apply_stage(const void* dy, const unsigned short* mk, const void* x, const float* gamma, const float* mean,
                       const float* rstd, const float* seg_sums, void* dx, void* dres, int dtype, long long P,
                       long long P_split, int C, cudaStream_t st) {
    const long long P0 = (P_split > 0 && P_split < P) ? P_split : P;
    const long long total4 = ge::cdivll(P, BPIX4) * (C / 4);
    const unsigned blocks4 = (unsigned)ge::cdivll(total4, 256);
    if (dtype == GE_DTYPE_F32)
        bn_apply_bwd4_kernel<float><<<blocks4, 256, 0, st>>>((const float*)dy, mk, (const float*)x, mean, rstd,
            gamma, seg_sums, (float*)dx, (float*)dres, P, P0, C);
    else
        bn_apply_bwd4_kernel<bf16><<<blocks4, 256, 0, st>>>((const bf16*)dy, mk, (const bf16*)x, mean, rstd,
            gamma, seg_sums, (bf16*)dx, (bf16*)dres, P, P0, C);
    GE_CHECK_LAUNCH("ge_bn_bwd(apply)");
    return GE_OK;
}

}  // namespace

extern "C" int ge_bn_set_path(int path) {
    GE_REQUIRE(path >= 0 && path <= 2, GE_ERR_ARG, "ge_bn_set_path: path must be 0/1 (streaming kernels) or 2 (cooperative where it fits)");
    g_bn_path = path;
    return GE_OK;
}

// partials [rows][2][C] followed by the backward's per-segment sums [2][2][C]
extern "C" size_t ge_bn_workspace_bytes(long long P, int C) {
    if (P <= 0 || C <= 0 || !bn_shape_ok(C)) return 0;
    return (bn_ws_rows(P, C) * 2 * C + 4 * (size_t)C) * sizeof(float);
}

extern "C" size_t ge_bn_relu_mask_bytes(long long P, int C) {
    if (P <= 0 || C <= 0 || C % 8 != 0) return 0;
    return (size_t)ge::cdivll(P, FPIX4) * (C / 4) * sizeof(unsigned short);
}

// x, residual (or NULL), out: [P,C] NHWC-flattened (P = N*H*W) in `dtype`; gamma, beta, running_* fp32 [C];
// save_mean, save_rstd fp32 [nseg][C] with nseg = 2 when 0 < P_split < P, else 1.  running_* may be NULL
// (track_running_stats=False); num_batches_tracked (int64, or NULL) is incremented by nseg on the device;
// relu_mask (ge_bn_relu_mask_bytes, or NULL) receives the 1-bit ReLU mask the backward consumes.
extern "C" int ge_bn_fwd_train(const void* x, const void* residual, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, long long* num_batches_tracked,
                               float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                               void* relu_mask, void* workspace, size_t workspace_bytes,
                               int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(x && gamma && beta && out && save_mean && save_rstd && workspace, GE_ERR_ARG, "ge_bn_fwd_train: null pointer");
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P, GE_ERR_ARG, "ge_bn_fwd_train: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_fwd_train: unsupported channel count C=%d (C%%8==0, C/8 | 256)", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_fwd_train: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_fwd_train: unsupported dtype %d", dtype);
    {   // single-launch path: the map fits the chip's shared memory
        const int es = dtype == GE_DTYPE_F32 ? 4 : 2;
        ge_bn_coop::Plan cp;
        // (not with a residual: its read in the apply phase has too little memory-level parallelism at one CTA per SM;
        //  measured 50 vs 34 us at [256,1024,7,7])
        if (g_bn_path == 2 && residual == nullptr && coop_supported() &&
            ge_bn_coop::coop_fit(P, P_split, C, es, ge::sm_count(), &cp) >= 1) {
            float* partc = static_cast<float*>(workspace);
            unsigned short* mkc = relu ? static_cast<unsigned short*>(relu_mask) : nullptr;
            const size_t smemc = (size_t)cp.max_pixels() * C * es + ge_bn_coop::red_bytes(C, es);
            void* args[] = {(void*)&x, (void*)&residual, (void*)&gamma, (void*)&beta, (void*)&running_mean, (void*)&running_var,
                            (void*)&num_batches_tracked, (void*)&momentum, (void*)&eps, (void*)&out, (void*)&save_mean,
                            (void*)&save_rstd, (void*)&mkc, (void*)&partc, (void*)&cp, (void*)&C, (void*)&relu};
            static size_t a0 = 0, a1 = 0;
            if (dtype == GE_DTYPE_F32)
                return coop_launch(ge_bn_coop::bn_fwd_coop_kernel<float>, cp.G, smemc, args, st, &a0, "ge_bn_fwd_train(coop)");
            return coop_launch(ge_bn_coop::bn_fwd_coop_kernel<bf16>, cp.G, smemc, args, st, &a1, "ge_bn_fwd_train(coop)");
        }
    }
    const SegPlan sp = bn_plan(P, P_split, C);
    const size_t smem = bn_smem(C);
    float* part = static_cast<float*>(workspace);
    // shift = running mean when tracked (close to the batch mean), else a zero shift
    const float* shift = running_mean != nullptr ? running_mean : save_mean;
    if (running_mean == nullptr) GE_CUDA(cudaMemsetAsync(save_mean, 0, (size_t)C * sizeof(float), st), "ge_bn_fwd_train(memset)");
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_fwd_train(attr)"); c0 = smem; }
        bn_partial_stats_kernel<float><<<sp.chunks, BN_THREADS, smem, st>>>((const float*)x, shift, part, sp, C);
    } else if (dtype == GE_DTYPE_BF16) {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_fwd_train(attr)"); c1 = smem; }
        bn_partial_stats_kernel<bf16><<<sp.chunks, BN_THREADS, smem, st>>>((const bf16*)x, shift, part, sp, C);
    } else { ge_set_error("ge_bn_fwd_train: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_bn_fwd_train(stats)");
    bn_finalize_stats_kernel<<<ge::cdiv(C, FCH), FCH * FLN, 0, st>>>(part, sp, shift, save_mean, save_rstd,
                                                               running_mean, running_var, num_batches_tracked, C, eps, momentum, 1);
    GE_CHECK_LAUNCH("ge_bn_fwd_train(finalize)");
    const unsigned blocks4 = (unsigned)ge::cdivll(ge::cdivll(P, FPIX4) * (C / 4), 256);
    unsigned short* mk = relu ? static_cast<unsigned short*>(relu_mask) : nullptr;
    if (dtype == GE_DTYPE_F32)
        bn_apply_fwd4_kernel<float, 0><<<blocks4, 256, 0, st>>>((const float*)x, (const float*)residual, save_mean, save_rstd,
                                                                 gamma, beta, (float*)out, mk, P, sp.P0, C, eps, relu);
    else
        bn_apply_fwd4_kernel<bf16, 0><<<blocks4, 256, 0, st>>>((const bf16*)x, (const bf16*)residual, save_mean, save_rstd,
                                                                gamma, beta, (bf16*)out, mk, P, sp.P0, C, eps, relu);
    GE_CHECK_LAUNCH("ge_bn_fwd_train(apply)");
    return GE_OK;
}

// Training-mode forward for an x whose per-CTA partial statistics already exist (written by the epilogue of
// ge_conv1x1_bn_stats, which produced x): finalize + apply only -- the statistics pass over x is gone.
// part fp32 [rows][2][C]: sums of (x - shift) and (x - shift)^2 with shift = running_mean as it is BEFORE this call
// (zero when running_mean is NULL); rows [0, rows_segment0) cover pixels [0, P_split), the rest [P_split, P).
static int bn_fwd_from_rows(const char* name, int count_scale, const void* x, const void* residual, const float* gamma, const float* beta,
                                       float* running_mean, float* running_var, long long* num_batches_tracked,
                                       float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                                       void* relu_mask, const float* part, int rows, int rows_segment0,
                                       int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(x && gamma && beta && out && save_mean && save_rstd && part, GE_ERR_ARG, "%s: null pointer", name);
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P && rows > 0, GE_ERR_ARG, "%s: bad dimension", name);
    GE_REQUIRE(C % 8 == 0, GE_ERR_SHAPE, "ge_bn_fwd_train_prestat/sync: C=%d must be a multiple of 8", C);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_fwd_train_prestat/sync: unsupported dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    SegPlan sp;
    sp.P = P;
    sp.P0 = (P_split > 0 && P_split < P) ? P_split : P;
    sp.chunks = rows;
    sp.chunks0 = sp.P0 < P ? rows_segment0 : rows;
    sp.ppc0 = sp.ppc1 = 0;
    GE_REQUIRE(sp.chunks0 >= 1 && (sp.P0 == P || sp.chunks0 < rows), GE_ERR_ARG, "%s: bad segment rows", name);
    const float* shift = running_mean != nullptr ? running_mean : save_mean;
    if (running_mean == nullptr) GE_CUDA(cudaMemsetAsync(save_mean, 0, (size_t)C * sizeof(float), st), "ge_bn_fwd_train_prestat(memset)");
    bn_finalize_stats_kernel<<<ge::cdiv(C, FCH), FCH * FLN, 0, st>>>(part, sp, shift, save_mean, save_rstd,
                                                               running_mean, running_var, num_batches_tracked, C, eps, momentum, count_scale);
    GE_CHECK_LAUNCH("ge_bn_fwd_train_prestat(finalize)");
    const unsigned blocks4 = (unsigned)ge::cdivll(ge::cdivll(P, FPIX4) * (C / 4), 256);
    unsigned short* mk = relu ? static_cast<unsigned short*>(relu_mask) : nullptr;
    if (dtype == GE_DTYPE_F32)
        bn_apply_fwd4_kernel<float, 0><<<blocks4, 256, 0, st>>>((const float*)x, (const float*)residual, save_mean, save_rstd,
                                                                 gamma, beta, (float*)out, mk, P, sp.P0, C, eps, relu);
    else
        bn_apply_fwd4_kernel<bf16, 0><<<blocks4, 256, 0, st>>>((const bf16*)x, (const bf16*)residual, save_mean, save_rstd,
                                                                gamma, beta, (bf16*)out, mk, P, sp.P0, C, eps, relu);
    GE_CHECK_LAUNCH("ge_bn_fwd_train_prestat(apply)");
    return GE_OK;
}

extern "C" int ge_bn_fwd_train_prestat(const void* x, const void* residual, const float* gamma, const float* beta,
                                       float* running_mean, float* running_var, long long* num_batches_tracked,
                                       float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                                       void* relu_mask, const float* part, int rows, int rows_segment0,
                                       int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    return bn_fwd_from_rows("ge_bn_fwd_train_prestat", 1, x, residual, gamma, beta, running_mean, running_var,
                            num_batches_tracked, momentum, eps, out, save_mean, save_rstd, relu_mask, part, rows,
                            rows_segment0, dtype, P, P_split, C, relu, stream);
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
                                                                 gamma, beta, (float*)out, nullptr, P, P, C, eps, relu);
    else if (dtype == GE_DTYPE_BF16)
        bn_apply_fwd4_kernel<bf16, 1><<<blocks4, 256, 0, st>>>((const bf16*)x, (const bf16*)residual, running_mean, running_var,
                                                                gamma, beta, (bf16*)out, nullptr, P, P, C, eps, relu);
    else { ge_set_error("ge_bn_fwd_eval: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_bn_fwd_eval");
    return GE_OK;
}

// Training-mode backward: mean / rstd = save_mean / save_rstd of the forward ([nseg][C]); relu_mask = the forward's
// mask (required when relu != 0).  dres (or NULL) receives the gradient of the residual input.  dgamma, dbeta fp32 [C]
// (overwritten).
extern "C" int ge_bn_bwd(const void* dy, const void* relu_mask, const void* x, const float* gamma,
                         const float* mean, const float* rstd, void* dx, void* dres,
                         float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                         int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, GE_ERR_ARG, "ge_bn_bwd: null pointer");
    GE_REQUIRE(!relu || relu_mask, GE_ERR_ARG, "ge_bn_bwd: the forward's ReLU mask is needed");
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P, GE_ERR_ARG, "ge_bn_bwd: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_bwd: unsupported channel count C=%d", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_bwd: workspace too small");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_bwd: unsupported dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = static_cast<float*>(workspace);
    float* seg_sums = part + bn_ws_rows(P, C) * 2 * C;
    const unsigned short* mk = relu ? static_cast<const unsigned short*>(relu_mask) : nullptr;
    {   // single-launch path
        const int es = dtype == GE_DTYPE_F32 ? 4 : 2;
        ge_bn_coop::Plan cp;
        const int fit = (g_bn_path == 2 && coop_supported()) ? ge_bn_coop::coop_fit(P, P_split, C, es, ge::sm_count(), &cp) : 0;
        if (fit >= 1) {
            const size_t smemc = (size_t)cp.max_pixels() * C * es * (fit == 2 ? 2 : 1) + ge_bn_coop::red_bytes(C, es);
            void* args[] = {(void*)&dy, (void*)&mk, (void*)&x, (void*)&gamma, (void*)&mean, (void*)&rstd, (void*)&dx, (void*)&dres,
                            (void*)&dgamma, (void*)&dbeta, (void*)&part, (void*)&seg_sums, (void*)&cp, (void*)&C};
            static size_t a[4] = {0, 0, 0, 0};
            if (dtype == GE_DTYPE_F32)
                return fit == 2 ? coop_launch(ge_bn_coop::bn_bwd_coop_kernel<float, true>, cp.G, smemc, args, st, &a[0], "ge_bn_bwd(coop)")
                                : coop_launch(ge_bn_coop::bn_bwd_coop_kernel<float, false>, cp.G, smemc, args, st, &a[1], "ge_bn_bwd(coop)");
            return fit == 2 ? coop_launch(ge_bn_coop::bn_bwd_coop_kernel<bf16, true>, cp.G, smemc, args, st, &a[2], "ge_bn_bwd(coop)")
                            : coop_launch(ge_bn_coop::bn_bwd_coop_kernel<bf16, false>, cp.G, smemc, args, st, &a[3], "ge_bn_bwd(coop)");
        }
    }
    if (int rc = bn_bwd_reduce_stage(dy, mk, x, mean, rstd, part, seg_sums, dgamma, dbeta, dtype, P, P_split, C, st)) return rc;
    return bn_bwd_apply_stage(dy, mk, x, gamma, mean, rstd, seg_sums, dx, dres, dtype, P, P_split, C, st);
}

// ---- SyncBatchNorm: the same kernels with the cross-rank exchange between the statistics and the apply stages ------
// The caller (functional._SyncBnAct) all-reduces the small per-segment sums with op=AVG between the two calls of each
// direction; every rank holds the same number of pixels per segment (the path shards evenly), so an average of the
// sums over ranks divided by the local pixel count is the global statistic.

// Stage 1 forward: sums fp32 [nseg][2][C] = per-segment sums of (x - shift) and (x - shift)^2 over THIS rank's pixels;
// shift = running_mean (identical on every rank) or NULL for a zero shift.  workspace: ge_bn_workspace_bytes(P, C).
extern "C" int ge_bn_sync_stats(const void* x, const float* shift, float* sums, void* workspace, size_t workspace_bytes,
                                int dtype, long long P, long long P_split, int C, ge_stream_t stream) {
    GE_REQUIRE(x && sums && workspace, GE_ERR_ARG, "ge_bn_sync_stats: null pointer");
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P, GE_ERR_ARG, "ge_bn_sync_stats: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_sync_stats: unsupported channel count C=%d (C%%8==0, C/8 | 256)", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_sync_stats: workspace too small");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_sync_stats: unsupported dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const SegPlan sp = bn_plan(P, P_split, C);
    const size_t smem = bn_smem(C);
    float* part = static_cast<float*>(workspace);
    float* zero = part + bn_ws_rows(P, C) * 2 * C;           // the [4][C] tail of the workspace
    if (shift == nullptr) {
        GE_CUDA(cudaMemsetAsync(zero, 0, (size_t)C * sizeof(float), st), "ge_bn_sync_stats(memset)");
        shift = zero;
    }
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_sync_stats(attr)"); c0 = smem; }
        bn_partial_stats_kernel<float><<<sp.chunks, BN_THREADS, smem, st>>>((const float*)x, shift, part, sp, C);
    } else {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(bn_partial_stats_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_bn_sync_stats(attr)"); c1 = smem; }
        bn_partial_stats_kernel<bf16><<<sp.chunks, BN_THREADS, smem, st>>>((const bf16*)x, shift, part, sp, C);
    }
    GE_CHECK_LAUNCH("ge_bn_sync_stats(partial)");
    bn_rows_to_segments_kernel<<<ge::cdiv(C, FCH), FCH * FLN, 0, st>>>(part, sp, sums, C);
    GE_CHECK_LAUNCH("ge_bn_sync_stats(segments)");
    return GE_OK;
}

// Stage 2 forward: sums_avg = the all-reduced (AVG over `world` ranks) output of ge_bn_sync_stats.  Everything else as
// ge_bn_fwd_train: save_mean / save_rstd [nseg][C], running statistics updated once per segment with the unbiased
// variance of P_segment * world pixels, num_batches_tracked += nseg, out, ReLU mask.
extern "C" int ge_bn_sync_fwd_apply(const void* x, const void* residual, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, long long* num_batches_tracked,
                                    float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                                    void* relu_mask, const float* sums_avg, int world,
                                    int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(world >= 1, GE_ERR_ARG, "ge_bn_sync_fwd_apply: world=%d", world);
    const int nseg = (P_split > 0 && P_split < P) ? 2 : 1;
    return bn_fwd_from_rows("ge_bn_sync_fwd_apply", world, x, residual, gamma, beta, running_mean, running_var,
                            num_batches_tracked, momentum, eps, out, save_mean, save_rstd, relu_mask, sums_avg, nseg, 1,
                            dtype, P, P_split, C, relu, stream);
}

// Stage 1 backward: seg_sums fp32 [2][nseg][C] = per-segment sums of dy*relu' and dy*relu'*xhat over this rank's pixels;
// dgamma / dbeta = this rank's parameter gradients (the gradient exchange averages them like every other parameter).
extern "C" int ge_bn_sync_bwd_reduce(const void* dy, const void* relu_mask, const void* x, const float* mean,
                                     const float* rstd, float* seg_sums, float* dgamma, float* dbeta,
                                     void* workspace, size_t workspace_bytes,
                                     int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(dy && x && mean && rstd && seg_sums && dgamma && dbeta && workspace, GE_ERR_ARG, "ge_bn_sync_bwd_reduce: null pointer");
    GE_REQUIRE(!relu || relu_mask, GE_ERR_ARG, "ge_bn_sync_bwd_reduce: the forward's ReLU mask is needed");
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P, GE_ERR_ARG, "ge_bn_sync_bwd_reduce: bad dimension");
    GE_REQUIRE(bn_shape_ok(C), GE_ERR_SHAPE, "ge_bn_sync_bwd_reduce: unsupported channel count C=%d", C);
    GE_REQUIRE(workspace_bytes >= ge_bn_workspace_bytes(P, C), GE_ERR_ARG, "ge_bn_sync_bwd_reduce: workspace too small");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_sync_bwd_reduce: unsupported dtype %d", dtype);
    const unsigned short* mk = relu ? static_cast<const unsigned short*>(relu_mask) : nullptr;
    return bn_bwd_reduce_stage(dy, mk, x, mean, rstd, static_cast<float*>(workspace), seg_sums, dgamma, dbeta, dtype, P, P_split,
                               C, (cudaStream_t)stream);
}

// Stage 2 backward: seg_sums_avg = the all-reduced (AVG) output of ge_bn_sync_bwd_reduce.
extern "C" int ge_bn_sync_bwd_apply(const void* dy, const void* relu_mask, const void* x, const float* gamma,
                                    const float* mean, const float* rstd, const float* seg_sums_avg, void* dx, void* dres,
                                    int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream) {
    GE_REQUIRE(dy && x && gamma && mean && rstd && seg_sums_avg && dx, GE_ERR_ARG, "ge_bn_sync_bwd_apply: null pointer");
    GE_REQUIRE(!relu || relu_mask, GE_ERR_ARG, "ge_bn_sync_bwd_apply: the forward's ReLU mask is needed");
    GE_REQUIRE(P > 0 && C > 0 && P_split >= 0 && P_split <= P, GE_ERR_ARG, "ge_bn_sync_bwd_apply: bad dimension");
    GE_REQUIRE(C % 4 == 0, GE_ERR_SHAPE, "ge_bn_sync_bwd_apply: unsupported channel count C=%d", C);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_bn_sync_bwd_apply: unsupported dtype %d", dtype);
    const unsigned short* mk = relu ? static_cast<const unsigned short*>(relu_mask) : nullptr;
    return bn_bwd_apply_stage(dy, mk, x, gamma, mean, rstd, seg_sums_avg, dx, dres, dtype, P, P_split, C, (cudaStream_t)stream);
}
