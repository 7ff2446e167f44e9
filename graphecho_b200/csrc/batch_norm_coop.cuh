// Single-launch (cooperative) fused BatchNorm for maps whose working set fits the chip's shared memory.
//
// The three-kernel path of batch_norm.cu reads x twice (statistics, then apply) and is launch-latency-bound on the
// small maps of ResNet layers 2-4 (15-30 us per call in a CUDA graph against 2-8 us of compulsory traffic).  Here one
// persistent CTA per SM loads its slice of x into shared memory ONCE (148 x ~195 KB = 28.8 MB of on-chip capacity),
// reduces its partial statistics, and after two grid barriers (partials -> finalize by a few CTAs -> everyone)
// normalises straight out of shared memory: x is read from HBM exactly once, `out` written once -- the compulsory
// traffic -- in one launch.  The backward keeps x and the masked dy on chip between the reduction of (sum dy, sum
// dy*xhat) and the computation of dx; when only one of the two fits, x stays and dy is re-read (from L2 at these sizes).
// Semantics (segments, running statistics, 1-bit ReLU mask layout) are those of batch_norm.cu.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace ge_bn_coop {

namespace cg = cooperative_groups;
using namespace ge;

constexpr int CT = 512;                 // threads per CTA
constexpr int kSliceCap = 192 * 1024;   // bytes of shared memory for the resident slice(s)

// pixel ranges per CTA: CTAs [0,g0) share segment 0 = [0,P0), CTAs [g0,G) share segment 1 = [P0,P); every range starts
// at a multiple of 4 pixels (the ReLU mask packs 4 pixels x 4 channels per word; P0 % 4 == 0 is required when two
// segments exist).
struct Plan {
    long long P, P0;
    int G, g0;
    long long pp0, pp1;                 // pixels per CTA in segment 0 / 1 (multiples of 4)
    __host__ __device__ __forceinline__ void range(int cta, long long& a, long long& b, int& seg) const {
        if (cta < g0) {
            seg = 0;
            a = (long long)cta * pp0;
            b = a + pp0 < P0 ? a + pp0 : P0;
        } else {
            seg = 1;
            a = P0 + (long long)(cta - g0) * pp1;
            b = a + pp1 < P ? a + pp1 : P;
        }
        if (a > b) a = b;
    }
    __host__ __device__ __forceinline__ int nseg() const { return P0 < P ? 2 : 1; }
    __host__ __device__ __forceinline__ long long max_pixels() const { return pp0 > pp1 ? pp0 : pp1; }
};

inline Plan make_plan(long long P, long long P_split, int G) {
    Plan pl;
    pl.P = P;
    pl.P0 = (P_split > 0 && P_split < P) ? P_split : P;
    pl.G = G;
    auto round4 = [](long long v) { return (v + 3) / 4 * 4; };
    if (pl.P0 < P) {
        long long g0 = (G * pl.P0 + P / 2) / P;
        if (g0 < 1) g0 = 1;
        if (g0 > G - 1) g0 = G - 1;
        pl.g0 = (int)g0;
        pl.pp0 = round4(cdivll(pl.P0, pl.g0));
        pl.pp1 = round4(cdivll(P - pl.P0, G - pl.g0));
    } else {
        pl.g0 = G;
        pl.pp0 = round4(cdivll(P, G));
        pl.pp1 = 4;
    }
    return pl;
}

template <typename T> struct VecT;                       // 16-byte vector of T
template <> struct VecT<__nv_bfloat16> { static constexpr int N = 8; };
template <> struct VecT<float> { static constexpr int N = 4; };

template <typename T, int N> __device__ __forceinline__ void ldv(const T* p, float (&f)[N]);
template <> __device__ __forceinline__ void ldv<__nv_bfloat16, 8>(const __nv_bfloat16* p, float (&f)[8]) { load8<__nv_bfloat16>(p, f); }
template <> __device__ __forceinline__ void ldv<float, 4>(const float* p, float (&f)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
template <typename T, int N> __device__ __forceinline__ void stv(T* p, const float (&f)[N]);
template <> __device__ __forceinline__ void stv<__nv_bfloat16, 8>(__nv_bfloat16* p, const float (&f)[8]) { store8<__nv_bfloat16>(p, f); }
template <> __device__ __forceinline__ void stv<float, 4>(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}

// Reduce the per-CTA partials of segment `seg` for channel c: sum over its CTAs of part[cta][which][c].
// Called by the finalize CTAs: 32 channels x 16 part-lanes per CTA of 512 threads.
__device__ __forceinline__ void reduce_cta_parts(const float* __restrict__ part, int first, int n, int C, int c, int lane_p,
                                                 float (*sh)[2][33], float& sa, float& sb) {
    float a = 0.f, b = 0.f;
    if (c < C)
        for (int q = first + lane_p; q < first + n; q += 16) {
            a += part[(size_t)q * 2 * C + c];
            b += part[(size_t)q * 2 * C + C + c];
        }
    __syncthreads();
    sh[lane_p][0][threadIdx.x & 31] = a;
    sh[lane_p][1][threadIdx.x & 31] = b;
    __syncthreads();
    sa = 0.f; sb = 0.f;
    if (lane_p == 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) { sa += sh[q][0][threadIdx.x & 31]; sb += sh[q][1][threadIdx.x & 31]; }
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T>
__global__ void __launch_bounds__(CT, 1)
bn_fwd_coop_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ nbt, float momentum, float eps, T* __restrict__ out,
                   float* __restrict__ save_mean, float* __restrict__ save_rstd, unsigned short* __restrict__ mask,
                   float* __restrict__ part, Plan pl, int C, int relu) {
    constexpr int V = VecT<T>::N;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ float fsh[16][2][33];
    cg::grid_group grid = cg::this_grid();
    const int cvecs = C / V, nPL = CT / cvecs;
    const int tid = threadIdx.x, cv = tid % cvecs, plane = tid / cvecs, cc = cv * V;
    T* xs = reinterpret_cast<T*>(smraw);                                     // [pixels][C]
    long long p0, p1;
    int seg;
    pl.range(blockIdx.x, p0, p1, seg);
    const int npix = (int)(p1 - p0);
    float* red = reinterpret_cast<float*>(smraw + (size_t)pl.max_pixels() * C * sizeof(T));   // [nPL][C] x 2

    // ---- phase 1: slice -> shared memory, partial sums of (x - shift), (x - shift)^2
    const float* shift_src = running_mean != nullptr ? running_mean : nullptr;
    float shift[V], a[V], b[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { shift[u] = shift_src ? shift_src[cc + u] : 0.f; a[u] = 0.f; b[u] = 0.f; }
#pragma unroll 4
    for (int p = plane; p < npix; p += nPL) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + (p0 + p) * C + cc);
        *reinterpret_cast<uint4*>(xs + (size_t)p * C + cc) = raw;
        float v[V];
        ldv<T, V>(reinterpret_cast<const T*>(&raw), v);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const float d = v[u] - shift[u];
            a[u] += d;
            b[u] = fmaf(d, d, b[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < V; ++u) { red[plane * C + cc + u] = a[u]; red[(size_t)nPL * C + plane * C + cc + u] = b[u]; }
    __syncthreads();
    for (int c = tid; c < C; c += CT) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += red[q * C + c]; sb += red[(size_t)nPL * C + q * C + c]; }
        part[(size_t)blockIdx.x * 2 * C + c] = sa;
        part[(size_t)blockIdx.x * 2 * C + C + c] = sb;
    }
    grid.sync();

    // ---- phase 2: CTAs [0, C/32) finalize 32 channels each (both segments, running statistics in segment order)
    if (blockIdx.x * 32 < C) {
        const int c = blockIdx.x * 32 + (tid & 31), lane_p = tid >> 5;
        const bool owner = lane_p == 0 && c < C;
        const int nseg = pl.nseg();
        const float sft = (c < C && shift_src) ? shift_src[c] : 0.f;
        float rm = 0.f, rv = 0.f;
        if (owner && running_mean != nullptr) { rm = running_mean[c]; rv = running_var[c]; }
        for (int s = 0; s < nseg; ++s) {
            const int first = s == 0 ? 0 : pl.g0, n = s == 0 ? pl.g0 : pl.G - pl.g0;
            const long long Ps = s == 0 ? pl.P0 : pl.P - pl.P0;
            float sa, sb;
            reduce_cta_parts(part, first, n, C, c, lane_p, fsh, sa, sb);
            if (owner) {
                const float inv = 1.f / (float)Ps;
                const float md = sa * inv;
                const float var = fmaxf(sb * inv - md * md, 0.f);
                const float mean = md + sft;
                save_mean[s * C + c] = mean;
                save_rstd[s * C + c] = 1.f / sqrtf(var + eps);
                const float unbiased = Ps > 1 ? var * ((float)Ps / (float)(Ps - 1)) : var;
                rm = (1.f - momentum) * rm + momentum * mean;
                rv = (1.f - momentum) * rv + momentum * unbiased;
            }
        }
        if (owner && running_mean != nullptr) { running_mean[c] = rm; running_var[c] = rv; }
        if (nbt != nullptr && blockIdx.x == 0 && tid == 0) *nbt += nseg;
    }
    grid.sync();

    // ---- phase 3: normalise out of shared memory; a thread owns V channels of 4 consecutive pixels per step
    float sc[V], sh[V];
#pragma unroll
    for (int u = 0; u < V; ++u) {
        const float r = save_rstd[seg * C + cc + u], m = save_mean[seg * C + cc + u];
        sc[u] = r * gamma[cc + u];
        sh[u] = beta[cc + u] - m * sc[u];
    }
    const int c4 = C >> 2;
    const int ngrp = (npix + 3) >> 2;
    for (int g = plane; g < ngrp; g += nPL) {
        unsigned bits[V / 4];
#pragma unroll
        for (int w = 0; w < V / 4; ++w) bits[w] = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = g * 4 + q;
            if (p >= npix) break;
            float v[V], o[V];
            ldv<T, V>(xs + (size_t)p * C + cc, v);
            float rf[V];
            if (res != nullptr) ldv<T, V>(res + (p0 + p) * C + cc, rf);
#pragma unroll
            for (int u = 0; u < V; ++u) {
                o[u] = fmaf(v[u], sc[u], sh[u]);
                if (res != nullptr) o[u] += rf[u];
                if (relu) {
                    if (o[u] > 0.f) bits[u >> 2] |= 1u << (q * 4 + (u & 3));
                    o[u] = fmaxf(o[u], 0.f);
                }
            }
            stv<T, V>(out + (p0 + p) * C + cc, o);
        }
        if (mask != nullptr) {
            unsigned short* mw = mask + ((p0 >> 2) + g) * c4 + (cc >> 2);
#pragma unroll
            for (int w = 0; w < V / 4; ++w) mw[w] = (unsigned short)bits[w];
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// KEEP_DY: the masked dy slice stays in shared memory next to x (working set 2 maps); otherwise dy (+ mask) is re-read.
template <typename T, bool KEEP_DY>
__global__ void __launch_bounds__(CT, 1)
bn_bwd_coop_kernel(const T* __restrict__ dy, const unsigned short* __restrict__ mask, const T* __restrict__ x,
                   const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                   T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ part, float* __restrict__ seg_sums, Plan pl, int C) {
    constexpr int V = VecT<T>::N;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ float fsh[16][2][33];
    cg::grid_group grid = cg::this_grid();
    const int cvecs = C / V, nPL = CT / cvecs, c4 = C >> 2;
    const int tid = threadIdx.x, cv = tid % cvecs, plane = tid / cvecs, cc = cv * V;
    const size_t slice_elems = (size_t)pl.max_pixels() * C;
    T* xs = reinterpret_cast<T*>(smraw);
    T* gs = xs + slice_elems;                                                  // only when KEEP_DY
    float* red = reinterpret_cast<float*>(smraw + slice_elems * sizeof(T) * (KEEP_DY ? 2 : 1));
    long long p0, p1;
    int seg;
    pl.range(blockIdx.x, p0, p1, seg);
    const int npix = (int)(p1 - p0);
    const int ngrp = (npix + 3) >> 2;

    float m[V], r[V], a[V], b[V];
#pragma unroll
    for (int u = 0; u < V; ++u) { m[u] = mean[seg * C + cc + u]; r[u] = rstd[seg * C + cc + u]; a[u] = 0.f; b[u] = 0.f; }
    // ---- phase 1: masked dy and x -> shared memory; S1 = sum dyr, S2 = sum dyr * xhat
    for (int g = plane; g < ngrp; g += nPL) {
        unsigned bits[V / 4];
#pragma unroll
        for (int w = 0; w < V / 4; ++w)
            bits[w] = mask != nullptr ? (unsigned)mask[((p0 >> 2) + g) * c4 + (cc >> 2) + w] : 0xFFFFu;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = g * 4 + q;
            if (p >= npix) break;
            float gv[V], xv[V];
            ldv<T, V>(dy + (p0 + p) * C + cc, gv);
            const uint4 raw = *reinterpret_cast<const uint4*>(x + (p0 + p) * C + cc);
            *reinterpret_cast<uint4*>(xs + (size_t)p * C + cc) = raw;
            ldv<T, V>(reinterpret_cast<const T*>(&raw), xv);
#pragma unroll
            for (int u = 0; u < V; ++u) {
                gv[u] = (bits[u >> 2] >> (q * 4 + (u & 3))) & 1u ? gv[u] : 0.f;
                a[u] += gv[u];
                b[u] = fmaf(gv[u], (xv[u] - m[u]) * r[u], b[u]);
            }
            if (KEEP_DY) stv<T, V>(gs + (size_t)p * C + cc, gv);
        }
    }
#pragma unroll
    for (int u = 0; u < V; ++u) { red[plane * C + cc + u] = a[u]; red[(size_t)nPL * C + plane * C + cc + u] = b[u]; }
    __syncthreads();
    for (int c = tid; c < C; c += CT) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += red[q * C + c]; sb += red[(size_t)nPL * C + q * C + c]; }
        part[(size_t)blockIdx.x * 2 * C + c] = sa;
        part[(size_t)blockIdx.x * 2 * C + C + c] = sb;
    }
    grid.sync();

    // ---- phase 2: per-segment sums, dgamma, dbeta
    const int nseg = pl.nseg();
    if (blockIdx.x * 32 < C) {
        const int c = blockIdx.x * 32 + (tid & 31), lane_p = tid >> 5;
        const bool owner = lane_p == 0 && c < C;
        float ta = 0.f, tb = 0.f;
        for (int s = 0; s < nseg; ++s) {
            const int first = s == 0 ? 0 : pl.g0, n = s == 0 ? pl.g0 : pl.G - pl.g0;
            float sa, sb;
            reduce_cta_parts(part, first, n, C, c, lane_p, fsh, sa, sb);
            if (owner) {
                seg_sums[s * C + c] = sa;
                seg_sums[(nseg + s) * C + c] = sb;
                ta += sa; tb += sb;
            }
        }
        if (owner) { dbeta[c] = ta; dgamma[c] = tb; }
    }
    grid.sync();

    // ---- phase 3: dx = k0*dyr - k1 - (x - m)*k2
    float k0[V], k1[V], k2[V];
    const float invP = 1.f / (float)(seg == 0 ? pl.P0 : pl.P - pl.P0);
#pragma unroll
    for (int u = 0; u < V; ++u) {
        k0[u] = r[u] * gamma[cc + u];
        k1[u] = k0[u] * seg_sums[seg * C + cc + u] * invP;
        k2[u] = k0[u] * seg_sums[(nseg + seg) * C + cc + u] * invP * r[u];
    }
    for (int g = plane; g < ngrp; g += nPL) {
        unsigned bits[V / 4];
        if (!KEEP_DY) {
#pragma unroll
            for (int w = 0; w < V / 4; ++w)
                bits[w] = mask != nullptr ? (unsigned)mask[((p0 >> 2) + g) * c4 + (cc >> 2) + w] : 0xFFFFu;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = g * 4 + q;
            if (p >= npix) break;
            float gv[V], xv[V], o[V];
            ldv<T, V>(xs + (size_t)p * C + cc, xv);
            if (KEEP_DY) {
                ldv<T, V>(gs + (size_t)p * C + cc, gv);
            } else {
                ldv<T, V>(dy + (p0 + p) * C + cc, gv);
#pragma unroll
                for (int u = 0; u < V; ++u) gv[u] = (bits[u >> 2] >> (q * 4 + (u & 3))) & 1u ? gv[u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < V; ++u) o[u] = k0[u] * gv[u] - k1[u] - (xv[u] - m[u]) * k2[u];
            if (dres != nullptr) stv<T, V>(dres + (p0 + p) * C + cc, gv);
            stv<T, V>(dx + (p0 + p) * C + cc, o);
        }
    }
}

// Host-side eligibility: C a power-of-two-ish multiple of the vector width that divides the CTA, segment boundary on a
// mask word, and the slice(s) within the shared-memory budget.  Returns the number of resident maps that fit
// (0 = not eligible, 1 = x only, 2 = x and dy).
inline int coop_fit(long long P, long long P_split, int C, int es, int G, Plan* out) {
    const int V = 16 / es;
    if (C % V != 0) return 0;
    const int cvecs = C / V;
    if (cvecs > CT || CT % cvecs != 0) return 0;
    if (P_split > 0 && P_split < P && (P_split % 4) != 0) return 0;
    if (G < 2 || P < 4LL * G) return 0;
    const Plan pl = make_plan(P, P_split, G);
    const size_t slice = (size_t)pl.max_pixels() * C * es;
    if (out) *out = pl;
    if (2 * slice <= (size_t)kSliceCap) return 2;
    if (slice <= (size_t)kSliceCap) return 1;
    return 0;
}

inline size_t red_bytes(int C, int es) {
    const int V = 16 / es;
    return (size_t)2 * (CT / (C / V)) * C * sizeof(float);
}

}  // namespace ge_bn_coop
