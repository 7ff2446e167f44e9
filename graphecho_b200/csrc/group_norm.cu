// K7(ii) — GroupNorm + ReLU (+ bilinear up-sampling) on NHWC feature maps, forward and backward.
//
// One family of kernels serves both users on the hot path:
//   * FPN semantic head: GroupNorm(C,C) [one channel per group] -> ReLU -> _upsample
//     (/root/reference/models/fpnseg.py:354-355, 428-442)
//   * Discriminator towers: GroupNorm(32,256) -> ReLU (fpnseg.py:455-466), where PyTorch eager under
//     autocast up-casts to fp32, runs a contiguous-NCHW kernel and casts back: ~1.6 GB of traffic per
//     layer at p2 against 0.3 GB here.
// Every thread moves 8 consecutive channels (one 128-bit access in bf16); statistics are reduced per
// sample by one CTA; the backward of a same-size call recomputes relu'(y)*dout on the fly instead of
// staging it.  HBM-bound: forward = 2 reads + 1 write of the map, backward = 3 reads + 1 write.
#include "common.cuh"
#include "bilinear.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using bf16 = __nv_bfloat16;
using namespace ge;

constexpr int GN_THREADS = 512;

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) { load8<float>(p, f); }

// sum over the `cpg` consecutive lanes of a warp that hold one group's channels
__device__ __forceinline__ float group_lane_sum(float v, int cpg) {
    for (int o = cpg >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- statistics: mean / rstd of every group, written expanded per channel [N,C] -----------------
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
group_stats_kernel(const T* __restrict__ x, const float* __restrict__ pre_bias, float* __restrict__ mean,
                   float* __restrict__ rstd, float* __restrict__ chan_sum, int HW, int C, float eps, int cpg) {
    extern __shared__ __align__(16) float sm[];      // [nPL][C] sums, [nPL][C] squares, [C] shift
    const int c8 = C >> 3, nPL = GN_THREADS / c8;
    float* ssum = sm;
    float* ssq = sm + (size_t)nPL * C;
    float* sshift = ssq + (size_t)nPL * C;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int co = tid % c8, pl = tid / c8;
    const T* xb = x + (size_t)n * HW * C + co * 8;
    float shift[8], a[8], b[8];
    load8<T>(xb, shift);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = 0.f; b[u] = 0.f; }
    if (pl < nPL) {
#pragma unroll 4
        for (int p = pl; p < HW; p += nPL) {
            float v[8];
            load8<T>(xb + (size_t)p * C, v);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float d = v[u] - shift[u];
                a[u] += d;
                b[u] = fmaf(d, d, b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            ssum[pl * C + co * 8 + u] = a[u];
            ssq[pl * C + co * 8 + u] = b[u];
        }
        if (pl == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) sshift[co * 8 + u] = shift[u];
        }
    }
    __syncthreads();
    const float inv = 1.f / (float)HW;
    for (int c = tid; c < C; c += GN_THREADS) {      // C % 32 == 0 when cpg > 1: whole warps stay together
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += ssum[q * C + c]; sb += ssq[q * C + c]; }
        // Statistics of x + pre_bias (a per-channel constant, e.g. the bias of the convolution that produced x,
        // folded here so that neither the add nor its gradient reduction is a separate pass).  The stored mean is
        // the EFFECTIVE one, mean(x + b) - b[c]: every consumer normalises the raw x with it.
        const float pb = pre_bias != nullptr ? pre_bias[c] : 0.f;
        const float sh = sshift[c] + pb;
        if (chan_sum != nullptr) chan_sum[(size_t)n * C + c] = fmaf((float)HW, sshift[c], sa);
        float m, var;
        if (cpg == 1) {
            const float md = sa * inv;
            var = fmaxf(sb * inv - md * md, 0.f);
            m = md + sh;
        } else {
            // raw moments of this channel, then pooled over the group's channels
            float S = fmaf((float)HW, sh, sa);
            float Q = sb + 2.f * sh * sa + (float)HW * sh * sh;
            S = group_lane_sum(S, cpg);
            Q = group_lane_sum(Q, cpg);
            const float ginv = inv / (float)cpg;
            m = S * ginv;
            var = fmaxf(Q * ginv - m * m, 0.f);
        }
        m -= pb;
        mean[(size_t)n * C + c] = m;
        rstd[(size_t)n * C + c] = 1.f / sqrtf(var + eps);
    }
}

// folded affine, as PyTorch's GroupNorm kernels do: y = x*(rstd*gamma) + (beta - mean*rstd*gamma)
__device__ __forceinline__ void affine8(const float* mean, const float* rstd, const float* gamma, const float* beta,
                                        float (&sc)[8], float (&sh)[8]) {
    float m[8], r[8], g[8], b[8];
    load8f(mean, m); load8f(rstd, r); load8f(gamma, g); load8f(beta, b);
#pragma unroll
    for (int u = 0; u < 8; ++u) { sc[u] = r[u] * g[u]; sh[u] = b[u] - m[u] * sc[u]; }
}

// ---- forward: out = bilinear_up(relu(gn(x)))  (same size = identity) ----------------------------
// A thread owns one channel octet of GPIX consecutive output pixels of one sample (per-(n,c) constants
// loaded once, GPIX independent 128-bit loads in flight).
constexpr int GPIX = 4;

template <typename T>
__global__ void __launch_bounds__(256)
gn_relu_up_fwd_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ out,
                      int h, int w, int H, int W, int C, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c8 = C >> 3;
    const int cc = (int)(e % c8) * 8;
    const int HW = H * W;
    const int groups = (HW + GPIX - 1) / GPIX;
    const long long pg = e / c8;
    const int n = (int)(pg / groups);
    const int q0 = (int)(pg - (long long)n * groups) * GPIX;
    float sc[8], sh[8];
    affine8(mean + (size_t)n * C + cc, rstd + (size_t)n * C + cc, gamma + cc, beta + cc, sc, sh);
    const T* xb = x + (size_t)n * h * w * C + cc;
    T* ob = out + (size_t)n * HW * C + cc;
    if (h == H && w == W) {
        float v[GPIX][8];
#pragma unroll
        for (int q = 0; q < GPIX; ++q)
            if (q0 + q < HW) load8<T>(xb + (size_t)(q0 + q) * C, v[q]);
#pragma unroll
        for (int q = 0; q < GPIX; ++q) {
            if (q0 + q >= HW) break;
            float res[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) res[u] = fmaxf(fmaf(v[q][u], sc[u], sh[u]), 0.f);
            store8<T>(ob + (size_t)(q0 + q) * C, res);
        }
        return;
    }
    const float scy = ac_scale(h, H), scx = ac_scale(w, W);
    for (int q = 0; q < GPIX; ++q) {
        const int pix = q0 + q;
        if (pix >= HW) break;
        const int oy = pix / W, ox = pix - oy * W;
        int y0, y1, x0, x1; float ly, lx;
        src_coord(oy, scy, h, y0, y1, ly);
        src_coord(ox, scx, w, x0, x1, lx);
        const float hy = 1.f - ly, hx = 1.f - lx;
        float a[8], b[8], c[8], d[8], res[8];
        load8<T>(xb + ((size_t)y0 * w + x0) * C, a);
        load8<T>(xb + ((size_t)y0 * w + x1) * C, b);
        load8<T>(xb + ((size_t)y1 * w + x0) * C, c);
        load8<T>(xb + ((size_t)y1 * w + x1) * C, d);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float na = fmaxf(fmaf(a[u], sc[u], sh[u]), 0.f), nb = fmaxf(fmaf(b[u], sc[u], sh[u]), 0.f);
            const float nc = fmaxf(fmaf(c[u], sc[u], sh[u]), 0.f), nd = fmaxf(fmaf(d[u], sc[u], sh[u]), 0.f);
            res[u] = hy * (hx * na + lx * nb) + ly * (hx * nc + lx * nd);
        }
        store8<T>(ob + (size_t)pix * C, res);
    }
}

// gradient arriving at source pixel (sy,sx): identity or the adjoint of the bilinear gather
template <typename T>
__device__ __forceinline__ void upstream8(const T* db, int sy, int sx, int h, int w, int H, int W, int C,
                                          float scy, float scx, float (&g)[8]) {
    if (h == H && w == W) {
        load8<T>(db + ((size_t)sy * W + sx) * C, g);
        return;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = 0.f;
    int ylo, yhi, xlo, xhi;
    dst_range(sy, scy, H, ylo, yhi);
    dst_range(sx, scx, W, xlo, xhi);
    for (int oy = ylo; oy <= yhi; ++oy) {
        const float wy = tap_weight(oy, sy, scy, h);
        if (wy == 0.f) continue;
        for (int ox = xlo; ox <= xhi; ++ox) {
            const float wx = tap_weight(ox, sx, scx, w);
            if (wx == 0.f) continue;
            float v[8];
            load8<T>(db + ((size_t)oy * W + ox) * C, v);
            const float ww = wy * wx;
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = fmaf(ww, v[u], g[u]);
        }
    }
}

// ---- backward phase 1: per-sample sums S1 = sum dyh, S2 = sum dyh*xhat, and the group terms -------
// dyh = relu'(y) * upstream.  For an up-sampling call dyh is staged (fp32) for phase 2; for a
// same-size call phase 2 recomputes it.
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_relu_up_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                             const float* __restrict__ mean, const float* __restrict__ rstd,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             float* __restrict__ dyh, float* __restrict__ S1, float* __restrict__ S2,
                             float* __restrict__ A1, float* __restrict__ A2,
                             int h, int w, int H, int W, int C, int cpg) {
    extern __shared__ __align__(16) float sm[];
    const int c8 = C >> 3, nPL = GN_THREADS / c8, hw = h * w;
    float* s1 = sm;
    float* s2 = sm + (size_t)nPL * C;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int co = tid % c8, pl = tid / c8, cc = co * 8;
    float a1[8], a2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a1[u] = 0.f; a2[u] = 0.f; }
    if (pl < nPL) {
        float sc[8], sh[8], m[8], r[8];
        affine8(mean + (size_t)n * C + cc, rstd + (size_t)n * C + cc, gamma + cc, beta + cc, sc, sh);
        load8f(mean + (size_t)n * C + cc, m);
        load8f(rstd + (size_t)n * C + cc, r);
        const T* xb = x + (size_t)n * hw * C + cc;
        const T* db = dout + (size_t)n * H * W * C + cc;
        const bool ident = (h == H && w == W);
        const float scy = ac_scale(h, H), scx = ac_scale(w, W);
        for (int p = pl; p < hw; p += nPL) {
            const int sy = p / w, sx = p - sy * w;
            float g[8], xv[8], d[8];
            upstream8<T>(db, sy, sx, h, w, H, W, C, scy, scx, g);
            load8<T>(xb + (size_t)p * C, xv);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float yh = fmaf(xv[u], sc[u], sh[u]);
                d[u] = (yh > 0.f) ? g[u] : 0.f;
                a1[u] += d[u];
                a2[u] = fmaf(d[u], (xv[u] - m[u]) * r[u], a2[u]);
            }
            if (!ident) store8<float>(dyh + ((size_t)n * hw + p) * C + cc, d);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1[pl * C + cc + u] = a1[u]; s2[pl * C + cc + u] = a2[u]; }
    }
    __syncthreads();
    const float inv = 1.f / ((float)hw * (float)cpg);
    for (int c = tid; c < C; c += GN_THREADS) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += s1[q * C + c]; sb += s2[q * C + c]; }
        S1[(size_t)n * C + c] = sa;
        S2[(size_t)n * C + c] = sb;
        const float gm = gamma[c];
        float ga = gm * sa, gb = gm * sb;
        if (cpg > 1) { ga = group_lane_sum(ga, cpg); gb = group_lane_sum(gb, cpg); }
        A1[(size_t)n * C + c] = ga * inv;
        A2[(size_t)n * C + c] = gb * inv;
    }
}

// ---- backward phase 2: dx = rstd * (gamma*dyh - A1 - xhat*A2) ------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_relu_up_bwd_apply_kernel(const T* __restrict__ dout, const float* __restrict__ dyh, const T* __restrict__ x,
                            const float* __restrict__ mean, const float* __restrict__ rstd,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            const float* __restrict__ A1, const float* __restrict__ A2, T* __restrict__ dx,
                            int hw, int C, int ident, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c8 = C >> 3;
    const int cc = (int)(e % c8) * 8;
    const int groups = (hw + GPIX - 1) / GPIX;
    const long long pg = e / c8;
    const int n = (int)(pg / groups);
    const int q0 = (int)(pg - (long long)n * groups) * GPIX;
    float sc[8], sh[8], m[8], r[8], g[8], a1[8], a2[8];
    affine8(mean + (size_t)n * C + cc, rstd + (size_t)n * C + cc, gamma + cc, beta + cc, sc, sh);
    load8f(mean + (size_t)n * C + cc, m);
    load8f(rstd + (size_t)n * C + cc, r);
    load8f(gamma + cc, g);
    load8f(A1 + (size_t)n * C + cc, a1);
    load8f(A2 + (size_t)n * C + cc, a2);
    const size_t base = (size_t)n * hw * C + cc;
    float xv[GPIX][8], d[GPIX][8];
#pragma unroll
    for (int q = 0; q < GPIX; ++q) {
        if (q0 + q >= hw) continue;
        load8<T>(x + base + (size_t)(q0 + q) * C, xv[q]);
        if (ident) load8<T>(dout + base + (size_t)(q0 + q) * C, d[q]);
        else load8f(dyh + base + (size_t)(q0 + q) * C, d[q]);
    }
#pragma unroll
    for (int q = 0; q < GPIX; ++q) {
        if (q0 + q >= hw) break;
        float res[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float dv = d[q][u];
            if (ident) dv = (fmaf(xv[q][u], sc[u], sh[u]) > 0.f) ? dv : 0.f;
            const float xh = (xv[q][u] - m[u]) * r[u];
            res[u] = r[u] * (g[u] * dv - a1[u] - xh * a2[u]);
        }
        store8<T>(dx + base + (size_t)(q0 + q) * C, res);
    }
}


// Same-size variant of phase 1: no bilinear adjoint in the loop, four channels per thread, four pixels of both
// streams in flight; 64 registers so that two 512-thread CTAs share an SM and all samples of a 256-frame batch
// run in one wave (the general kernel: 91 registers, one CTA per SM, two uneven waves).
template <typename T>
__global__ void __launch_bounds__(GN_THREADS, 2)
gn_relu_bwd_reduce_same_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                               const float* __restrict__ mean, const float* __restrict__ rstd,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ S1, float* __restrict__ S2,
                               float* __restrict__ A1, float* __restrict__ A2, int hw, int C, int cpg) {
    extern __shared__ __align__(16) float sm[];
    const int c4 = C >> 2, nPL = GN_THREADS / c4;
    float* s1 = sm;
    float* s2 = sm + (size_t)nPL * C;
    const int n = blockIdx.x, tid = threadIdx.x;
    const int co = tid % c4, pl = tid / c4, cc = co * 4;
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    if (pl < nPL) {
        const size_t so = (size_t)n * C + cc;
        const float4 m4 = *reinterpret_cast<const float4*>(mean + so), r4 = *reinterpret_cast<const float4*>(rstd + so);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc), b4 = *reinterpret_cast<const float4*>(beta + cc);
        const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
        float sc[4], sh[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { sc[u] = r[u] * g[u]; sh[u] = b[u] - m[u] * sc[u]; }
        const T* xb = x + (size_t)n * hw * C + cc;
        const T* db = dout + (size_t)n * hw * C + cc;
        for (int p0 = pl; p0 < hw; p0 += 4 * nPL) {
            Vec4<T> xv[4], dv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + q * nPL;
                if (p < hw) { xv[q].load(xb + (size_t)p * C); dv[q].load(db + (size_t)p * C); }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (p0 + q * nPL >= hw) break;
                float xf[4], df[4];
                xv[q].get(xf); dv[q].get(df);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float d = (fmaf(xf[u], sc[u], sh[u]) > 0.f) ? df[u] : 0.f;
                    a1[u] += d;
                    a2[u] = fmaf(d, (xf[u] - m[u]) * r[u], a2[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s1[pl * C + cc + u] = a1[u]; s2[pl * C + cc + u] = a2[u]; }
    }
    __syncthreads();
    const float inv = 1.f / ((float)hw * (float)cpg);
    for (int c = tid; c < C; c += GN_THREADS) {
        float sa = 0.f, sb = 0.f;
        for (int q = 0; q < nPL; ++q) { sa += s1[q * C + c]; sb += s2[q * C + c]; }
        S1[(size_t)n * C + c] = sa;
        S2[(size_t)n * C + c] = sb;
        const float gm = gamma[c];
        float ga = gm * sa, gb = gm * sb;
        if (cpg > 1) { ga = group_lane_sum(ga, cpg); gb = group_lane_sum(gb, cpg); }
        A1[(size_t)n * C + c] = ga * inv;
        A2[(size_t)n * C + c] = gb * inv;
    }
}

// Same-size variant of phase 2 (the discriminator towers and the same-size head branches: 19 of the 23 calls of a
// config-2 step).  A thread owns FOUR channels (one 64-bit bf16 access) of GPIX consecutive pixels: half the
// per-thread state of the octet kernel above (134 registers, 11 % occupancy, 20 % of HBM peak in the round-1
// ncu capture), so three times as many warps are in flight.  dx = c0*dyh - c1 - (x-m)*c2 with the per-(n,c)
// constants folded once.
template <typename T>
__global__ void __launch_bounds__(256)
gn_relu_bwd_apply_same_kernel(const T* __restrict__ dout, const T* __restrict__ x,
                              const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              const float* __restrict__ A1, const float* __restrict__ A2, T* __restrict__ dx,
                              int hw, int C, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c4 = C >> 2;
    const int cc = (int)(e % c4) * 4;
    const int groups = (hw + GPIX - 1) / GPIX;
    const long long pg = e / c4;
    const int n = (int)(pg / groups);
    const int q0 = (int)(pg - (long long)n * groups) * GPIX;
    const size_t sc_off = (size_t)n * C + cc;
    const float4 m4 = *reinterpret_cast<const float4*>(mean + sc_off), r4 = *reinterpret_cast<const float4*>(rstd + sc_off);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + cc), b4 = *reinterpret_cast<const float4*>(beta + cc);
    const float4 p4 = *reinterpret_cast<const float4*>(A1 + sc_off), s4 = *reinterpret_cast<const float4*>(A2 + sc_off);
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    const float a1[4] = {p4.x, p4.y, p4.z, p4.w}, a2[4] = {s4.x, s4.y, s4.z, s4.w};
    float c0[4], c1[4], c2[4], sh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        c0[u] = r[u] * g[u];                 // also the folded scale of the forward affine
        c1[u] = r[u] * a1[u];
        c2[u] = r[u] * r[u] * a2[u];
        sh[u] = b[u] - m[u] * c0[u];
    }
    const size_t base = (size_t)n * hw * C + cc;
    Vec4<T> xv[GPIX], dv[GPIX];
#pragma unroll
    for (int q = 0; q < GPIX; ++q)
        if (q0 + q < hw) { xv[q].load(x + base + (size_t)(q0 + q) * C); dv[q].load(dout + base + (size_t)(q0 + q) * C); }
#pragma unroll
    for (int q = 0; q < GPIX; ++q) {
        if (q0 + q >= hw) break;
        float xf[4], df[4], res[4];
        xv[q].get(xf); dv[q].get(df);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float d = (fmaf(xf[u], c0[u], sh[u]) > 0.f) ? df[u] : 0.f;
            res[u] = c0[u] * d - c1[u] - (xf[u] - m[u]) * c2[u];
        }
        Vec4<T> o;
        o.set(res);
        o.store(dx + base + (size_t)(q0 + q) * C);
    }
}

bool gn_shape_ok(int C, int cpg) {
    if (C % 8 != 0 || C / 8 > GN_THREADS || GN_THREADS % (C / 8) != 0) return false;
    if (cpg < 1 || C % cpg != 0) return false;
    if (cpg > 1 && (C % 32 != 0 || cpg > 32 || (cpg & (cpg - 1)) != 0)) return false;
    return true;
}

size_t gn_smem(int C) { return ((size_t)2 * (GN_THREADS / (C / 8)) * C + C) * sizeof(float); }

template <typename K>
int set_smem_once(K kernel, size_t bytes, size_t& cached, const char* name) {
    if (bytes > cached) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), name);
        cached = bytes;
    }
    return GE_OK;
}

}  // namespace

// mean, rstd: fp32 [N,C], the statistics of each group written for every one of its channels.
namespace {
int group_stats_impl(const void* x, const float* pre_bias, float* mean, float* rstd, float* chan_sum, int dtype,
                     int N, int HW, int C, int channels_per_group, float eps, ge_stream_t stream);
}

extern "C" int ge_group_stats(const void* x, float* mean, float* rstd, int dtype,
                              int N, int HW, int C, int channels_per_group, float eps, ge_stream_t stream) {
    return group_stats_impl(x, nullptr, mean, rstd, nullptr, dtype, N, HW, C, channels_per_group, eps, stream);
}

// Statistics of x + pre_bias[c] (pre_bias fp32 [C] or NULL).  mean receives the EFFECTIVE mean (group mean minus
// pre_bias[c]) so that ge_gn_relu_upsample_fwd/bwd run unchanged on the raw x; chan_sum (fp32 [N,C] or NULL) receives
// sum_hw x[n,:,c], from which the caller gets the gradient of pre_bias without a pass over the map.
extern "C" int ge_group_stats_bias(const void* x, const float* pre_bias, float* mean, float* rstd, float* chan_sum,
                                   int dtype, int N, int HW, int C, int channels_per_group, float eps,
                                   ge_stream_t stream) {
    return group_stats_impl(x, pre_bias, mean, rstd, chan_sum, dtype, N, HW, C, channels_per_group, eps, stream);
}

namespace {
int group_stats_impl(const void* x, const float* pre_bias, float* mean, float* rstd, float* chan_sum, int dtype,
                     int N, int HW, int C, int channels_per_group, float eps, ge_stream_t stream) {
    GE_REQUIRE(x && mean && rstd, GE_ERR_ARG, "ge_group_stats: null pointer");
    GE_REQUIRE(N > 0 && HW > 0 && C > 0, GE_ERR_ARG, "ge_group_stats: bad dimension");
    GE_REQUIRE(gn_shape_ok(C, channels_per_group), GE_ERR_SHAPE,
               "ge_group_stats: unsupported C=%d / channels_per_group=%d (C%%8==0, C/8 | 512, group size a power of two <= 32)",
               C, channels_per_group);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = gn_smem(C);
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        int rc = set_smem_once(group_stats_kernel<float>, smem, c0, "ge_group_stats(attr)");
        if (rc) return rc;
        group_stats_kernel<float><<<N, GN_THREADS, smem, st>>>((const float*)x, pre_bias, mean, rstd, chan_sum, HW, C, eps, channels_per_group);
    } else if (dtype == GE_DTYPE_BF16) {
        int rc = set_smem_once(group_stats_kernel<bf16>, smem, c1, "ge_group_stats(attr)");
        if (rc) return rc;
        group_stats_kernel<bf16><<<N, GN_THREADS, smem, st>>>((const bf16*)x, pre_bias, mean, rstd, chan_sum, HW, C, eps, channels_per_group);
    } else { ge_set_error("ge_group_stats: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_group_stats");
    return GE_OK;
}
}  // namespace

extern "C" int ge_gn_relu_upsample_fwd(const void* x, const float* mean, const float* rstd,
                                       const float* gamma, const float* beta, void* out, int dtype,
                                       int N, int h, int w, int H, int W, int C, ge_stream_t stream) {
    GE_REQUIRE(x && mean && rstd && gamma && beta && out, GE_ERR_ARG, "ge_gn_relu_upsample_fwd: null pointer");
    GE_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0, GE_ERR_ARG, "ge_gn_relu_upsample_fwd: bad dimension");
    GE_REQUIRE(C % 8 == 0, GE_ERR_SHAPE, "ge_gn_relu_upsample_fwd: C=%d must be a multiple of 8", C);
    const long long total8 = (long long)N * ge::cdiv(H * W, GPIX) * (C / 8);
    const unsigned blocks = (unsigned)ge::cdivll(total8, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        gn_relu_up_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, mean, rstd, gamma, beta, (float*)out, h, w, H, W, C, total8);
    else if (dtype == GE_DTYPE_BF16)
        gn_relu_up_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, mean, rstd, gamma, beta, (bf16*)out, h, w, H, W, C, total8);
    else { ge_set_error("ge_gn_relu_upsample_fwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_gn_relu_upsample_fwd");
    return GE_OK;
}

// dyh: fp32 scratch [N,h,w,C], only touched when (h,w) != (H,W) (may be NULL for a same-size call);
// S1,S2 [N,C]: dbeta = sum_n S1, dgamma = sum_n S2;  A1,A2 [N,C]: scratch.
extern "C" int ge_gn_relu_upsample_bwd(const void* dout, const void* x, const float* mean, const float* rstd,
                                       const float* gamma, const float* beta, float* dyh, float* S1, float* S2,
                                       float* A1, float* A2, void* dx, int dtype, int N, int h, int w, int H, int W,
                                       int C, int channels_per_group, ge_stream_t stream) {
    GE_REQUIRE(dout && x && mean && rstd && gamma && beta && S1 && S2 && A1 && A2 && dx, GE_ERR_ARG,
               "ge_gn_relu_upsample_bwd: null pointer");
    GE_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0, GE_ERR_ARG, "ge_gn_relu_upsample_bwd: bad dimension");
    GE_REQUIRE(gn_shape_ok(C, channels_per_group), GE_ERR_SHAPE,
               "ge_gn_relu_upsample_bwd: unsupported C=%d / channels_per_group=%d", C, channels_per_group);
    const int ident = (h == H && w == W) ? 1 : 0;
    GE_REQUIRE(ident || dyh, GE_ERR_ARG, "ge_gn_relu_upsample_bwd: dyh scratch is required for an up-sampling call");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = gn_smem(C);
    const long long total8 = (long long)N * ge::cdiv(h * w, GPIX) * (C / 8);
    const unsigned blocks = (unsigned)ge::cdivll(total8, 256);
    const long long total4 = total8 * 2;
    // the same-size reduce kernel: 4 channels per thread, [nPL][C] x 2 floats of shared memory (<= 48 KB by construction)
    const bool same_ok = C / 4 <= GN_THREADS && GN_THREADS % (C / 4) == 0;
    const size_t smem_same = same_ok ? (size_t)2 * (GN_THREADS / (C / 4)) * C * sizeof(float) : 0;
    const unsigned blocks4 = (unsigned)ge::cdivll(total4, 256);
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        int rc = set_smem_once(gn_relu_up_bwd_reduce_kernel<float>, smem, c0, "ge_gn_relu_upsample_bwd(attr)");
        if (rc) return rc;
        if (ident && same_ok)
            gn_relu_bwd_reduce_same_kernel<float><<<N, GN_THREADS, smem_same, st>>>((const float*)dout, (const float*)x, mean, rstd,
                gamma, beta, S1, S2, A1, A2, h * w, C, channels_per_group);
        else
            gn_relu_up_bwd_reduce_kernel<float><<<N, GN_THREADS, smem, st>>>((const float*)dout, (const float*)x, mean, rstd,
                gamma, beta, dyh, S1, S2, A1, A2, h, w, H, W, C, channels_per_group);
        GE_CHECK_LAUNCH("ge_gn_relu_upsample_bwd(reduce)");
        if (ident)
            gn_relu_bwd_apply_same_kernel<float><<<blocks4, 256, 0, st>>>((const float*)dout, (const float*)x, mean, rstd,
                gamma, beta, A1, A2, (float*)dx, h * w, C, total4);
        else
            gn_relu_up_bwd_apply_kernel<float><<<blocks, 256, 0, st>>>((const float*)dout, dyh, (const float*)x, mean, rstd,
                gamma, beta, A1, A2, (float*)dx, h * w, C, ident, total8);
    } else if (dtype == GE_DTYPE_BF16) {
        int rc = set_smem_once(gn_relu_up_bwd_reduce_kernel<bf16>, smem, c1, "ge_gn_relu_upsample_bwd(attr)");
        if (rc) return rc;
        if (ident && same_ok)
            gn_relu_bwd_reduce_same_kernel<bf16><<<N, GN_THREADS, smem_same, st>>>((const bf16*)dout, (const bf16*)x, mean, rstd,
                gamma, beta, S1, S2, A1, A2, h * w, C, channels_per_group);
        else
            gn_relu_up_bwd_reduce_kernel<bf16><<<N, GN_THREADS, smem, st>>>((const bf16*)dout, (const bf16*)x, mean, rstd,
                gamma, beta, dyh, S1, S2, A1, A2, h, w, H, W, C, channels_per_group);
        GE_CHECK_LAUNCH("ge_gn_relu_upsample_bwd(reduce)");
        if (ident)
            gn_relu_bwd_apply_same_kernel<bf16><<<blocks4, 256, 0, st>>>((const bf16*)dout, (const bf16*)x, mean, rstd,
                gamma, beta, A1, A2, (bf16*)dx, h * w, C, total4);
        else
            gn_relu_up_bwd_apply_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)dout, dyh, (const bf16*)x, mean, rstd,
                gamma, beta, A1, A2, (bf16*)dx, h * w, C, ident, total8);
    } else { ge_set_error("ge_gn_relu_upsample_bwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_gn_relu_upsample_bwd(apply)");
    return GE_OK;
}
