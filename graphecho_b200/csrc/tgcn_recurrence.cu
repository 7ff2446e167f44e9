// K6 -- the TGCN recurrence as ONE persistent launch (forward) and one (backward).
//
// Reference: models/TGCN.py:224-235 calls DyGraphConv2d.forward (:62-78) once per time step,
//     hidden_t = GELU( Conv1x1_{groups=4}( interleave[x_t ; max_k( hidden_{t-1}[:, nn_k(x_t -> hidden_{t-1})] - x_t )] ) )
// with the k-NN taken between the channel-normalised x_t (queries) and hidden_{t-1} (keys) (vig.py:369-381, 270-274),
// i.e. per step ~35 launches (normalise x2, bmm, topk, two materialised [B,C,N,k] gathers, grouped cuDNN conv, GELU)
// and the hidden state round-trips through HBM T times.  Everything that does not depend on the recurrent state (pyramid
// pooling, the 1x1-conv MLP, the position embedding) is hoisted out and batched over all frames (models/TGCN.py here);
// what remains is strictly sequential and tiny: per clip 64 nodes x 256 channels.
//
// Here one CTA owns one clip for the whole sequence: hidden (64 KB), x_t (64 KB) and the max-relative features (64 KB)
// live in shared memory, the 64x64 distance tile is built with fp32 FFMA (no TF32: near-ties must order like the fp32
// reference), a warp per query row picks the k nearest (ties -> lower index), the max-relative gather, the grouped 1x1
// convolution (weights streamed from L2, 128 KB) and the exact-erf GELU follow, and the new hidden state overwrites
// the old one in place.  The kernel records what the exact adjoint needs (hidden states, neighbour lists, arg-max
// slots, pre-activations); the backward kernel walks the sequence in reverse with dH resident in shared memory.
// Shapes: C = Cout = 256, N = 64 (clip_shape 8x8), k <= 16, dilation 1 -- the only configuration the reference trainers
// build (train_cardiac_uda.py:120); anything else takes the step-by-step path.
// Bounded by the latency of T dependent steps (8 CTAs busy); bytes per clip and step: 64 KB x_t in, 64 KB hidden out.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using namespace ge;

constexpr int TC = 256;       // channels = threads
constexpr int TN = 64;        // nodes
constexpr int TGRP = 4;       // conv groups
constexpr int TKMAX = 16;
constexpr int TQ = 2 * TC / TGRP;   // input channels per group of the interleaved feature (128)

__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float v) {
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    return cdf + v * pdf;
}

struct Smem {
    float* xs;     // [TC][TN] x_t
    float* hs;     // [TC][TN] hidden
    float* xj;     // [TC][TN] max-relative features
    float* dist;   // [TN][TN]
    float* nx;     // [TN] 1/max(|x_n|,eps), [TN] |x^_n|^2
    float* ny;     // same for hidden
    int* idx;      // [TN][TKMAX]
};

__device__ __forceinline__ Smem carve(unsigned char* raw) {
    Smem s;
    s.xs = reinterpret_cast<float*>(raw);
    s.hs = s.xs + TC * TN;
    s.xj = s.hs + TC * TN;
    s.dist = s.xj + TC * TN;
    s.nx = s.dist + TN * TN;
    s.ny = s.nx + 2 * TN;
    s.idx = reinterpret_cast<int*>(s.ny + 2 * TN);
    return s;
}
constexpr size_t kSmemBytes = (size_t)(3 * TC * TN + TN * TN + 4 * TN) * sizeof(float) + TN * TKMAX * sizeof(int);

// emb [B,T,C,N]; Wt [TQ][C] (transposed grouped-conv weight: Wt[q][o] = W[o][q]); bias [C];
// hidden_all [B,T,C,N], z_all [B,T,C,N], idx_all int32 [B,T,N,k], argk_all uint8 [B,T,C,N]
__global__ void __launch_bounds__(TC, 1)
tgcn_recurrence_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ Wt, const float* __restrict__ bias,
                           float* __restrict__ hidden_all, float* __restrict__ z_all, int* __restrict__ idx_all,
                           unsigned char* __restrict__ argk_all, int T, int k) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ float red[4][TN];
    const Smem s = carve(raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < TC * TN; e += TC) s.hs[e] = 0.f;                       // hidden_0 = 0 (TGCN.py:230)
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const size_t base = ((size_t)b * T + t) * TC * TN;
        // ---- x_t -> shared memory
        for (int e = tid * 4; e < TC * TN; e += TC * 4)
            *reinterpret_cast<float4*>(s.xs + e) = *reinterpret_cast<const float4*>(emb + base + e);
        __syncthreads();
        // ---- norms of x_t and hidden: 4 threads per node, 64 channels each
        {
            const int n = tid & (TN - 1), part = tid >> 6;
            float sx = 0.f, sy = 0.f;
            for (int c = part * 64; c < part * 64 + 64; ++c) {
                const float a = s.xs[c * TN + n], h = s.hs[c * TN + n];
                sx = fmaf(a, a, sx); sy = fmaf(h, h, sy);
            }
            red[part][n] = sx;
            __syncthreads();
            float invx = 0.f;
            if (part == 0) {
                const float tot = red[0][n] + red[1][n] + red[2][n] + red[3][n];
                invx = 1.f / fmaxf(sqrtf(tot), 1e-12f);
                s.nx[n] = invx;
            }
            __syncthreads();
            red[part][n] = sy;
            __syncthreads();
            if (part == 0) {
                const float tot = red[0][n] + red[1][n] + red[2][n] + red[3][n];
                s.ny[n] = 1.f / fmaxf(sqrtf(tot), 1e-12f);
            }
            __syncthreads();
            // squared norms of the NORMALISED vectors (what the reference's x_square / y_square are)
            float qx = 0.f, qy = 0.f;
            const float ix = s.nx[n], iy = s.ny[n];
            for (int c = part * 64; c < part * 64 + 64; ++c) {
                const float a = s.xs[c * TN + n] * ix, h = s.hs[c * TN + n] * iy;
                qx = fmaf(a, a, qx); qy = fmaf(h, h, qy);
            }
            red[part][n] = qx;
            __syncthreads();
            if (part == 0) s.nx[TN + n] = red[0][n] + red[1][n] + red[2][n] + red[3][n];
            __syncthreads();
            red[part][n] = qy;
            __syncthreads();
            if (part == 0) s.ny[TN + n] = red[0][n] + red[1][n] + red[2][n] + red[3][n];
            __syncthreads();
        }
        // ---- distance tile: thread (ti, tj) owns a 4x4 block of (query i, key j)
        {
            const int ti = tid >> 4, tj = tid & 15;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) acc[a][c2] = 0.f;
            const float4 ix4 = *reinterpret_cast<const float4*>(s.nx + 4 * ti), iy4 = *reinterpret_cast<const float4*>(s.ny + 4 * tj);
            const float ix[4] = {ix4.x, ix4.y, ix4.z, ix4.w}, iy[4] = {iy4.x, iy4.y, iy4.z, iy4.w};
#pragma unroll 4
            for (int c = 0; c < TC; ++c) {
                const float4 xa = *reinterpret_cast<const float4*>(s.xs + c * TN + 4 * ti);
                const float4 ya = *reinterpret_cast<const float4*>(s.hs + c * TN + 4 * tj);
                const float xv[4] = {xa.x * ix[0], xa.y * ix[1], xa.z * ix[2], xa.w * ix[3]};
                const float yv[4] = {ya.x * iy[0], ya.y * iy[1], ya.z * iy[2], ya.w * iy[3]};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) acc[a][c2] = fmaf(xv[a], yv[c2], acc[a][c2]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2)     // (|x^|^2 + (-2 x^.y^)) + |y^|^2     [vig.py:270-274]
                    s.dist[(4 * ti + a) * TN + 4 * tj + c2] = (s.nx[TN + 4 * ti + a] + (-2.f * acc[a][c2])) + s.ny[TN + 4 * tj + c2];
        }
        __syncthreads();
        // ---- k nearest keys per query row, ascending distance, ties -> lower index: one warp per row
        for (int i = warp; i < TN; i += TC / 32) {
            float d0 = s.dist[i * TN + lane], d1 = s.dist[i * TN + lane + 32];
            for (int r = 0; r < k; ++r) {
                float best = d0 <= d1 ? d0 : d1;
                int bj = d0 <= d1 ? lane : lane + 32;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float od = __shfl_xor_sync(kFull, best, o);
                    const int oj = __shfl_xor_sync(kFull, bj, o);
                    if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
                }
                if (lane == 0) s.idx[i * TKMAX + r] = bj;
                if (bj == lane) d0 = INFINITY;
                if (bj == lane + 32) d1 = INFINITY;
            }
        }
        __syncthreads();
        // ---- max-relative features: thread (node i, 64-channel block)
        {
            const int i = tid & (TN - 1), cb = (tid >> 6) * 64;
            int nb[TKMAX];
#pragma unroll
            for (int r = 0; r < TKMAX; ++r) nb[r] = r < k ? s.idx[i * TKMAX + r] : 0;
            if (cb == 0)
                for (int r = 0; r < k; ++r) idx_all[(((size_t)b * T + t) * TN + i) * k + r] = nb[r];
            for (int c = cb; c < cb + 64; ++c) {
                const float xi = s.xs[c * TN + i];
                float best = -INFINITY;
                int slot = 0;
#pragma unroll
                for (int r = 0; r < TKMAX; ++r) {
                    if (r < k) {
                        const float v = s.hs[c * TN + nb[r]] - xi;          // y_j - x_i   (vig.py:100-102)
                        if (v > best) { best = v; slot = r; }
                    }
                }
                s.xj[c * TN + i] = best;
                argk_all[base + c * TN + i] = (unsigned char)slot;
            }
        }
        __syncthreads();
        // ---- grouped 1x1 conv (+bias) + GELU: thread = output channel o, all 64 nodes
        {
            const int o = tid, g = o / (TC / TGRP), c0 = g * (TC / TGRP);
            float acc[TN];
#pragma unroll
            for (int i = 0; i < TN; ++i) acc[i] = 0.f;
            for (int q = 0; q < TQ; ++q) {
                const float w = Wt[q * TC + o];
                const float* f = ((q & 1) ? s.xj : s.xs) + (c0 + (q >> 1)) * TN;    // interleaved [x_c, xj_c]  (vig.py:104)
#pragma unroll
                for (int i4 = 0; i4 < TN; i4 += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(f + i4);
                    acc[i4] = fmaf(w, v.x, acc[i4]); acc[i4 + 1] = fmaf(w, v.y, acc[i4 + 1]);
                    acc[i4 + 2] = fmaf(w, v.z, acc[i4 + 2]); acc[i4 + 3] = fmaf(w, v.w, acc[i4 + 3]);
                }
            }
            const float bo = bias[o];
            __syncthreads();                       // every thread is done reading the OLD hidden (gather) before it is overwritten
#pragma unroll
            for (int i4 = 0; i4 < TN; i4 += 4) {
                float z[4], h[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { z[u] = acc[i4 + u] + bo; h[u] = gelu_exact(z[u]); }
                *reinterpret_cast<float4*>(z_all + base + o * TN + i4) = make_float4(z[0], z[1], z[2], z[3]);
                *reinterpret_cast<float4*>(hidden_all + base + o * TN + i4) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(s.hs + o * TN + i4) = make_float4(h[0], h[1], h[2], h[3]);
            }
        }
        __syncthreads();
    }
}

// Reverse sweep.  dH [B,C,N] = gradient of the final hidden state; d_emb [B,T,C,N]; dWt_part [B][TQ][C], db_part [B][C]
// (zero-filled by the caller, summed over b by the caller); scratch [B][C][N] fp32.
__global__ void __launch_bounds__(TC, 1)
tgcn_recurrence_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ W, const float* __restrict__ hidden_all,
                           const float* __restrict__ z_all, const int* __restrict__ idx_all,
                           const unsigned char* __restrict__ argk_all, const float* __restrict__ dH,
                           float* __restrict__ d_emb, float* __restrict__ dWt_part, float* __restrict__ db_part,
                           float* __restrict__ scratch, int T, int k) {
    extern __shared__ __align__(16) unsigned char raw[];
    const Smem s = carve(raw);
    float* dz = s.hs;                 // [TC][TN]: dH at loop entry, dz after the GELU adjoint, dH_{t-1} at loop exit
    const int b = blockIdx.x, tid = threadIdx.x;
    float* sc = scratch + (size_t)b * TC * TN;
    float* dWt = dWt_part + (size_t)b * TQ * TC;
    for (int e = tid * 4; e < TC * TN; e += TC * 4)
        *reinterpret_cast<float4*>(dz + e) = *reinterpret_cast<const float4*>(dH + (size_t)b * TC * TN + e);
    float dbias = 0.f;
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
        const size_t base = ((size_t)b * T + t) * TC * TN;
        const float* hp = t > 0 ? hidden_all + base - (size_t)TC * TN : nullptr;      // hidden_{t-1} (zeros for t = 0)
        // ---- dz = dH * gelu'(z); x_t -> xs; xj recomputed from hidden_{t-1}, the neighbour list and the arg-max slots
        for (int e = tid * 4; e < TC * TN; e += TC * 4) {
            const float4 z = *reinterpret_cast<const float4*>(z_all + base + e);
            float4 d = *reinterpret_cast<float4*>(dz + e);
            d.x *= gelu_grad(z.x); d.y *= gelu_grad(z.y); d.z *= gelu_grad(z.z); d.w *= gelu_grad(z.w);
            *reinterpret_cast<float4*>(dz + e) = d;
            *reinterpret_cast<float4*>(s.xs + e) = *reinterpret_cast<const float4*>(emb + base + e);
        }
        for (int e = tid; e < TN * k; e += TC) s.idx[(e / k) * TKMAX + (e % k)] = idx_all[((size_t)b * T + t) * TN * k + e];
        __syncthreads();
        {
            const int i = tid & (TN - 1), cb = (tid >> 6) * 64;
            for (int c = cb; c < cb + 64; ++c) {
                const int j = s.idx[i * TKMAX + argk_all[base + c * TN + i]];
                const float yv = hp ? hp[c * TN + j] : 0.f;
                s.xj[c * TN + i] = yv - s.xs[c * TN + i];
            }
        }
        __syncthreads();
        // ---- weight / bias gradients: thread = output channel o;  dWt[q][o] += sum_i dz[o][i] * feat[q][i]
        {
            const int o = tid, g = o / (TC / TGRP), c0 = g * (TC / TGRP);
            float row[TN];
            float sb = 0.f;
#pragma unroll
            for (int i4 = 0; i4 < TN; i4 += 4) {
                const float4 v = *reinterpret_cast<const float4*>(dz + o * TN + i4);
                row[i4] = v.x; row[i4 + 1] = v.y; row[i4 + 2] = v.z; row[i4 + 3] = v.w;
                sb += (v.x + v.y) + (v.z + v.w);
            }
            dbias += sb;
            for (int q = 0; q < TQ; ++q) {
                const float* f = ((q & 1) ? s.xj : s.xs) + (c0 + (q >> 1)) * TN;
                float a = 0.f;
#pragma unroll
                for (int i4 = 0; i4 < TN; i4 += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(f + i4);
                    a = fmaf(row[i4], v.x, a); a = fmaf(row[i4 + 1], v.y, a);
                    a = fmaf(row[i4 + 2], v.z, a); a = fmaf(row[i4 + 3], v.w, a);
                }
                dWt[q * TC + o] += a;
            }
        }
        // ---- feature gradients: thread = input channel c;  dfx = sum_o W[o][2c'] dz[o][:],  dfj = sum_o W[o][2c'+1] dz[o][:]
        {
            const int c = tid, g = c / (TC / TGRP), cl = c - g * (TC / TGRP), o0 = g * (TC / TGRP);
            for (int half = 0; half < 2; ++half) {
                float fx[TN / 2], fj[TN / 2];
#pragma unroll
                for (int i = 0; i < TN / 2; ++i) { fx[i] = 0.f; fj[i] = 0.f; }
                for (int o = o0; o < o0 + TC / TGRP; ++o) {
                    const float2 w = *reinterpret_cast<const float2*>(W + (size_t)o * TQ + 2 * cl);
                    const float* dr = dz + o * TN + half * (TN / 2);
#pragma unroll
                    for (int i4 = 0; i4 < TN / 2; i4 += 4) {
                        const float4 v = *reinterpret_cast<const float4*>(dr + i4);
                        fx[i4] = fmaf(w.x, v.x, fx[i4]); fx[i4 + 1] = fmaf(w.x, v.y, fx[i4 + 1]);
                        fx[i4 + 2] = fmaf(w.x, v.z, fx[i4 + 2]); fx[i4 + 3] = fmaf(w.x, v.w, fx[i4 + 3]);
                        fj[i4] = fmaf(w.y, v.x, fj[i4]); fj[i4 + 1] = fmaf(w.y, v.y, fj[i4 + 1]);
                        fj[i4 + 2] = fmaf(w.y, v.z, fj[i4 + 2]); fj[i4 + 3] = fmaf(w.y, v.w, fj[i4 + 3]);
                    }
                }
#pragma unroll
                for (int i4 = 0; i4 < TN / 2; i4 += 4) {
                    const int i = half * (TN / 2) + i4;
                    // x_i enters as itself and as -x_i inside (y_j - x_i)
                    *reinterpret_cast<float4*>(d_emb + base + c * TN + i) =
                        make_float4(fx[i4] - fj[i4], fx[i4 + 1] - fj[i4 + 1], fx[i4 + 2] - fj[i4 + 2], fx[i4 + 3] - fj[i4 + 3]);
                    *reinterpret_cast<float4*>(sc + c * TN + i) = make_float4(fj[i4], fj[i4 + 1], fj[i4 + 2], fj[i4 + 3]);
                }
            }
        }
        __syncthreads();              // every reader of dz is done: the buffer becomes dH_{t-1}
        {
            const int c = tid;
            float* row = dz + c * TN;
#pragma unroll
            for (int i4 = 0; i4 < TN; i4 += 4) *reinterpret_cast<float4*>(row + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t > 0)
                for (int i = 0; i < TN; ++i) {            // y_j received the gradient of the arg-max slot
                    const int j = s.idx[i * TKMAX + argk_all[base + c * TN + i]];
                    row[j] += sc[c * TN + i];
                }
        }
        __syncthreads();
    }
    db_part[(size_t)b * TC + tid] = dbias;
}

}  // namespace

extern "C" int ge_tgcn_recurrence_supported(int C, int Cout, int N, int k, int dilation, int groups) {
    return (C == TC && Cout == TC && N == TN && k >= 1 && k <= TKMAX && dilation == 1 && groups == TGRP) ? 1 : 0;
}

// emb fp32 [B,T,C,N] (embedded frames + position); Wt fp32 [2C/groups, Cout] = transposed weight of the grouped 1x1
// conv (grapher.gconv.nn.0.weight [Cout, 2C/groups, 1, 1]); bias fp32 [Cout].  Outputs (all saved for the backward):
// hidden_all, z_all fp32 [B,T,C,N] (hidden_all[:, T-1] is the result), idx_all int32 [B,T,N,k], argk_all uint8 [B,T,C,N].
extern "C" int ge_tgcn_recurrence_fwd(const float* emb, const float* Wt, const float* bias, float* hidden_all,
                                      float* z_all, int* idx_all, unsigned char* argk_all,
                                      int B, int T, int C, int N, int k, ge_stream_t stream) {
    GE_REQUIRE(emb && Wt && bias && hidden_all && z_all && idx_all && argk_all, GE_ERR_ARG, "ge_tgcn_recurrence_fwd: null pointer");
    GE_REQUIRE(B > 0 && T > 0, GE_ERR_ARG, "ge_tgcn_recurrence_fwd: bad dimension");
    GE_REQUIRE(ge_tgcn_recurrence_supported(C, C, N, k, 1, TGRP), GE_ERR_SHAPE,
               "ge_tgcn_recurrence_fwd: supports C = 256, N = 64, k <= 16 (got C=%d N=%d k=%d)", C, N, k);
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        GE_CUDA(cudaFuncSetAttribute(tgcn_recurrence_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes), "ge_tgcn_recurrence_fwd(attr)");
        attr = true;
    }
    tgcn_recurrence_fwd_kernel<<<B, TC, kSmemBytes, st>>>(emb, Wt, bias, hidden_all, z_all, idx_all, argk_all, T, k);
    GE_CHECK_LAUNCH("ge_tgcn_recurrence_fwd");
    return GE_OK;
}

// W fp32 [Cout, 2C/groups]; dH fp32 [B,C,N]; d_emb fp32 [B,T,C,N]; dWt_part fp32 [B, 2C/groups, Cout] and db_part fp32
// [B, Cout] must be ZERO-FILLED (the caller sums them over B); scratch fp32 [B,C,N].
extern "C" int ge_tgcn_recurrence_bwd(const float* emb, const float* W, const float* hidden_all, const float* z_all,
                                      const int* idx_all, const unsigned char* argk_all, const float* dH,
                                      float* d_emb, float* dWt_part, float* db_part, float* scratch,
                                      int B, int T, int C, int N, int k, ge_stream_t stream) {
    GE_REQUIRE(emb && W && hidden_all && z_all && idx_all && argk_all && dH && d_emb && dWt_part && db_part && scratch,
               GE_ERR_ARG, "ge_tgcn_recurrence_bwd: null pointer");
    GE_REQUIRE(B > 0 && T > 0, GE_ERR_ARG, "ge_tgcn_recurrence_bwd: bad dimension");
    GE_REQUIRE(ge_tgcn_recurrence_supported(C, C, N, k, 1, TGRP), GE_ERR_SHAPE,
               "ge_tgcn_recurrence_bwd: supports C = 256, N = 64, k <= 16 (got C=%d N=%d k=%d)", C, N, k);
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr = false;
    if (!attr) {
        GE_CUDA(cudaFuncSetAttribute(tgcn_recurrence_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes), "ge_tgcn_recurrence_bwd(attr)");
        attr = true;
    }
    tgcn_recurrence_bwd_kernel<<<B, TC, kSmemBytes, st>>>(emb, W, hidden_all, z_all, idx_all, argk_all, dH, d_emb, dWt_part,
                                                          db_part, scratch, T, k);
    GE_CHECK_LAUNCH("ge_tgcn_recurrence_bwd");
    return GE_OK;
}
