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
// Work per problem: H*N1*N2 relu-coupled terms, 2 issue slots each (max + fma, see the forward kernel) —
// instruction-issue bound, not HBM bound: algorithmic bytes are 4*H*(N1+N2) + 4*N1*N2.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int TI = 32;       // output tile rows per CTA
constexpr int TJ = 32;       // output tile cols per CTA
constexpr int FWD_THREADS = 256;
constexpr int FWD_WARPS = FWD_THREADS / 32;
constexpr int KC = 256;      // hidden channels staged per pass (67 KB of shared memory: 3 CTAs per SM)
constexpr int RED_LD = 40;   // padded row stride of the cross-warp reduction tile

// Forward.  relu(a + b) = a + max(b, -a) exactly (both branches are the same fp32 operations), so
//   M_ij = b2 + sum_k w_k a_ik + sum_k w_k max(b_jk, -a_ik):
// the pairwise part costs one FMNMX (ALU pipe) + one FFMA (FMA pipe) per (i, j, k) instead of add + max + fma, and
// the row term is a per-row dot product (O(N1 H)).  Each CTA: 32x32 outputs, the hidden axis staged through shared
// memory in passes of KC channels (A stored negated), the 8 warps split each pass (split-K inside the CTA, so a
// single 250x250 problem still yields 64 CTAs x 8 warps), every lane owns an 8x4 register tile with rows li+4r /
// cols lj+8c so that the 128-bit shared loads are bank-conflict free with a row stride of KC+4 floats.
__global__ void __launch_bounds__(FWD_THREADS, 3)
affinity_pairwise_fwd_kernel(const float* __restrict__ A, const float* __restrict__ B,
                             const float* __restrict__ w2, const float* __restrict__ b2,
                             float* __restrict__ M, int N1, int N2, int H) {
    extern __shared__ __align__(16) float smem[];
    constexpr int ld = KC + 4;
    float* nAs = smem;                // [TI][ld]   -a
    float* Bs = nAs + TI * ld;        // [TJ][ld]
    float* ws = Bs + TJ * ld;         // [KC]

    const int b = blockIdx.z;
    A += (size_t)b * N1 * H;
    B += (size_t)b * N2 * H;
    M += (size_t)b * N1 * N2;
    const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int li = lane >> 3, lj = lane & 7;

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    float rowc[4] = {0.f, 0.f, 0.f, 0.f};     // lane 0 of warp w: rows w, w+8, w+16, w+24

    for (int k0 = 0; k0 < H; k0 += KC) {
        const int kc = min(KC, H - k0);       // multiple of 32 (host guarantees H % 32 == 0)
        const int h4 = kc >> 2;
        if (k0 > 0) __syncthreads();          // previous pass fully consumed
        for (int e = tid; e < TI * h4; e += FWD_THREADS) {
            const int r = e / h4, k4 = e - r * h4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (i0 + r < N1) va = __ldg(reinterpret_cast<const float4*>(A + (size_t)(i0 + r) * H + k0) + k4);
            if (j0 + r < N2) vb = __ldg(reinterpret_cast<const float4*>(B + (size_t)(j0 + r) * H + k0) + k4);
            *reinterpret_cast<float4*>(nAs + r * ld + 4 * k4) = make_float4(-va.x, -va.y, -va.z, -va.w);
            *reinterpret_cast<float4*>(Bs + r * ld + 4 * k4) = vb;
        }
        for (int k = tid; k < kc; k += FWD_THREADS) ws[k] = __ldg(w2 + k0 + k);
        __syncthreads();
        // row term of this pass
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const float* ar = nAs + (warp + 8 * rr) * ld;
            float sdot = 0.f;
            for (int k = lane; k < kc; k += 32) sdot = fmaf(ws[k], ar[k], sdot);
            rowc[rr] -= ge::warp_sum(sdot);
        }
        const int kslice = kc / FWD_WARPS;    // multiple of 4
        const int kbeg = warp * kslice, kend = kbeg + kslice;
        for (int k = kbeg; k < kend; k += 4) {
            float4 a4[8], b4[4];
#pragma unroll
            for (int r = 0; r < 8; ++r) a4[r] = *reinterpret_cast<const float4*>(nAs + (li + 4 * r) * ld + k);
#pragma unroll
            for (int c = 0; c < 4; ++c) b4[c] = *reinterpret_cast<const float4*>(Bs + (lj + 8 * c) * ld + k);
            const float4 w4 = *reinterpret_cast<const float4*>(ws + k);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[r][c] = fmaf(fmaxf(b4[c].x, a4[r].x), w4.x, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].y, a4[r].y), w4.y, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].z, a4[r].z), w4.z, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].w, a4[r].w), w4.w, acc[r][c]);
                }
        }
    }
    __syncthreads();
    float* red = smem;  // reuse: [FWD_WARPS][TI * RED_LD], then [TI] row terms behind it
    float* rcs = smem + FWD_WARPS * TI * RED_LD;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            red[warp * (TI * RED_LD) + (li + 4 * r) * RED_LD + lj + 8 * c] = acc[r][c];
    if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) rcs[warp + 8 * rr] = rowc[rr];
    }
    __syncthreads();
    const float bias = __ldg(b2);
    for (int o = tid; o < TI * TJ; o += FWD_THREADS) {
        const int i = o >> 5, j = o & 31;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < FWD_WARPS; ++w) s += red[w * (TI * RED_LD) + i * RED_LD + j];
        if (i0 + i < N1 && j0 + j < N2) M[(size_t)(i0 + i) * N2 + j0 + j] = s + rcs[i] + bias;
    }
}

// Forward for batches that fill the machine anyway: 64x64 outputs per CTA, no split-K (each of the 8 warps owns a
// 32x16 sub-tile, lanes 8x2 register tiles), the hidden axis in passes of 128 channels (same 67 KB of shared memory, 3
// CTAs/SM).  Per (i,j,k) term the tile staging moves half the bytes of the 32x32 kernel (the L2 -> shared-memory stream
// was 4.3 GB at 512 problems) and the cross-warp reduction is gone.
constexpr int BI = 64, BJ = 64, BKC = 128;

__global__ void __launch_bounds__(FWD_THREADS, 3)
affinity_pairwise_fwd64_kernel(const float* __restrict__ A, const float* __restrict__ B,
                               const float* __restrict__ w2, const float* __restrict__ b2,
                               float* __restrict__ M, int N1, int N2, int H) {
    extern __shared__ __align__(16) float smem[];
    constexpr int ld = BKC + 4;
    float* nAs = smem;                // [BI][ld]   -a
    float* Bs = nAs + BI * ld;        // [BJ][ld]
    float* ws = Bs + BJ * ld;         // [BKC]
    float* rcs = ws + BKC;            // [BI]

    const int b = blockIdx.z;
    A += (size_t)b * N1 * H;
    B += (size_t)b * N2 * H;
    M += (size_t)b * N1 * N2;
    const int i0 = blockIdx.y * BI, j0 = blockIdx.x * BJ;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 2, wc = warp & 3;            // 2 x 4 warps
    const int li = lane >> 3, lj = lane & 7;
    const int rbase = wr * 32 + li, cbase = wc * 16 + lj;

    float acc[8][2];
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
    float rowc[8];                                      // rows warp + 8 rr
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) rowc[rr] = 0.f;

    for (int k0 = 0; k0 < H; k0 += BKC) {
        const int kc = min(BKC, H - k0);                // multiple of 32
        const int h4 = kc >> 2;
        if (k0 > 0) __syncthreads();
        for (int e = tid; e < BI * h4; e += FWD_THREADS) {
            const int r = e / h4, k4 = e - r * h4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (i0 + r < N1) va = __ldg(reinterpret_cast<const float4*>(A + (size_t)(i0 + r) * H + k0) + k4);
            if (j0 + r < N2) vb = __ldg(reinterpret_cast<const float4*>(B + (size_t)(j0 + r) * H + k0) + k4);
            *reinterpret_cast<float4*>(nAs + r * ld + 4 * k4) = make_float4(-va.x, -va.y, -va.z, -va.w);
            *reinterpret_cast<float4*>(Bs + r * ld + 4 * k4) = vb;
        }
        for (int k = tid; k < kc; k += FWD_THREADS) ws[k] = __ldg(w2 + k0 + k);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const float* ar = nAs + (warp + 8 * rr) * ld;
            float sdot = 0.f;
            for (int k = lane; k < kc; k += 32) sdot = fmaf(ws[k], ar[k], sdot);
            rowc[rr] -= ge::warp_sum(sdot);
        }
        for (int k = 0; k < kc; k += 4) {
            float4 a4[8], b4[2];
#pragma unroll
            for (int r = 0; r < 8; ++r) a4[r] = *reinterpret_cast<const float4*>(nAs + (rbase + 4 * r) * ld + k);
#pragma unroll
            for (int c = 0; c < 2; ++c) b4[c] = *reinterpret_cast<const float4*>(Bs + (cbase + 8 * c) * ld + k);
            const float4 w4 = *reinterpret_cast<const float4*>(ws + k);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    acc[r][c] = fmaf(fmaxf(b4[c].x, a4[r].x), w4.x, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].y, a4[r].y), w4.y, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].z, a4[r].z), w4.z, acc[r][c]);
                    acc[r][c] = fmaf(fmaxf(b4[c].w, a4[r].w), w4.w, acc[r][c]);
                }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) rcs[warp + 8 * rr] = rowc[rr];
    }
    __syncthreads();
    const float bias = __ldg(b2);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = rbase + 4 * r;
        if (i0 + i >= N1) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int j = cbase + 8 * c;
            if (j0 + j < N2) M[(size_t)(i0 + i) * N2 + j0 + j] = acc[r][c] + rcs[i] + bias;
        }
    }
}

// Backward sweep: one thread per hidden channel k, a tile of RT rows of P per CTA, a range of rows of Q.
//   TRANS=false: P=A (rows i), Q=B, g(r,q) = dM[(i0+r)*N2 + q]
//   TRANS=true : P=B (rows j), Q=A, g(r,q) = dM[q*N2 + (j0+r)]
// The ReLU gate is b_jk > -a_ik (<=> a + b > 0, exact), one FSETP; the gated upstream gradient is added to the
// P-side accumulator and (BOTH) to the Q-side accumulator with predicated FADDs: 3 (BOTH) or 2 issue slots per
// (i, j, k).  Accumulators are UNWEIGHTED sums s_ik = sum_j g_ij [a_ik + b_jk > 0]; then
//   dA_ik = w_k s_ik,  dB_jk = w_k t_jk,  dw_k = sum_i a_ik s_ik + sum_j b_jk t_jk
// (g relu(a+b) = g [gate] (a + b) splits over the two sides), so dw costs O((N1+N2) H) instead of a third
// accumulation per element.
//   direct:  the CTA sees all of Q (one launch per side): writes w_k s to the output and its dw partial.
//   partial: small batches split Q over CTAs to fill the machine; the CTA writes its s partial and (BOTH) its t
//            partial per row tile; affinity_bwd_rows_kernel sums them.
constexpr int RT = 32;
constexpr int QT = 8;        // Q rows per inner pass (4 in BOTH mode: 32 + 32 + 4 + 4 live accumulators/operands)

template <bool BOTH>
__device__ __forceinline__ void gate_add(float& dp, float& dq, float bq, float nap, float g) {
    if (BOTH)
        asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %2, %3;\n\t@p add.f32 %0, %0, %4;\n\t@p add.f32 %1, %1, %4;\n\t}"
            : "+f"(dp), "+f"(dq) : "f"(bq), "f"(nap), "f"(g));
    else
        asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.f32 %0, %0, %3;\n\t}"
            : "+f"(dp) : "f"(bq), "f"(nap), "f"(g));
}

template <bool TRANS, bool BOTH>
__global__ void __launch_bounds__(512, 1)
affinity_pairwise_bwd_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                             const float* __restrict__ w2, const float* __restrict__ dM,
                             float* __restrict__ outP, float* __restrict__ outQ_part,
                             float* __restrict__ dw_part, float* __restrict__ db_part,
                             int NP, int NQ, int H, int N1, int N2, int QR, int direct) {
    extern __shared__ __align__(16) float g[];   // [RT][gld]
    __shared__ float scratch[32];
    const int it = blockIdx.x, js = blockIdx.y, b = blockIdx.z;
    const int n_it = gridDim.x, n_js = gridDim.y;
    P += (size_t)b * NP * H;
    Q += (size_t)b * NQ * H;
    dM += (size_t)b * N1 * N2;
    const int r0 = it * RT;
    const int q0 = js * QR;
    const int nq = max(0, min(QR, NQ - q0));
    const int nq8 = (nq + QT - 1) / QT * QT;
    const int gld = nq8 + 4;
    float gsum = 0.f;
    if (TRANS) {
        for (int e = threadIdx.x; e < RT * nq8; e += blockDim.x) {
            const int q = e / RT, r = e - q * RT;
            float v = 0.f;
            if (r0 + r < NP && q < nq) v = dM[(size_t)(q0 + q) * N2 + (r0 + r)];
            g[r * gld + q] = v;
        }
    } else {
        for (int e = threadIdx.x; e < RT * nq8; e += blockDim.x) {
            const int r = e / nq8, q = e - r * nq8;
            float v = 0.f;
            if (r0 + r < NP && q < nq) v = dM[(size_t)(r0 + r) * N2 + (q0 + q)];
            g[r * gld + q] = v;
            gsum += v;
        }
    }
    if (db_part) {   // block_sum synchronises
        gsum = ge::block_sum(gsum, scratch);
        if (threadIdx.x == 0) db_part[((size_t)b * n_it + it) * n_js + js] = gsum;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        float na[RT], dp[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            na[r] = (r0 + r < NP) ? -__ldg(P + (size_t)(r0 + r) * H + k) : 0.f;
            dp[r] = 0.f;
        }
        constexpr int QS = BOTH ? 4 : 8;
        for (int qc = 0; qc < nq8; qc += QS) {
            float bq[QS], dq[QS];
#pragma unroll
            for (int u = 0; u < QS; ++u) {
                bq[u] = (qc + u < nq) ? __ldg(Q + (size_t)(q0 + qc + u) * H + k) : 0.f;
                dq[u] = 0.f;
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const float4 ga = *reinterpret_cast<const float4*>(g + r * gld + qc);
                gate_add<BOTH>(dp[r], dq[0], bq[0], na[r], ga.x);
                gate_add<BOTH>(dp[r], dq[1], bq[1], na[r], ga.y);
                gate_add<BOTH>(dp[r], dq[2], bq[2], na[r], ga.z);
                gate_add<BOTH>(dp[r], dq[3], bq[3], na[r], ga.w);
                if (QS == 8) {
                    const float4 gb = *reinterpret_cast<const float4*>(g + r * gld + qc + 4);
                    gate_add<BOTH>(dp[r], dq[QS - 4], bq[QS - 4], na[r], gb.x);
                    gate_add<BOTH>(dp[r], dq[QS - 3], bq[QS - 3], na[r], gb.y);
                    gate_add<BOTH>(dp[r], dq[QS - 2], bq[QS - 2], na[r], gb.z);
                    gate_add<BOTH>(dp[r], dq[QS - 1], bq[QS - 1], na[r], gb.w);
                }
            }
            if (BOTH) {
#pragma unroll
                for (int u = 0; u < QS; ++u)
                    if (qc + u < nq)
                        outQ_part[(((size_t)b * n_it + it) * NQ + q0 + qc + u) * H + k] = dq[u];
            }
        }
        if (direct) {
            const float wk = __ldg(w2 + k);
            float accW = 0.f;
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                if (r0 + r < NP) outP[((size_t)b * NP + r0 + r) * H + k] = wk * dp[r];
                accW = fmaf(-na[r], dp[r], accW);
            }
            dw_part[((size_t)b * n_it + it) * H + k] = accW;
        } else {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r0 + r < NP) outP[(((size_t)b * n_js + js) * NP + r0 + r) * H + k] = dp[r];
        }
    }
}

// Partial mode: sums the per-CTA partials of FR rows of [A ; B] per CTA, scales by w, and forms the dw partial.
constexpr int FR = 2;
__global__ void __launch_bounds__(512)
affinity_bwd_rows_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ w2,
                         const float* __restrict__ dA_part, const float* __restrict__ dB_part,
                         float* __restrict__ dA, float* __restrict__ dB, float* __restrict__ dw_part,
                         int N1, int N2, int H, int n_js, int n_it) {
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
        const float wk = __ldg(w2 + k);
        float accW = 0.f;
#pragma unroll
        for (int rr = 0; rr < FR; ++rr) {
            const int row = blockIdx.x * FR + rr;
            if (row < N1) {
                float s = 0.f;
#pragma unroll 8
                for (int p = 0; p < n_js; ++p) s += __ldg(dA_part + (((size_t)b * n_js + p) * N1 + row) * H + k);
                const size_t o = ((size_t)b * N1 + row) * H + k;
                dA[o] = wk * s;
                accW = fmaf(__ldg(A + o), s, accW);
            } else if (row - N1 < N2) {
                const int j = row - N1;
                float s = 0.f;
#pragma unroll 8
                for (int p = 0; p < n_it; ++p) s += __ldg(dB_part + (((size_t)b * n_it + p) * N2 + j) * H + k);
                const size_t o = ((size_t)b * N2 + j) * H + k;
                dB[o] = wk * s;
                accW = fmaf(__ldg(B + o), s, accW);
            }
        }
        dw_part[((size_t)b * gridDim.x + blockIdx.x) * H + k] = accW;
    }
}

// dw2[k] = sum of the partials (one CTA per 32 channels, 16 warps split the partials); db2 = sum of the per-CTA
// sums of dM (CTA 0).
__global__ void __launch_bounds__(512)
affinity_bwd_finalize_kernel(const float* __restrict__ dw_part, int nparts, int H,
                             const float* __restrict__ db_part, int ndb,
                             float* __restrict__ dw2, float* __restrict__ db2) {
    __shared__ float red[16][33];
    __shared__ float scratch[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (k < H)
        for (int p = warp; p < nparts; p += 16) s += dw_part[(size_t)p * H + k];
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && k < H) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) t += red[w][lane];
        dw2[k] = t;
    }
    if (blockIdx.x == 0) {
        float t = 0.f;
        for (int e = threadIdx.x; e < ndb; e += blockDim.x) t += db_part[e];
        t = ge::block_sum(t, scratch);
        if (threadIdx.x == 0) db2[0] = t;
    }
}

struct BwdPlan {
    bool partial;      // split Q over CTAs (small batches)
    int n_itA, n_itB;  // row tiles of A / B
    int QR, n_js;      // partial mode: columns of B per CTA, number of splits
    size_t dw_parts, db_parts, dA_part, dB_part;   // floats
    size_t total() const { return dw_parts + db_parts + dA_part + dB_part; }
};

BwdPlan bwd_plan(int batch, int N1, int N2, int H) {
    BwdPlan p;
    p.n_itA = ge::cdiv(N1, RT);
    p.n_itB = ge::cdiv(N2, RT);
    const long long ctas = (long long)batch * p.n_itA;
    const int sms = ge::sm_count();
    p.partial = ctas < 2LL * sms;
    if (p.partial) {
        int js = (int)ge::cdivll(2LL * sms, ctas);
        js = js < 1 ? 1 : js;
        p.QR = ge::cdiv(ge::cdiv(N2, js), QT) * QT;
        p.n_js = ge::cdiv(N2, p.QR);
        p.dw_parts = (size_t)batch * ge::cdiv(N1 + N2, FR) * H;
        p.db_parts = (size_t)batch * p.n_itA * p.n_js;
        p.dA_part = (size_t)batch * p.n_js * N1 * H;
        p.dB_part = (size_t)batch * p.n_itA * N2 * H;
    } else {
        p.QR = 0;
        p.n_js = 1;
        p.dw_parts = (size_t)batch * (p.n_itA + p.n_itB) * H;
        p.db_parts = (size_t)batch * p.n_itA;
        p.dA_part = p.dB_part = 0;
    }
    p.db_parts = (p.db_parts + 3) & ~(size_t)3;
    return p;
}

template <typename K>
int set_smem(K kernel, size_t bytes, size_t& cached, const char* name) {
    if (bytes > cached) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), name);
        cached = bytes;
    }
    return GE_OK;
}

}  // namespace

extern "C" int ge_affinity_pairwise_fwd(const float* A, const float* B, const float* w2, const float* b2,
                                        float* M, int batch, int N1, int N2, int H, ge_stream_t stream) {
    GE_REQUIRE(A && B && w2 && b2 && M, GE_ERR_ARG, "ge_affinity_pairwise_fwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && H > 0, GE_ERR_ARG, "ge_affinity_pairwise_fwd: non-positive dimension");
    GE_REQUIRE(H % 32 == 0, GE_ERR_SHAPE, "ge_affinity_pairwise_fwd: hidden width H=%d must be a multiple of 32", H);
    const size_t smem = ((size_t)(TI + TJ) * (KC + 4) + KC) * sizeof(float);
    const size_t red = ((size_t)FWD_WARPS * TI * RED_LD + TI) * sizeof(float);
    const size_t bytes = smem > red ? smem : red;
    static size_t cached = 0;
    if (int rc = set_smem(affinity_pairwise_fwd_kernel, bytes, cached, "ge_affinity_pairwise_fwd(attr)")) return rc;
    // 64x64 tiles once they alone fill every SM three times over (3 CTAs/SM); the split-K 32x32 kernel otherwise (a single
    // 250x250 problem: 64 CTAs x 8 warps instead of 16 CTAs)
    const long long big_ctas = (long long)batch * ge::cdiv(N1, BI) * ge::cdiv(N2, BJ);
    if (big_ctas >= 3LL * ge::sm_count()) {
        const size_t bytes64 = ((size_t)(BI + BJ) * (BKC + 4) + BKC + BI) * sizeof(float);
        static size_t cached64 = 0;
        if (int rc = set_smem(affinity_pairwise_fwd64_kernel, bytes64, cached64, "ge_affinity_pairwise_fwd(attr)")) return rc;
        dim3 grid64(ge::cdiv(N2, BJ), ge::cdiv(N1, BI), batch);
        affinity_pairwise_fwd64_kernel<<<grid64, FWD_THREADS, bytes64, (cudaStream_t)stream>>>(A, B, w2, b2, M, N1, N2, H);
        GE_CHECK_LAUNCH("ge_affinity_pairwise_fwd");
        return GE_OK;
    }
    dim3 grid(ge::cdiv(N2, TJ), ge::cdiv(N1, TI), batch);
    affinity_pairwise_fwd_kernel<<<grid, FWD_THREADS, bytes, (cudaStream_t)stream>>>(A, B, w2, b2, M, N1, N2, H);
    GE_CHECK_LAUNCH("ge_affinity_pairwise_fwd");
    return GE_OK;
}

extern "C" size_t ge_affinity_pairwise_bwd_workspace_bytes(int batch, int N1, int N2, int H) {
    if (batch <= 0 || N1 <= 0 || N2 <= 0 || H <= 0) return 0;
    return bwd_plan(batch, N1, N2, H).total() * sizeof(float);
}

extern "C" int ge_affinity_pairwise_bwd(const float* A, const float* B, const float* w2, const float* dM,
                                        float* dA, float* dB, float* dw2, float* db2,
                                        void* workspace, size_t workspace_bytes,
                                        int batch, int N1, int N2, int H, ge_stream_t stream) {
    GE_REQUIRE(A && B && w2 && dM && dA && dB && dw2 && db2 && workspace, GE_ERR_ARG,
               "ge_affinity_pairwise_bwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && H > 0, GE_ERR_ARG, "ge_affinity_pairwise_bwd: non-positive dimension");
    GE_REQUIRE(H % 32 == 0, GE_ERR_SHAPE, "ge_affinity_pairwise_bwd: hidden width H=%d must be a multiple of 32", H);
    const BwdPlan p = bwd_plan(batch, N1, N2, H);
    GE_REQUIRE(workspace_bytes >= p.total() * sizeof(float), GE_ERR_ARG, "ge_affinity_pairwise_bwd: workspace too small");
    const int qmax = p.partial ? p.QR : (N1 > N2 ? N1 : N2);
    const size_t gmax = (size_t)RT * (ge::cdiv(qmax, QT) * QT + 4) * sizeof(float);
    GE_REQUIRE(gmax <= 200 * 1024, GE_ERR_CAPACITY,
               "ge_affinity_pairwise_bwd: N=%d too large for the shared dM tile", N1 > N2 ? N1 : N2);
    cudaStream_t st = (cudaStream_t)stream;
    float* dw_part = static_cast<float*>(workspace);
    float* db_part = dw_part + p.dw_parts;
    float* dA_part = db_part + p.db_parts;
    float* dB_part = dA_part + p.dA_part;
    const int threads = H >= 512 ? 512 : H;
    int n_dw_parts, n_db_parts;
    if (p.partial) {
        const size_t smem = (size_t)RT * (p.QR + 4) * sizeof(float);
        static size_t cached = 0;
        if (int rc = set_smem(affinity_pairwise_bwd_kernel<false, true>, smem, cached, "ge_affinity_pairwise_bwd(attr)")) return rc;
        dim3 grid(p.n_itA, p.n_js, batch);
        affinity_pairwise_bwd_kernel<false, true><<<grid, threads, smem, st>>>(
            A, B, w2, dM, dA_part, dB_part, nullptr, db_part, N1, N2, H, N1, N2, p.QR, 0);
        GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(dA+dB partials)");
        dim3 grid2(ge::cdiv(N1 + N2, FR), batch);
        affinity_bwd_rows_kernel<<<grid2, threads, 0, st>>>(A, B, w2, dA_part, dB_part, dA, dB, dw_part,
                                                           N1, N2, H, p.n_js, p.n_itA);
        GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(rows)");
        n_dw_parts = batch * ge::cdiv(N1 + N2, FR);
        n_db_parts = batch * p.n_itA * p.n_js;
    } else {
        {
            const int qr = ge::cdiv(N2, QT) * QT;
            const size_t smem = (size_t)RT * (qr + 4) * sizeof(float);
            static size_t cached = 0;
            if (int rc = set_smem(affinity_pairwise_bwd_kernel<false, false>, smem, cached, "ge_affinity_pairwise_bwd(attr)")) return rc;
            dim3 grid(p.n_itA, 1, batch);
            affinity_pairwise_bwd_kernel<false, false><<<grid, threads, smem, st>>>(
                A, B, w2, dM, dA, nullptr, dw_part, db_part, N1, N2, H, N1, N2, qr, 1);
            GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(dA)");
        }
        {
            const int qr = ge::cdiv(N1, QT) * QT;
            const size_t smem = (size_t)RT * (qr + 4) * sizeof(float);
            static size_t cached = 0;
            if (int rc = set_smem(affinity_pairwise_bwd_kernel<true, false>, smem, cached, "ge_affinity_pairwise_bwd(attr)")) return rc;
            dim3 grid(p.n_itB, 1, batch);
            affinity_pairwise_bwd_kernel<true, false><<<grid, threads, smem, st>>>(
                B, A, w2, dM, dB, nullptr, dw_part + (size_t)batch * p.n_itA * H, nullptr, N2, N1, H, N1, N2, qr, 1);
            GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(dB)");
        }
        n_dw_parts = batch * (p.n_itA + p.n_itB);
        n_db_parts = batch * p.n_itA;
    }
    affinity_bwd_finalize_kernel<<<ge::cdiv(H, 32), 512, 0, st>>>(dw_part, n_dw_parts, H, db_part, n_db_parts, dw2, db2);
    GE_CHECK_LAUNCH("ge_affinity_pairwise_bwd(finalize)");
    return GE_OK;
}
