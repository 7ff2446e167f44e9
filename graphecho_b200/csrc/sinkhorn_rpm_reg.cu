// K4 (register-resident path) — instance-norm -> slack-padded Sinkhorn (n_iters row/col passes) -> exp
// and its exact unrolled adjoint, with the matrix held in REGISTERS for the whole loop.
//
// Same operator as sinkhorn_rpm.cu (/root/reference/models/graph_matching.py:574-575, 637-676), restated
// in the scaling (exponent) domain.  With z = instnorm(M), K = exp(z), u = exp(-r), v = exp(-c) the
// reference iteration  r_i <- LSE_j(z_ij - c_j U {0}),  c_j <- LSE_i(z_ij - r_i U {0})  is exactly
//     u_i <- 1 / (1 + sum_j K_ij v_j)          (row pass; the 1 is the slack column)
//     v_j <- 1 / (1 + sum_i K_ij u_i)          (column pass; the 1 is the slack row)
// and the result is P_ij = K_ij u_i v_j.  Because of the slack entries u, v stay in (0, 1], so the only
// overflow hazard is K itself: the kernel measures max z first and, when it exceeds kZMax, leaves the
// problem to the log-domain kernel (stats[2] = 1) — bit-for-bit the round-1 behaviour for such inputs.
// Inside the loop there is no exp/log at all: two matrix-vector products per iteration, 2 FMAs per
// matrix element, K read from registers (a TR x TC tile per thread), so the loop is bounded by the FP32
// pipe and the reduction/synchronisation latency, not by MUFU or shared-memory bandwidth.
//
// Layout: 16 warps per CTA; warp w owns TR rows, lane l owns the columns {128 q + 4 l .. +3}, q < QC
// (each 128-bit access of a warp is one contiguous 512-byte row segment).  Rows are split over the CTAs
// of a thread-block cluster; the column pass reduces 16 warps through shared memory, then the CTAs
// through distributed shared memory (one cluster barrier per iteration).
//
// Backward: the adjoint of the unrolled loop, replayed from the saved u_t, v_t.  All softmax weights of
// the log-domain form are K_ij u_i v_j products, so dz = K o F with F accumulated as rank-1 updates:
//     x_j = gc_j v^t_j ;  gr_i -= u^t_i (K x)_i ;  y_i = gr_i u^t_i ;  F += u^t x^T + y (v^{t-1})^T ;
//     gc_j = -v^{t-1}_j (K^T y)_j
// again two matrix-vector products (+ one rank-2 update) per iteration, no exp.
#include "common.cuh"
#include "sinkhorn_rpm_reg.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int FW = ge::rpmreg::kWarps;
constexpr int FT = FW * 32;
constexpr float IN_EPS = 1e-5f;      // nn.InstanceNorm2d default eps (graph_matching.py:177)
constexpr float kZMax = 80.f;        // exp(80) * 512 terms stays finite in fp32

// Sum TR per-lane values across the 32 lanes of a warp with TR/2 + TR/4 + .. + 1 + log2(32/TR) shuffles
// (halving exchange, then butterflies).  On return every lane holds the complete sum of row `row`.
template <int TR>
__device__ __forceinline__ float lane_transpose_reduce(float (&v)[TR], int lane, int& row) {
    row = 0;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int off = 16 >> s;
        const int nb = TR >> s;
        if (nb > 1) {
            const int n = nb >> 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int k = 0; k < n; ++k) {
                const float send = upper ? v[k] : v[k + n];
                const float keep = upper ? v[k + n] : v[k];
                v[k] = keep + __shfl_xor_sync(ge::kFull, send, off);
            }
            if (upper) row += n;
        } else {
            v[0] += __shfl_xor_sync(ge::kFull, v[0], off);
        }
    }
    return v[0];
}

struct Xchg {
    float sum0, sum1, mx;
};

// Cluster-wide (sum, sum, max) of per-thread values.  Deterministic (rank order), identical in every CTA.
__device__ __forceinline__ Xchg cluster_reduce3(cg::cluster_group& cl, unsigned cs, float* s_red, float* s_xc,
                                                int& parity, float a, float b, float m) {
    Xchg r;
    r.sum0 = ge::block_sum(a, s_red);
    r.sum1 = ge::block_sum(b, s_red);
    r.mx = ge::block_max(m, s_red);
    if (cs == 1) return r;
    float* mine = s_xc + parity * 4;
    if (threadIdx.x == 0) { mine[0] = r.sum0; mine[1] = r.sum1; mine[2] = r.mx; }
    cl.sync();
    r.sum0 = 0.f; r.sum1 = 0.f; r.mx = -INFINITY;
    for (unsigned q = 0; q < cs; ++q) {
        const float* peer = cl.map_shared_rank(mine, q);
        r.sum0 += peer[0];
        r.sum1 += peer[1];
        r.mx = fmaxf(r.mx, peer[2]);
    }
    parity ^= 1;
    return r;
}

template <int TR, int QC>
struct Tile {
    static constexpr int TC = 4 * QC;
    static constexpr int NC = 128 * QC;
};

// Column reduction of per-thread partials q[TC] (sum over the TR rows of the thread): 16 warps through
// s_part, then the cluster through s_x.  Returns, in threads tid < NC, the cluster-wide sum of column tid.
template <int QC>
__device__ __forceinline__ float column_reduce(cg::cluster_group& cl, unsigned cs, const float (&q)[4 * QC],
                                               float* s_part, float* s_x, int& parity) {
    constexpr int NC = 128 * QC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c4 = 0; c4 < QC; ++c4)
        *reinterpret_cast<float4*>(s_part + warp * NC + c4 * 128 + 4 * lane) =
            make_float4(q[4 * c4], q[4 * c4 + 1], q[4 * c4 + 2], q[4 * c4 + 3]);
    __syncthreads();
    float tot = 0.f;
    if (tid < NC) {
#pragma unroll
        for (int w = 0; w < FW; ++w) tot += s_part[w * NC + tid];
    }
    if (cs > 1) {
        float* mine = s_x + parity * NC;
        if (tid < NC) mine[tid] = tot;
        cl.sync();
        if (tid < NC) {
            tot = 0.f;
            for (unsigned r = 0; r < cs; ++r) tot += cl.map_shared_rank(mine, r)[tid];
        }
        parity ^= 1;
    }
    return tot;
}

// Loads the thread's TR x TC tile of a row-major [N1, N2] matrix (0 outside).
template <int TR, int QC>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int N2, int grow0, int rpw, int rows_end,
                                          bool vec, float (&t)[TR][4 * QC]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        const int gr = grow0 + r;
        const bool rv = r < rpw && gr < rows_end;
#pragma unroll
        for (int c4 = 0; c4 < QC; ++c4) {
            const int col = c4 * 128 + 4 * lane;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rv && col < N2) {
                const float* p = src + (size_t)gr * N2 + col;
                if (vec) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    v.x = __ldg(p);
                    if (col + 1 < N2) v.y = __ldg(p + 1);
                    if (col + 2 < N2) v.z = __ldg(p + 2);
                    if (col + 3 < N2) v.w = __ldg(p + 3);
                }
            }
            t[r][4 * c4] = v.x; t[r][4 * c4 + 1] = v.y; t[r][4 * c4 + 2] = v.z; t[r][4 * c4 + 3] = v.w;
        }
    }
}

template <int TR, int QC>
__device__ __forceinline__ void store_tile(float* __restrict__ dst, int N2, int grow0, int rpw, int rows_end,
                                           bool vec, const float (&t)[TR][4 * QC]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        const int gr = grow0 + r;
        if (!(r < rpw && gr < rows_end)) continue;
#pragma unroll
        for (int c4 = 0; c4 < QC; ++c4) {
            const int col = c4 * 128 + 4 * lane;
            if (col >= N2) continue;
            float* p = dst + (size_t)gr * N2 + col;
            if (vec) {
                *reinterpret_cast<float4*>(p) = make_float4(t[r][4 * c4], t[r][4 * c4 + 1], t[r][4 * c4 + 2], t[r][4 * c4 + 3]);
            } else {
                p[0] = t[r][4 * c4];
                if (col + 1 < N2) p[1] = t[r][4 * c4 + 1];
                if (col + 2 < N2) p[2] = t[r][4 * c4 + 2];
                if (col + 3 < N2) p[3] = t[r][4 * c4 + 3];
            }
        }
    }
}

template <int TR, int QC>
__device__ __forceinline__ bool elem_valid(int r, int c, int grow0, int rpw, int rows_end, int N2) {
    const int lane = threadIdx.x & 31;
    const int col = (c >> 2) * 128 + 4 * lane + (c & 3);
    return r < rpw && grow0 + r < rows_end && col < N2;
}

template <int TR, int QC>
__global__ void __launch_bounds__(FT, (TR * 4 * QC <= 32) ? 2 : 1)
rpm_reg_fwd_kernel(const float* __restrict__ M, float* __restrict__ P, float* __restrict__ hist_u,
                   float* __restrict__ hist_v, float* __restrict__ stats, int N1, int N2, int n_iters,
                   int apply_instnorm) {
    constexpr int TC = 4 * QC, NC = 128 * QC;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks(), rank = cl.block_rank();
    const int prob = blockIdx.x / cs;
    const int Rb = ge::cdiv(N1, (int)cs);            // rows of this CTA (balanced over the cluster)
    const int rpw = ge::cdiv(Rb, FW);                // rows per warp, <= TR (host guarantees)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_begin = rank * Rb;
    const int rows_end = min(N1, row_begin + Rb);
    const int grow0 = row_begin + warp * rpw;        // first global row of this warp

    __shared__ __align__(16) float s_part[FW * NC];
    __shared__ __align__(16) float s_x[2 * NC];
    __shared__ __align__(16) float s_v[NC];
    __shared__ __align__(16) float s_u[FW * TR];
    __shared__ float s_red[32];
    __shared__ float s_xc[8];

    M += (size_t)prob * N1 * N2;
    P += (size_t)prob * N1 * N2;
    hist_u += (size_t)prob * max(n_iters, 1) * N1;
    hist_v += (size_t)prob * max(n_iters, 1) * N2;
    stats += (size_t)prob * 4;
    const bool vec = (N2 & 3) == 0;

    float K[TR][TC];
    load_tile<TR, QC>(M, N2, grow0, rpw, rows_end, vec, K);

    // ---- instance-norm statistics (two-pass, from registers) and the overflow guard ----
    float lsum = 0.f, lmax = -INFINITY;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c)
            if (elem_valid<TR, QC>(r, c, grow0, rpw, rows_end, N2)) {
                lsum += K[r][c];
                lmax = fmaxf(lmax, K[r][c]);
            }
    int xpar = 0;
    float mean = 0.f, rstd = 1.f, zmax;
    if (apply_instnorm) {
        Xchg a = cluster_reduce3(cl, cs, s_red, s_xc, xpar, lsum, 0.f, lmax);
        const float inv_n = 1.f / ((float)N1 * (float)N2);
        mean = a.sum0 * inv_n;
        float lsq = 0.f;
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c)
                if (elem_valid<TR, QC>(r, c, grow0, rpw, rows_end, N2)) {
                    const float d = K[r][c] - mean;
                    lsq += d * d;
                }
        Xchg b = cluster_reduce3(cl, cs, s_red, s_xc, xpar, lsq, 0.f, 0.f);
        rstd = 1.f / sqrtf(b.sum0 * inv_n + IN_EPS);
        zmax = (a.mx - mean) * rstd;
    } else {
        Xchg a = cluster_reduce3(cl, cs, s_red, s_xc, xpar, 0.f, 0.f, lmax);
        zmax = a.mx;
    }
    const bool defer = !(zmax <= kZMax);             // also true for NaN: the log-domain kernel decides
    if (rank == 0 && tid == 0) {
        stats[0] = mean; stats[1] = rstd; stats[2] = defer ? 1.f : 0.f; stats[3] = 0.f;
    }
    if (defer) {
        if (cs > 1) cl.sync();                       // peers may still be reading s_xc
        return;
    }
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c)
            K[r][c] = elem_valid<TR, QC>(r, c, grow0, rpw, rows_end, N2) ? expf((K[r][c] - mean) * rstd) : 0.f;

    float vreg[TC], ureg[TR];
#pragma unroll
    for (int c = 0; c < TC; ++c) vreg[c] = 1.f;
#pragma unroll
    for (int r = 0; r < TR; ++r) ureg[r] = 1.f;
    int cpar = 0;

    for (int t = 0; t < n_iters; ++t) {
        // row pass  [graph_matching.py:661-664]
        float p[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < TC; ++c) acc = fmaf(K[r][c], vreg[c], acc);
            p[r] = acc;
        }
        int row;
        const float tot = lane_transpose_reduce<TR>(p, lane, row);
        const float uval = __frcp_rn(1.f + tot);
        if ((lane & (32 / TR - 1)) == 0) {
            s_u[warp * TR + row] = uval;
            if (row < rpw && grow0 + row < rows_end) hist_u[(size_t)t * N1 + grow0 + row] = uval;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < TR; ++r) ureg[r] = s_u[warp * TR + r];
        // column pass  [graph_matching.py:666-669]
        float q[TC];
#pragma unroll
        for (int c = 0; c < TC; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < TR; ++r) acc = fmaf(K[r][c], ureg[r], acc);
            q[c] = acc;
        }
        const float ctot = column_reduce<QC>(cl, cs, q, s_part, s_x, cpar);
        if (tid < NC) {
            const float vval = __frcp_rn(1.f + ctot);
            s_v[tid] = vval;
            if (rank == 0 && tid < N2) hist_v[(size_t)t * N2 + tid] = vval;
        }
        __syncthreads();
#pragma unroll
        for (int c4 = 0; c4 < QC; ++c4) {
            const float4 v4 = *reinterpret_cast<const float4*>(s_v + c4 * 128 + 4 * lane);
            vreg[4 * c4] = v4.x; vreg[4 * c4 + 1] = v4.y; vreg[4 * c4 + 2] = v4.z; vreg[4 * c4 + 3] = v4.w;
        }
    }
    // ---- crop + exp  [graph_matching.py:575, 676]:  P = K u v ----
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c) K[r][c] *= ureg[r] * vreg[c];
    store_tile<TR, QC>(P, N2, grow0, rpw, rows_end, vec, K);
    if (cs > 1) cl.sync();                           // peers may still be reading this CTA's partials
}

template <int TR, int QC>
__global__ void __launch_bounds__(FT, 1)
rpm_reg_bwd_kernel(const float* __restrict__ M, const float* __restrict__ G, const float* __restrict__ hist_u,
                   const float* __restrict__ hist_v, const float* __restrict__ stats, float* __restrict__ dM,
                   int N1, int N2, int n_iters, int apply_instnorm) {
    constexpr int TC = 4 * QC, NC = 128 * QC;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks(), rank = cl.block_rank();
    const int prob = blockIdx.x / cs;
    if (stats[(size_t)prob * 4 + 2] != 0.f) return;  // the forward deferred to the log-domain kernel
    const int Rb = ge::cdiv(N1, (int)cs);
    const int rpw = ge::cdiv(Rb, FW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_begin = rank * Rb;
    const int rows_end = min(N1, row_begin + Rb);
    const int grow0 = row_begin + warp * rpw;
    constexpr int RL = FW * TR;                      // local row slots

    extern __shared__ __align__(16) float dyn[];     // hu [T][RL] | hv [T][NC]
    float* hu = dyn;
    float* hv = dyn + (size_t)n_iters * RL;
    __shared__ __align__(16) float s_part[FW * NC];
    __shared__ __align__(16) float s_x[2 * NC];
    __shared__ __align__(16) float s_gc[NC];
    __shared__ __align__(16) float s_y[RL];
    __shared__ float s_red[32];
    __shared__ float s_xc[8];

    M += (size_t)prob * N1 * N2;
    G += (size_t)prob * N1 * N2;
    dM += (size_t)prob * N1 * N2;
    hist_u += (size_t)prob * max(n_iters, 1) * N1;
    hist_v += (size_t)prob * max(n_iters, 1) * N2;
    stats += (size_t)prob * 4;
    const float mean = stats[0], rstd = stats[1];
    const bool vec = (N2 & 3) == 0;

    // saved scaling vectors -> shared memory (slot = warp * TR + r, so a warp's rows are contiguous)
    for (int e = tid; e < n_iters * RL; e += FT) {
        const int t = e / RL, s = e - t * RL;
        const int w = s / TR, r = s - w * TR;
        const int gr = row_begin + w * rpw + r;
        hu[e] = (r < rpw && gr < rows_end) ? hist_u[(size_t)t * N1 + gr] : 1.f;
    }
    for (int e = tid; e < n_iters * NC; e += FT) {
        const int t = e / NC, j = e - t * NC;
        hv[e] = (j < N2) ? hist_v[(size_t)t * N2 + j] : 1.f;
    }

    float K[TR][TC], F[TR][TC];
    load_tile<TR, QC>(M, N2, grow0, rpw, rows_end, vec, K);
    load_tile<TR, QC>(G, N2, grow0, rpw, rows_end, vec, F);
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c)
            K[r][c] = elem_valid<TR, QC>(r, c, grow0, rpw, rows_end, N2) ? expf((K[r][c] - mean) * rstd) : 0.f;
    __syncthreads();

    int cpar = 0;
    if (n_iters > 0) {
        float vcur[TC], ureg[TR];
        const float* hvT = hv + (size_t)(n_iters - 1) * NC;
        const float* huT = hu + (size_t)(n_iters - 1) * RL;
#pragma unroll
        for (int c4 = 0; c4 < QC; ++c4) {
            const float4 v4 = *reinterpret_cast<const float4*>(hvT + c4 * 128 + 4 * lane);
            vcur[4 * c4] = v4.x; vcur[4 * c4 + 1] = v4.y; vcur[4 * c4 + 2] = v4.z; vcur[4 * c4 + 3] = v4.w;
        }
#pragma unroll
        for (int r = 0; r < TR; ++r) ureg[r] = huT[warp * TR + r];
        // F0 = G u_T v_T  (so that E = K o F = G o P);  gr = -rowsum(E), gc = -colsum(E)
        float p[TR], q[TC];
#pragma unroll
        for (int c = 0; c < TC; ++c) q[c] = 0.f;
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < TC; ++c) {
                F[r][c] *= ureg[r] * vcur[c];
                const float e = K[r][c] * F[r][c];
                acc += e;
                q[c] += e;
            }
            p[r] = acc;
        }
        int row;
        float gr = -lane_transpose_reduce<TR>(p, lane, row);   // adjoint of r_T, held per row-lane group
        {
            const float ctot = column_reduce<QC>(cl, cs, q, s_part, s_x, cpar);
            if (tid < NC) s_gc[tid] = -ctot;
            __syncthreads();
        }
        for (int t = n_iters - 1; t >= 0; --t) {
            const float* hut = hu + (size_t)t * RL;
            float vprev[TC], x[TC];
#pragma unroll
            for (int c4 = 0; c4 < QC; ++c4) {
                float4 v4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (t > 0) v4 = *reinterpret_cast<const float4*>(hv + (size_t)(t - 1) * NC + c4 * 128 + 4 * lane);
                vprev[4 * c4] = v4.x; vprev[4 * c4 + 1] = v4.y; vprev[4 * c4 + 2] = v4.z; vprev[4 * c4 + 3] = v4.w;
                const float4 g4 = *reinterpret_cast<const float4*>(s_gc + c4 * 128 + 4 * lane);
                x[4 * c4] = g4.x * vcur[4 * c4]; x[4 * c4 + 1] = g4.y * vcur[4 * c4 + 1];
                x[4 * c4 + 2] = g4.z * vcur[4 * c4 + 2]; x[4 * c4 + 3] = g4.w * vcur[4 * c4 + 3];
            }
#pragma unroll
            for (int r = 0; r < TR; ++r) ureg[r] = hut[warp * TR + r];
            // adjoint of the column pass c_t = LSE_i(z - r_t):  gr_i -= u_i (K x)_i
#pragma unroll
            for (int r = 0; r < TR; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < TC; ++c) acc = fmaf(K[r][c], x[c], acc);
                p[r] = acc;
            }
            const float srow = lane_transpose_reduce<TR>(p, lane, row);
            const float urow = hut[warp * TR + row];
            gr = fmaf(-urow, srow, gr);
            if ((lane & (32 / TR - 1)) == 0) s_y[warp * TR + row] = gr * urow;
            __syncwarp();
            float yreg[TR];
#pragma unroll
            for (int r = 0; r < TR; ++r) yreg[r] = s_y[warp * TR + r];
            // F += u x^T + y vprev^T ;  adjoint of the row pass r_t = LSE_j(z - c_{t-1}):  gc_j = -vprev_j (K^T y)_j
#pragma unroll
            for (int c = 0; c < TC; ++c) q[c] = 0.f;
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int c = 0; c < TC; ++c) {
                    F[r][c] = fmaf(ureg[r], x[c], F[r][c]);
                    F[r][c] = fmaf(yreg[r], vprev[c], F[r][c]);
                    q[c] = fmaf(K[r][c], yreg[r], q[c]);
                }
            const float ctot = column_reduce<QC>(cl, cs, q, s_part, s_x, cpar);
            if (tid < NC) {
                const float vp = (t > 0) ? hv[(size_t)(t - 1) * NC + tid] : 1.f;
                s_gc[tid] = -vp * ctot;
            }
            __syncthreads();
            gr = 0.f;
#pragma unroll
            for (int c = 0; c < TC; ++c) vcur[c] = vprev[c];
            __syncwarp();                            // s_y is rewritten next iteration
        }
    }
    // dz = K o F  (n_iters == 0: F = G, P = K)
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c) F[r][c] *= K[r][c];

    if (apply_instnorm) {
        // instance-norm adjoint: dM = rstd * (dz - mean(dz) - z * mean(dz o z)); z from a second read of M
        load_tile<TR, QC>(M, N2, grow0, rpw, rows_end, vec, K);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c) {
                const bool ok = elem_valid<TR, QC>(r, c, grow0, rpw, rows_end, N2);
                K[r][c] = ok ? (K[r][c] - mean) * rstd : 0.f;
                a += F[r][c];
                b = fmaf(F[r][c], K[r][c], b);
            }
        int xpar = 0;
        Xchg s = cluster_reduce3(cl, cs, s_red, s_xc, xpar, a, b, 0.f);
        const float inv_n = 1.f / ((float)N1 * (float)N2);
        const float m1 = s.sum0 * inv_n, m2 = s.sum1 * inv_n;
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int c = 0; c < TC; ++c) F[r][c] = rstd * (F[r][c] - m1 - K[r][c] * m2);
    }
    store_tile<TR, QC>(dM, N2, grow0, rpw, rows_end, vec, F);
    if (cs > 1) cl.sync();
}

template <typename Kern, typename... Args>
int launch(Kern kernel, const char* name, int cs, int batch, size_t smem, cudaStream_t st, Args... args) {
    static size_t max_smem_set = 0;                  // per template instantiation
    if (smem > max_smem_set) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        max_smem_set = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * batch));
    cfg.blockDim = dim3(FT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GE_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...), name);
    ge_count_launches(1);
    return GE_OK;
}

}  // namespace

namespace ge { namespace rpmreg {

bool fits(int N1, int N2, int n_iters) {
    if (N2 > 256 || N1 > kMaxCluster * kWarps * 4) return false;
    const int nc = N2 <= 128 ? 128 : 256;
    return (size_t)n_iters * (kWarps * 4 + nc) * sizeof(float) <= 160 * 1024;
}

int fwd(const float* M, float* P, float* hist_u, float* hist_v, float* stats, int batch, int N1, int N2,
        int n_iters, int apply_instnorm, int rows_per_thread, cudaStream_t st) {
    const char* name = "ge_sinkhorn_rpm_fwd[reg]";
    const bool wide = N2 > 128;
    // 8 rows per thread halves the cluster size (fewer exchanges) at 1 CTA/SM; 4 rows keep 2 CTAs/SM
    int tr = rows_per_thread;
    if (tr != 4 && tr != 8) tr = 4;
    if (tr == 8 && cdiv(N1, kWarps * 8) > kMaxCluster) tr = 4;
    const int cs = cdiv(N1, kWarps * tr);
    if (tr == 8) {
        if (wide) return launch(rpm_reg_fwd_kernel<8, 2>, name, cs, batch, 0, st, M, P, hist_u, hist_v, stats, N1, N2, n_iters, apply_instnorm);
        return launch(rpm_reg_fwd_kernel<8, 1>, name, cs, batch, 0, st, M, P, hist_u, hist_v, stats, N1, N2, n_iters, apply_instnorm);
    }
    if (wide) return launch(rpm_reg_fwd_kernel<4, 2>, name, cs, batch, 0, st, M, P, hist_u, hist_v, stats, N1, N2, n_iters, apply_instnorm);
    return launch(rpm_reg_fwd_kernel<4, 1>, name, cs, batch, 0, st, M, P, hist_u, hist_v, stats, N1, N2, n_iters, apply_instnorm);
}

int bwd(const float* M, const float* G, const float* hist_u, const float* hist_v, const float* stats, float* dM,
        int batch, int N1, int N2, int n_iters, int apply_instnorm, cudaStream_t st) {
    const char* name = "ge_sinkhorn_rpm_bwd[reg]";
    const bool wide = N2 > 128;
    const int cs = cdiv(N1, kWarps * 4);
    const size_t smem = (size_t)n_iters * (kWarps * 4 + (wide ? 256 : 128)) * sizeof(float);
    if (wide) return launch(rpm_reg_bwd_kernel<4, 2>, name, cs, batch, smem, st, M, G, hist_u, hist_v, stats, dM, N1, N2, n_iters, apply_instnorm);
    return launch(rpm_reg_bwd_kernel<4, 1>, name, cs, batch, smem, st, M, G, hist_u, hist_v, stats, dM, N1, N2, n_iters, apply_instnorm);
}

} }  // namespace ge::rpmreg
