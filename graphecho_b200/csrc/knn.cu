// K1 — dense dilated k-NN graph build (ViG).
//
// Replaces DenseDilatedKnnGraph.forward -> (xy_)dense_knn_matrix -> (xy_)pairwise_distance ->
// torch.topk -> DenseDilated  (/root/reference/models/vig.py:232-381): ~8 launches and a
// materialised [B,N,M] distance tensor + full top-k.  Here: one pre-pass for the L2 norms
// (F.normalize semantics: x / max(||x||, 1e-12)) and one tiled kernel that never writes the
// distance matrix: a CTA owns 64 query points, streams 64-key tiles, accumulates the fp32
// inner products with FFMA (no TF32: near-ties must order like the fp32 reference), forms
//   dist = (|x^|^2 + (-2 x^.y^)) + |y^|^2 (+ relative_pos)        [vig.py:270-274, 325-326]
// and keeps a sorted running top-K per query row in registers (K = k*dilation <= 64).
// Ties are broken towards the lower key index (deterministic; torch.topk's tie order is
// implementation-defined).  Output is int64 [2,B,N,k]: [0] = neighbour index (sorted by
// distance, every dilation-th entry), [1] = centre index                [vig.py:328-329, 353].
// Work: 2*B*N*M*C flops; algorithmic bytes 4*B*C*(N+M) + 16*B*N*k.
#include "common.cuh"
#include "knn_tc.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int TQ = 64;         // query rows per CTA
constexpr int TK = 64;         // keys per tile
constexpr int KC = 16;         // channels per staged chunk
constexpr int KNN_THREADS = 256;
constexpr int XLD = TQ + 4;    // padded leading dim of the staged chunks
constexpr int DLD = TK + 1;    // padded leading dim of the distance tile

// xn[b,:,n] = x[b,:,n] / max(||x[b,:,n]||_2, 1e-12)  (F.normalize, vig.py:372-373);  sq[b][n] = sum_c xn^2
__global__ void __launch_bounds__(128)
knn_norm_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ sq, int C, int N) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* xb = x + (size_t)b * C * N + n;
    float* ob = xn + (size_t)b * C * N + n;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = xb[(size_t)c * N];
        s = fmaf(v, v, s);
    }
    const float d = fmaxf(sqrtf(s), 1e-12f);
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = xb[(size_t)c * N] / d;
        ob[(size_t)c * N] = v;
        q = fmaf(v, v, q);
    }
    sq[(size_t)b * N + n] = q;
}

struct TopList {  // sorted ascending; lane l holds slots l and l+32
    float d0, d1;
    int i0, i1;
};

__device__ __forceinline__ void list_insert(TopList& L, float d, int idx, int K, int lane) {
    // position = number of entries <= d  (earlier = lower index wins ties)
    const unsigned m0 = __ballot_sync(ge::kFull, L.d0 <= d);
    const unsigned m1 = __ballot_sync(ge::kFull, L.d1 <= d);
    const int p = __popc(m0) + __popc(m1);
    if (p >= K) return;  // uniform across the warp
    const float up0d = __shfl_up_sync(ge::kFull, L.d0, 1);
    const int up0i = __shfl_up_sync(ge::kFull, L.i0, 1);
    const float up1d = __shfl_up_sync(ge::kFull, L.d1, 1);
    const int up1i = __shfl_up_sync(ge::kFull, L.i1, 1);
    const float last0d = __shfl_sync(ge::kFull, L.d0, 31);
    const int last0i = __shfl_sync(ge::kFull, L.i0, 31);
    const int s1 = lane + 32;
    if (s1 > p) {
        L.d1 = (lane == 0) ? last0d : up1d;
        L.i1 = (lane == 0) ? last0i : up1i;
    } else if (s1 == p) {
        L.d1 = d; L.i1 = idx;
    }
    if (lane > p) {
        L.d0 = up0d; L.i0 = up0i;
    } else if (lane == p) {
        L.d0 = d; L.i0 = idx;
    }
}

__device__ __forceinline__ float list_kth(const TopList& L, int K) {
    const int s = K - 1;
    return (s < 32) ? __shfl_sync(ge::kFull, L.d0, s) : __shfl_sync(ge::kFull, L.d1, s - 32);
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_graph_kernel(const float* __restrict__ x, const float* __restrict__ y,
                 const float* __restrict__ xsq, const float* __restrict__ ysq,
                 const float* __restrict__ rel, long long* __restrict__ out,
                 int B, int C, int N, int M, int K, int dilation) {
    __shared__ __align__(16) float Xs[KC][XLD];
    __shared__ __align__(16) float Ys[KC][XLD];
    __shared__ float Ds[TQ][DLD];
    __shared__ float s_xsq[TQ], s_ysq[TK];

    const int b = blockIdx.y, i0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const float* xb = x + (size_t)b * C * N;
    const float* yb = y + (size_t)b * C * M;

    if (tid < TQ) {
        const int i = i0 + tid;
        s_xsq[tid] = (i < N) ? xsq[(size_t)b * N + i] : 0.f;
    }

    TopList lists[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        lists[r].d0 = INFINITY; lists[r].d1 = INFINITY;
        lists[r].i0 = 0; lists[r].i1 = 0;
    }

    for (int j0 = 0; j0 < M; j0 += TK) {
        __syncthreads();  // previous tile's Ds / s_y* fully consumed
        if (tid < TK) {
            const int j = j0 + tid;
            s_ysq[tid] = (j < M) ? ysq[(size_t)b * M + j] : 0.f;
        }
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;

        for (int c0 = 0; c0 < C; c0 += KC) {
            __syncthreads();
            for (int e = tid; e < KC * TQ; e += KNN_THREADS) {
                const int kc = e / TQ, p = e - kc * TQ;
                const int c = c0 + kc;
                float xv = 0.f, yv = 0.f;
                if (c < C) {
                    if (i0 + p < N) xv = xb[(size_t)c * N + i0 + p];
                    if (j0 + p < M) yv = yb[(size_t)c * M + j0 + p];
                }
                Xs[kc][p] = xv;
                Ys[kc][p] = yv;
            }
            __syncthreads();
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                const float4 a4 = *reinterpret_cast<const float4*>(&Xs[kc][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Ys[kc][tx * 4]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
            }
        }
        // distances into the shared tile
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int il = ty * 4 + a, jl = tx * 4 + c;
                float d = (s_xsq[il] + (-2.f * acc[a][c])) + s_ysq[jl];
                const int i = i0 + il, j = j0 + jl;
                if (rel != nullptr && i < N && j < M) d += rel[(size_t)i * M + j];
                if (j >= M) d = INFINITY;
                Ds[il][jl] = d;
            }
        __syncthreads();
        // running top-K: warp w owns rows 8w..8w+7
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int il = warp * 8 + r;
            TopList& L = lists[r];
            float thr = list_kth(L, K);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int jl = lane + 32 * h;
                const float d = Ds[il][jl];
                unsigned pass = __ballot_sync(ge::kFull, d < thr);
                while (pass) {
                    const int src = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const float dc = __shfl_sync(ge::kFull, d, src);
                    if (dc < thr) {  // thr may have tightened since the ballot (uniform)
                        list_insert(L, dc, j0 + src + 32 * h, K, lane);
                        thr = list_kth(L, K);
                    }
                }
            }
        }
    }
    // write every dilation-th neighbour
    const int kout = K / dilation;
    long long* out0 = out + (size_t)b * N * kout;
    long long* out1 = out + (size_t)B * N * kout + (size_t)b * N * kout;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + warp * 8 + r;
        if (i >= N) continue;
        const TopList& L = lists[r];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int s = lane + 32 * h;
            if (s < K && (s % dilation) == 0) {
                const int o = s / dilation;
                out0[(size_t)i * kout + o] = (h == 0) ? L.i0 : L.i1;
                out1[(size_t)i * kout + o] = i;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Fast path (N%4 == 0, M%4 == 0, C%32 == 0, query tile fits shared memory): the 64-query tile of the
// normalised x stays RESIDENT in shared memory for the whole CTA ([C][68] floats), key tiles stream
// through a 2-stage cp.async ring in 32-channel chunks (16-byte copies, zero-fill past the edge), so the
// FFMA loop never waits on a synchronous global load and nothing is divided in the loop.
constexpr int FKC = 32;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N_)); }

__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_graph_fast_kernel(const float* __restrict__ x, const float* __restrict__ y,
                      const float* __restrict__ xsq, const float* __restrict__ ysq,
                      const float* __restrict__ rel, long long* __restrict__ out,
                      int B, int C, int N, int M, int K, int dilation) {
    extern __shared__ __align__(16) float smem[];
    float* Xr = smem;                              // [C][XLD]
    float* Yb = Xr + (size_t)C * XLD;              // [2][FKC][XLD]
    float* Ds = Yb + 2 * FKC * XLD;                // [TQ][DLD]
    float* s_xsq = Ds + TQ * DLD;                  // [TQ]
    float* s_ysq = s_xsq + TQ;                     // [TK]

    const int b = blockIdx.y, i0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const float* xb = x + (size_t)b * C * N;
    const float* yb = y + (size_t)b * C * M;
    const int nchunks = C / FKC;
    const int ntiles = (M + TK - 1) / TK;

    // resident query tile
    for (int e = tid; e < C * (TQ / 4); e += KNN_THREADS) {
        const int c = e / (TQ / 4), p4 = (e - c * (TQ / 4)) * 4;
        cp_async16(Xr + c * XLD + p4, xb + (size_t)c * N + i0 + p4, i0 + p4 < N);
    }
    cp_async_commit();
    if (tid < TQ) s_xsq[tid] = (i0 + tid < N) ? xsq[(size_t)b * N + i0 + tid] : 0.f;

    auto issue = [&](int step) {            // step = tile * nchunks + chunk
        const int tile = step / nchunks, ch = step - tile * nchunks;
        const int j0 = tile * TK, c0 = ch * FKC;
        float* dst = Yb + (step & 1) * FKC * XLD;
        for (int e = tid; e < FKC * (TK / 4); e += KNN_THREADS) {
            const int kc = e / (TK / 4), p4 = (e - kc * (TK / 4)) * 4;
            cp_async16(dst + kc * XLD + p4, yb + (size_t)(c0 + kc) * M + j0 + p4, j0 + p4 < M);
        }
        cp_async_commit();
    };

    TopList lists[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        lists[r].d0 = INFINITY; lists[r].d1 = INFINITY;
        lists[r].i0 = 0; lists[r].i1 = 0;
    }

    const int nsteps = ntiles * nchunks;
    issue(0);
    float acc[4][4];
    for (int step = 0; step < nsteps; ++step) {
        const int tile = step / nchunks, ch = step - tile * nchunks;
        if (ch == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
        }
        if (step + 1 < nsteps) { issue(step + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                       // chunk `step` (and the resident tile) visible to all
        const float* Xc = Xr + (size_t)ch * FKC * XLD;
        const float* Yc = Yb + (step & 1) * FKC * XLD;
#pragma unroll
        for (int kc = 0; kc < FKC; ++kc) {
            const float4 a4 = *reinterpret_cast<const float4*>(Xc + kc * XLD + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Yc + kc * XLD + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
        }
        if (ch == nchunks - 1) {
            const int j0 = tile * TK;
            if (tid < TK) s_ysq[tid] = (j0 + tid < M) ? ysq[(size_t)b * M + j0 + tid] : 0.f;
            __syncthreads();                   // s_ysq ready; previous tile's Ds consumed
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int il = ty * 4 + a, jl = tx * 4 + c;
                    float d = (s_xsq[il] + (-2.f * acc[a][c])) + s_ysq[jl];
                    const int i = i0 + il, j = j0 + jl;
                    if (rel != nullptr && i < N && j < M) d += rel[(size_t)i * M + j];
                    if (j >= M) d = INFINITY;
                    Ds[il * DLD + jl] = d;
                }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int il = warp * 8 + r;
                TopList& L = lists[r];
                float thr = list_kth(L, K);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jl = lane + 32 * h;
                    const float d = Ds[il * DLD + jl];
                    unsigned pass = __ballot_sync(ge::kFull, d < thr);
                    while (pass) {
                        const int src = __ffs(pass) - 1;
                        pass &= pass - 1;
                        const float dc = __shfl_sync(ge::kFull, d, src);
                        if (dc < thr) {
                            list_insert(L, dc, j0 + src + 32 * h, K, lane);
                            thr = list_kth(L, K);
                        }
                    }
                }
            }
        }
        __syncthreads();                       // everyone done with buffer (step & 1) before it is refilled
    }
    const int kout = K / dilation;
    long long* out0 = out + (size_t)b * N * kout;
    long long* out1 = out + (size_t)B * N * kout + (size_t)b * N * kout;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + warp * 8 + r;
        if (i >= N) continue;
        const TopList& L = lists[r];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int sidx = lane + 32 * h;
            if (sidx < K && (sidx % dilation) == 0) {
                const int o = sidx / dilation;
                out0[(size_t)i * kout + o] = (h == 0) ? L.i0 : L.i1;
                out1[(size_t)i * kout + o] = i;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Wide path (K <= 32): 128 queries x 64 keys per CTA step, 8x4 register tiles (3 shared-memory vector
// loads per 32 FFMA, so the FFMA pipe -- not shared-memory bandwidth -- is the limiter), both operands
// streamed through a 2-stage cp.async ring in 32-channel chunks, running top-K lists in shared memory
// (one slot per lane).
constexpr int WQ = 128;
constexpr int WXLD = WQ + 4;

struct List1 { float d; int i; };   // sorted ascending, lane l holds slot l

__device__ __forceinline__ void list1_insert(List1& L, float d, int idx, int K, int lane) {
    const int p = __popc(__ballot_sync(ge::kFull, L.d <= d));
    if (p >= K) return;
    const float upd = __shfl_up_sync(ge::kFull, L.d, 1);
    const int upi = __shfl_up_sync(ge::kFull, L.i, 1);
    if (lane > p) { L.d = upd; L.i = upi; }
    else if (lane == p) { L.d = d; L.i = idx; }
}

__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_graph_wide_kernel(const float* __restrict__ x, const float* __restrict__ y,
                      const float* __restrict__ xsq, const float* __restrict__ ysq,
                      const float* __restrict__ rel, long long* __restrict__ out,
                      int B, int C, int N, int M, int K, int dilation, int LS) {
    extern __shared__ __align__(16) float smem[];
    float* Xb = smem;                              // [2][FKC][WXLD]
    float* Yb = Xb + 2 * FKC * WXLD;               // [2][FKC][XLD]
    float* Ds = Yb + 2 * FKC * XLD;                // [WQ][DLD]
    float* s_xsq = Ds + WQ * DLD;                  // [WQ]
    float* s_ysq = s_xsq + WQ;                     // [TK]
    float* Ld = s_ysq + TK;                        // [WQ][LS] list distances (LS = 16 or 32 slots)
    int* Li = reinterpret_cast<int*>(Ld + WQ * LS);// [WQ][LS] list indices

    const int b = blockIdx.y, i0 = blockIdx.x * WQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;        // rows ty*8.., cols tx*4..
    const float* xb = x + (size_t)b * C * N;
    const float* yb = y + (size_t)b * C * M;
    const int nchunks = C / FKC;
    const int ntiles = (M + TK - 1) / TK;

    if (tid < WQ) s_xsq[tid] = (i0 + tid < N) ? xsq[(size_t)b * N + i0 + tid] : 0.f;
    for (int e = tid; e < WQ * LS; e += KNN_THREADS) { Ld[e] = INFINITY; Li[e] = 0; }

    auto issue = [&](int step) {
        const int tile = step / nchunks, ch = step - tile * nchunks;
        const int j0 = tile * TK, c0 = ch * FKC;
        float* dx = Xb + (step & 1) * FKC * WXLD;
        float* dy = Yb + (step & 1) * FKC * XLD;
        for (int e = tid; e < FKC * (WQ / 4); e += KNN_THREADS) {
            const int kc = e / (WQ / 4), p4 = (e - kc * (WQ / 4)) * 4;
            cp_async16(dx + kc * WXLD + p4, xb + (size_t)(c0 + kc) * N + i0 + p4, i0 + p4 < N);
        }
        for (int e = tid; e < FKC * (TK / 4); e += KNN_THREADS) {
            const int kc = e / (TK / 4), p4 = (e - kc * (TK / 4)) * 4;
            cp_async16(dy + kc * XLD + p4, yb + (size_t)(c0 + kc) * M + j0 + p4, j0 + p4 < M);
        }
        cp_async_commit();
    };

    const int nsteps = ntiles * nchunks;
    issue(0);
    float acc[8][4];
    for (int step = 0; step < nsteps; ++step) {
        const int tile = step / nchunks, ch = step - tile * nchunks;
        if (ch == 0) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
        }
        if (step + 1 < nsteps) { issue(step + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const float* Xc = Xb + (step & 1) * FKC * WXLD + ty * 8;
        const float* Yc = Yb + (step & 1) * FKC * XLD + tx * 4;
#pragma unroll 8
        for (int kc = 0; kc < FKC; ++kc) {
            const float4 a0 = *reinterpret_cast<const float4*>(Xc + kc * WXLD);
            const float4 a1 = *reinterpret_cast<const float4*>(Xc + kc * WXLD + 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Yc + kc * XLD);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
        }
        if (ch == nchunks - 1) {
            const int j0 = tile * TK;
            if (tid < TK) s_ysq[tid] = (j0 + tid < M) ? ysq[(size_t)b * M + j0 + tid] : 0.f;
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int il = ty * 8 + a, jl = tx * 4 + c;
                    float d = (s_xsq[il] + (-2.f * acc[a][c])) + s_ysq[jl];
                    const int i = i0 + il, j = j0 + jl;
                    if (rel != nullptr && i < N && j < M) d += rel[(size_t)i * M + j];
                    if (j >= M) d = INFINITY;
                    Ds[il * DLD + jl] = d;
                }
            __syncthreads();
            for (int r = 0; r < WQ / 8; ++r) {      // warp w owns rows 16w .. 16w+15
                const int il = warp * (WQ / 8) + r;
                List1 L;
                L.d = lane < LS ? Ld[il * LS + lane] : INFINITY;
                L.i = lane < LS ? Li[il * LS + lane] : 0;
                float thr = __shfl_sync(ge::kFull, L.d, K - 1);
                bool dirty = false;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float d = Ds[il * DLD + lane + 32 * h];
                    unsigned pass = __ballot_sync(ge::kFull, d < thr);
                    while (pass) {
                        const int src = __ffs(pass) - 1;
                        pass &= pass - 1;
                        const float dc = __shfl_sync(ge::kFull, d, src);
                        if (dc < thr) {
                            list1_insert(L, dc, j0 + src + 32 * h, K, lane);
                            thr = __shfl_sync(ge::kFull, L.d, K - 1);
                            dirty = true;
                        }
                    }
                }
                if (dirty && lane < LS) { Ld[il * LS + lane] = L.d; Li[il * LS + lane] = L.i; }
            }
        }
        __syncthreads();
    }
    const int kout = K / dilation;
    long long* out0 = out + (size_t)b * N * kout;
    long long* out1 = out + (size_t)B * N * kout + (size_t)b * N * kout;
    for (int e = tid; e < WQ * LS; e += KNN_THREADS) {
        const int il = e / LS, sidx = e - il * LS;
        const int i = i0 + il;
        if (i < N && sidx < K && (sidx % dilation) == 0) {
            const int o = sidx / dilation;
            out0[(size_t)i * kout + o] = Li[e];
            out1[(size_t)i * kout + o] = i;
        }
    }
}

size_t knn_wide_smem(int LS) {
    return ((size_t)2 * FKC * WXLD + 2 * FKC * XLD + WQ * DLD + WQ + TK + 2 * WQ * LS) * sizeof(float);
}

size_t knn_fast_smem(int C) {
    return ((size_t)C * XLD + 2 * FKC * XLD + TQ * DLD + TQ + TK) * sizeof(float);
}

}  // namespace

namespace {
int g_knn_path = 0;   // 0 = choose, 1 = fp32 FFMA kernels only, 2 = require the tcgen05 kernel
}

extern "C" int ge_knn_graph_set_path(int path) {
    GE_REQUIRE(path >= 0 && path <= 2, GE_ERR_ARG, "ge_knn_graph_set_path: path must be 0, 1 or 2");
    g_knn_path = path;
    return GE_OK;
}

extern "C" size_t ge_knn_graph_workspace_bytes(int B, int C, int N, int M) {
    if (B <= 0 || C <= 0 || N <= 0 || M <= 0) return 0;
    // FFMA path: normalised copies of x and y + their squared norms; tcgen05 path: K-major TF32 hi/lo splits
    const size_t ffma = ((size_t)B * C * ((size_t)N + M) + (size_t)B * ((size_t)N + M)) * sizeof(float);
    const size_t tc = ge::knn_tc_workspace_bytes(B, C, N, M);
    return ffma > tc ? ffma : tc;
}

extern "C" int ge_knn_graph(const float* x, const float* y, const float* relative_pos, long long* edge_index,
                            void* workspace, size_t workspace_bytes,
                            int B, int C, int N, int M, int k, int dilation, ge_stream_t stream) {
    GE_REQUIRE(x && edge_index && workspace, GE_ERR_ARG, "ge_knn_graph: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0 && dilation > 0, GE_ERR_ARG, "ge_knn_graph: bad dimension");
    const int K = k * dilation;
    GE_REQUIRE(K <= 64, GE_ERR_SHAPE, "ge_knn_graph: k*dilation=%d exceeds the 64-entry register list", K);
    GE_REQUIRE(K <= M, GE_ERR_SHAPE, "ge_knn_graph: k*dilation=%d exceeds the number of keys %d", K, M);
    GE_REQUIRE(y != nullptr || N == M, GE_ERR_SHAPE, "ge_knn_graph: self-graph needs M == N");
    GE_REQUIRE(workspace_bytes >= ge_knn_graph_workspace_bytes(B, C, N, M), GE_ERR_ARG, "ge_knn_graph: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_ok = ge::knn_tc_applicable(B, C, N, M, K, relative_pos != nullptr) &&
                       reinterpret_cast<uintptr_t>(workspace) % 128 == 0;
    GE_REQUIRE(g_knn_path != 2 || tc_ok, GE_ERR_SHAPE,
               "ge_knn_graph: the tcgen05 path was required but does not cover B=%d C=%d N=%d M=%d K=%d", B, C, N, M, K);
    if (tc_ok && g_knn_path != 1) return ge::knn_tc_run(x, y, 0, edge_index, workspace, B, C, N, M, K, dilation, st);
    float* xn = static_cast<float*>(workspace);
    float* yn = xn + (size_t)B * C * N;
    float* xsq = yn + (size_t)B * C * M;
    float* ysq = xsq + (size_t)B * N;
    knn_norm_kernel<<<dim3(ge::cdiv(N, 128), B), 128, 0, st>>>(x, xn, xsq, C, N);
    GE_CHECK_LAUNCH("ge_knn_graph(norm x)");
    if (y != nullptr) {
        knn_norm_kernel<<<dim3(ge::cdiv(M, 128), B), 128, 0, st>>>(y, yn, ysq, C, M);
        GE_CHECK_LAUNCH("ge_knn_graph(norm y)");
    } else {
        yn = xn; ysq = xsq;
    }
    const size_t smem = knn_fast_smem(C);
    const bool fast = (N % 4 == 0) && (M % 4 == 0) && (C % FKC == 0) && smem <= 110 * 1024 &&
                      (reinterpret_cast<uintptr_t>(xn) % 16 == 0);
    const bool wide = (N % 4 == 0) && (M % 4 == 0) && (C % FKC == 0) && K <= 32 && N >= WQ &&
                      (reinterpret_cast<uintptr_t>(xn) % 16 == 0);
    if (wide) {
        static size_t wcached = 0;
        const int LS = K <= 16 ? 16 : 32;
        const size_t wsmem = knn_wide_smem(LS);
        if (wsmem > wcached) {
            GE_CUDA(cudaFuncSetAttribute(knn_graph_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem),
                    "ge_knn_graph(attr)");
            wcached = wsmem;
        }
        knn_graph_wide_kernel<<<dim3(ge::cdiv(N, WQ), B), KNN_THREADS, wsmem, st>>>(xn, yn, xsq, ysq, relative_pos,
                                                                                    edge_index, B, C, N, M, K, dilation, LS);
    } else if (fast) {
        static size_t cached = 0;
        if (smem > cached) {
            GE_CUDA(cudaFuncSetAttribute(knn_graph_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "ge_knn_graph(attr)");
            cached = smem;
        }
        knn_graph_fast_kernel<<<dim3(ge::cdiv(N, TQ), B), KNN_THREADS, smem, st>>>(xn, yn, xsq, ysq, relative_pos,
                                                                                   edge_index, B, C, N, M, K, dilation);
    } else {
        knn_graph_kernel<<<dim3(ge::cdiv(N, TQ), B), KNN_THREADS, 0, st>>>(xn, yn, xsq, ysq, relative_pos,
                                                                           edge_index, B, C, N, M, K, dilation);
    }
    GE_CHECK_LAUNCH("ge_knn_graph");
    return GE_OK;
}

// Node-major (channels-last) entry: x [B,N,C], y [B,M,C] or NULL, fp32 or bf16 -- the layout the FPN feature maps
// already have, so the Grapher needs no [B,C,N] transpose.  Tensor-core path only.
extern "C" int ge_knn_graph_nmajor_supported(int B, int C, int N, int M, int k, int dilation) {
    return ge::knn_tc_applicable(B, C, N, M, k * dilation, false) && C % 8 == 0 ? 1 : 0;
}

extern "C" int ge_knn_graph_nmajor(const void* x, const void* y, int dtype, long long* edge_index,
                                   void* workspace, size_t workspace_bytes,
                                   int B, int C, int N, int M, int k, int dilation, ge_stream_t stream) {
    GE_REQUIRE(x && edge_index && workspace, GE_ERR_ARG, "ge_knn_graph_nmajor: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0 && dilation > 0, GE_ERR_ARG, "ge_knn_graph_nmajor: bad dimension");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_knn_graph_nmajor: unsupported dtype %d", dtype);
    const int K = k * dilation;
    GE_REQUIRE(K <= M, GE_ERR_SHAPE, "ge_knn_graph_nmajor: k*dilation=%d exceeds the number of keys %d", K, M);
    GE_REQUIRE(y != nullptr || N == M, GE_ERR_SHAPE, "ge_knn_graph_nmajor: self-graph needs M == N");
    GE_REQUIRE(ge_knn_graph_nmajor_supported(B, C, N, M, k, dilation), GE_ERR_SHAPE,
               "ge_knn_graph_nmajor: the tcgen05 path does not cover B=%d C=%d N=%d M=%d K=%d", B, C, N, M, K);
    GE_REQUIRE(workspace_bytes >= ge::knn_tc_workspace_bytes(B, C, N, M), GE_ERR_ARG, "ge_knn_graph_nmajor: workspace too small");
    GE_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 128 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
               (y == nullptr || reinterpret_cast<uintptr_t>(y) % 16 == 0), GE_ERR_ARG, "ge_knn_graph_nmajor: misaligned pointer");
    return ge::knn_tc_run(x, y, dtype == GE_DTYPE_F32 ? 1 : 2, edge_index, workspace, B, C, N, M, K, dilation, (cudaStream_t)stream);
}
