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
#include "../../include/graphecho_b200.h"

namespace {

constexpr int TQ = 64;         // query rows per CTA
constexpr int TK = 64;         // keys per tile
constexpr int KC = 16;         // channels per staged chunk
constexpr int KNN_THREADS = 256;
constexpr int XLD = TQ + 4;    // padded leading dim of the staged chunks
constexpr int DLD = TK + 1;    // padded leading dim of the distance tile

// den[b][n] = max(||x[b,:,n]||_2, 1e-12);  sq[b][n] = sum_c (x/den)^2
__global__ void __launch_bounds__(128)
knn_norm_kernel(const float* __restrict__ x, float* __restrict__ den, float* __restrict__ sq, int C, int N) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* xb = x + (size_t)b * C * N + n;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = xb[(size_t)c * N];
        s = fmaf(v, v, s);
    }
    const float d = fmaxf(sqrtf(s), 1e-12f);
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = xb[(size_t)c * N] / d;
        q = fmaf(v, v, q);
    }
    den[(size_t)b * N + n] = d;
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
                 const float* __restrict__ xden, const float* __restrict__ xsq,
                 const float* __restrict__ yden, const float* __restrict__ ysq,
                 const float* __restrict__ rel, long long* __restrict__ out,
                 int B, int C, int N, int M, int K, int dilation) {
    __shared__ __align__(16) float Xs[KC][XLD];
    __shared__ __align__(16) float Ys[KC][XLD];
    __shared__ float Ds[TQ][DLD];
    __shared__ float s_xden[TQ], s_xsq[TQ], s_yden[TK], s_ysq[TK];

    const int b = blockIdx.y, i0 = blockIdx.x * TQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const float* xb = x + (size_t)b * C * N;
    const float* yb = y + (size_t)b * C * M;

    if (tid < TQ) {
        const int i = i0 + tid;
        s_xden[tid] = (i < N) ? xden[(size_t)b * N + i] : 1.f;
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
            s_yden[tid] = (j < M) ? yden[(size_t)b * M + j] : 1.f;
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
                    if (i0 + p < N) xv = xb[(size_t)c * N + i0 + p] / s_xden[p];
                    if (j0 + p < M) yv = yb[(size_t)c * M + j0 + p] / s_yden[p];
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

}  // namespace

extern "C" size_t ge_knn_graph_workspace_bytes(int B, int C, int N, int M) {
    (void)C;
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return (size_t)2 * B * ((size_t)N + M) * sizeof(float);
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
    float* xden = static_cast<float*>(workspace);
    float* xsq = xden + (size_t)B * N;
    float* yden = xsq + (size_t)B * N;
    float* ysq = yden + (size_t)B * M;
    knn_norm_kernel<<<dim3(ge::cdiv(N, 128), B), 128, 0, st>>>(x, xden, xsq, C, N);
    GE_CHECK_LAUNCH("ge_knn_graph(norm x)");
    if (y != nullptr) {
        knn_norm_kernel<<<dim3(ge::cdiv(M, 128), B), 128, 0, st>>>(y, yden, ysq, C, M);
        GE_CHECK_LAUNCH("ge_knn_graph(norm y)");
    } else {
        yden = xden; ysq = xsq; y = x;
    }
    knn_graph_kernel<<<dim3(ge::cdiv(N, TQ), B), KNN_THREADS, 0, st>>>(x, y, xden, xsq, yden, ysq, relative_pos,
                                                                       edge_index, B, C, N, M, K, dilation);
    GE_CHECK_LAUNCH("ge_knn_graph");
    return GE_OK;
}
