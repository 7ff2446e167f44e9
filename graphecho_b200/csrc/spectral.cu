// Spectral bipartition for GModule.update_seed, one CTA per problem, everything in shared memory.
//
// Replaces the reference's CPU round trip through sklearn.cluster.SpectralClustering(2,
// affinity='nearest_neighbors', n_neighbors=m, assign_labels='kmeans')
// (/root/reference/models/graph_matching.py:532-567; the reference's own TODO at :538 asks for a GPU
// version).  For n points (row 0 = the class seed):
//   1. squared distances (self distance forced smallest: include_self=True)
//   2. binary m-NN connectivity by rank counting, symmetrised: A = (C + C^T)/2
//   3. S = D^-1/2 A D^-1/2; its leading eigenvector is sqrt(deg); the second one (the Fiedler
//      direction of the normalised Laplacian) by deflated power iteration on (S + I)/2
//   4. embedding coordinate f = v2 / sqrt(deg); exact 1-D 2-means (best split of the sorted values)
//   5. keep[i-1] = point i falls on the seed's side.
// n <= 192: the n x n matrix lives in shared memory; 192 < n <= 512: block-wise distances and a bit-matrix
// connectivity (second kernel below); larger problems are handled by the caller.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int SP_THREADS = 512;
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int SP_DC = 32;
constexpr int SP_MAX_N = 192;

__global__ void __launch_bounds__(SP_THREADS)
spectral_bipartition_kernel(const float* __restrict__ pts, unsigned char* __restrict__ keep,
                            int n, int d, int knn, int iters) {
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                               // [n][n]   distances, then the normalised adjacency
    float* xs = S + (size_t)n * n;               // [n][SP_DC+1] staging
    float* v = xs + (size_t)n * (SP_DC + 1);     // [n]
    float* w = v + n;                            // [n]
    float* u1 = w + n;                           // [n]  leading eigenvector sqrt(deg)/|.|
    float* dinv = u1 + n;                        // [n]
    float* sorted = dinv + n;                    // [n]
    float* scratch = sorted + n;                 // [32]
    unsigned char* conn = reinterpret_cast<unsigned char*>(scratch + 32);   // [n][n]
    __shared__ int s_split;
    __shared__ int s_rank0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nn = n * n;

    // ---- 1. squared distances ----
    for (int e = tid; e < nn; e += SP_THREADS) S[e] = 0.f;
    for (int d0 = 0; d0 < d; d0 += SP_DC) {
        __syncthreads();
        for (int e = tid; e < n * SP_DC; e += SP_THREADS) {
            const int i = e / SP_DC, c = e - i * SP_DC;
            xs[i * (SP_DC + 1) + c] = (d0 + c < d) ? pts[(size_t)i * d + d0 + c] : 0.f;
        }
        __syncthreads();
        for (int e = tid; e < nn; e += SP_THREADS) {
            const int i = e / n, j = e - i * n;
            const float* a = xs + i * (SP_DC + 1);
            const float* b = xs + j * (SP_DC + 1);
            float acc = S[e];
#pragma unroll 8
            for (int c = 0; c < SP_DC; ++c) {
                const float t = a[c] - b[c];
                acc = fmaf(t, t, acc);
            }
            S[e] = acc;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += SP_THREADS) S[i * n + i] = -1.f;
    __syncthreads();
    // ---- 2. m-NN connectivity: conn[i][j] = rank of j in row i < knn ----
    for (int e = tid; e < nn; e += SP_THREADS) {
        const int i = e / n, j = e - i * n;
        const float dij = S[e];
        const float* row = S + i * n;
        int rank = 0;
        for (int l = 0; l < n; ++l) {
            const float dl = row[l];
            rank += (dl < dij || (dl == dij && l < j)) ? 1 : 0;
        }
        conn[e] = rank < knn ? 1 : 0;
    }
    __syncthreads();
    // ---- 3. symmetrise, degrees, S = D^-1/2 A D^-1/2 ----
    for (int e = tid; e < nn; e += SP_THREADS) {
        const int i = e / n, j = e - i * n;
        S[e] = 0.5f * (float)(conn[e] + conn[j * n + i]);
    }
    __syncthreads();
    for (int i = warp; i < n; i += SP_WARPS) {
        float acc = 0.f;
        for (int j = lane; j < n; j += 32) acc += S[i * n + j];
        acc = ge::warp_sum(acc);
        if (lane == 0) { dinv[i] = rsqrtf(fmaxf(acc, 1e-12f)); u1[i] = sqrtf(fmaxf(acc, 1e-12f)); }
    }
    __syncthreads();
    for (int e = tid; e < nn; e += SP_THREADS) {
        const int i = e / n, j = e - i * n;
        S[e] *= dinv[i] * dinv[j];
    }
    float nrm = 0.f;
    for (int i = tid; i < n; i += SP_THREADS) nrm += u1[i] * u1[i];
    nrm = ge::block_sum(nrm, scratch);
    const float inrm = rsqrtf(nrm);
    for (int i = tid; i < n; i += SP_THREADS) {
        u1[i] *= inrm;
        v[i] = __sinf(1.7f * (float)(i + 1)) + 0.01f * (float)i;      // deterministic, generic start
    }
    __syncthreads();
    // ---- 4. deflated power iteration on (S + I)/2 ----
    for (int it = 0; it < iters; ++it) {
        float dotp = 0.f;
        for (int i = tid; i < n; i += SP_THREADS) dotp += u1[i] * v[i];
        dotp = ge::block_sum(dotp, scratch);
        for (int i = tid; i < n; i += SP_THREADS) v[i] -= dotp * u1[i];
        __syncthreads();
        for (int i = warp; i < n; i += SP_WARPS) {
            float acc = 0.f;
            for (int j = lane; j < n; j += 32) acc = fmaf(S[i * n + j], v[j], acc);
            acc = ge::warp_sum(acc);
            if (lane == 0) w[i] = 0.5f * (acc + v[i]);
        }
        __syncthreads();
        float nn2 = 0.f;
        for (int i = tid; i < n; i += SP_THREADS) nn2 += w[i] * w[i];
        nn2 = ge::block_sum(nn2, scratch);
        const float s = rsqrtf(fmaxf(nn2, 1e-30f));
        for (int i = tid; i < n; i += SP_THREADS) v[i] = w[i] * s;
        __syncthreads();
    }
    // ---- 5. embedding coordinate, exact 1-D 2-means, side of the seed ----
    for (int i = tid; i < n; i += SP_THREADS) w[i] = v[i] * dinv[i];
    __syncthreads();
    for (int i = tid; i < n; i += SP_THREADS) {
        const float fi = w[i];
        int rank = 0;
        for (int l = 0; l < n; ++l) rank += (w[l] < fi || (w[l] == fi && l < i)) ? 1 : 0;
        sorted[rank] = fi;
        if (i == 0) s_rank0 = rank;
        reinterpret_cast<int*>(v)[i] = rank;     // v is free now: keep the ranks
    }
    __syncthreads();
    if (tid == 0) {
        // prefix sums in double: n <= 192, serial is fine
        double tot = 0.0, totsq = 0.0;
        for (int i = 0; i < n; ++i) { tot += sorted[i]; totsq += (double)sorted[i] * sorted[i]; }
        double cs = 0.0, cq = 0.0, best = 1e300;
        int arg = 0;
        for (int i = 0; i < n - 1; ++i) {
            cs += sorted[i]; cq += (double)sorted[i] * sorted[i];
            const double nl = i + 1, nr = n - nl;
            const double sse = (cq - cs * cs / nl) + ((totsq - cq) - (tot - cs) * (tot - cs) / nr);
            if (sse < best) { best = sse; arg = i; }
        }
        s_split = arg;
    }
    __syncthreads();
    const bool left0 = s_rank0 <= s_split;
    for (int i = 1 + tid; i < n; i += SP_THREADS) {
        const bool left = reinterpret_cast<int*>(v)[i] <= s_split;
        keep[i - 1] = (left == left0) ? 1 : 0;
    }
}


// ---------------------------------------------------------------------------------------------
// Large-problem variant (192 < n <= 512): the n x n fp32 matrix no longer fits shared memory, and it is not needed:
// distances are produced one 32-row block at a time (only their per-row ranks matter), the connectivity is kept
// as a bit matrix (n^2/8 bytes), and the normalised adjacency S_ij = (c_ij + c_ji)/2 * dinv_i * dinv_j is rebuilt
// from the bits inside the power iteration.  Same algorithm and tie rules as the kernel above; slower per point
// (a few ms on its one SM) but it runs on the seed stream, off the step's critical path, and replaces the
// ~200-launch torch.linalg.eigh route per class bank.
constexpr int SPB_MAX_N = 512;
constexpr int SPB_RB = 32;                      // rows per distance block

__global__ void __launch_bounds__(SP_THREADS)
spectral_bipartition_big_kernel(const float* __restrict__ pts, unsigned char* __restrict__ keep,
                                int n, int d, int knn, int iters) {
    extern __shared__ __align__(16) float sm[];
    const int W = (n + 31) >> 5;
    unsigned* bits = reinterpret_cast<unsigned*>(sm);          // [n][W]  conn[i][j]
    float* blk = sm + (size_t)n * W;                           // [SPB_RB][n] distances of the current row block
    float* xs = blk + (size_t)SPB_RB * n;                      // [n][SP_DC+1] staging
    float* v = xs + (size_t)n * (SP_DC + 1);                   // [n]
    float* w = v + n;                                          // [n]
    float* u1 = w + n;                                         // [n]
    float* dinv = u1 + n;                                      // [n]
    float* sorted = dinv + n;                                  // [n]
    float* tj = sorted + n;                                    // [n]  dinv_j * v_j
    float* scratch = tj + n;                                   // [32]
    __shared__ int s_split;
    __shared__ int s_rank0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < n * W; e += SP_THREADS) bits[e] = 0u;
    // ---- 1+2. distances and m-NN connectivity, one block of rows at a time ----
    for (int i0 = 0; i0 < n; i0 += SPB_RB) {
        const int rows = min(SPB_RB, n - i0);
        for (int e = tid; e < rows * n; e += SP_THREADS) blk[e] = 0.f;
        for (int d0 = 0; d0 < d; d0 += SP_DC) {
            __syncthreads();
            for (int e = tid; e < n * SP_DC; e += SP_THREADS) {
                const int i = e / SP_DC, c = e - i * SP_DC;
                xs[i * (SP_DC + 1) + c] = (d0 + c < d) ? pts[(size_t)i * d + d0 + c] : 0.f;
            }
            __syncthreads();
            for (int e = tid; e < rows * n; e += SP_THREADS) {
                const int il = e / n, j = e - il * n;
                const float* a = xs + (i0 + il) * (SP_DC + 1);
                const float* b = xs + j * (SP_DC + 1);
                float acc = blk[e];
#pragma unroll 8
                for (int c = 0; c < SP_DC; ++c) {
                    const float t = a[c] - b[c];
                    acc = fmaf(t, t, acc);
                }
                blk[e] = acc;
            }
        }
        __syncthreads();
        for (int il = tid; il < rows; il += SP_THREADS) blk[il * n + i0 + il] = -1.f;      // include_self
        __syncthreads();
        for (int e = tid; e < rows * n; e += SP_THREADS) {
            const int il = e / n, j = e - il * n;
            const float dij = blk[e];
            const float* row = blk + il * n;
            int rank = 0;
            for (int l = 0; l < n; ++l) {
                const float dl = row[l];
                rank += (dl < dij || (dl == dij && l < j)) ? 1 : 0;
            }
            if (rank < knn) atomicOr(&bits[(i0 + il) * W + (j >> 5)], 1u << (j & 31));
        }
        __syncthreads();
    }
    auto adj2 = [&](int i, int j) -> float {       // 2 * A_ij = conn_ij + conn_ji
        return (float)(((bits[i * W + (j >> 5)] >> (j & 31)) & 1u) + ((bits[j * W + (i >> 5)] >> (i & 31)) & 1u));
    };
    // ---- 3. degrees ----
    for (int i = warp; i < n; i += SP_WARPS) {
        float acc = 0.f;
        for (int j = lane; j < n; j += 32) acc += 0.5f * adj2(i, j);
        acc = ge::warp_sum(acc);
        if (lane == 0) { dinv[i] = rsqrtf(fmaxf(acc, 1e-12f)); u1[i] = sqrtf(fmaxf(acc, 1e-12f)); }
    }
    __syncthreads();
    float nrm = 0.f;
    for (int i = tid; i < n; i += SP_THREADS) nrm += u1[i] * u1[i];
    nrm = ge::block_sum(nrm, scratch);
    const float inrm = rsqrtf(nrm);
    for (int i = tid; i < n; i += SP_THREADS) {
        u1[i] *= inrm;
        v[i] = __sinf(1.7f * (float)(i + 1)) + 0.01f * (float)i;
    }
    __syncthreads();
    // ---- 4. deflated power iteration on (S + I)/2, S rebuilt from the bit matrix ----
    for (int it = 0; it < iters; ++it) {
        float dotp = 0.f;
        for (int i = tid; i < n; i += SP_THREADS) dotp += u1[i] * v[i];
        dotp = ge::block_sum(dotp, scratch);
        for (int i = tid; i < n; i += SP_THREADS) {
            v[i] -= dotp * u1[i];
            tj[i] = dinv[i] * v[i];
        }
        __syncthreads();
        for (int i = warp; i < n; i += SP_WARPS) {
            float acc = 0.f;
            for (int j = lane; j < n; j += 32) acc = fmaf(adj2(i, j), tj[j], acc);
            acc = ge::warp_sum(acc);
            if (lane == 0) w[i] = 0.5f * (0.5f * dinv[i] * acc + v[i]);
        }
        __syncthreads();
        float nn2 = 0.f;
        for (int i = tid; i < n; i += SP_THREADS) nn2 += w[i] * w[i];
        nn2 = ge::block_sum(nn2, scratch);
        const float sc = rsqrtf(fmaxf(nn2, 1e-30f));
        for (int i = tid; i < n; i += SP_THREADS) v[i] = w[i] * sc;
        __syncthreads();
    }
    // ---- 5. embedding coordinate, exact 1-D 2-means, side of the seed ----
    for (int i = tid; i < n; i += SP_THREADS) w[i] = v[i] * dinv[i];
    __syncthreads();
    for (int i = tid; i < n; i += SP_THREADS) {
        const float fi = w[i];
        int rank = 0;
        for (int l = 0; l < n; ++l) rank += (w[l] < fi || (w[l] == fi && l < i)) ? 1 : 0;
        sorted[rank] = fi;
        if (i == 0) s_rank0 = rank;
        reinterpret_cast<int*>(v)[i] = rank;
    }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0, totsq = 0.0;
        for (int i = 0; i < n; ++i) { tot += sorted[i]; totsq += (double)sorted[i] * sorted[i]; }
        double cs = 0.0, cq = 0.0, best = 1e300;
        int arg = 0;
        for (int i = 0; i < n - 1; ++i) {
            cs += sorted[i]; cq += (double)sorted[i] * sorted[i];
            const double nl = i + 1, nr = n - nl;
            const double sse = (cq - cs * cs / nl) + ((totsq - cq) - (tot - cs) * (tot - cs) / nr);
            if (sse < best) { best = sse; arg = i; }
        }
        s_split = arg;
    }
    __syncthreads();
    const bool left0 = s_rank0 <= s_split;
    for (int i = 1 + tid; i < n; i += SP_THREADS) {
        const bool left = reinterpret_cast<int*>(v)[i] <= s_split;
        keep[i - 1] = (left == left0) ? 1 : 0;
    }
}

size_t spb_smem(int n) {
    const size_t W = (n + 31) / 32;
    return ((size_t)n * W + (size_t)SPB_RB * n + (size_t)n * (SP_DC + 1) + 6 * (size_t)n + 32) * sizeof(float) + 16;
}

size_t sp_smem(int n) {
    return ((size_t)n * n + (size_t)n * (SP_DC + 1) + 5 * (size_t)n + 32) * sizeof(float) + (size_t)n * n + 16;
}

}  // namespace

extern "C" int ge_spectral_bipartition_max_points(void) { return SPB_MAX_N; }

// pts fp32 [n,d] (row 0 = seed), keep uint8 [n-1]; n_neighbors as sklearn's (self included).
extern "C" int ge_spectral_bipartition(const float* pts, unsigned char* keep, int n, int d, int n_neighbors,
                                       int iterations, ge_stream_t stream) {
    GE_REQUIRE(pts && keep, GE_ERR_ARG, "ge_spectral_bipartition: null pointer");
    GE_REQUIRE(n >= 2 && d > 0 && n_neighbors > 0 && iterations > 0, GE_ERR_ARG, "ge_spectral_bipartition: bad dimension");
    GE_REQUIRE(n <= SPB_MAX_N, GE_ERR_CAPACITY, "ge_spectral_bipartition: n=%d exceeds the in-shared-memory limit %d", n, SPB_MAX_N);
    if (n > SP_MAX_N) {
        const size_t bsmem = spb_smem(n);
        static size_t bcached = 0;
        if (bsmem > bcached) {
            GE_CUDA(cudaFuncSetAttribute(spectral_bipartition_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem),
                    "ge_spectral_bipartition(attr)");
            bcached = bsmem;
        }
        const int kb = n_neighbors < n ? n_neighbors : n;
        spectral_bipartition_big_kernel<<<1, SP_THREADS, bsmem, (cudaStream_t)stream>>>(pts, keep, n, d, kb, iterations);
        GE_CHECK_LAUNCH("ge_spectral_bipartition");
        return GE_OK;
    }
    const size_t smem = sp_smem(n);
    static size_t cached = 0;
    if (smem > cached) {
        GE_CUDA(cudaFuncSetAttribute(spectral_bipartition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "ge_spectral_bipartition(attr)");
        cached = smem;
    }
    const int k = n_neighbors < n ? n_neighbors : n;
    spectral_bipartition_kernel<<<1, SP_THREADS, smem, (cudaStream_t)stream>>>(pts, keep, n, d, k, iterations);
    GE_CHECK_LAUNCH("ge_spectral_bipartition");
    return GE_OK;
}
