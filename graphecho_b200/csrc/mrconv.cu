// K2 — max-relative graph aggregation (the gather half of MRConv2d).
//
// Replaces  batched_index_select x2 -> max_k(x_j - x_i) -> interleaving cat/reshape
// (/root/reference/models/vig.py:96-104, 209-229): the reference materialises two
// [B,C,N,k] gathers (k-fold read/write amplification) plus transposes.  Here one pass reads
// x (and y) once and writes the channel-interleaved [B,2C,N] tensor
//     out[b,2c,n] = x[b,c,n]      out[b,2c+1,n] = max_k ( y[b,c,idx0[b,n,k]] - x[b,c,idx1[b,n,k]] )
// that the grouped 1x1 conv (BasicConv, vig.py:476-500) consumes, and records the arg-max
// neighbour slot (uint8) so the backward is a pure scatter.
// Algorithmic bytes: 4*B*C*(N+M) + 8*B*N*k in, 8*B*C*N out.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

constexpr int MR_THREADS = 128;

__global__ void __launch_bounds__(MR_THREADS)
mrconv_gather_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                         const long long* __restrict__ idx0, const long long* __restrict__ idx1,
                         float* __restrict__ out, unsigned char* __restrict__ argk,
                         int C, int N, int M, int k) {
    extern __shared__ int s_idx[];          // [2][MR_THREADS][k]  (neighbour, centre)
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * MR_THREADS;
    const int n = n0 + threadIdx.x;
    int* nb = s_idx;
    int* ct = s_idx + MR_THREADS * k;
    for (int e = threadIdx.x; e < MR_THREADS * k; e += MR_THREADS) {
        const int p = e / k, kk = e - p * k;
        int v0 = 0, v1 = n0 + p;
        if (n0 + p < N) {
            v0 = (int)idx0[((size_t)b * N + n0 + p) * k + kk];
            if (idx1 != nullptr) v1 = (int)idx1[((size_t)b * N + n0 + p) * k + kk];
        }
        nb[e] = v0;
        ct[e] = v1;
    }
    __syncthreads();
    if (n >= N) return;
    const float* xb = x + (size_t)b * C * N;
    const float* yb = (y != nullptr) ? y + (size_t)b * C * M : xb;
    float* ob = out + (size_t)b * 2 * C * N;
    unsigned char* ab = argk + (size_t)b * C * N;
    const int* mynb = nb + threadIdx.x * k;
    const int* myct = ct + threadIdx.x * k;
    const int cper = (C + gridDim.z - 1) / gridDim.z;
    const int cbeg = blockIdx.z * cper, cend = min(C, cbeg + cper);
    for (int c = cbeg; c < cend; ++c) {
        const float* xc = xb + (size_t)c * N;
        const float* yc = yb + (size_t)c * M;
        float best = -INFINITY;
        int arg = 0;
        if (idx1 == nullptr) {
            const float xi = xc[n];
            for (int kk = 0; kk < k; ++kk) {
                const float v = __ldg(yc + mynb[kk]) - xi;
                if (v > best) { best = v; arg = kk; }
            }
        } else {
            for (int kk = 0; kk < k; ++kk) {
                const float v = __ldg(yc + mynb[kk]) - __ldg(xc + myct[kk]);
                if (v > best) { best = v; arg = kk; }
            }
        }
        ob[(size_t)(2 * c) * N + n] = xc[n];
        ob[(size_t)(2 * c + 1) * N + n] = best;
        ab[(size_t)c * N + n] = (unsigned char)arg;
    }
}

// dx = d_out[even channels] (- g at the centre when the centre is the point itself)
__global__ void __launch_bounds__(256)
mrconv_gather_bwd_init_kernel(const float* __restrict__ dout, float* __restrict__ dx,
                              int C, int N, int identity_centre, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int n = (int)(e % N);
    const long long bc = e / N;           // b*C + c
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    const float* db = dout + (size_t)b * 2 * C * N;
    float v = db[(size_t)(2 * c) * N + n];
    if (identity_centre) v -= db[(size_t)(2 * c + 1) * N + n];
    dx[e] = v;
}

__global__ void __launch_bounds__(256)
mrconv_gather_bwd_scatter_kernel(const float* __restrict__ dout, const long long* __restrict__ idx0,
                                 const long long* __restrict__ idx1, const unsigned char* __restrict__ argk,
                                 float* __restrict__ dx, float* __restrict__ dy,
                                 int C, int N, int M, int k, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int n = (int)(e % N);
    const long long bc = e / N;
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    const float g = dout[(size_t)b * 2 * C * N + (size_t)(2 * c + 1) * N + n];
    if (g == 0.f) return;
    const int kk = argk[e];
    const long long j = idx0[((size_t)b * N + n) * k + kk];
    atomicAdd(dy + ((size_t)b * C + c) * M + j, g);
    if (idx1 != nullptr) {
        const long long ic = idx1[((size_t)b * N + n) * k + kk];
        atomicAdd(dx + ((size_t)b * C + c) * N + ic, -g);
    }
}

}  // namespace

extern "C" int ge_mrconv_gather_fwd(const float* x, const float* y, const long long* idx_nbr,
                                    const long long* idx_ctr, float* out, unsigned char* argk,
                                    int B, int C, int N, int M, int k, ge_stream_t stream) {
    GE_REQUIRE(x && idx_nbr && out && argk, GE_ERR_ARG, "ge_mrconv_gather_fwd: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0, GE_ERR_ARG, "ge_mrconv_gather_fwd: bad dimension");
    GE_REQUIRE(k <= 255, GE_ERR_SHAPE, "ge_mrconv_gather_fwd: k=%d > 255", k);
    GE_REQUIRE(y != nullptr || N == M, GE_ERR_SHAPE, "ge_mrconv_gather_fwd: self-graph needs M == N");
    const size_t smem = (size_t)2 * MR_THREADS * k * sizeof(int);
    GE_REQUIRE(smem <= 200 * 1024, GE_ERR_CAPACITY, "ge_mrconv_gather_fwd: k too large");
    { static size_t ge_max_smem__ = 0; if ((size_t)(smem) > ge_max_smem__) { GE_CUDA(cudaFuncSetAttribute(mrconv_gather_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)), "ge_mrconv_gather_fwd(attr)"); ge_max_smem__ = (size_t)(smem); } }
    // channel chunks in grid.z so that small graphs still fill the machine
    const long long ctas = (long long)ge::cdiv(N, MR_THREADS) * B;
    int zc = 1;
    while (zc < 8 && ctas * zc < (long long)ge::sm_count() * 16 && C / (zc * 2) >= 16) zc *= 2;
    mrconv_gather_fwd_kernel<<<dim3(ge::cdiv(N, MR_THREADS), B, zc), MR_THREADS, smem, (cudaStream_t)stream>>>(
        x, y, idx_nbr, idx_ctr, out, argk, C, N, M, k);
    GE_CHECK_LAUNCH("ge_mrconv_gather_fwd");
    return GE_OK;
}

// dx [B,C,N] is overwritten; dy [B,C,M] must be ZERO-FILLED by the caller when y was given
// (pass dy = NULL for a self-graph: neighbour gradients are then accumulated into dx).
extern "C" int ge_mrconv_gather_bwd(const float* dout, const long long* idx_nbr, const long long* idx_ctr,
                                    const unsigned char* argk, float* dx, float* dy,
                                    int B, int C, int N, int M, int k, ge_stream_t stream) {
    GE_REQUIRE(dout && idx_nbr && argk && dx, GE_ERR_ARG, "ge_mrconv_gather_bwd: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0, GE_ERR_ARG, "ge_mrconv_gather_bwd: bad dimension");
    GE_REQUIRE(dy != nullptr || N == M, GE_ERR_SHAPE, "ge_mrconv_gather_bwd: self-graph needs M == N");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)B * C * N;
    const unsigned blocks = (unsigned)ge::cdivll(total, 256);
    mrconv_gather_bwd_init_kernel<<<blocks, 256, 0, st>>>(dout, dx, C, N, idx_ctr == nullptr ? 1 : 0, total);
    GE_CHECK_LAUNCH("ge_mrconv_gather_bwd(init)");
    mrconv_gather_bwd_scatter_kernel<<<blocks, 256, 0, st>>>(dout, idx_nbr, idx_ctr, argk, dx,
                                                             dy != nullptr ? dy : dx, C, N, M, k, total);
    GE_CHECK_LAUNCH("ge_mrconv_gather_bwd(scatter)");
    return GE_OK;
}
