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


// ---------------------------------------------------------------------------------------------
// Node-major (channels-last) variant for the Grapher: x [B,N,C], y [B,M,C] (or x), out [B,N,2C] in the map's own
// dtype (bf16 under autocast).  The centre of every edge is the point itself (what DenseDilatedKnnGraph emits).
// One warp per point: a lane owns 8 consecutive channels, so the point's row and its k neighbour rows are read
// as full 128-bit coalesced rows (the [B,C,N] layout reads one scattered scalar per channel and neighbour).
template <typename T>
__global__ void __launch_bounds__(256)
mr_gather_nmajor_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const long long* __restrict__ idx0,
                            T* __restrict__ out, unsigned char* __restrict__ argk,
                            int C, int N, int M, int k, long long points) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= points) return;
    const long long b = p / N;
    const int my_idx = lane < k ? (int)idx0[p * k + lane] : 0;
    const T* ybase = y + b * (long long)M * C;
    for (int c0 = 0; c0 < C; c0 += 256) {                    // all lanes run every trip: the shuffles need the whole warp
        const int c = c0 + lane * 8;
        const bool live = c < C;
        float xi[8], best[8];
        int arg[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { xi[u] = 0.f; best[u] = -INFINITY; arg[u] = 0; }
        if (live) ge::load8<T>(x + p * C + c, xi);
        for (int kk = 0; kk < k; ++kk) {
            const int j = __shfl_sync(ge::kFull, my_idx, kk);
            if (live) {
                float v[8];
                ge::load8<T>(ybase + (long long)j * C + c, v);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float d = v[u] - xi[u];
                    if (d > best[u]) { best[u] = d; arg[u] = kk; }
                }
            }
        }
        if (!live) continue;
        float o0[8], o1[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            o0[2 * u] = xi[u]; o0[2 * u + 1] = best[u];
            o1[2 * u] = xi[4 + u]; o1[2 * u + 1] = best[4 + u];
        }
        ge::store8<T>(out + p * 2 * C + 2 * c, o0);
        ge::store8<T>(out + p * 2 * C + 2 * c + 8, o1);
        unsigned long long packed = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) packed |= (unsigned long long)(arg[u] & 0xff) << (8 * u);
        *reinterpret_cast<unsigned long long*>(argk + p * C + c) = packed;
    }
}

// dx[p][c] = dout[p][2c] - dout[p][2c+1]  (fp32 accumulator that the scatter below adds neighbour terms into)
template <typename T>
__global__ void __launch_bounds__(256)
mr_gather_nmajor_bwd_init_kernel(const T* __restrict__ dout, float* __restrict__ dx, int C, long long octets) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= octets) return;
    const int c8 = C >> 3;
    const long long p = e / c8;
    const int c = (int)(e - p * c8) * 8;
    float a[8], b[8], o[8];
    ge::load8<T>(dout + p * 2 * C + 2 * c, a);
    ge::load8<T>(dout + p * 2 * C + 2 * c + 8, b);
#pragma unroll
    for (int u = 0; u < 4; ++u) { o[u] = a[2 * u] - a[2 * u + 1]; o[4 + u] = b[2 * u] - b[2 * u + 1]; }
    ge::store8<float>(dx + p * C + c, o);
}

template <typename T>
__global__ void __launch_bounds__(256)
mr_gather_nmajor_bwd_scatter_kernel(const T* __restrict__ dout, const long long* __restrict__ idx0,
                                    const unsigned char* __restrict__ argk, float* __restrict__ dtarget,
                                    int C, int N, int M, int k, long long points) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= points) return;
    const long long b = p / N;
    const int my_idx = lane < k ? (int)idx0[p * k + lane] : 0;
    float* tb = dtarget + b * (long long)M * C;
    for (int c0 = 0; c0 < C; c0 += 256) {                    // all lanes run every trip: the shuffles need the whole warp
        const int c = c0 + lane * 8;
        const bool live = c < C;
        float a[8], bq[8];
        unsigned long long packed = 0;
        if (live) {
            ge::load8<T>(dout + p * 2 * C + 2 * c, a);
            ge::load8<T>(dout + p * 2 * C + 2 * c + 8, bq);
            packed = *reinterpret_cast<const unsigned long long*>(argk + p * C + c);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int kk = (int)((packed >> (8 * u)) & 0xff);
            const int j = __shfl_sync(ge::kFull, my_idx, kk);
            const float g = u < 4 ? a[2 * u + 1] : bq[2 * (u - 4) + 1];
            if (live && g != 0.f) atomicAdd(tb + (long long)j * C + c + u, g);
        }
    }
}

// Self-graph backward in ONE launch, gather form, no floating-point atomics (fp32 shared-memory atomics are CAS spin
// loops on sm_100: the first, scatter-into-shared version of this kernel spent its time in ATOMS.CAST.SPIN).
// A CTA owns (frame b, 64-channel slab).  It stages the slab's dout_xj values (bf16/fp32) and arg-max slots in shared
// memory, builds the REVERSE neighbour lists of the frame there (integer atomics), and then every (node j, 8 channels)
// item sums its own term dout_x - dout_xj and the dout_xj of every node whose arg-max neighbour in that channel is j.
// dx is written once, in the activation dtype.  Replaces init (205 MB fp32 written) + global-atomic scatter (474 MB
// read, ncu) + a cast pass.
constexpr int MRB_CS = 32;          // channels per slab (2 CTAs per SM at N = 784: the phases of one overlap the other)
constexpr int MRB_THREADS = 512;
constexpr int MRB_U = 2;            // items in flight per thread

template <typename T>
__global__ void __launch_bounds__(MRB_THREADS, 2)
mr_gather_nmajor_bwd_self_kernel(const T* __restrict__ dout, const long long* __restrict__ idx0,
                                 const unsigned char* __restrict__ argk, T* __restrict__ dx, int C, int N, int k) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* gj = reinterpret_cast<T*>(raw);                                             // [N][64] dout_xj of the slab
    unsigned char* ak = raw + (size_t)N * MRB_CS * sizeof(T);                      // [N][64] arg-max slots
    int* cnt = reinterpret_cast<int*>(ak + (size_t)N * MRB_CS);                    // [N+1] incoming-edge counts -> offsets
    int* fill = cnt + (N + 1);                                                     // [N] fill cursors
    unsigned short* src = reinterpret_cast<unsigned short*>(fill + N);             // [N*k] source node of an incoming edge
    unsigned char* slot = reinterpret_cast<unsigned char*>(src + (size_t)N * k);   // [N*k] its neighbour slot
    __shared__ int wsum[MRB_THREADS / 32];
    const int b = blockIdx.y, c0 = blockIdx.x * MRB_CS, tid = threadIdx.x;
    const long long p0 = (long long)b * N;
    constexpr int OCT = MRB_CS / 8;
    const int total = N * OCT;
    for (int i = tid; i <= N; i += MRB_THREADS) cnt[i] = 0;
    __syncthreads();
    // ---- stage dout_xj + arg-max slots of the slab; count incoming edges.  MRB_U items per thread in flight: one CTA
    // per SM has to keep ~64 KB of loads outstanding on its own to stream at HBM rate.
    for (int base = tid; base < total; base += MRB_THREADS * MRB_U) {
        float a[MRB_U][8], bq[MRB_U][8];
        unsigned long long pk[MRB_U];
#pragma unroll
        for (int u = 0; u < MRB_U; ++u) {
            const int e = base + u * MRB_THREADS;
            if (e < total) {
                const int i = e / OCT, c = c0 + (e - i * OCT) * 8;
                ge::load8<T>(dout + (p0 + i) * 2 * C + 2 * c, a[u]);
                ge::load8<T>(dout + (p0 + i) * 2 * C + 2 * c + 8, bq[u]);
                pk[u] = *reinterpret_cast<const unsigned long long*>(argk + (p0 + i) * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < MRB_U; ++u) {
            const int e = base + u * MRB_THREADS;
            if (e < total) {
                const int i = e / OCT, cl = (e - i * OCT) * 8;
                float g[8], own[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    g[q] = a[u][2 * q + 1]; g[4 + q] = bq[u][2 * q + 1];
                    own[q] = a[u][2 * q] - a[u][2 * q + 1]; own[4 + q] = bq[u][2 * q] - bq[u][2 * q + 1];
                }
                ge::store8<T>(gj + (size_t)i * MRB_CS + cl, g);
                *reinterpret_cast<unsigned long long*>(ak + (size_t)i * MRB_CS + cl) = pk[u];
                ge::store8<T>(dx + (p0 + i) * C + c0 + cl, own);          // own term parked in the output (read back below)
            }
        }
    }
    for (int e = tid; e < N * k; e += MRB_THREADS) atomicAdd(&cnt[(int)idx0[p0 * k + e] + 1], 1);
    __syncthreads();
    // ---- exclusive scan of the counts (cnt[j+1] held the count of j): block scan over N + 1 entries
    {
        const int per = (N + 1 + MRB_THREADS - 1) / MRB_THREADS;
        const int lo = tid * per, hi = min(N + 1, lo + per);
        int s_ = 0;
        for (int i = lo; i < hi; ++i) s_ += cnt[i];
        int incl = s_;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t_ = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t_; }
        if ((tid & 31) == 31) wsum[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
        int run = base + incl - s_;
        for (int i = lo; i < hi; ++i) { run += cnt[i]; cnt[i] = run; }        // inclusive prefix: cnt[j] = offset of node j's list end... 
    }
    __syncthreads();
    // after the inclusive scan over the shifted counts, cnt[j] = number of edges into nodes < j  (list of j = [cnt[j], cnt[j+1]))
    for (int i = tid; i < N; i += MRB_THREADS) fill[i] = cnt[i];
    __syncthreads();
    for (int e = tid; e < N * k; e += MRB_THREADS) {
        const int j = (int)idx0[p0 * k + e];
        const int pos = atomicAdd(&fill[j], 1);
        src[pos] = (unsigned short)(e / k);
        slot[pos] = (unsigned char)(e - (e / k) * k);
    }
    __syncthreads();
    // ---- gather: item = (node j, 8 channels)
    for (int base = tid; base < total; base += MRB_THREADS * MRB_U) {
        float o[MRB_U][8];
#pragma unroll
        for (int u = 0; u < MRB_U; ++u) {
            const int e = base + u * MRB_THREADS;
            if (e < total) {
                const int j = e / OCT, cl = (e - j * OCT) * 8;
                ge::load8<T>(dx + (p0 + j) * C + c0 + cl, o[u]);           // own term (this thread wrote it: L2 hit)
            }
        }
#pragma unroll
        for (int u = 0; u < MRB_U; ++u) {
            const int e = base + u * MRB_THREADS;
            if (e < total) {
                const int j = e / OCT, cl = (e - j * OCT) * 8;
                for (int t_ = cnt[j]; t_ < cnt[j + 1]; ++t_) {
                    const int i = src[t_];
                    const unsigned kk = slot[t_];
                    const unsigned long long pk = *reinterpret_cast<const unsigned long long*>(ak + (size_t)i * MRB_CS + cl);
                    float g[8];
                    ge::load8<T>(gj + (size_t)i * MRB_CS + cl, g);
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[u][q] += (((pk >> (8 * q)) & 0xffull) == kk) ? g[q] : 0.f;
                }
                ge::store8<T>(dx + (p0 + j) * C + c0 + cl, o[u]);
            }
        }
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

// ---- node-major (channels-last) entries: x [B,N,C], y [B,M,C] or NULL, out [B,N,2C] in `dtype`; centre = the point
extern "C" int ge_mrconv_gather_nmajor_fwd(const void* x, const void* y, const long long* idx_nbr, void* out,
                                           unsigned char* argk, int dtype, int B, int C, int N, int M, int k,
                                           ge_stream_t stream) {
    GE_REQUIRE(x && idx_nbr && out && argk, GE_ERR_ARG, "ge_mrconv_gather_nmajor_fwd: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0, GE_ERR_ARG, "ge_mrconv_gather_nmajor_fwd: bad dimension");
    GE_REQUIRE(k <= 32 && C % 8 == 0, GE_ERR_SHAPE, "ge_mrconv_gather_nmajor_fwd: needs k <= 32 and C %% 8 == 0 (k=%d C=%d)", k, C);
    GE_REQUIRE(y != nullptr || N == M, GE_ERR_SHAPE, "ge_mrconv_gather_nmajor_fwd: self-graph needs M == N");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_mrconv_gather_nmajor_fwd: unsupported dtype %d", dtype);
    const long long points = (long long)B * N;
    const unsigned blocks = (unsigned)ge::cdivll(points, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        mr_gather_nmajor_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (const float*)(y ? y : x), idx_nbr,
                                                                   (float*)out, argk, C, N, M, k, points);
    else
        mr_gather_nmajor_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)(y ? y : x),
                                                                           idx_nbr, (__nv_bfloat16*)out, argk, C, N, M, k, points);
    GE_CHECK_LAUNCH("ge_mrconv_gather_nmajor_fwd");
    return GE_OK;
}

// dx [B,N,C] fp32 is overwritten; dy [B,M,C] fp32 must be ZERO-FILLED by the caller when y was given (NULL = self-graph).
extern "C" int ge_mrconv_gather_nmajor_bwd(const void* dout, const long long* idx_nbr, const unsigned char* argk,
                                           float* dx, float* dy, int dtype, int B, int C, int N, int M, int k,
                                           ge_stream_t stream) {
    GE_REQUIRE(dout && idx_nbr && argk && dx, GE_ERR_ARG, "ge_mrconv_gather_nmajor_bwd: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && k > 0, GE_ERR_ARG, "ge_mrconv_gather_nmajor_bwd: bad dimension");
    GE_REQUIRE(k <= 32 && C % 8 == 0, GE_ERR_SHAPE, "ge_mrconv_gather_nmajor_bwd: needs k <= 32 and C %% 8 == 0");
    GE_REQUIRE(dy != nullptr || N == M, GE_ERR_SHAPE, "ge_mrconv_gather_nmajor_bwd: self-graph needs M == N");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_mrconv_gather_nmajor_bwd: unsupported dtype %d", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const long long points = (long long)B * N, octets = points * (C / 8);
    float* target = dy != nullptr ? dy : dx;
    if (dtype == GE_DTYPE_F32) {
        mr_gather_nmajor_bwd_init_kernel<float><<<(unsigned)ge::cdivll(octets, 256), 256, 0, st>>>((const float*)dout, dx, C, octets);
        GE_CHECK_LAUNCH("ge_mrconv_gather_nmajor_bwd(init)");
        mr_gather_nmajor_bwd_scatter_kernel<float><<<(unsigned)ge::cdivll(points, 8), 256, 0, st>>>((const float*)dout, idx_nbr, argk, target, C, N, M, k, points);
    } else {
        mr_gather_nmajor_bwd_init_kernel<__nv_bfloat16><<<(unsigned)ge::cdivll(octets, 256), 256, 0, st>>>((const __nv_bfloat16*)dout, dx, C, octets);
        GE_CHECK_LAUNCH("ge_mrconv_gather_nmajor_bwd(init)");
        mr_gather_nmajor_bwd_scatter_kernel<__nv_bfloat16><<<(unsigned)ge::cdivll(points, 8), 256, 0, st>>>((const __nv_bfloat16*)dout, idx_nbr, argk, target, C, N, M, k, points);
    }
    GE_CHECK_LAUNCH("ge_mrconv_gather_nmajor_bwd(scatter)");
    return GE_OK;
}

// Self-graph backward with the gradient slab resident in shared memory (one launch, dx written once in the activation
// dtype).  Returns GE_ERR_CAPACITY when the slab does not fit (N * 64 * 4 + N * k * 2 bytes > 220 KB): the caller then
// takes ge_mrconv_gather_nmajor_bwd.  dx [B,N,C] in `dtype`, overwritten.
extern "C" int ge_mrconv_gather_nmajor_bwd_self(const void* dout, const long long* idx_nbr, const unsigned char* argk,
                                                void* dx, int dtype, int B, int C, int N, int k, ge_stream_t stream) {
    GE_REQUIRE(dout && idx_nbr && argk && dx, GE_ERR_ARG, "ge_mrconv_gather_nmajor_bwd_self: null pointer");
    GE_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, GE_ERR_ARG, "ge_mrconv_gather_nmajor_bwd_self: bad dimension");
    GE_REQUIRE(k <= 32 && C % MRB_CS == 0 && N <= 65535 && B <= 65535, GE_ERR_SHAPE,
               "ge_mrconv_gather_nmajor_bwd_self: needs k <= 32, C %% 32 == 0, N <= 65535");
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_mrconv_gather_nmajor_bwd_self: unsupported dtype %d", dtype);
    const size_t es_ = dtype == GE_DTYPE_F32 ? 4 : 2;
    const size_t smem = (size_t)N * MRB_CS * (es_ + 1) + (size_t)(2 * N + 1) * sizeof(int) + (size_t)N * k * 3 + 16;
    GE_REQUIRE(smem <= 110 * 1024, GE_ERR_CAPACITY, "ge_mrconv_gather_nmajor_bwd_self: N=%d does not fit shared memory", N);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(C / MRB_CS, B);
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(mr_gather_nmajor_bwd_self_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_mrconv_gather_nmajor_bwd_self(attr)"); c0 = smem; }
        mr_gather_nmajor_bwd_self_kernel<float><<<grid, MRB_THREADS, smem, st>>>((const float*)dout, idx_nbr, argk, (float*)dx, C, N, k);
    } else {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(mr_gather_nmajor_bwd_self_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_mrconv_gather_nmajor_bwd_self(attr)"); c1 = smem; }
        mr_gather_nmajor_bwd_self_kernel<__nv_bfloat16><<<grid, MRB_THREADS, smem, st>>>((const __nv_bfloat16*)dout, idx_nbr, argk, (__nv_bfloat16*)dx, C, N, k);
    }
    GE_CHECK_LAUNCH("ge_mrconv_gather_nmajor_bwd_self");
    return GE_OK;
}
