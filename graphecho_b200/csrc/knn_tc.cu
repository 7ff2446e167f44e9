// K1 on the 5th-generation tensor cores: dense dilated k-NN graph build (ViG Grapher, N >= 128).
//
// Same contract as the FFMA kernels in knn.cu (DenseDilatedKnnGraph.forward ->
// (xy_)dense_knn_matrix -> (xy_)pairwise_distance -> topk, /root/reference/models/vig.py:232-381),
// but the [N,M] inner-product GEMM -- 2*B*N*M*C flops, the whole cost of the op -- runs on
// tcgen05.mma (kind::f16) with the accumulators in TMEM and operands staged by TMA:
//
//   * pre-pass (knn_split_kernel): L2-normalise every point (F.normalize, vig.py:372-373), transpose
//     to K-major [B,N,C] and split each fp32 value v (|v| <= 1) into two fp16 terms
//         hi = fp16(v),   lo = fp16((v - hi) * 2^11)          =>  v = hi + lo * 2^-11  (+- 2^-22 |v|)
//     The inner product is accumulated in fp32 in TWO TMEM accumulators,
//         D1 = sum hi.hi'      D2 = sum (lo.hi' + hi.lo')      x.y = D1 + 2^-11 * D2,
//     (products of fp16 values are exact in fp32); the dropped lo.lo' term and the rounding of lo are
//     both ~2^-22 relative, i.e. fp32-level, so near-ties order like the fp32 reference (tests accept
//     either order only where two distances differ by < 2e-6).  Half the operand bytes and twice the
//     tensor rate of a 3xTF32 split at the same accuracy.
//   * main kernel (knn_tc_kernel): one CTA per (batch, 128-query tile).  The query tile (hi and lo,
//     <= 128 KB) is loaded ONCE by TMA and stays resident in shared memory; key tiles (hi, lo) stream
//     through a 3-stage mbarrier ring in 64-channel (128-byte, SWIZZLE_128B) slabs.  Warp 0 = TMA
//     producer, warp 1 = MMA issuer (one elected thread, 12 tcgen05.mma per slab), accumulators
//     double-buffered in TMEM (2 x 2 x 128 columns) so the next key tile's GEMM overlaps the selection
//     of the current one.  Warps 2.. = selection: tcgen05.ld gives every thread ONE query row of the
//     distance tile, so each thread keeps its own sorted top-K list in registers, forming
//       dist = (|x^|^2 + (-2 x^.y^)) + |y^|^2                               [vig.py:270-274]
//     exactly as the reference orders the additions.  SPLIT warps share a row (alternating 16-column
//     chunks) and their lists are merged through shared memory at the end.  Ties -> lower key index.
//   * output int64 [2,B,N,k]: [0] = neighbour (sorted by distance, every dilation-th entry),
//     [1] = centre index                                                    [vig.py:328-329, 353].
//
// Work: 2*B*N*M*C algorithmic flops (x3 on the tensor pipe); bytes 4*B*C*(N+M) + 16*B*N*k.
#include "common.cuh"
#include "knn_tc.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int TC_BM = 128;                         // query rows per CTA (UMMA M)
constexpr int TC_BK = 64;                          // channels per slab = one 128-byte swizzle row of fp16
constexpr int TC_STAGES = 3;
constexpr int TC_MAXKB = 4;                        // resident query tile: up to 4 slabs (C <= 256)
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;   // one [128][64] fp16 operand slab = 16 KB
constexpr int TC_A_BYTES = 2 * TC_MAXKB * TC_TILE_BYTES;   // resident query hi + lo
constexpr int TC_STAGE_BYTES = 2 * TC_TILE_BYTES;  // key hi, key lo
constexpr int TC_TMEM_COLS = 512;                  // 2 buffers x (D1, D2) x 128 columns
constexpr size_t TC_SMEM = (size_t)TC_A_BYTES + (size_t)TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + 1024 /*key norms*/;
// threads = producer warp + MMA warp + 4 x SPLIT selection warps (SPLIT warps share each TMEM lane quarter)
constexpr int tc_threads(int split) { return 64 + 128 * split; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;   // stride byte offset
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// one lane of a converged warp (the others skip); keeps the surrounding control flow warp-uniform so the
// descriptors stay in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Pre-pass: normalise, transpose to [B,N,C], split into fp16 hi / scaled lo, squared norms of the fp32
// normalised vectors.  CTA = 32 points x all channels through a padded shared tile.
__global__ void __launch_bounds__(256)
knn_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                 float* __restrict__ sq, int C, int N, int sq_stride) {
    extern __shared__ float tile[];              // [C][33]
    const int b = blockIdx.y, n0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* xb = x + (size_t)b * C * N;
    for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = (n0 + lane < N) ? xb[(size_t)c * N + n0 + lane] : 0.f;
    __syncthreads();
    for (int q = 0; q < 4; ++q) {                // warp w owns points 4w .. 4w+3
        const int p = warp * 4 + q, n = n0 + p;
        if (n >= N) break;                       // uniform per warp
        float s = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = tile[c * 33 + p]; s = fmaf(v, v, s); }
        s = ge::warp_sum(s);
        const float d = fmaxf(sqrtf(s), 1e-12f);
        __half2* oh = reinterpret_cast<__half2*>(hi + ((size_t)b * N + n) * C);
        __half2* ol = reinterpret_cast<__half2*>(lo + ((size_t)b * N + n) * C);
        float qsum = 0.f;
        for (int c2 = lane; c2 < C / 2; c2 += 32) {          // two adjacent channels per lane
            const float v0 = tile[(2 * c2) * 33 + p] / d, v1 = tile[(2 * c2 + 1) * 33 + p] / d;
            qsum = fmaf(v0, v0, qsum);
            qsum = fmaf(v1, v1, qsum);
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const float r0 = (v0 - __half2float(h0)) * 2048.f, r1 = (v1 - __half2float(h1)) * 2048.f;
            oh[c2] = __halves2half2(h0, h1);
            ol[c2] = __halves2half2(__float2half_rn(r0), __float2half_rn(r1));
        }
        qsum = ge::warp_sum(qsum);
        if (lane == 0) sq[(size_t)b * sq_stride + n] = qsum;
    }
}

// Same pre-pass for node-major input x [B,N,C] (channels-last feature maps): rows are already contiguous, so
// one warp owns one point (8 channels per lane per trip, 128-bit bf16 / 2 x 128-bit fp32 loads), no transpose.
template <typename T>
__global__ void __launch_bounds__(256)
knn_split_nmajor_kernel(const T* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo,
                        float* __restrict__ sq, int C, int N, long long points, int sq_stride) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= points) return;
    const T* row = x + p * C;
    float s = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        float v[8];
        ge::load8<T>(row + c, v);
#pragma unroll
        for (int u = 0; u < 8; ++u) s = fmaf(v[u], v[u], s);
    }
    s = ge::warp_sum(s);
    const float d = fmaxf(sqrtf(s), 1e-12f);
    float qsum = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        float v[8];
        ge::load8<T>(row + c, v);
        __half2 h2[4], l2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float v0 = v[2 * u] / d, v1 = v[2 * u + 1] / d;
            qsum = fmaf(v0, v0, qsum);
            qsum = fmaf(v1, v1, qsum);
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            h2[u] = __halves2half2(h0, h1);
            l2[u] = __halves2half2(__float2half_rn((v0 - __half2float(h0)) * 2048.f),
                                   __float2half_rn((v1 - __half2float(h1)) * 2048.f));
        }
        *reinterpret_cast<uint4*>(hi + p * C + c) = *reinterpret_cast<const uint4*>(h2);
        *reinterpret_cast<uint4*>(lo + p * C + c) = *reinterpret_cast<const uint4*>(l2);
    }
    qsum = ge::warp_sum(qsum);
    if (lane == 0) sq[(p / N) * sq_stride + (p % N)] = qsum;
}

__global__ void knn_fill_pad_kernel(float* __restrict__ sq, int M, int Mpad, int B) {
    const int per = Mpad - M;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per * B) return;
    const int b = t / per, j = M + (t - b * per);
    sq[(size_t)b * Mpad + j] = INFINITY;
}

// ---------------------------------------------------------------------------------------------
// Per-thread sorted list (ascending) of KL slots held in registers.  Only the last K slots are live: the
// first KL-K hold -inf sentinels that nothing can displace, so the admission threshold is always the
// statically indexed d[KL-1] (a run-time d[K-1] would push the whole list into local memory).
template <int KL>
struct RegList {
    float d[KL];
    int i[KL];
    __device__ __forceinline__ void init(int K) {
#pragma unroll
        for (int s = 0; s < KL; ++s) { d[s] = (s < KL - K) ? -INFINITY : INFINITY; i[s] = 0; }
    }
    __device__ __forceinline__ float thr() const { return d[KL - 1]; }
    // insert (dist, idx) after every entry <= dist (earlier = lower key index wins ties)
    __device__ __forceinline__ void insert(float dist, int idx) {
#pragma unroll
        for (int s = KL - 1; s > 0; --s) {
            const bool up = d[s - 1] > dist;
            const bool put = !up && d[s] > dist;
            i[s] = up ? i[s - 1] : (put ? idx : i[s]);
            d[s] = up ? d[s - 1] : (put ? dist : d[s]);
        }
        if (d[0] > dist) { d[0] = dist; i[0] = idx; }
    }
    // merge variant: order by (dist, idx) so that equal distances keep the lower key index first
    __device__ __forceinline__ void insert_lex(float dist, int idx) {
#pragma unroll
        for (int s = KL - 1; s > 0; --s) {
            const bool up = d[s - 1] > dist || (d[s - 1] == dist && i[s - 1] > idx);
            const bool put = !up && (d[s] > dist || (d[s] == dist && i[s] > idx));
            i[s] = up ? i[s - 1] : (put ? idx : i[s]);
            d[s] = up ? d[s - 1] : (put ? dist : d[s]);
        }
        if (d[0] > dist || (d[0] == dist && i[0] > idx)) { d[0] = dist; i[0] = idx; }
    }
};

// dist[e] for a lane-dependent e without dynamic register indexing (binary select tree)
__device__ __forceinline__ float select16(const float (&v)[16], int e) {
    const bool b0 = e & 1, b1 = e & 2, b2 = e & 4, b3 = e & 8;
    float a[8], c[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = b0 ? v[2 * k + 1] : v[2 * k];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = b1 ? a[2 * k + 1] : a[2 * k];
    const float g0 = b2 ? c[1] : c[0], g1 = b2 ? c[3] : c[2];
    return b3 ? g1 : g0;
}

template <int KL, int TC_SPLIT>
__global__ void __launch_bounds__(tc_threads(TC_SPLIT), 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
              const __grid_constant__ CUtensorMap tm_yh, const __grid_constant__ CUtensorMap tm_yl,
              const float* __restrict__ xsq, const float* __restrict__ ysq_pad, long long* __restrict__ out,
              int B, int N, int M, int Mpad, int BN, int nkb, int K, int dilation, int xsq_stride) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;         // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t stage_base = base + TC_A_BYTES;                       // [0, TC_A_BYTES): resident query slabs (hi 0..3, lo 4..7)
    const uint32_t bar_base = stage_base + TC_STAGES * TC_STAGE_BYTES;
    // barriers: full[3], empty[3], tmem_full[2], tmem_empty[2], a_full; then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + 2 + s); };
    const uint32_t afull_bar = bar_base + 8u * (2 * TC_STAGES + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 5);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* ys_s = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [2][128] key norms of the live tiles

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, i0 = blockIdx.x * TC_BM;
    const int ntiles = Mpad / BN;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * TC_SPLIT); }
        mbar_init(afull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // selection-warp state lives at function scope: the lists are merged after the role branches join
    const int q = warp & 3;                                         // TMEM lane quarter this warp may read
    const int split = (warp >= 2) ? (warp - 2) >> 2 : 0;
    const int row = q * 32 + lane, i = i0 + row;
    const float xs = (warp >= 2 && i < N) ? xsq[(size_t)b * xsq_stride + i] : INFINITY;   // rows past N never admit a key
    const float* ysq = ysq_pad + (size_t)b * Mpad;
    RegList<KL> L;
    L.init(K);

    if (warp == 0) {
        // ===== TMA producer (whole warp runs the loop, one elected lane issues) =====
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_xh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_xl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_yh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_yl) : "memory");
            // resident query tile: nkb slabs of hi and of lo, one barrier
            mbar_expect_tx(afull_bar, 2u * (uint32_t)nkb * TC_TILE_BYTES);
            for (int kb = 0; kb < nkb; ++kb) {
                tma_load_3d(base + kb * TC_TILE_BYTES, &tm_xh, afull_bar, kb * TC_BK, i0, b);
                tma_load_3d(base + (TC_MAXKB + kb) * TC_TILE_BYTES, &tm_xl, afull_bar, kb * TC_BK, i0, b);
            }
        }
        __syncwarp();
        const uint32_t half_tx = (uint32_t)BN * 128u;                 // key hi slab; the lo slab follows it contiguously
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < ntiles; ++t) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(s), 2u * half_tx);
                    const uint32_t dst = stage_base + s * TC_STAGE_BYTES;
                    tma_load_3d(dst, &tm_yh, full_bar(s), kb * TC_BK, t * BN, b);
                    tma_load_3d(dst + half_tx, &tm_yl, full_bar(s), kb * TC_BK, t * BN, b);
                }
                __syncwarp();
                if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp runs the loop, one elected lane issues) =====
        // Per 16-channel step two MMAs instead of three: the key hi and lo slabs are contiguous, so
        //   [D1 | D2] += A_hi . [B_hi ; B_lo]^T   (N = 2*BN)        D2 += A_lo . B_hi^T   (N = BN)
        // instruction descriptors: D fp32, A/B fp16, both K-major, M = 128
        const uint32_t idesc_n = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const uint32_t idesc_2n = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        mbar_wait(afull_bar, 0u);
        int s = 0;
        uint32_t ph = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            mbar_wait(tempty_bar(buf), (((uint32_t)t >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d1 = tmem_base + (uint32_t)buf * 256u, d2 = d1 + (uint32_t)BN;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_h = base + kb * TC_TILE_BYTES, a_l = base + (TC_MAXKB + kb) * TC_TILE_BYTES;
                const uint32_t b_h = stage_base + s * TC_STAGE_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {           // UMMA_K = 16 fp16 = 32 bytes
                        const uint64_t dah = tc_smem_desc(a_h + k * 32), dal = tc_smem_desc(a_l + k * 32);
                        const uint64_t dbh = tc_smem_desc(b_h + k * 32);
                        tc_mma_f16(d1, dah, dbh, idesc_2n, (kb | k) != 0 ? 1u : 0u);
                        tc_mma_f16(d2, dal, dbh, idesc_n, 1u);
                    }
                    tc_commit(empty_bar(s));                        // frees the smem stage when these MMAs retire
                    if (kb == nkb - 1) tc_commit(tfull_bar(buf));   // accumulators of tile t complete
                }
                __syncwarp();
                if (++s == TC_STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== selection warps: thread <-> one query row of the accumulator; the TC_SPLIT warps of a
        // lane quarter take alternating 16-column chunks and keep separate lists, merged at the end =====
        const int nchunks = BN >> 4;
        const int sel_tid = tid - 64;                               // 0 .. 128*TC_SPLIT-1
        float ynext = (sel_tid < BN) ? ysq[sel_tid] : 0.f;          // |y^|^2 of key tile 0 (one key per thread)
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            const int j0 = t * BN;
            // stage this tile's key norms in shared memory (double-buffered), prefetch the next tile's
            if (sel_tid < BN) ys_s[buf * 128 + sel_tid] = ynext;
            asm volatile("bar.sync 1, %0;" ::"r"(128 * TC_SPLIT) : "memory");
            if (sel_tid < BN && t + 1 < ntiles) ynext = ysq[j0 + BN + sel_tid];
            mbar_wait(tfull_bar(buf), ((uint32_t)t >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 256u;
            uint32_t r[16], r2[16];
            int c = (split + t) % TC_SPLIT;
            if (c < nchunks) {
                tc_ld16(taddr + (uint32_t)(c * 16), r);
                tc_ld16(taddr + (uint32_t)BN + (uint32_t)(c * 16), r2);
            }
            while (c < nchunks) {
                float dist[16];
                const float4* yp = reinterpret_cast<const float4*>(ys_s + buf * 128 + c * 16);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 v = yp[e];
                    dist[4 * e] = v.x; dist[4 * e + 1] = v.y; dist[4 * e + 2] = v.z; dist[4 * e + 3] = v.w;
                }
                tc_ld_wait();
                const float thr0 = L.thr();
                unsigned mask = 0;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float ip = fmaf(__uint_as_float(r2[e]), 1.f / 2048.f, __uint_as_float(r[e]));
                    dist[e] = (xs + (-2.f * ip)) + dist[e];                              // padded keys carry +inf
                    mask |= (dist[e] < thr0) ? (1u << e) : 0u;
                }
                const int cn = c + TC_SPLIT;
                if (cn < nchunks) {                                  // next chunk's accumulators fly while this one is selected
                    tc_ld16(taddr + (uint32_t)(cn * 16), r);
                    tc_ld16(taddr + (uint32_t)BN + (uint32_t)(cn * 16), r2);
                }
                // Every lane inserts ITS OWN next candidate per round (rounds = the largest candidate count of a
                // lane, not the number of columns in which some lane has one).  The round is branch-free and
                // warp-uniform: a lane without a candidate inserts +inf, which changes nothing.
                const int rounds = __reduce_max_sync(ge::kFull, __popc(mask));
                for (int rd = 0; rd < rounds; ++rd) {
                    const int e = __ffs(mask) - 1;
                    const float v = (mask != 0u) ? select16(dist, e & 15) : INFINITY;
                    mask &= mask - 1u;
                    L.insert(v, j0 + c * 16 + e);
                }
                c = cn;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
    }
    tc_fence_before();
    __syncthreads();                     // all MMAs retired (every tile was consumed): the stage buffers are free
    // merge the TC_SPLIT partial lists of every row through shared memory ([split-1][slot][row], conflict-free)
    float* md = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
    int* mi = reinterpret_cast<int*>(md + (TC_SPLIT - 1) * KL * TC_BM);
    if (warp >= 2 && split > 0) {
#pragma unroll
        for (int s = 0; s < KL; ++s) {
            md[((split - 1) * KL + s) * TC_BM + row] = L.d[s];
            mi[((split - 1) * KL + s) * TC_BM + row] = L.i[s];
        }
    }
    __syncthreads();
    if (warp >= 2 && split == 0) {
        for (int o = 0; o < TC_SPLIT - 1; ++o) {
#pragma unroll
            for (int s = 0; s < KL; ++s) {
                const float dv = md[(o * KL + s) * TC_BM + row];
                const int iv = mi[(o * KL + s) * TC_BM + row];
                if (dv > -INFINITY && dv <= L.thr()) L.insert_lex(dv, iv);
            }
        }
        if (i < N) {
            const int kout = K / dilation;
            long long* out0 = out + ((size_t)b * N + i) * kout;
            long long* out1 = out + (size_t)B * N * kout + ((size_t)b * N + i) * kout;
#pragma unroll
            for (int s = 0; s < KL; ++s) {
                const int o = s - (KL - K);                         // rank among the live slots
                if (o >= 0 && (o % dilation) == 0) {
                    out0[o / dilation] = L.i[s];
                    out1[o / dilation] = i;
                }
            }
        }
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [B][rows][C] fp16, box = {64 channels, box_rows, 1}, 128-byte swizzle, out-of-bounds rows / channels read as zero
bool make_map(CUtensorMap* map, const __half* ptr, int B, int rows, int C, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)rows * C * 2};
    const cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_bn(int M) {                     // key-tile width (UMMA N): least padding, then widest
    int best = 128, best_pad = ge::cdiv(M, 128) * 128;
    for (int bn = 112; bn >= 64; bn -= 16) {
        const int pad = ge::cdiv(M, bn) * bn;
        if (pad < best_pad) { best = bn; best_pad = pad; }
    }
    return best;
}

template <int KL, int SPLIT>
int launch_tc(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& yh, const CUtensorMap& yl,
              const float* xsq, int xsq_stride, const float* ysq, long long* out, int B, int N, int M, int Mpad, int BN,
              int nkb, int K, int dilation, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        GE_CUDA(cudaFuncSetAttribute(knn_tc_kernel<KL, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM),
                "ge_knn_graph(tc attr)");
        attr_done = true;
    }
    knn_tc_kernel<KL, SPLIT><<<dim3(ge::cdiv(N, TC_BM), B), tc_threads(SPLIT), TC_SMEM, st>>>(xh, xl, yh, yl, xsq, ysq, out, B, N, M, Mpad,
                                                                                BN, nkb, K, dilation, xsq_stride);
    GE_CHECK_LAUNCH("ge_knn_graph(tcgen05)");
    return GE_OK;
}

}  // namespace

namespace ge {

bool knn_tc_applicable(int B, int C, int N, int M, int K, bool has_rel) {
    return !has_rel && C % 32 == 0 && C <= TC_MAXKB * TC_BK && K <= 32 && N >= 128 && M >= 128 && B <= 65535 &&
           encode_fn() != nullptr;
}

size_t knn_tc_workspace_bytes(int B, int C, int N, int M) {
    const size_t mpad = (size_t)M + 128;
    return 2 * (size_t)B * C * ((size_t)N + M) * sizeof(__half) + ((size_t)B * N + (size_t)B * mpad) * sizeof(float) + 256;
}

namespace {
// layout 0: x [B,C,N] fp32 (the reference's graph layout);  1: [B,N,C] fp32;  2: [B,N,C] bf16 (node-major)
int launch_split(const void* x, int layout, __half* hi, __half* lo, float* sq, int B, int C, int N, int sq_stride,
                 cudaStream_t st) {
    if (layout == 0) {
        const size_t split_smem = (size_t)C * 33 * sizeof(float);
        static size_t split_cached = 48 * 1024;
        if (split_smem > split_cached) {
            GE_CUDA(cudaFuncSetAttribute(knn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem),
                    "ge_knn_graph(split attr)");
            split_cached = split_smem;
        }
        knn_split_kernel<<<dim3(ge::cdiv(N, 32), B), 256, split_smem, st>>>(static_cast<const float*>(x), hi, lo, sq, C, N, sq_stride);
    } else {
        const long long points = (long long)B * N;
        const unsigned blocks = (unsigned)ge::cdivll(points, 8);
        if (layout == 1)
            knn_split_nmajor_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), hi, lo, sq, C, N, points, sq_stride);
        else
            knn_split_nmajor_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), hi, lo, sq, C, N, points, sq_stride);
    }
    GE_CHECK_LAUNCH("ge_knn_graph(split)");
    return GE_OK;
}
}  // namespace

int knn_tc_run(const void* x, const void* y, int layout, long long* edge_index, void* workspace,
               int B, int C, int N, int M, int K, int dilation, cudaStream_t st) {
    const int BN = pick_bn(M);
    const int Mpad = ge::cdiv(M, BN) * BN;
    __half* xh = static_cast<__half*>(workspace);
    __half* xl = xh + (size_t)B * N * C;
    __half* yh = xl + (size_t)B * N * C;
    __half* yl = yh + (y ? (size_t)B * M * C : 0);
    float* ysq = reinterpret_cast<float*>(yl + (y ? (size_t)B * M * C : 0));   // [B][Mpad], padded keys = +inf
    float* xsq = ysq + (size_t)B * Mpad;                     // [B][N]
    int rc;
    if (y != nullptr) {
        if ((rc = launch_split(x, layout, xh, xl, xsq, B, C, N, N, st)) != GE_OK) return rc;
        if ((rc = launch_split(y, layout, yh, yl, ysq, B, C, M, Mpad, st)) != GE_OK) return rc;
    } else {
        // self-graph: keys are the queries; the padded norm row doubles as the query norms
        if ((rc = launch_split(x, layout, xh, xl, ysq, B, C, N, Mpad, st)) != GE_OK) return rc;
        yh = xh; yl = xl;
    }
    if (Mpad > M) {
        const int n = (Mpad - M) * B;
        knn_fill_pad_kernel<<<ge::cdiv(n, 256), 256, 0, st>>>(ysq, M, Mpad, B);
        GE_CHECK_LAUNCH("ge_knn_graph(pad)");
    }
    CUtensorMap mxh, mxl, myh, myl;
    const bool ok = make_map(&mxh, xh, B, N, C, TC_BM) && make_map(&mxl, xl, B, N, C, TC_BM) &&
                    make_map(&myh, yh, B, M, C, BN) && make_map(&myl, yl, B, M, C, BN);
    GE_REQUIRE(ok, GE_ERR_SHAPE, "ge_knn_graph: cuTensorMapEncodeTiled failed (B=%d C=%d N=%d M=%d)", B, C, N, M);
    const float* xsq_p = (y != nullptr) ? xsq : ysq;
    const int xs = (y != nullptr) ? N : Mpad;
    const int nkb = ge::cdiv(C, TC_BK);
    if (K <= 9) return launch_tc<9, 2>(mxh, mxl, myh, myl, xsq_p, xs, ysq, edge_index, B, N, M, Mpad, BN, nkb, K, dilation, st);
    if (K <= 18) return launch_tc<18, 2>(mxh, mxl, myh, myl, xsq_p, xs, ysq, edge_index, B, N, M, Mpad, BN, nkb, K, dilation, st);
    return launch_tc<32, 2>(mxh, mxl, myh, myl, xsq_p, xs, ysq, edge_index, B, N, M, Mpad, BN, nkb, K, dilation, st);
}

}  // namespace ge
