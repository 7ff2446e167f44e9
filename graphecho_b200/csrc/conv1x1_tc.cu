// f3 (first slice) -- 1x1 convolution on an NHWC map as a tcgen05 GEMM whose epilogue already produces the BatchNorm
// statistics of its output.
//
// Reference ops: Bottleneck.forward's conv1 -> bn1 and conv3 -> bn3 (models/fpnseg.py:192-212) and the Grapher's
// fc1 = Conv2d(C, C, 1) + BatchNorm2d (models/vig.py:402-405): y = x W^T with x [P, K] (P = N*H*W pixels, NHWC = K-major),
// W [Cout, K]; the reference then reads y again to take the batch statistics.  On a B200 these GEMMs are memory-bound
// (K, Cout <= 1024: 2*K*Cout/(K+Cout) flop per byte << the tensor/HBM balance), so what matters is bytes:
//   * A (x) streams through a TMA/mbarrier ring exactly once per output-column tile; the weight tile stays RESIDENT in
//     shared memory for the whole persistent CTA (one CTA per SM loops over 128-row tiles);
//   * one elected thread issues tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators double-buffered in TMEM) so
//     the epilogue of tile i overlaps the loads and MMAs of tile i+1;
//   * epilogue warps read the accumulator rows with tcgen05.ld, store bf16 y, and reduce sum(y - shift) and
//     sum((y - shift)^2) per output channel from the fp32 ACCUMULATORS with a 16-shuffle butterfly per 16 columns; per-CTA
//     partial rows go to the same workspace layout the BatchNorm finalize kernel reads -> the separate statistics pass
//     over y (1 of the 3 passes of the fused BatchNorm forward) disappears.
// Segments (per-domain statistics, see batch_norm.cu): CTAs [0,g0) own the tiles of segment 0, the rest those of
// segment 1; a tile never contributes rows of the other segment to a partial.
// Algorithmic bytes: 2*P*(K + Cout) + 2*K*Cout; flops 2*P*K*Cout.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/graphecho_b200.h"
#include <cuda_bf16.h>

namespace {

using namespace ge_tc;
using bf16 = __nv_bfloat16;

constexpr int CV_BM = 128;                 // rows per tile (UMMA M)
constexpr int CV_BK = 64;                  // K per slab = one 128-byte swizzle row of bf16
constexpr int CV_A_SLAB = CV_BM * CV_BK * 2;        // 16 KB
constexpr int CV_MAX_STAGES = 8;
constexpr int CV_W_BUDGET = 128 * 1024;    // resident weight tile
constexpr int CV_SMEM_BUDGET = 220 * 1024;
constexpr int CV_EPI_WARPS = 8;            // two warps per TMEM lane quarter, alternating 16-column chunks
constexpr int CV_THREADS = 64 + 32 * CV_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

struct ConvPlan {
    long long P, P0;                        // rows, rows of segment 0 (== P: one segment)
    int G, g0;                              // CTAs along x, CTAs of segment 0
    int N, K, nkb, stages;
};

template <int BN>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  bf16* __restrict__ out, const float* __restrict__ shift, float* __restrict__ part, ConvPlan pl) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_base = base;                                              // nkb slabs of [BN][64] bf16
    const uint32_t a_base = w_base + (uint32_t)pl.nkb * BN * 128u;              // ring of [128][64] bf16
    const uint32_t bar_base = a_base + (uint32_t)pl.stages * CV_A_SLAB;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (CV_MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * CV_MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * CV_MAX_STAGES + 2 + s); };
    const uint32_t wfull_bar = bar_base + 8u * (2 * CV_MAX_STAGES + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * CV_MAX_STAGES + 5);
    uint8_t* gen = smem_raw + (bar_base + 256u - smem_u32(smem_raw));           // generic-pointer view past the barriers
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* shift_s = reinterpret_cast<float*>(gen);                             // [BN]
    float* stat_s = shift_s + BN;                                               // [4][2][BN]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.y * BN;
    // this CTA's tiles: segment, first row of the segment, rows of the segment, tile stride
    const int seg = ((int)blockIdx.x < pl.g0) ? 0 : 1;
    const long long seg_lo = seg == 0 ? 0 : pl.P0, seg_hi = seg == 0 ? pl.P0 : pl.P;
    const int cta_in_seg = seg == 0 ? (int)blockIdx.x : (int)blockIdx.x - pl.g0;
    const int ctas_in_seg = seg == 0 ? pl.g0 : pl.G - pl.g0;
    const int seg_tiles = (int)((seg_hi - seg_lo + CV_BM - 1) / CV_BM);
    const int my_tiles = cta_in_seg < seg_tiles ? (seg_tiles - cta_in_seg + ctas_in_seg - 1) / ctas_in_seg : 0;
    constexpr uint32_t TMEM_COLS = BN <= 64 ? 128u : (BN <= 128 ? 256u : 512u);

    if (tid == 0) {
        for (int s = 0; s < pl.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), CV_EPI_WARPS); }
        mbar_init(wfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < BN; j += CV_THREADS) shift_s[j] = shift != nullptr ? shift[n0 + j] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
            mbar_expect_tx(wfull_bar, (uint32_t)pl.nkb * BN * 128u);
            for (int kb = 0; kb < pl.nkb; ++kb) tma_load_2d(w_base + (uint32_t)kb * BN * 128u, &tmW, wfull_bar, kb * CV_BK, n0);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const long long row0 = seg_lo + (long long)(cta_in_seg + it * ctas_in_seg) * CV_BM;
            for (int kb = 0; kb < pl.nkb; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(s), CV_A_SLAB);
                    tma_load_2d(a_base + (uint32_t)s * CV_A_SLAB, &tmA, full_bar(s), kb * CV_BK, (int)row0);
                }
                __syncwarp();
                if (++s == pl.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = BN
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CV_BM >> 4) << 24);
        mbar_wait(wfull_bar, 0u);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            mbar_wait(tempty_bar(buf), (((uint32_t)it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + (uint32_t)buf * BN;
            for (int kb = 0; kb < pl.nkb; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_s = a_base + (uint32_t)s * CV_A_SLAB, b_s = w_base + (uint32_t)kb * BN * 128u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < CV_BK / 16; ++k)              // UMMA_K = 16 bf16 = 32 bytes
                        tc_mma_f16(d, tc_smem_desc(a_s + k * 32), tc_smem_desc(b_s + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(empty_bar(s));
                    if (kb == pl.nkb - 1) tc_commit(tfull_bar(buf));
                }
                __syncwarp();
                if (++s == pl.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue warps: thread <-> one accumulator row.  Two warps per TMEM lane quarter, each owning a contiguous
        // half of the tile's columns, processed in groups of GC 16-column chunks (<= 128 bytes of bf16 per row): the
        // converted rows are staged in a per-warp, XOR-swizzled shared-memory tile and leave as full 128-byte lines
        // (a warp store covers 4 whole rows instead of 32 partial ones). =====
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;
        constexpr int NCH = BN / 16;
        constexpr int EPW = CV_EPI_WARPS / 4;
        constexpr int MYCH = NCH / EPW;                                // chunks per warp
        constexpr int GC = MYCH < 4 ? MYCH : 4;                        // chunks per staging group
        constexpr int SEGS = GC * 2;                                   // 16-byte segments per staged row (<= 8)
        constexpr int RPI = 32 / SEGS;                                 // rows per copy-out instruction
        uint8_t* stg = gen + (size_t)BN * 9 * sizeof(float) + (size_t)(warp - 2) * 4096;     // [32 rows][128 B]
        float acc_s[MYCH], acc_q[MYCH];
#pragma unroll
        for (int c = 0; c < MYCH; ++c) { acc_s[c] = 0.f; acc_q[c] = 0.f; }
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const long long row0 = seg_lo + (long long)(cta_in_seg + it * ctas_in_seg) * CV_BM + q * 32;
            const bool valid = row0 + lane < seg_hi;
            mbar_wait(tfull_bar(buf), ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * BN + (uint32_t)(half * MYCH * 16);
            uint32_t r[16];
            tc_ld16(taddr, r);
#pragma unroll
            for (int cc = 0; cc < MYCH; ++cc) {
                const int c = half * MYCH + cc;                        // chunk index within the tile
                tc_ld_wait();
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
                if (cc + 1 < MYCH) tc_ld16(taddr + (uint32_t)((cc + 1) * 16), r);     // next chunk flies while this one is processed
                {
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                        pk[e] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    const int sg = (cc % GC) * 2;
                    uint8_t* rowp = stg + lane * 128;
                    *reinterpret_cast<uint4*>(rowp + (((sg) ^ (lane & (SEGS - 1))) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(rowp + (((sg + 1) ^ (lane & (SEGS - 1))) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
                if (part != nullptr) {
                    // column sums over the warp's 32 rows: recursive halving, 16 + 16 shuffles for 16 columns x (sum, sumsq)
                    float a[16], b[16];
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 sh = *reinterpret_cast<const float4*>(shift_s + c * 16 + e4 * 4);
                        const float s4[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float d = valid ? v[e4 * 4 + u] - s4[u] : 0.f;
                            a[e4 * 4 + u] = d;
                            b[e4 * 4 + u] = d * d;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float sa = b4 ? a[j] : a[j + 8], ka = b4 ? a[j + 8] : a[j];
                        const float sb = b4 ? b[j] : b[j + 8], kb_ = b4 ? b[j + 8] : b[j];
                        a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, 16);
                        b[j] = kb_ + __shfl_xor_sync(0xffffffffu, sb, 16);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float sa = b3 ? a[j] : a[j + 4], ka = b3 ? a[j + 4] : a[j];
                        const float sb = b3 ? b[j] : b[j + 4], kb_ = b3 ? b[j + 4] : b[j];
                        a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, 8);
                        b[j] = kb_ + __shfl_xor_sync(0xffffffffu, sb, 8);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float sa = b2 ? a[j] : a[j + 2], ka = b2 ? a[j + 2] : a[j];
                        const float sb = b2 ? b[j] : b[j + 2], kb_ = b2 ? b[j + 2] : b[j];
                        a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, 4);
                        b[j] = kb_ + __shfl_xor_sync(0xffffffffu, sb, 4);
                    }
                    {
                        const float sa = b1 ? a[0] : a[1], ka = b1 ? a[1] : a[0];
                        const float sb = b1 ? b[0] : b[1], kb_ = b1 ? b[1] : b[0];
                        a[0] = ka + __shfl_xor_sync(0xffffffffu, sa, 2);
                        b[0] = kb_ + __shfl_xor_sync(0xffffffffu, sb, 2);
                    }
                    acc_s[cc] += a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
                    acc_q[cc] += b[0] + __shfl_xor_sync(0xffffffffu, b[0], 1);
                }
                if ((cc % GC) == GC - 1) {
                    // copy the staged [32 rows][GC*32 B] block out: lane -> (row = i*RPI + lane / SEGS, segment = lane % SEGS)
                    __syncwarp();
                    const int sgm = lane % SEGS, rsub = lane / SEGS;
                    const int colbase = n0 + (half * MYCH + cc - (GC - 1)) * 16;
#pragma unroll
                    for (int i = 0; i < 32 / RPI; ++i) {
                        const int rr = i * RPI + rsub;
                        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((sgm ^ (rr & (SEGS - 1))) << 4));
                        if (row0 + rr < seg_hi)
                            *reinterpret_cast<uint4*>(out + (row0 + rr) * pl.N + colbase + sgm * 8) = val;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        if (part != nullptr) {
            // lane l holds column col(l) = 8*b4 + 4*b3 + 2*b2 + b1 of every chunk (lanes l and l^1 hold the same sum)
            const int col = (b4 ? 8 : 0) + (b3 ? 4 : 0) + (b2 ? 2 : 0) + (b1 ? 1 : 0);
            if ((lane & 1) == 0) {
#pragma unroll
                for (int cc = 0; cc < MYCH; ++cc) {
                    const int c = half * MYCH + cc;
                    stat_s[(q * 2 + 0) * BN + c * 16 + col] = acc_s[cc];
                    stat_s[(q * 2 + 1) * BN + c * 16 + col] = acc_q[cc];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * CV_EPI_WARPS) : "memory");
            for (int j = tid - 64; j < BN; j += 32 * CV_EPI_WARPS) {
                const float sa = (stat_s[0 * BN + j] + stat_s[2 * BN + j]) + (stat_s[4 * BN + j] + stat_s[6 * BN + j]);
                const float sb = (stat_s[1 * BN + j] + stat_s[3 * BN + j]) + (stat_s[5 * BN + j] + stat_s[7 * BN + j]);
                part[(size_t)blockIdx.x * 2 * pl.N + n0 + j] = sa;
                part[(size_t)blockIdx.x * 2 * pl.N + pl.N + n0 + j] = sb;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

int pick_bn(int N, int K) {
    for (int bn = 256; bn >= 64; bn >>= 1)
        if (N % bn == 0 && (size_t)bn * K * 2 <= (size_t)CV_W_BUDGET) return bn;
    return 0;
}

size_t conv_smem(int bn, int K, int stages) {
    return (size_t)bn * K * 2 + (size_t)stages * CV_A_SLAB + 1024 /*align*/ + 256 /*barriers*/ + (size_t)bn * 9 * sizeof(float) +
           (size_t)CV_EPI_WARPS * 4096 /*output staging*/;
}

bool make_plan(long long P, long long P_split, int N, int K, ConvPlan* pl, int* bn_out, dim3* grid) {
    if (P <= 0 || K <= 0 || N <= 0 || K % CV_BK != 0 || N % 64 != 0) return false;
    const int bn = pick_bn(N, K);
    if (bn == 0 || encode_fn() == nullptr) return false;
    int stages = CV_MAX_STAGES;
    while (stages > 2 && conv_smem(bn, K, stages) > (size_t)CV_SMEM_BUDGET) --stages;
    if (conv_smem(bn, K, stages) > (size_t)CV_SMEM_BUDGET) return false;
    const int gy = N / bn;
    int G = ge::sm_count() / gy;
    if (G < 2) return false;
    const long long tiles = ge::cdivll(P, CV_BM);
    if (G > tiles) G = (int)(tiles < 2 ? 2 : tiles);
    pl->P = P;
    pl->P0 = (P_split > 0 && P_split < P) ? P_split : P;
    pl->G = G;
    pl->g0 = G;
    if (pl->P0 < P) {
        long long g0 = ((long long)G * pl->P0 + P / 2) / P;
        if (g0 < 1) g0 = 1;
        if (g0 > G - 1) g0 = G - 1;
        pl->g0 = (int)g0;
    }
    pl->N = N; pl->K = K; pl->nkb = K / CV_BK; pl->stages = stages;
    *bn_out = bn;
    *grid = dim3((unsigned)G, (unsigned)gy);
    return true;
}

}  // namespace

// Can ge_conv1x1_bn_stats run this shape (bf16, K % 64 == 0, N % 64 == 0, the weight tile fits shared memory)?
extern "C" int ge_conv1x1_tc_supported(long long P, int K, int N) {
    ConvPlan pl;
    int bn;
    dim3 grid;
    return make_plan(P, 0, N, K, &pl, &bn, &grid) ? 1 : 0;
}

// rows of the partial-statistics workspace = CTAs along x of the launch for this shape (0 = unsupported)
extern "C" int ge_conv1x1_tc_partial_rows(long long P, long long P_split, int K, int N, int* rows_segment0) {
    ConvPlan pl;
    int bn;
    dim3 grid;
    if (!make_plan(P, P_split, N, K, &pl, &bn, &grid)) return 0;
    if (rows_segment0) *rows_segment0 = pl.g0;
    return pl.G;
}

// y [P,N] bf16 = x [P,K] bf16 * W[N,K]^T bf16 (fp32 accumulate).  part (or NULL) receives per-CTA partial rows
// [rows][2][N] fp32 of sum(y - shift[n]) and sum((y - shift[n])^2) over each CTA's pixels (shift may be NULL = 0), rows
// [0, rows_segment0) covering pixels [0, P_split) and the rest [P_split, P): exactly what ge_bn_fwd_train_prestat reads.
extern "C" int ge_conv1x1_bn_stats(const void* x, const void* w, void* y, const float* shift, float* part,
                                   long long P, long long P_split, int K, int N, ge_stream_t stream) {
    GE_REQUIRE(x && w && y, GE_ERR_ARG, "ge_conv1x1_bn_stats: null pointer");
    ConvPlan pl;
    int bn;
    dim3 grid;
    GE_REQUIRE(make_plan(P, P_split, N, K, &pl, &bn, &grid), GE_ERR_SHAPE,
               "ge_conv1x1_bn_stats: unsupported shape P=%lld K=%d N=%d (K %% 64, N %% 64, weight tile <= 128 KB)", P, K, N);
    CUtensorMap mA, mW;
    GE_REQUIRE(make_map_2d_bf16(&mA, x, P, K, CV_BM) && make_map_2d_bf16(&mW, w, N, K, bn), GE_ERR_SHAPE,
               "ge_conv1x1_bn_stats: cuTensorMapEncodeTiled failed (P=%lld K=%d N=%d)", P, K, N);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = conv_smem(bn, K, pl.stages);
    static size_t cached[3] = {0, 0, 0};
#define GE_CV_LAUNCH(BNV, slot)                                                                                              \
    do {                                                                                                                     \
        if (smem > cached[slot]) {                                                                                           \
            GE_CUDA(cudaFuncSetAttribute(conv1x1_tc_kernel<BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),     \
                    "ge_conv1x1_bn_stats(attr)");                                                                            \
            cached[slot] = smem;                                                                                             \
        }                                                                                                                    \
        conv1x1_tc_kernel<BNV><<<grid, CV_THREADS, smem, st>>>(mA, mW, (bf16*)y, shift, part, pl);                             \
    } while (0)
    if (bn == 256) GE_CV_LAUNCH(256, 0);
    else if (bn == 128) GE_CV_LAUNCH(128, 1);
    else GE_CV_LAUNCH(64, 2);
#undef GE_CV_LAUNCH
    GE_CHECK_LAUNCH("ge_conv1x1_bn_stats");
    return GE_OK;
}
