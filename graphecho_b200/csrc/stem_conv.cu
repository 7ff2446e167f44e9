// Stem convolution of the ResNet backbone: Conv2d(1, 64, kernel 7, stride 2, padding 3, bias=False) on gray frames
// (/root/reference/models/fpnseg.py:229, 251: `self.conv1`), forward and weight gradient, as direct FP32-pipe kernels.
//
// cuDNN has no tensor-core path for one input channel: in the training step the library spends 0.46 ms forward
// (layout conversion 0.12 + an sm80 TF32 implicit GEMM 0.24 + an output conversion 0.10) and 0.64 ms on the weight
// gradient (a strided copy 0.21, two layout conversions 0.20, the wgrad kernel 0.22) for 2.5 GFLOP each way -- 4.8 % of
// the step (profiles/r2i_launches_step.md).  The operation is tiny per output (49 taps) and the output is the big
// tensor (103 MB at 256 frames): a direct kernel that keeps the 49 x 64 filter and the input rows of its tile in shared
// memory, computes 4 pixels x 16 channels per thread and writes full 128-byte NHWC lines is bound by the FP32 pipe
// (2.5 GFLOP -> ~70 us) and the one compulsory write.
//
// Numerics: in bf16 mode (output type bf16, what autocast runs) the input and the filter are rounded to bf16 first and
// products are accumulated in fp32, as the library's bf16 convolution does; in fp32 mode everything is fp32.
// The input needs no gradient (it is the data); the backward is the weight gradient only.
#include "common.cuh"
#include <algorithm>
#include <type_traits>
#include "../../include/graphecho_b200.h"

namespace {

constexpr int SC_K = 7, SC_TAPS = 49, SC_C = 64, SC_RT = 4;   // filter size, taps, output channels, output rows per tile
constexpr int SC_THREADS = 256;

__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <typename T> __device__ __forceinline__ void store16(T* p, const float (&f)[16]);
template <> __device__ __forceinline__ void store16<float>(float* p, const float (&f)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(p + 4 * q) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
}
template <> __device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[16]) {
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = f[u]; b[u] = f[8 + u]; }
    ge::store8<__nv_bfloat16>(p, a);
    ge::store8<__nv_bfloat16>(p + 8, b);
}
template <typename T> __device__ __forceinline__ void load16(const T* p, float (&f)[16]);
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&f)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
        f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
    }
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[16]) {
    float a[8], b[8];
    ge::load8<__nv_bfloat16>(p, a);
    ge::load8<__nv_bfloat16>(p + 8, b);
#pragma unroll
    for (int u = 0; u < 8; ++u) { f[u] = a[u]; f[8 + u] = b[u]; }
}

// input rows of one tile -> shared memory: patch[(2 RT + 5)][pw], pw = 2 Wo + 8; patch row r = input row 2 oy0 - 3 + r,
// patch column c = input column c - 3 (zero outside the image; one spare zero column on the right)
template <bool BF16>
__device__ __forceinline__ void stage_patch(const float* __restrict__ xf, float* patch, int H, int W, int oy0, int pw) {
    constexpr int PR = 2 * SC_RT + 5;
    for (int e = threadIdx.x; e < PR * pw; e += blockDim.x) {
        const int r = e / pw, c = e - r * pw;
        const int iy = 2 * oy0 - 3 + r, ix = c - 3;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xf + (size_t)iy * W + ix);
        patch[e] = BF16 ? round_bf16(v) : v;
    }
}

// ---------------------------------------------------------------------------------------------- forward
// CTA = (frame, SC_RT output rows); work item = 4 consecutive output pixels x 16 channels.
template <typename T>
__global__ void __launch_bounds__(SC_THREADS)
stem_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                     int H, int W, int Ho, int Wo) {
    constexpr bool BF16 = !std::is_same<T, float>::value;
    extern __shared__ __align__(16) float sm[];
    float* ws = sm;                              // [49][64]  (tap-major, channel contiguous)
    float* patch = sm + SC_TAPS * SC_C;          // [2 RT + 5][pw]
    const int pw = 2 * Wo + 8;
    const int f = blockIdx.y, oy0 = blockIdx.x * SC_RT;
    for (int e = threadIdx.x; e < SC_TAPS * SC_C; e += SC_THREADS) {
        const int tap = e / SC_C, c = e - tap * SC_C;
        const float v = __ldg(w + (size_t)c * SC_TAPS + tap);     // weight [64][1][7][7]
        ws[e] = BF16 ? round_bf16(v) : v;
    }
    stage_patch<BF16>(x + (size_t)f * H * W, patch, H, W, oy0, pw);
    __syncthreads();
    const int quads = Wo >> 2;                   // Wo % 4 == 0 (host)
    const int items = SC_RT * quads * 4;
    for (int it = threadIdx.x; it < items; it += SC_THREADS) {
        const int g = it & 3, q = (it >> 2) % quads, r = (it >> 2) / quads;
        const int oy = oy0 + r;
        if (oy >= Ho) continue;
        float acc[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[j][u] = 0.f;
#pragma unroll 1
        for (int ky = 0; ky < SC_K; ++ky) {
            const float* prow = patch + (2 * r + ky) * pw + 8 * q;      // input column 2 (4q) - 3 + 3 = 8q in patch coords
            float xin[13];
#pragma unroll
            for (int u = 0; u < 13; ++u) xin[u] = prow[u];
#pragma unroll
            for (int kx = 0; kx < SC_K; ++kx) {
                const float* wt = ws + (ky * SC_K + kx) * SC_C + 16 * g;
                float wv[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 v = *reinterpret_cast<const float4*>(wt + 4 * q4);
                    wv[4 * q4] = v.x; wv[4 * q4 + 1] = v.y; wv[4 * q4 + 2] = v.z; wv[4 * q4 + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int u = 0; u < 16; ++u) acc[j][u] = fmaf(xin[2 * j + kx], wv[u], acc[j][u]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            store16<T>(y + (((size_t)f * Ho + oy) * Wo + 4 * q + j) * SC_C + 16 * g, acc[j]);
    }
}

// ---------------------------------------------------------------------------------------------- weight gradient
// Persistent CTAs over (frame, SC_RT-row) tiles.  Thread = (tap group tg of 14: one ky and kx 0..3 or 4..6, channel
// group g of 4 x 16 channels, output row ps of the tile): 64 accumulators live across all tiles of the CTA; per output
// pixel 4 input values + 16 upstream gradients from shared memory and 64 FFMAs.  Per-CTA partial -> part[cta][49][64].
template <typename T>
__global__ void __launch_bounds__(SC_THREADS)
stem_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ part,
                       int F, int H, int W, int Ho, int Wo) {
    constexpr bool BF16 = !std::is_same<T, float>::value;
    extern __shared__ __align__(16) float sm[];
    const int pw = 2 * Wo + 8;
    constexpr int PR = 2 * SC_RT + 5;
    float* patch = sm;                                                   // [PR][pw]
    T* dyt = reinterpret_cast<T*>(sm + ((PR * pw + 3) & ~3));            // [SC_RT][Wo][64]
    const int tid = threadIdx.x;
    const int tg = tid % 14, g = (tid / 14) & 3, ps = tid / 56;          // 14 x 4 x 4 = 224 active threads
    const bool active = tid < 224;
    const int ky = tg >> 1, kx0 = (tg & 1) * 4;
    float acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[j][u] = 0.f;
    const int row_tiles = ge::cdiv(Ho, SC_RT);
    const int tiles = F * row_tiles;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int f = tile / row_tiles, oy0 = (tile - f * row_tiles) * SC_RT;
        __syncthreads();                                                 // previous tile fully consumed
        stage_patch<BF16>(x + (size_t)f * H * W, patch, H, W, oy0, pw);
        const int rows = min(SC_RT, Ho - oy0);
        const int n16 = rows * Wo * SC_C / 8;                            // 16-byte chunks (bf16: 8 elements; fp32: 4 -> x2 below)
        const T* src = dy + ((size_t)f * Ho + oy0) * Wo * SC_C;
        if (BF16) {
            for (int e = tid; e < n16; e += SC_THREADS)
                reinterpret_cast<uint4*>(dyt)[e] = __ldg(reinterpret_cast<const uint4*>(src) + e);
        } else {
            for (int e = tid; e < 2 * n16; e += SC_THREADS)
                reinterpret_cast<uint4*>(dyt)[e] = __ldg(reinterpret_cast<const uint4*>(src) + e);
        }
        __syncthreads();
        if (active && ps < rows) {
            const float* prow = patch + (2 * ps + ky) * pw + kx0;
            const T* drow = dyt + (size_t)ps * Wo * SC_C + 16 * g;
#pragma unroll 2
            for (int ox = 0; ox < Wo; ++ox) {
                float d[16];
                load16<T>(drow + (size_t)ox * SC_C, d);
                const float x0 = prow[2 * ox], x1 = prow[2 * ox + 1], x2 = prow[2 * ox + 2], x3 = prow[2 * ox + 3];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    acc[0][u] = fmaf(x0, d[u], acc[0][u]);
                    acc[1][u] = fmaf(x1, d[u], acc[1][u]);
                    acc[2][u] = fmaf(x2, d[u], acc[2][u]);
                    acc[3][u] = fmaf(x3, d[u], acc[3][u]);
                }
            }
        }
    }
    // reduce the 4 row-threads of each (tap group, channel group) through shared memory, write the CTA's partial
    __syncthreads();
    float* red = sm;                                                     // [4][49][64]
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kx = kx0 + j;
            if (kx < SC_K) {
                float* dst = red + ((size_t)ps * SC_TAPS + ky * SC_K + kx) * SC_C + 16 * g;
#pragma unroll
                for (int u = 0; u < 16; ++u) dst[u] = acc[j][u];
            }
        }
    }
    __syncthreads();
    float* out = part + (size_t)blockIdx.x * SC_TAPS * SC_C;
    for (int e = tid; e < SC_TAPS * SC_C; e += SC_THREADS)
        out[e] = red[e] + red[SC_TAPS * SC_C + e] + red[2 * SC_TAPS * SC_C + e] + red[3 * SC_TAPS * SC_C + e];
}

// dw[c][tap] = sum over the CTA partials part[cta][tap][c]
__global__ void __launch_bounds__(256)
stem_conv_wgrad_finalize_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;                 // tap * 64 + c
    if (e >= SC_TAPS * SC_C) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 4 <= nparts; p += 4) {
        s0 += part[(size_t)p * SC_TAPS * SC_C + e];
        s1 += part[(size_t)(p + 1) * SC_TAPS * SC_C + e];
        s2 += part[(size_t)(p + 2) * SC_TAPS * SC_C + e];
        s3 += part[(size_t)(p + 3) * SC_TAPS * SC_C + e];
    }
    for (; p < nparts; ++p) s0 += part[(size_t)p * SC_TAPS * SC_C + e];
    const int tap = e / SC_C, c = e - tap * SC_C;
    dw[(size_t)c * SC_TAPS + tap] = (s0 + s1) + (s2 + s3);
}

size_t fwd_smem(int Wo) { return ((size_t)SC_TAPS * SC_C + (size_t)(2 * SC_RT + 5) * (2 * Wo + 8)) * sizeof(float); }
size_t wgrad_smem(int Wo, int es) {
    const size_t patch = (((size_t)(2 * SC_RT + 5) * (2 * Wo + 8) + 3) & ~(size_t)3) * sizeof(float);
    const size_t tile = (size_t)SC_RT * Wo * SC_C * es;
    const size_t red = (size_t)4 * SC_TAPS * SC_C * sizeof(float);
    return std::max(patch + tile, red);
}
int wgrad_ctas(int F, int Ho) { return std::min(F * ge::cdiv(Ho, SC_RT), 2 * ge::sm_count()); }

}  // namespace

extern "C" int ge_stem_conv_supported(int H, int W) {
    return (H > 0 && W > 0 && H % 2 == 0 && W % 8 == 0 && W <= 512) ? 1 : 0;
}

// x fp32 [F,1,H,W]; w fp32 [64,1,7,7]; y [F,H/2,W/2,64] NHWC in `dtype` (bf16: operands rounded to bf16, fp32 accumulate).
extern "C" int ge_stem_conv_fwd(const float* x, const float* w, void* y, int F, int H, int W, int dtype, ge_stream_t stream) {
    GE_REQUIRE(x && w && y, GE_ERR_ARG, "ge_stem_conv_fwd: null pointer");
    GE_REQUIRE(F > 0 && ge_stem_conv_supported(H, W), GE_ERR_SHAPE, "ge_stem_conv_fwd: unsupported frame size %dx%d (H even, W %% 8 == 0, W <= 512)", H, W);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_stem_conv_fwd: unsupported dtype %d", dtype);
    const int Ho = H / 2, Wo = W / 2;
    const size_t smem = fwd_smem(Wo);
    dim3 grid(ge::cdiv(Ho, SC_RT), F);
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(stem_conv_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_stem_conv_fwd(attr)"); c0 = smem; }
        stem_conv_fwd_kernel<float><<<grid, SC_THREADS, smem, (cudaStream_t)stream>>>(x, w, (float*)y, H, W, Ho, Wo);
    } else {
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(stem_conv_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_stem_conv_fwd(attr)"); c1 = smem; }
        stem_conv_fwd_kernel<__nv_bfloat16><<<grid, SC_THREADS, smem, (cudaStream_t)stream>>>(x, w, (__nv_bfloat16*)y, H, W, Ho, Wo);
    }
    GE_CHECK_LAUNCH("ge_stem_conv_fwd");
    return GE_OK;
}

extern "C" size_t ge_stem_conv_wgrad_workspace_bytes(int F, int H, int W) {
    if (F <= 0 || !ge_stem_conv_supported(H, W)) return 0;
    return (size_t)wgrad_ctas(F, H / 2) * SC_TAPS * SC_C * sizeof(float);
}

// dy [F,H/2,W/2,64] NHWC in `dtype`; dw fp32 [64,1,7,7] (overwritten); workspace: ge_stem_conv_wgrad_workspace_bytes.
extern "C" int ge_stem_conv_wgrad(const float* x, const void* dy, float* dw, void* workspace, size_t workspace_bytes,
                                  int F, int H, int W, int dtype, ge_stream_t stream) {
    GE_REQUIRE(x && dy && dw && workspace, GE_ERR_ARG, "ge_stem_conv_wgrad: null pointer");
    GE_REQUIRE(F > 0 && ge_stem_conv_supported(H, W), GE_ERR_SHAPE, "ge_stem_conv_wgrad: unsupported frame size %dx%d", H, W);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_stem_conv_wgrad: unsupported dtype %d", dtype);
    GE_REQUIRE(workspace_bytes >= ge_stem_conv_wgrad_workspace_bytes(F, H, W), GE_ERR_ARG, "ge_stem_conv_wgrad: workspace too small");
    const int Ho = H / 2, Wo = W / 2;
    const int ctas = wgrad_ctas(F, Ho);
    float* part = static_cast<float*>(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    static size_t c0 = 0, c1 = 0;
    if (dtype == GE_DTYPE_F32) {
        const size_t smem = wgrad_smem(Wo, 4);
        GE_REQUIRE(smem <= 220 * 1024, GE_ERR_CAPACITY, "ge_stem_conv_wgrad: frame width %d too large for the shared tile", W);
        if (smem > c0) { GE_CUDA(cudaFuncSetAttribute(stem_conv_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_stem_conv_wgrad(attr)"); c0 = smem; }
        stem_conv_wgrad_kernel<float><<<ctas, SC_THREADS, smem, st>>>(x, (const float*)dy, part, F, H, W, Ho, Wo);
    } else {
        const size_t smem = wgrad_smem(Wo, 2);
        GE_REQUIRE(smem <= 220 * 1024, GE_ERR_CAPACITY, "ge_stem_conv_wgrad: frame width %d too large for the shared tile", W);
        if (smem > c1) { GE_CUDA(cudaFuncSetAttribute(stem_conv_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ge_stem_conv_wgrad(attr)"); c1 = smem; }
        stem_conv_wgrad_kernel<__nv_bfloat16><<<ctas, SC_THREADS, smem, st>>>(x, (const __nv_bfloat16*)dy, part, F, H, W, Ho, Wo);
    }
    GE_CHECK_LAUNCH("ge_stem_conv_wgrad");
    stem_conv_wgrad_finalize_kernel<<<ge::cdiv(SC_TAPS * SC_C, 256), 256, 0, st>>>(part, ctas, dw);
    GE_CHECK_LAUNCH("ge_stem_conv_wgrad(finalize)");
    return GE_OK;
}
