// 3x3 / stride 2 / pad 1 max pooling on NHWC maps, forward and backward — the ResNet stem's nn.MaxPool2d(3, 2, 1)
// (/root/reference/models/fpnseg.py:232, 254).
//
// Forward: a thread owns 8 channels of one output pixel, reads its (up to) 9 input pixels as 128-bit rows and
// records the arg-max tap (0..8, first maximum in row-major window order like ATen) as one byte per element.
// Backward: GATHER form -- a thread owns 8 channels of one INPUT pixel and sums the gradients of the (at most 4)
// windows that selected it: no atomics, deterministic, every tensor read or written exactly once.
// HBM-bound: forward reads H*W + writes (H/2)*(W/2) elements (+1 byte each), backward the reverse.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using bf16 = __nv_bfloat16;

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, unsigned char* __restrict__ arg,
                      int H, int W, int Ho, int Wo, int C, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c8 = C >> 3;
    const int cc = (int)(e % c8) * 8;
    long long p = e / c8;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const long long n = p / Ho;
    float best[8];
    int tap[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { best[u] = -INFINITY; tap[u] = 0; }
    const T* xb = x + n * H * W * C + cc;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int iy = oy * 2 - 1 + dy;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int ix = ox * 2 - 1 + dx;
            if (ix < 0 || ix >= W) continue;
            float v[8];
            ge::load8<T>(xb + ((long long)iy * W + ix) * C, v);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (v[u] > best[u] || v[u] != v[u]) { best[u] = v[u]; tap[u] = dy * 3 + dx; }
        }
    }
    const long long o = ((n * Ho + oy) * Wo + ox) * C + cc;
    ge::store8<T>(out + o, best);
    unsigned long long packed = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) packed |= (unsigned long long)tap[u] << (8 * u);
    *reinterpret_cast<unsigned long long*>(arg + o) = packed;
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const T* __restrict__ dout, const unsigned char* __restrict__ arg, T* __restrict__ dx,
                      int H, int W, int Ho, int Wo, int C, long long total) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int c8 = C >> 3;
    const int cc = (int)(e % c8) * 8;
    long long p = e / c8;
    const int ix = (int)(p % W); p /= W;
    const int iy = (int)(p % H);
    const long long n = p / H;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // windows (oy, ox) with oy*2-1 <= iy <= oy*2+1
    const int oy0 = max(0, iy / 2), oy1 = min(Ho - 1, (iy + 1) / 2);
    const int ox0 = max(0, ix / 2), ox1 = min(Wo - 1, (ix + 1) / 2);
    for (int oy = oy0; oy <= oy1; ++oy) {
        const int dy = iy - (oy * 2 - 1);
        if (dy < 0 || dy > 2) continue;
        for (int ox = ox0; ox <= ox1; ++ox) {
            const int dxx = ix - (ox * 2 - 1);
            if (dxx < 0 || dxx > 2) continue;
            const long long o = ((n * Ho + oy) * Wo + ox) * C + cc;
            const unsigned long long packed = *reinterpret_cast<const unsigned long long*>(arg + o);
            float g[8];
            ge::load8<T>(dout + o, g);
            const unsigned want = (unsigned)(dy * 3 + dxx);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (((packed >> (8 * u)) & 0xffu) == want) acc[u] += g[u];
        }
    }
    ge::store8<T>(dx + ((n * H + iy) * W + ix) * C + cc, acc);
}

}  // namespace

// x [N,H,W,C] -> out [N,Ho,Wo,C] with Ho = (H-1)/2+1, Wo = (W-1)/2+1; arg uint8 [N,Ho,Wo,C] (window tap 0..8).
extern "C" int ge_maxpool3s2_fwd(const void* x, void* out, unsigned char* arg, int dtype,
                                 int N, int H, int W, int C, ge_stream_t stream) {
    GE_REQUIRE(x && out && arg, GE_ERR_ARG, "ge_maxpool3s2_fwd: null pointer");
    GE_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, GE_ERR_ARG, "ge_maxpool3s2_fwd: bad dimension");
    GE_REQUIRE(C % 8 == 0, GE_ERR_SHAPE, "ge_maxpool3s2_fwd: C=%d must be a multiple of 8", C);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    const unsigned blocks = (unsigned)ge::cdivll(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        maxpool3s2_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)out, arg, H, W, Ho, Wo, C, total);
    else if (dtype == GE_DTYPE_BF16)
        maxpool3s2_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)out, arg, H, W, Ho, Wo, C, total);
    else { ge_set_error("ge_maxpool3s2_fwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_maxpool3s2_fwd");
    return GE_OK;
}

extern "C" int ge_maxpool3s2_bwd(const void* dout, const unsigned char* arg, void* dx, int dtype,
                                 int N, int H, int W, int C, ge_stream_t stream) {
    GE_REQUIRE(dout && arg && dx, GE_ERR_ARG, "ge_maxpool3s2_bwd: null pointer");
    GE_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, GE_ERR_ARG, "ge_maxpool3s2_bwd: bad dimension");
    GE_REQUIRE(C % 8 == 0, GE_ERR_SHAPE, "ge_maxpool3s2_bwd: C=%d must be a multiple of 8", C);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const long long total = (long long)N * H * W * (C / 8);
    const unsigned blocks = (unsigned)ge::cdivll(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        maxpool3s2_bwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)dout, arg, (float*)dx, H, W, Ho, Wo, C, total);
    else if (dtype == GE_DTYPE_BF16)
        maxpool3s2_bwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)dout, arg, (bf16*)dx, H, W, Ho, Wo, C, total);
    else { ge_set_error("ge_maxpool3s2_bwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_maxpool3s2_bwd");
    return GE_OK;
}
