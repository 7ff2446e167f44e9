// K6 (phase 1) — TGCN pyramid aggregation: avg_pool2d(r_l) of each FPN level + channel concat.
//
// Replaces the per-timestep  F.avg_pool2d x3 + torch.cat  of TGCN.DyGraphConv2d.forward
// (/root/reference/models/TGCN.py:62-70), which runs inside the python loop over t.  The
// pooling does not depend on the recurrent state, so it is hoisted out of the recurrence and
// done for all b*t frames in one launch per level: each level is read exactly once
// (5.57 MB/frame at 256^2: 4*256*(64^2+32^2+16^2+8^2)) and the pooled [BT, sum C_l, Ho, Wo]
// tensor is written channels_last, ready for the 1x1-conv GEMM of the MLP.
// Inputs are NHWC-dense views (channel stride 1) with an arbitrary frame stride.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using bf16 = __nv_bfloat16;

// thread per (frame, oy, ox, 4 channels)
template <typename T>
__global__ void __launch_bounds__(256)
pool_concat_fwd_kernel(const T* __restrict__ in, long long frame_stride, float* __restrict__ out,
                       int H, int W, int C, int r, int Ho, int Wo, int Ctot, int coff, long long total4) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const int c4 = C >> 2;
    const int cc = (int)(e % c4) * 4;
    long long p = e / c4;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const long long f = p / Ho;
    const T* ib = in + f * frame_stride + cc;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < r; ++dy)
        for (int dx = 0; dx < r; ++dx) {
            ge::Vec4<T> v; v.load(ib + ((size_t)(oy * r + dy) * W + (ox * r + dx)) * C);
            float fv[4]; v.get(fv);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += fv[q];
        }
    const float div = (float)(r * r);
    float4 o = make_float4(acc[0] / div, acc[1] / div, acc[2] / div, acc[3] / div);
    *reinterpret_cast<float4*>(out + ((f * Ho + oy) * Wo + ox) * Ctot + coff + cc) = o;
}

// thread per (frame, y, x, 4 channels) of the level's gradient
template <typename T>
__global__ void __launch_bounds__(256)
pool_concat_bwd_kernel(const float* __restrict__ dout, T* __restrict__ din,
                       int H, int W, int C, int r, int Ho, int Wo, int Ctot, int coff, long long total4) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total4) return;
    const int c4 = C >> 2;
    const int cc = (int)(e % c4) * 4;
    long long p = e / c4;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const long long f = p / H;
    const int oy = y / r, ox = x / r;
    float res[4] = {0.f, 0.f, 0.f, 0.f};
    if (oy < Ho && ox < Wo) {
        const float4 g = *reinterpret_cast<const float4*>(dout + ((f * Ho + oy) * Wo + ox) * Ctot + coff + cc);
        const float div = (float)(r * r);
        res[0] = g.x / div; res[1] = g.y / div; res[2] = g.z / div; res[3] = g.w / div;
    }
    ge::Vec4<T> o; o.set(res);
    o.store(din + ((f * H + y) * W + x) * C + cc);
}

}  // namespace

// in: level feature map, NHWC-dense per frame ([H,W,C], channel stride 1), frames `frame_stride`
// elements apart; out: fp32 [frames, Ho, Wo, Ctot] (channels_last), this level fills channels
// [coff, coff+C).  Ho = H / r, Wo = W / r (floor, as avg_pool2d).
extern "C" int ge_tgcn_pool_concat_fwd(const void* in, long long frame_stride, float* out, int dtype,
                                       long long frames, int H, int W, int C, int r, int Ctot, int coff,
                                       ge_stream_t stream) {
    GE_REQUIRE(in && out, GE_ERR_ARG, "ge_tgcn_pool_concat_fwd: null pointer");
    GE_REQUIRE(frames > 0 && H > 0 && W > 0 && C > 0 && r > 0 && Ctot >= coff + C && coff >= 0, GE_ERR_ARG,
               "ge_tgcn_pool_concat_fwd: bad dimension");
    GE_REQUIRE(C % 4 == 0 && coff % 4 == 0 && Ctot % 4 == 0, GE_ERR_SHAPE, "ge_tgcn_pool_concat_fwd: channels must be multiples of 4");
    const int Ho = H / r, Wo = W / r;
    GE_REQUIRE(Ho > 0 && Wo > 0, GE_ERR_SHAPE, "ge_tgcn_pool_concat_fwd: pooling window larger than the map");
    const long long total4 = frames * Ho * Wo * (C / 4);
    const unsigned blocks = (unsigned)ge::cdivll(total4, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        pool_concat_fwd_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, frame_stride, out, H, W, C, r, Ho, Wo, Ctot, coff, total4);
    else if (dtype == GE_DTYPE_BF16)
        pool_concat_fwd_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)in, frame_stride, out, H, W, C, r, Ho, Wo, Ctot, coff, total4);
    else { ge_set_error("ge_tgcn_pool_concat_fwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_tgcn_pool_concat_fwd");
    return GE_OK;
}

// dout: fp32 [frames, Ho, Wo, Ctot]; din: [frames, H, W, C] NHWC-dense, fully overwritten.
extern "C" int ge_tgcn_pool_concat_bwd(const float* dout, void* din, int dtype,
                                       long long frames, int H, int W, int C, int r, int Ctot, int coff,
                                       ge_stream_t stream) {
    GE_REQUIRE(dout && din, GE_ERR_ARG, "ge_tgcn_pool_concat_bwd: null pointer");
    GE_REQUIRE(frames > 0 && H > 0 && W > 0 && C > 0 && r > 0 && Ctot >= coff + C && coff >= 0, GE_ERR_ARG,
               "ge_tgcn_pool_concat_bwd: bad dimension");
    GE_REQUIRE(C % 4 == 0 && coff % 4 == 0 && Ctot % 4 == 0, GE_ERR_SHAPE, "ge_tgcn_pool_concat_bwd: channels must be multiples of 4");
    const int Ho = H / r, Wo = W / r;
    const long long total4 = frames * H * W * (C / 4);
    const unsigned blocks = (unsigned)ge::cdivll(total4, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32)
        pool_concat_bwd_kernel<float><<<blocks, 256, 0, st>>>(dout, (float*)din, H, W, C, r, Ho, Wo, Ctot, coff, total4);
    else if (dtype == GE_DTYPE_BF16)
        pool_concat_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(dout, (bf16*)din, H, W, C, r, Ho, Wo, Ctot, coff, total4);
    else { ge_set_error("ge_tgcn_pool_concat_bwd: unsupported dtype %d", dtype); return GE_ERR_DTYPE; }
    GE_CHECK_LAUNCH("ge_tgcn_pool_concat_bwd");
    return GE_OK;
}
