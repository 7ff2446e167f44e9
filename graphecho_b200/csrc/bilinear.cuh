// Bilinear (align_corners=True) coordinate helpers shared by the NHWC feature-map kernels.
#pragma once
#include "common.cuh"

namespace ge {

// align_corners=True source coordinate (PyTorch area_pixel_compute_scale / source index)
__device__ __forceinline__ void src_coord(int dst, float scale, int in, int& i0, int& i1, float& l1) {
    const float s = scale * (float)dst;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = s - (float)i0;
}
__host__ __device__ __forceinline__ float ac_scale(int in, int out) {
    return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
}
// range of destination indices whose taps can touch source index s
__device__ __forceinline__ void dst_range(int s, float scale, int out, int& lo, int& hi) {
    if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
    lo = (int)floorf((float)(s - 1) / scale) - 1;
    hi = (int)ceilf((float)(s + 1) / scale) + 1;
    if (lo < 0) lo = 0;
    if (hi > out - 1) hi = out - 1;
}
// weight of destination index d onto source index s along one axis
__device__ __forceinline__ float tap_weight(int d, int s, float scale, int in) {
    int i0, i1; float l1;
    src_coord(d, scale, in, i0, i1, l1);
    float w = 0.f;
    if (i0 == s) w += 1.f - l1;
    if (i1 == s) w += l1;
    return w;
}

}  // namespace ge
