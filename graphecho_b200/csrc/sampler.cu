// f1 -- node sampler, device half: per-location labels of every pyramid level and the (positive, negative) counts the
// host needs, in ONE launch per domain.
//
// Reference: PrototypeComputation.prepare_targets / compute_targets_for_locations (models/graph_matching.py:874-959)
// over GModule.compute_locations (:609-635): a location of level l sits at (x*s_l + s_l/2, y*s_l + s_l/2) with the
// reference's strides s = 8,16,32,64,128 (twice the true pyramid strides -- reproduced); it takes
// the class of the smallest-area class box that contains it (strictly) and whose largest side distance lies in the
// level's size range, else 0.  The reference loops over images in Python and builds [L,K,4] temporaries per image; the
// torch-vectorised version of round 1 was ~60 small launches per domain.  Integer/geometry work: bit-exact.
#include "common.cuh"
#include <algorithm>
#include "../../include/graphecho_b200.h"

namespace {

constexpr int SMP_MAXL = 5;
constexpr int SMP_MAXK = 8;
constexpr float SMP_INF = 100000000.f;

struct Levels {
    int n;
    int h[SMP_MAXL], w[SMP_MAXL], stride[SMP_MAXL];
    float lo[SMP_MAXL], hi[SMP_MAXL];
    long long off[SMP_MAXL + 1];       // offsets of the levels in the flat label array, per image-major level block
};

// labels: int64, level l occupies [off[l]*B .. off[l+1]*B) as [B][h_l*w_l]; counts int32 [L][2] (zero-filled by the caller)
__global__ void __launch_bounds__(256)
sampler_labels_kernel(const float* __restrict__ boxes, long long* __restrict__ labels, int* __restrict__ counts,
                      Levels lv, int B, int K) {
    __shared__ int cnt[SMP_MAXL][2];
    if (threadIdx.x < SMP_MAXL * 2) cnt[threadIdx.x >> 1][threadIdx.x & 1] = 0;
    __syncthreads();
    const long long total = lv.off[lv.n] * B;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int l = 0;
        while (l + 1 < lv.n && e >= lv.off[l + 1] * B) ++l;
        const long long r = e - lv.off[l] * B;
        const int hw = lv.h[l] * lv.w[l];
        const int b = (int)(r / hw), loc = (int)(r - (long long)b * hw);
        const int yy = loc / lv.w[l], xx = loc - yy * lv.w[l];
        const float x = (float)(xx * lv.stride[l]) + (float)(lv.stride[l] / 2);
        const float y = (float)(yy * lv.stride[l]) + (float)(lv.stride[l] / 2);
        float amin = SMP_INF;
        int arg = 0;
        for (int k = 0; k < K; ++k) {
            const float4 bx = *reinterpret_cast<const float4*>(boxes + ((size_t)b * K + k) * 4);
            const float dl = x - bx.x, dt = y - bx.y, dr = bx.z - x, db = bx.w - y;
            const float mn = fminf(fminf(dl, dt), fminf(dr, db)), mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
            float area = (bx.w - bx.y) * (bx.z - bx.x);
            if (!(mn > 0.f) || !(mx >= lv.lo[l] && mx <= lv.hi[l])) area = SMP_INF;
            if (area < amin) { amin = area; arg = k; }          // first minimum wins (torch.min)
        }
        const int lab = amin == SMP_INF ? 0 : arg;
        labels[e] = lab;
        atomicAdd(&cnt[l][lab > 0 ? 0 : 1], 1);
    }
    __syncthreads();
    if (threadIdx.x < lv.n * 2 && cnt[threadIdx.x >> 1][threadIdx.x & 1] != 0)
        atomicAdd(&counts[threadIdx.x], cnt[threadIdx.x >> 1][threadIdx.x & 1]);
}

}  // namespace

// boxes fp32 [B,K,4] (xmin,ymin,xmax,ymax per class plane, ge_mask_boxes); level geometry: heights / widths / location
// strides and the [lo, hi] size range of each of the `levels` pyramid levels; labels int64 [B * sum(h_l*w_l)] (level-major,
// image-major inside a level -- the layout torch.split(..., dim=1) of the reference produces after flattening);
// counts int32 [levels][2] = (#label > 0, #label == 0), MUST be zero-filled.
extern "C" int ge_sampler_labels(const float* boxes, long long* labels, int* counts, const int* heights, const int* widths,
                                 const int* strides, const float* size_lo, const float* size_hi, int levels, int B, int K,
                                 ge_stream_t stream) {
    GE_REQUIRE(boxes && labels && counts && heights && widths && strides && size_lo && size_hi, GE_ERR_ARG, "ge_sampler_labels: null pointer");
    GE_REQUIRE(levels >= 1 && levels <= SMP_MAXL && B > 0 && K >= 1 && K <= SMP_MAXK, GE_ERR_SHAPE,
               "ge_sampler_labels: levels=%d (<= %d), K=%d (<= %d), B=%d", levels, SMP_MAXL, K, SMP_MAXK, B);
    Levels lv;
    lv.n = levels;
    lv.off[0] = 0;
    for (int l = 0; l < levels; ++l) {
        GE_REQUIRE(heights[l] > 0 && widths[l] > 0 && strides[l] > 0, GE_ERR_ARG, "ge_sampler_labels: bad level geometry");
        lv.h[l] = heights[l]; lv.w[l] = widths[l]; lv.stride[l] = strides[l];
        lv.lo[l] = size_lo[l]; lv.hi[l] = size_hi[l];
        lv.off[l + 1] = lv.off[l] + (long long)heights[l] * widths[l];
    }
    const long long total = lv.off[levels] * B;
    const int grid = (int)std::min<long long>(ge::cdivll(total, 256), (long long)ge::sm_count() * 8);
    sampler_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, labels, counts, lv, B, K);
    GE_CHECK_LAUNCH("ge_sampler_labels");
    return GE_OK;
}
