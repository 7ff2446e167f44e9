// f4 -- the full-resolution passes on either side of the network in the trainers' inner loop:
//   * seg_loss = DiceLoss()(pred, masks) + BCEWithLogitsLoss()(pred, masks)          (train_cardiac_uda.py:228,
//     train_camus_echo.py:212; utils/losses.py:64-95): softmax over classes, per (frame, class) dice terms
//     sum(p*t), sum(p^2 + t^2), BCE-with-logits mean -- ONE pass over logits and masks forward (the reference runs
//     ~25 element-wise / reduction kernels), one pass backward that writes dlogits directly;
//   * score_maps = where(sigmoid(pred_target) > 0.5, 1, 0) (:235 / :219) followed by GModule.find_bbox /
//     masks_to_boxes (models/graph_matching.py:702-746): the bounding box of every (image, class) plane of the
//     thresholded map -- computed straight from the logits (sigmoid(x) > 0.5 <=> x > 0) without materialising the
//     int64 score map, one CTA per plane.
// HBM-bound streaming kernels over NCHW fp32 logits (the segmentation tail writes them in that layout).
// Algorithmic bytes: forward 8 B per logit (logit + mask), backward 12 B per logit.
#include "common.cuh"
#include "../../include/graphecho_b200.h"

namespace {

using namespace ge;
constexpr int SL_THREADS = 256;
constexpr int SL_MAXC = 8;

// partial [F][chunks][3*nc + 1]: per class (sum p*t, sum p^2, sum t^2), then the BCE sum of the chunk
template <int NC>
__global__ void __launch_bounds__(SL_THREADS)
seg_loss_partial_kernel(const float* __restrict__ logits, const float* __restrict__ target, float* __restrict__ part,
                        int HW, int chunks) {
    __shared__ float scratch[32];
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int per = cdiv(HW, chunks);
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    const float* x = logits + (size_t)n * NC * HW;
    const float* t = target + (size_t)n * NC * HW;
    float spt[NC], spp[NC], stt[NC], bce = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) { spt[c] = 0.f; spp[c] = 0.f; stt[c] = 0.f; }
    for (int p = p0 + threadIdx.x; p < p1; p += SL_THREADS) {
        float xv[NC], tv[NC], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) { xv[c] = x[(size_t)c * HW + p]; tv[c] = t[(size_t)c * HW + p]; mx = fmaxf(mx, xv[c]); }
        float e[NC], den = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) { e[c] = expf(xv[c] - mx); den += e[c]; }
        const float inv = 1.f / den;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float pr = e[c] * inv;
            spt[c] = fmaf(pr, tv[c], spt[c]);
            spp[c] = fmaf(pr, pr, spp[c]);
            stt[c] = fmaf(tv[c], tv[c], stt[c]);
            // BCEWithLogits: max(x,0) - x*t + log(1 + exp(-|x|))
            bce += fmaxf(xv[c], 0.f) - xv[c] * tv[c] + log1pf(expf(-fabsf(xv[c])));
        }
    }
    float* out = part + ((size_t)n * chunks + chunk) * (3 * NC + 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float a = block_sum(spt[c], scratch), b = block_sum(spp[c], scratch), d = block_sum(stt[c], scratch);
        if (threadIdx.x == 0) { out[3 * c] = a; out[3 * c + 1] = b; out[3 * c + 2] = d; }
    }
    const float bs = block_sum(bce, scratch);
    if (threadIdx.x == 0) out[3 * NC] = bs;
}

// numden [F][nc][2] = (sum p*t + smooth, sum p^2 + sum t^2 + smooth); loss[0] = dice + bce, loss[1] = dice, loss[2] = bce
__global__ void __launch_bounds__(SL_THREADS)
seg_loss_finalize_kernel(const float* __restrict__ part, float* __restrict__ numden, float* __restrict__ loss,
                         int F, int nc, int chunks, long long nel, float smooth) {
    __shared__ float scratch[32];
    float dice = 0.f, bce = 0.f;
    const int stride = 3 * nc + 1;
    for (int i = threadIdx.x; i < F * nc; i += SL_THREADS) {
        const int n = i / nc, c = i - n * nc;
        float a = 0.f, b = 0.f, d = 0.f;
        for (int k = 0; k < chunks; ++k) {
            const float* q = part + ((size_t)n * chunks + k) * stride + 3 * c;
            a += q[0]; b += q[1]; d += q[2];
        }
        const float num = a + smooth, den = b + d + smooth;
        numden[2 * i] = num;
        numden[2 * i + 1] = den;
        dice += 1.f - num / den;                      // BinaryDiceLoss (utils/losses.py:49-58), reduction over frames below
    }
    for (int i = threadIdx.x; i < F * chunks; i += SL_THREADS) bce += part[(size_t)i * stride + 3 * nc];
    dice = block_sum(dice, scratch);
    bce = block_sum(bce, scratch);
    if (threadIdx.x == 0) {
        const float dl = dice / (float)F / (float)nc;   // mean over frames per class, summed over classes / nc (:86-95)
        const float bl = bce / (float)nel;
        loss[0] = dl + bl;
        loss[1] = dl;
        loss[2] = bl;
    }
}

// dlogits = gout * ( d dice / dx + d bce / dx )
template <int NC>
__global__ void __launch_bounds__(SL_THREADS)
seg_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, const float* __restrict__ numden,
                    const float* __restrict__ gout, float* __restrict__ dlogits, int F, int HW, long long nel) {
    const int n = blockIdx.y;
    const float g = gout[0];
    const float wd = g / ((float)F * (float)NC), wb = g / (float)nel;
    float num[NC], rden[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { num[c] = numden[2 * (n * NC + c)]; rden[c] = 1.f / numden[2 * (n * NC + c) + 1]; }
    const float* x = logits + (size_t)n * NC * HW;
    const float* t = target + (size_t)n * NC * HW;
    float* dx = dlogits + (size_t)n * NC * HW;
    for (int p = blockIdx.x * SL_THREADS + threadIdx.x; p < HW; p += gridDim.x * SL_THREADS) {
        float xv[NC], tv[NC], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < NC; ++c) { xv[c] = x[(size_t)c * HW + p]; tv[c] = t[(size_t)c * HW + p]; mx = fmaxf(mx, xv[c]); }
        float pr[NC], den = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) { pr[c] = expf(xv[c] - mx); den += pr[c]; }
        const float inv = 1.f / den;
        float gp[NC], dot = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            pr[c] *= inv;
            // d(1 - num/den)/dp = -t/den + num * 2p / den^2
            gp[c] = wd * (-tv[c] * rden[c] + num[c] * 2.f * pr[c] * rden[c] * rden[c]);
            dot = fmaf(pr[c], gp[c], dot);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float sg = 1.f / (1.f + expf(-xv[c]));
            dx[(size_t)c * HW + p] = pr[c] * (gp[c] - dot) + wb * (sg - tv[c]);
        }
    }
}

// boxes [planes][4] = (xmin, ymin, xmax, ymax) of the "on" pixels of every plane; an empty plane -> (0, 0, W, H)
// (graph_matching.py:728-732).  MODE 0: on = value != 0; MODE 1: on = value > 0 (sigmoid(value) > 0.5).
template <typename T, int MODE>
__global__ void __launch_bounds__(SL_THREADS)
mask_boxes_kernel(const T* __restrict__ maps, float* __restrict__ boxes, int H, int W) {
    __shared__ float scratch[32];
    const T* m = maps + (size_t)blockIdx.x * H * W;
    int xmin = W, xmax = -1, ymin = H, ymax = -1;
    for (int p = threadIdx.x; p < H * W; p += SL_THREADS) {
        const T v = m[p];
        const bool on = MODE == 0 ? (v != (T)0) : (v > (T)0);
        if (on) {
            const int y = p / W, xx = p - y * W;
            xmin = min(xmin, xx); xmax = max(xmax, xx);
            ymin = min(ymin, y); ymax = max(ymax, y);
        }
    }
    const float fxmin = block_min((float)xmin, scratch), fymin = block_min((float)ymin, scratch);
    const float fxmax = block_max((float)xmax, scratch), fymax = block_max((float)ymax, scratch);
    if (threadIdx.x == 0) {
        float* b = boxes + (size_t)blockIdx.x * 4;
        if (fxmax < 0.f) { b[0] = 0.f; b[1] = 0.f; b[2] = (float)W; b[3] = (float)H; }
        else { b[0] = fxmin; b[1] = fymin; b[2] = fxmax; b[3] = fymax; }
    }
}

int seg_chunks(int F, int HW) {
    const int want = cdiv(HW, SL_THREADS * 4);                 // >= 4 pixels per thread
    const int cap = max(1, (sm_count() * 4) / max(F, 1));       // ~4 CTAs per SM over the whole batch
    return max(1, min(want, max(cap, 1)));
}

}  // namespace

extern "C" size_t ge_seg_loss_workspace_bytes(int F, int nc, int HW) {
    if (F <= 0 || nc <= 0 || nc > SL_MAXC || HW <= 0) return 0;
    return (size_t)F * seg_chunks(F, HW) * (3 * nc + 1) * sizeof(float);
}

// logits, target fp32 NCHW [F,nc,H*W]; numden fp32 [F,nc,2] (saved for the backward); loss fp32 [3] = (dice + bce, dice, bce)
extern "C" int ge_seg_loss_fwd(const float* logits, const float* target, float* numden, float* loss, void* workspace,
                               size_t workspace_bytes, int F, int nc, int HW, float smooth, ge_stream_t stream) {
    GE_REQUIRE(logits && target && numden && loss && workspace, GE_ERR_ARG, "ge_seg_loss_fwd: null pointer");
    GE_REQUIRE(F > 0 && HW > 0 && nc > 0, GE_ERR_ARG, "ge_seg_loss_fwd: bad dimension");
    GE_REQUIRE(nc <= SL_MAXC, GE_ERR_SHAPE, "ge_seg_loss_fwd: at most %d classes (got %d)", SL_MAXC, nc);
    GE_REQUIRE(workspace_bytes >= ge_seg_loss_workspace_bytes(F, nc, HW), GE_ERR_ARG, "ge_seg_loss_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = seg_chunks(F, HW);
    float* part = static_cast<float*>(workspace);
    dim3 grid(chunks, F);
    switch (nc) {
#define GE_SL_CASE(K) case K: seg_loss_partial_kernel<K><<<grid, SL_THREADS, 0, st>>>(logits, target, part, HW, chunks); break;
        GE_SL_CASE(1) GE_SL_CASE(2) GE_SL_CASE(3) GE_SL_CASE(4) GE_SL_CASE(5) GE_SL_CASE(6) GE_SL_CASE(7) GE_SL_CASE(8)
#undef GE_SL_CASE
    }
    GE_CHECK_LAUNCH("ge_seg_loss_fwd(partial)");
    seg_loss_finalize_kernel<<<1, SL_THREADS, 0, st>>>(part, numden, loss, F, nc, chunks, (long long)F * nc * HW, smooth);
    GE_CHECK_LAUNCH("ge_seg_loss_fwd(finalize)");
    return GE_OK;
}

// gout: device scalar = d(total)/d(loss[0]); dlogits fp32 NCHW [F,nc,H*W] (overwritten)
extern "C" int ge_seg_loss_bwd(const float* logits, const float* target, const float* numden, const float* gout,
                               float* dlogits, int F, int nc, int HW, ge_stream_t stream) {
    GE_REQUIRE(logits && target && numden && gout && dlogits, GE_ERR_ARG, "ge_seg_loss_bwd: null pointer");
    GE_REQUIRE(F > 0 && HW > 0 && nc > 0, GE_ERR_ARG, "ge_seg_loss_bwd: bad dimension");
    GE_REQUIRE(nc <= SL_MAXC, GE_ERR_SHAPE, "ge_seg_loss_bwd: at most %d classes (got %d)", SL_MAXC, nc);
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = max(1, min(cdiv(HW, SL_THREADS * 2), max(1, (sm_count() * 8) / F)));
    dim3 grid(gx, F);
    const long long nel = (long long)F * nc * HW;
    switch (nc) {
#define GE_SL_CASE(K) case K: seg_loss_bwd_kernel<K><<<grid, SL_THREADS, 0, st>>>(logits, target, numden, gout, dlogits, F, HW, nel); break;
        GE_SL_CASE(1) GE_SL_CASE(2) GE_SL_CASE(3) GE_SL_CASE(4) GE_SL_CASE(5) GE_SL_CASE(6) GE_SL_CASE(7) GE_SL_CASE(8)
#undef GE_SL_CASE
    }
    GE_CHECK_LAUNCH("ge_seg_loss_bwd");
    return GE_OK;
}

// maps [planes, H, W] (dtype: GE_DTYPE_F32, or GE_DTYPE_I64 = 2 for int64 score maps); boxes fp32 [planes, 4].
// mode 0: on = value != 0 (masks_to_boxes on masks / thresholded score maps); mode 1: on = value > 0 (raw logits,
// i.e. the box of sigmoid(x) > 0.5 without materialising the score map).
extern "C" int ge_mask_boxes(const void* maps, float* boxes, int dtype, int planes, int H, int W, int mode, ge_stream_t stream) {
    GE_REQUIRE(maps && boxes, GE_ERR_ARG, "ge_mask_boxes: null pointer");
    GE_REQUIRE(planes > 0 && H > 0 && W > 0 && (mode == 0 || mode == 1), GE_ERR_ARG, "ge_mask_boxes: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GE_DTYPE_F32) {
        if (mode == 0) mask_boxes_kernel<float, 0><<<planes, SL_THREADS, 0, st>>>((const float*)maps, boxes, H, W);
        else mask_boxes_kernel<float, 1><<<planes, SL_THREADS, 0, st>>>((const float*)maps, boxes, H, W);
    } else if (dtype == 2) {
        if (mode == 0) mask_boxes_kernel<long long, 0><<<planes, SL_THREADS, 0, st>>>((const long long*)maps, boxes, H, W);
        else mask_boxes_kernel<long long, 1><<<planes, SL_THREADS, 0, st>>>((const long long*)maps, boxes, H, W);
    } else {
        ge_set_error("ge_mask_boxes: unsupported dtype %d (fp32 or int64)", dtype);
        return GE_ERR_DTYPE;
    }
    GE_CHECK_LAUNCH("ge_mask_boxes");
    return GE_OK;
}
