// K4b -- the 'o2o' matching loss on the Sinkhorn-normalised affinity (GModule._forward_aff,
// /root/reference/models/graph_matching.py:572-590 with BCEFocalLoss :23-45), forward and backward as one launch each.
//
//   same_ij  = (label1_i == label2_j)                                   (one_hot @ one_hot^T == 1)
//   idx_i    = argmax_j P_ij * same_ij ,  tp_i = P[i, idx_i]            (row-wise best same-class entry)
//   tp_loss  = mean_i( -alpha (1 - tp_i)^gamma log tp_i ) / N1
//   fp_loss  = mean_{diff}( -(1 - alpha) P^gamma log(1 - P) ) / sum_{diff} P     (the sum is detached)
//   loss     = tp_loss + fp_loss
// The torch version is ~25 element-wise / reduction launches forward and ~40 backward on [N1,N2] matrices of a few
// hundred rows: pure launch latency on the graph module's host-driven stream.  Here: one CTA walks the matrix once
// (a warp per row: argmax with first-index ties, the three different-class sums), and the backward is one
// element-wise pass.  fp32 throughout, deterministic (fixed reduction order).
#include "common.cuh"
#include <algorithm>
#include "../../include/graphecho_b200.h"

namespace {

constexpr int ML_THREADS = 1024;

__device__ __forceinline__ float powg(float x, float g) { return g == 2.f ? x * x : powf(x, g); }

// stats: [0] sum_{diff} P, [1] #diff, [2] sum_{diff} fp_elem, [3] sum_i tp_elem
__global__ void __launch_bounds__(ML_THREADS)
matching_loss_fwd_kernel(const float* __restrict__ P, const float* __restrict__ lab1, const float* __restrict__ lab2,
                         float* __restrict__ loss, int* __restrict__ idx, float* __restrict__ stats,
                         int N1, int N2, float alpha, float gamma) {
    __shared__ float scratch[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s_p = 0.f, s_n = 0.f, s_fp = 0.f, s_tp = 0.f;
    for (int i = warp; i < N1; i += ML_THREADS / 32) {
        const float li = lab1[i];
        const float* row = P + (size_t)i * N2;
        float best = 0.f;            // (P * same).max(-1): products are >= 0, so an all-different row gives (0, index 0)
        int bj = N2;
        for (int j = lane; j < N2; j += 32) {
            const float p = row[j];
            if (lab2[j] == li) {
                if (p > best) { best = p; bj = j; }
            } else {
                s_p += p;
                s_n += 1.f;
                s_fp += -(1.f - alpha) * powg(p, gamma) * logf(1.f - p);
            }
        }
        // first index among equal maxima; rows without a positive same-class entry take index 0
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(ge::kFull, best, o);
            const int oj = __shfl_xor_sync(ge::kFull, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (bj >= N2) bj = 0;
        if (lane == 0) {
            idx[i] = bj;
            const float tp = row[bj];
            s_tp += -alpha * powg(1.f - tp, gamma) * logf(tp);
        }
    }
    s_p = ge::block_sum(s_p, scratch);
    s_n = ge::block_sum(s_n, scratch);
    s_fp = ge::block_sum(s_fp, scratch);
    s_tp = ge::block_sum(s_tp, scratch);
    if (threadIdx.x == 0) {
        stats[0] = s_p; stats[1] = s_n; stats[2] = s_fp; stats[3] = s_tp;
        loss[0] = s_tp / (float)N1 / (float)N1 + s_fp / s_n / s_p;
    }
}

__global__ void __launch_bounds__(256)
matching_loss_bwd_kernel(const float* __restrict__ P, const float* __restrict__ lab1, const float* __restrict__ lab2,
                         const int* __restrict__ idx, const float* __restrict__ stats, const float* __restrict__ gout,
                         float* __restrict__ dP, int N1, int N2, float alpha, float gamma) {
    const float g = gout[0];
    const float c_fp = g / stats[1] / stats[0];
    const float c_tp = g / (float)N1 / (float)N1;
    const long long total = (long long)N1 * N2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / N2), j = (int)(e - (long long)i * N2);
        const float p = P[e];
        float d = 0.f;
        if (lab2[j] != lab1[i]) {
            const float q = 1.f - p;
            // d/dp [ -(1-a) p^g log(1-p) ]
            const float dpow = gamma == 2.f ? 2.f * p : gamma * powf(p, gamma - 1.f);
            d = c_fp * -(1.f - alpha) * (dpow * logf(q) - powg(p, gamma) / q);
        }
        if (j == idx[i]) {
            const float q = 1.f - p;
            // d/dp [ -a (1-p)^g log p ]
            const float dpow = gamma == 2.f ? 2.f * q : gamma * powf(q, gamma - 1.f);
            d += c_tp * -alpha * (-dpow * logf(p) + powg(q, gamma) / p);
        }
        dP[e] = d;
    }
}

}  // namespace

// P [N1,N2] fp32 in (0,1); lab1 [N1], lab2 [N2] fp32 class labels (exact small integers); loss [1]; idx int32 [N1]
// and stats fp32 [4] are saved for the backward.
extern "C" int ge_matching_loss_fwd(const float* P, const float* lab1, const float* lab2, float* loss, int* idx,
                                    float* stats, int N1, int N2, float alpha, float gamma, ge_stream_t stream) {
    GE_REQUIRE(P && lab1 && lab2 && loss && idx && stats, GE_ERR_ARG, "ge_matching_loss_fwd: null pointer");
    GE_REQUIRE(N1 > 0 && N2 > 0, GE_ERR_ARG, "ge_matching_loss_fwd: non-positive dimension");
    matching_loss_fwd_kernel<<<1, ML_THREADS, 0, (cudaStream_t)stream>>>(P, lab1, lab2, loss, idx, stats, N1, N2, alpha, gamma);
    GE_CHECK_LAUNCH("ge_matching_loss_fwd");
    return GE_OK;
}

// gout [1] = dLoss/dloss on the device; dP [N1,N2] receives dLoss/dP.
extern "C" int ge_matching_loss_bwd(const float* P, const float* lab1, const float* lab2, const int* idx,
                                    const float* stats, const float* gout, float* dP, int N1, int N2,
                                    float alpha, float gamma, ge_stream_t stream) {
    GE_REQUIRE(P && lab1 && lab2 && idx && stats && gout && dP, GE_ERR_ARG, "ge_matching_loss_bwd: null pointer");
    GE_REQUIRE(N1 > 0 && N2 > 0, GE_ERR_ARG, "ge_matching_loss_bwd: non-positive dimension");
    const long long total = (long long)N1 * N2;
    const int grid = (int)std::min<long long>(ge::cdivll(total, 256), (long long)ge::sm_count() * 4);
    matching_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, lab1, lab2, idx, stats, gout, dP, N1, N2, alpha, gamma);
    GE_CHECK_LAUNCH("ge_matching_loss_bwd");
    return GE_OK;
}
