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

// ---------------------------------------------------------------------------------------------------------------
// f1 -- node sampler, second half: positive / negative picks and the row gather as ONE launch (all levels, both
// domains), plus the scatter that is its backward.
//
// Reference: PrototypeComputation.__call__ (models/graph_matching.py:978-1013) per level: the positives are the
// locations with label > 0 in index order, every `step`-th of them when there are more than num_nodes_per_class
// (step = n_pos // 100); the negatives are the locations with label == 0, all of them when there are fewer negatives
// than positives, else the ones whose rank is floor(linspace(0, n_neg - 2, num_pos // 8)) (numpy, float64); the node
// array is [negatives of level 0..L-1 | positives of level 0..L-1].  The torch version was ~14 small launches per
// level and domain (nonzero, slices, index arithmetic, index_select, a host->device copy of the pick list) and as many
// in the backward; here one CTA per (domain, level) ranks the labels with a block scan, decides membership in closed
// form, and gathers the rows of its level.  Index work: bit-exact (the linspace is evaluated in fp64 as numpy does).
namespace {

constexpr int SG_THREADS = 1024;
constexpr int SG_MAX_ENTRIES = 10;

struct GatherEntry {
    const void* feat;              // [images, h, w, C] NHWC
    const long long* labels;       // [n]
    float* nodes;                  // [n_nodes_domain, C]
    long long* node_labels;        // [n_nodes_domain]
    long long* src_row;            // [n_nodes_domain] row of feat each node came from (for the backward)
    long long n, shift;
    int n_neg_all, step, n_pos_pick, n_neg_pick, neg_all, out_pos, out_neg;
};
struct GatherParams {
    GatherEntry e[SG_MAX_ENTRIES];
};

// rank -> index of the pick with that rank in floor(linspace(0, n_neg - 2, m)), or -1
__device__ __forceinline__ int linspace_pick_index(int r, int m, int n_neg_all) {
    if (m <= 0) return -1;
    if (m == 1) return r == 0 ? 0 : -1;
    const double stop = (double)(n_neg_all - 2);
    const double step = stop / (double)(m - 1);
    int i = step > 0.0 ? (int)floor((double)r / step) : 0;
    for (int c = i - 1; c <= i + 1; ++c) {
        if (c < 0 || c >= m) continue;
        const double y = (c == m - 1) ? stop : __dmul_rn((double)c, step);
        if ((long long)floor(y) == (long long)r) return c;
    }
    return -1;
}

template <typename T>
__global__ void __launch_bounds__(SG_THREADS)
sampler_gather_kernel(GatherParams prm, int C) {
    const GatherEntry g = prm.e[blockIdx.x];
    __shared__ int warp_tot[32];
    __shared__ int run_pos, run_neg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { run_pos = 0; run_neg = 0; }
    __syncthreads();
    for (long long base = 0; base < g.n; base += SG_THREADS * 4) {
        const long long e0 = base + (long long)tid * 4;
        int lab[4];
        int cnt = 0;                                   // positives in the low half, negatives in the high half
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            lab[u] = (e0 + u < g.n) ? (int)g.labels[e0 + u] : -1;
            cnt += lab[u] > 0 ? 1 : (lab[u] == 0 ? (1 << 16) : 0);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(ge::kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(ge::kFull, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;                   // exclusive warp offsets
            if (lane == 31) { warp_tot[31] = wi - w; }
        }
        __syncthreads();
        const int excl = incl - cnt + warp_tot[warp];
        int rp = run_pos + (excl & 0xffff), rn = run_neg + (excl >> 16);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (lab[u] > 0) {
                const int r = rp++;
                int slot = -1;
                if (g.step > 1) { if (r % g.step == 0) slot = r / g.step; } else slot = r;
                if (slot >= 0 && slot < g.n_pos_pick) {
                    g.src_row[g.out_pos + slot] = e0 + u + g.shift;
                    g.node_labels[g.out_pos + slot] = lab[u];
                }
            } else if (lab[u] == 0) {
                const int r = rn++;
                const int slot = g.neg_all ? r : linspace_pick_index(r, g.n_neg_pick, g.n_neg_all);
                if (slot >= 0 && slot < g.n_neg_pick) {
                    g.src_row[g.out_neg + slot] = e0 + u + g.shift;
                    g.node_labels[g.out_neg + slot] = 0;
                }
            }
        }
        __syncthreads();                               // everyone has read run_* and warp_tot
        if (tid == SG_THREADS - 1) {
            run_pos += (excl & 0xffff) + (cnt & 0xffff);
            run_neg += (excl >> 16) + (cnt >> 16);
        }
        __syncthreads();
    }
    // row gather of this level's picks (written by this CTA: block-level visibility is enough)
    __threadfence_block();
    __syncthreads();
    const T* feat = static_cast<const T*>(g.feat);
    const int c4n = C >> 2;
    const int total = g.n_neg_pick + g.n_pos_pick;
    for (int w = warp; w < total; w += SG_THREADS / 32) {
        const int slot = w < g.n_neg_pick ? g.out_neg + w : g.out_pos + (w - g.n_neg_pick);
        const long long row = g.src_row[slot];
        for (int c4 = lane; c4 < c4n; c4 += 32) {
            ge::Vec4<T> v;
            v.load(feat + row * C + 4 * c4);
            float f[4];
            v.get(f);
            *reinterpret_cast<float4*>(g.nodes + (size_t)slot * C + 4 * c4) = make_float4(f[0], f[1], f[2], f[3]);
        }
    }
}

struct ScatterEntry {
    void* dfeat;                   // zero-filled [images, h, w, C]
    const float* dnodes;           // [n_nodes_domain, C]
    const long long* src_row;
    int n_pos_pick, n_neg_pick, out_pos, out_neg;
};
struct ScatterParams {
    ScatterEntry e[SG_MAX_ENTRIES];
};

template <typename T>
__global__ void __launch_bounds__(64)
sampler_scatter_kernel(ScatterParams prm, int C) {
    const ScatterEntry g = prm.e[blockIdx.y];
    const int total = g.n_neg_pick + g.n_pos_pick;
    T* dfeat = static_cast<T*>(g.dfeat);
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int slot = w < g.n_neg_pick ? g.out_neg + w : g.out_pos + (w - g.n_neg_pick);
        const long long row = g.src_row[slot];
        for (int c4 = threadIdx.x; c4 < (C >> 2); c4 += 64) {
            const float4 d = *reinterpret_cast<const float4*>(g.dnodes + (size_t)slot * C + 4 * c4);
            const float f[4] = {d.x, d.y, d.z, d.w};
            ge::Vec4<T> v;
            v.set(f);
            v.store(dfeat + row * C + 4 * c4);
        }
    }
}

}  // namespace

// One entry per (domain, level), n_entries <= 10.  Per entry: feats NHWC feature map of the level, labels int64 [n_loc]
// (ge_sampler_labels), shift = batch_offset * h * w (row of the first labelled image in feats), n_neg_all = number of
// label == 0 locations, step = positive stride (>= 1), n_pos_pick / n_neg_pick = nodes kept, neg_all = 1 when every
// negative is kept (else the floor(linspace) ranks), out_pos / out_neg = first slot of the level's positives / negatives
// in the domain's arrays nodes fp32 [n_nodes, C], node_labels int64 [n_nodes], src_row int64 [n_nodes] (all three per
// entry: entries of one domain pass the same bases).  C % 4 == 0.
extern "C" int ge_sampler_gather(const void* const* feats, const long long* const* labels, float* const* nodes,
                                 long long* const* node_labels, long long* const* src_row, const long long* n_loc,
                                 const long long* shift, const int* n_neg_all, const int* step, const int* n_pos_pick,
                                 const int* n_neg_pick, const int* neg_all, const int* out_pos, const int* out_neg,
                                 int n_entries, int C, int dtype, ge_stream_t stream) {
    GE_REQUIRE(feats && labels && nodes && node_labels && src_row && n_loc && shift && n_neg_all && step && n_pos_pick &&
               n_neg_pick && neg_all && out_pos && out_neg, GE_ERR_ARG, "ge_sampler_gather: null pointer");
    GE_REQUIRE(n_entries >= 1 && n_entries <= SG_MAX_ENTRIES, GE_ERR_SHAPE, "ge_sampler_gather: n_entries=%d (1..%d)", n_entries, SG_MAX_ENTRIES);
    GE_REQUIRE(C > 0 && C % 4 == 0, GE_ERR_SHAPE, "ge_sampler_gather: C=%d must be a multiple of 4", C);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_sampler_gather: unsupported dtype %d", dtype);
    GatherParams prm;
    for (int i = 0; i < n_entries; ++i) {
        GE_REQUIRE(feats[i] && labels[i] && n_loc[i] > 0 && step[i] >= 1 && n_pos_pick[i] >= 0 && n_neg_pick[i] >= 0, GE_ERR_ARG,
                   "ge_sampler_gather: bad entry %d", i);
        GE_REQUIRE((n_pos_pick[i] + n_neg_pick[i] == 0) || (nodes[i] && node_labels[i] && src_row[i]), GE_ERR_ARG,
                   "ge_sampler_gather: entry %d has picks but no output arrays", i);
        GE_REQUIRE(n_loc[i] < (1LL << 31), GE_ERR_SHAPE, "ge_sampler_gather: level too large");
        GatherEntry& g = prm.e[i];
        g.feat = feats[i]; g.labels = labels[i]; g.nodes = nodes[i]; g.node_labels = node_labels[i]; g.src_row = src_row[i];
        g.n = n_loc[i]; g.shift = shift[i]; g.n_neg_all = n_neg_all[i]; g.step = step[i]; g.n_pos_pick = n_pos_pick[i];
        g.n_neg_pick = n_neg_pick[i]; g.neg_all = neg_all[i]; g.out_pos = out_pos[i]; g.out_neg = out_neg[i];
    }
    if (dtype == GE_DTYPE_F32)
        sampler_gather_kernel<float><<<n_entries, SG_THREADS, 0, (cudaStream_t)stream>>>(prm, C);
    else
        sampler_gather_kernel<__nv_bfloat16><<<n_entries, SG_THREADS, 0, (cudaStream_t)stream>>>(prm, C);
    GE_CHECK_LAUNCH("ge_sampler_gather");
    return GE_OK;
}

// Backward of ge_sampler_gather: dfeats[i] (ZERO-FILLED by the caller, NHWC like feats[i]) receives the rows of dnodes[i]
// at src_row[i] for the entry's slots.  The picked locations are distinct, so this is a plain scatter.
extern "C" int ge_sampler_scatter(void* const* dfeats, const float* const* dnodes, const long long* const* src_row,
                                  const int* n_pos_pick, const int* n_neg_pick, const int* out_pos, const int* out_neg,
                                  int n_entries, int C, int dtype, ge_stream_t stream) {
    GE_REQUIRE(dfeats && dnodes && src_row && n_pos_pick && n_neg_pick && out_pos && out_neg, GE_ERR_ARG, "ge_sampler_scatter: null pointer");
    GE_REQUIRE(n_entries >= 1 && n_entries <= SG_MAX_ENTRIES, GE_ERR_SHAPE, "ge_sampler_scatter: n_entries=%d (1..%d)", n_entries, SG_MAX_ENTRIES);
    GE_REQUIRE(C > 0 && C % 4 == 0, GE_ERR_SHAPE, "ge_sampler_scatter: C=%d must be a multiple of 4", C);
    GE_REQUIRE(dtype == GE_DTYPE_F32 || dtype == GE_DTYPE_BF16, GE_ERR_DTYPE, "ge_sampler_scatter: unsupported dtype %d", dtype);
    ScatterParams prm;
    int most = 0;
    for (int i = 0; i < n_entries; ++i) {
        const int tot = n_pos_pick[i] + n_neg_pick[i];
        GE_REQUIRE(tot == 0 || (dfeats[i] && dnodes[i] && src_row[i]), GE_ERR_ARG, "ge_sampler_scatter: bad entry %d", i);
        ScatterEntry& g = prm.e[i];
        g.dfeat = dfeats[i]; g.dnodes = dnodes[i]; g.src_row = src_row[i];
        g.n_pos_pick = n_pos_pick[i]; g.n_neg_pick = n_neg_pick[i]; g.out_pos = out_pos[i]; g.out_neg = out_neg[i];
        most = std::max(most, tot);
    }
    if (most == 0) return GE_OK;
    dim3 grid((unsigned)most, (unsigned)n_entries);
    if (dtype == GE_DTYPE_F32)
        sampler_scatter_kernel<float><<<grid, 64, 0, (cudaStream_t)stream>>>(prm, C);
    else
        sampler_scatter_kernel<__nv_bfloat16><<<grid, 64, 0, (cudaStream_t)stream>>>(prm, C);
    GE_CHECK_LAUNCH("ge_sampler_scatter");
    return GE_OK;
}
