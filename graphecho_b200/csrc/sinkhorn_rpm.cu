// K4 — instance-norm -> slack-padded log-domain Sinkhorn (n_iters row/col passes) -> exp,
// forward and reverse-mode backward, each as ONE launch with the matrix resident on-chip.
//
// Replaces GModule._forward_aff's  InstNorm_layer(M) ; sinkhorn_rpm(M, n_iters=20) ; exp()
// (/root/reference/models/graph_matching.py:574-575, 637-676).  The reference materialises
// the (N1+1)x(N2+1) padded matrix and rewrites it 2*n_iters times (~250 launches, autograd
// keeps 40 copies).  Here the padded matrix is never formed: with z = instnorm(M),
//   L_t[i][j] = z[i][j] - r_i - c_j,   slack column j=N2: L = -r_i,  slack row i=N1: L = -c_j
// and one reference iteration is exactly
//   r_i <- logsumexp_j( z_ij - c_j  U {0} )      (row pass, rows[:-1], all columns)
//   c_j <- logsumexp_i( z_ij - r_i  U {0} )      (col pass, cols[:-1], all rows)
// so only the two potential vectors change.  z is split by rows over the CTAs of a thread
// block cluster (shared memory); the row pass is CTA-local, the column pass exchanges
// per-CTA (max, sum) pairs through distributed shared memory: one cluster barrier per
// iteration, zero HBM traffic inside the loop.  Algorithmic bytes: 8*N1*N2 for the whole
// loop (+ 4*n_iters*(N1+N2) of saved potentials for the backward).
//
// Backward: exact adjoint of the n_iters unrolled iterations (not implicit differentiation,
// which differs at 20 unconverged iterations).  It replays the iterations in reverse from
// the saved potentials, recomputing the softmax weights exp(z - r_t - c_t) on the fly.
#include "common.cuh"
#include "../../include/graphecho_b200.h"
#include "sinkhorn_rpm_reg.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int RPM_THREADS = 512;
constexpr int RPM_WARPS = RPM_THREADS / 32;
constexpr float IN_EPS = 1e-5f;   // nn.InstanceNorm2d default eps (graph_matching.py:177)

struct RpmSmem {
    float* z;        // [R][ld]
    float* e;        // [R][ld]   (backward only)
    float* c;        // [N2p] column potentials / column adjoints (replicated in every CTA)
    float* r;        // [R]   row potentials / row adjoints (local rows)
    float* part;     // [2 parity][2][N2p] per-CTA column partials read by peers
    float* xchg;     // [2 parity][4] scalar exchange
    float* scratch;  // [32]
    float* hc;       // [N2p] saved c_t      (backward only)
    float* hcp;      // [N2p] saved c_{t-1}  (backward only)
    float* hr;       // [R]   saved r_t      (backward only)
};

__device__ __forceinline__ RpmSmem carve(float* base, int R, int ld, int N2p, bool bwd) {
    RpmSmem s;
    s.z = base; base += (size_t)R * ld;
    s.e = base; if (bwd) base += (size_t)R * ld;
    s.c = base; base += N2p;
    s.r = base; base += ((R + 3) & ~3);
    s.part = base; base += 4 * N2p;
    s.xchg = base; base += 8;
    s.scratch = base; base += 32;
    s.hc = base; if (bwd) base += N2p;
    s.hcp = base; if (bwd) base += N2p;
    s.hr = base;
    return s;
}

__host__ __device__ inline size_t rpm_smem_floats(int R, int ld, int N2p, bool bwd) {
    return (size_t)R * ld * (bwd ? 2 : 1) + N2p + ((R + 3) & ~3) + 4 * N2p + 8 + 32 +
           (bwd ? 2 * N2p + ((R + 3) & ~3) : 0);
}

// Cluster-wide sum of up to 4 scalars held by thread 0 of each CTA.  Deterministic
// (rank order), bit-identical in every CTA.  Uses parity double-buffering: a slot is
// rewritten only after another cluster barrier has separated it from its readers.
__device__ __forceinline__ void cluster_sum4(cg::cluster_group& cl, float* xchg, int& parity,
                                             float (&v)[4], int n) {
    float* mine = xchg + parity * 4;
    if (threadIdx.x == 0)
        for (int k = 0; k < n; ++k) mine[k] = v[k];
    cl.sync();
    const unsigned cs = cl.num_blocks();
    for (int k = 0; k < n; ++k) v[k] = 0.f;
    for (unsigned q = 0; q < cs; ++q) {
        const float* peer = cl.map_shared_rank(mine, q);
        for (int k = 0; k < n; ++k) v[k] += peer[k];
    }
    parity ^= 1;
}

__global__ void __launch_bounds__(RPM_THREADS, 1)
sinkhorn_rpm_fwd_kernel(const float* __restrict__ M, float* __restrict__ P,
                        float* __restrict__ hist_r, float* __restrict__ hist_c,
                        float* __restrict__ stats, int N1, int N2, int n_iters, int apply_instnorm, int gate) {
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks(), rank = cl.block_rank();
    const int prob = blockIdx.x / cs;
    // gate: run only the problems the register-resident kernel left to this one (stats[2] == 1)
    if (gate && stats[(size_t)prob * 4 + 2] == 0.f) return;
    const int R = ge::cdiv(N1, (int)cs);
    const int row0 = rank * R;
    const int rows = max(0, min(R, N1 - row0));
    const int ld = (N2 + 3) & ~3, N2p = ld;

    extern __shared__ __align__(16) float smem[];
    RpmSmem s = carve(smem, R, ld, N2p, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    M += (size_t)prob * N1 * N2;
    P += (size_t)prob * N1 * N2;
    hist_r += (size_t)prob * n_iters * N1;
    hist_c += (size_t)prob * n_iters * N2;
    stats += (size_t)prob * 4;

    // ---- load rows, instance-norm statistics (two-pass, cluster-wide) ----
    float lsum = 0.f;
    for (int e = tid; e < rows * N2; e += RPM_THREADS) {
        const int i = e / N2, j = e - i * N2;
        const float v = M[(size_t)(row0 + i) * N2 + j];
        s.z[i * ld + j] = v;
        lsum += v;
    }
    int parity = 0;
    float mean = 0.f, rstd = 1.f;
    if (apply_instnorm) {
        float v4[4];
        v4[0] = ge::block_sum(lsum, s.scratch);
        cluster_sum4(cl, s.xchg, parity, v4, 1);
        const float inv_n = 1.f / ((float)N1 * (float)N2);
        mean = v4[0] * inv_n;
        float lsq = 0.f;
        for (int e = tid; e < rows * N2; e += RPM_THREADS) {
            const int i = e / N2, j = e - i * N2;
            const float d = s.z[i * ld + j] - mean;
            lsq += d * d;
        }
        v4[0] = ge::block_sum(lsq, s.scratch);
        cluster_sum4(cl, s.xchg, parity, v4, 1);
        rstd = 1.f / sqrtf(v4[0] * inv_n + IN_EPS);
        for (int e = tid; e < rows * N2; e += RPM_THREADS) {
            const int i = e / N2, j = e - i * N2;
            s.z[i * ld + j] = (s.z[i * ld + j] - mean) * rstd;
        }
    }
    if (rank == 0 && tid == 0) { stats[0] = mean; stats[1] = rstd; stats[2] = 1.f; stats[3] = 0.f; }   // 1: log-domain history
    for (int j = tid; j < N2p; j += RPM_THREADS) s.c[j] = 0.f;
    for (int i = tid; i < R; i += RPM_THREADS) s.r[i] = 0.f;
    __syncthreads();

    // ---- n_iters x (row pass, column pass) ----
    for (int t = 0; t < n_iters; ++t) {
        // row pass: r_i = LSE_j(z_ij - c_j  U {0})   [graph_matching.py:661-664]
        for (int i = warp; i < rows; i += RPM_WARPS) {
            const float* zi = s.z + i * ld;
            float m = 0.f;  // slack-column entry
            for (int j = lane; j < N2; j += 32) m = fmaxf(m, zi[j] - s.c[j]);
            m = ge::warp_max(m);
            float acc = 0.f;
            for (int j = lane; j < N2; j += 32) acc += __expf(zi[j] - s.c[j] - m);
            acc = ge::warp_sum(acc) + __expf(-m);
            if (lane == 0) {
                const float rv = m + logf(acc);
                s.r[i] = rv;
                hist_r[(size_t)t * N1 + row0 + i] = rv;
            }
        }
        __syncthreads();
        // column pass partials over the local rows   [graph_matching.py:666-669]
        float* pm = s.part + (t & 1) * 2 * N2p;
        float* ps = pm + N2p;
        for (int j = tid; j < N2; j += RPM_THREADS) {
            float m = (rank == 0) ? 0.f : -INFINITY;  // slack-row entry lives in rank 0's partial
            for (int i = 0; i < rows; ++i) m = fmaxf(m, s.z[i * ld + j] - s.r[i]);
            float acc = (rank == 0) ? __expf(-m) : 0.f;
            for (int i = 0; i < rows; ++i) acc += __expf(s.z[i * ld + j] - s.r[i] - m);
            pm[j] = m;
            ps[j] = acc;
        }
        cl.sync();
        for (int j = tid; j < N2; j += RPM_THREADS) {
            float mq[16], sq[16];
            float mx = -INFINITY;
#pragma unroll
            for (unsigned q = 0; q < 16; ++q) {
                mq[q] = -INFINITY;
                sq[q] = 0.f;
                if (q < cs) {
                    const float* peer = cl.map_shared_rank(pm, q);
                    mq[q] = peer[j];
                    sq[q] = peer[N2p + j];
                }
                mx = fmaxf(mx, mq[q]);
            }
            float acc = 0.f;
#pragma unroll
            for (unsigned q = 0; q < 16; ++q)
                if (q < cs) acc += sq[q] * __expf(mq[q] - mx);
            const float cv = mx + logf(acc);
            s.c[j] = cv;
            if (rank == 0) hist_c[(size_t)t * N2 + j] = cv;
        }
        __syncthreads();
    }
    // ---- crop + exp  [graph_matching.py:575, 676] ----
    for (int e = tid; e < rows * N2; e += RPM_THREADS) {
        const int i = e / N2, j = e - i * N2;
        P[(size_t)(row0 + i) * N2 + j] = expf(s.z[i * ld + j] - s.r[i] - s.c[j]);
    }
    cl.sync();  // peers may still be reading this CTA's partials
}

__global__ void __launch_bounds__(RPM_THREADS, 1)
sinkhorn_rpm_bwd_kernel(const float* __restrict__ M, const float* __restrict__ G,
                        const float* __restrict__ hist_r, const float* __restrict__ hist_c,
                        const float* __restrict__ stats, float* __restrict__ dM,
                        int N1, int N2, int n_iters, int apply_instnorm, int gate) {
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks(), rank = cl.block_rank();
    const int prob = blockIdx.x / cs;
    if (gate && stats[(size_t)prob * 4 + 2] == 0.f) return;
    const int R = ge::cdiv(N1, (int)cs);
    const int row0 = rank * R;
    const int rows = max(0, min(R, N1 - row0));
    const int ld = (N2 + 3) & ~3, N2p = ld;

    extern __shared__ __align__(16) float smem[];
    RpmSmem s = carve(smem, R, ld, N2p, true);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    M += (size_t)prob * N1 * N2;
    G += (size_t)prob * N1 * N2;
    dM += (size_t)prob * N1 * N2;
    hist_r += (size_t)prob * n_iters * N1;
    hist_c += (size_t)prob * n_iters * N2;
    stats += (size_t)prob * 4;
    const float mean = stats[0], rstd = stats[1];

    // z (recomputed), E = G o P with P = exp(z - r_T - c_T); gr_i = -sum_j E_ij; partial gc_j
    const float* rT = hist_r + (size_t)(n_iters - 1) * N1;
    const float* cT = hist_c + (size_t)(n_iters - 1) * N2;
    for (int e = tid; e < rows * N2; e += RPM_THREADS) {
        const int i = e / N2, j = e - i * N2;
        const size_t gidx = (size_t)(row0 + i) * N2 + j;
        const float zv = apply_instnorm ? (M[gidx] - mean) * rstd : M[gidx];
        s.z[i * ld + j] = zv;
        float pv = 0.f, ev = 0.f;
        if (n_iters > 0) pv = expf(zv - rT[row0 + i] - cT[j]); else pv = expf(zv);
        ev = G[gidx] * pv;
        s.e[i * ld + j] = ev;
    }
    __syncthreads();
    int parity = 0;   // column-partial parity; one cluster barrier per use
    if (n_iters > 0) {
        for (int i = warp; i < rows; i += RPM_WARPS) {
            float acc = 0.f;
            for (int j = lane; j < N2; j += 32) acc += s.e[i * ld + j];
            acc = ge::warp_sum(acc);
            if (lane == 0) s.r[i] = -acc;     // adjoint of r_T so far
        }
        float* pp = s.part + parity * 2 * N2p;
        for (int j = tid; j < N2; j += RPM_THREADS) {
            float acc = 0.f;
            for (int i = 0; i < rows; ++i) acc += s.e[i * ld + j];
            pp[j] = acc;
        }
        cl.sync();
        for (int j = tid; j < N2; j += RPM_THREADS) {
            float acc = 0.f;
            for (unsigned q = 0; q < cs; ++q) acc += cl.map_shared_rank(pp, q)[j];
            s.c[j] = -acc;                    // adjoint of c_T
        }
        parity ^= 1;
        __syncthreads();
    }

    for (int t = n_iters - 1; t >= 0; --t) {
        for (int j = tid; j < N2; j += RPM_THREADS) {
            s.hc[j] = hist_c[(size_t)t * N2 + j];
            s.hcp[j] = (t > 0) ? hist_c[(size_t)(t - 1) * N2 + j] : 0.f;
        }
        for (int i = tid; i < rows; i += RPM_THREADS) s.hr[i] = hist_r[(size_t)t * N1 + row0 + i];
        __syncthreads();
        const float* rt = s.hr;
        const float* ct = s.hc;
        const float* cprev = s.hcp;
        // adjoint of the column pass c_t = LSE_i(z - r_t): weights W = exp(z - r_t - c_t)
        //   E += gc_j W ;  gr_i -= sum_j gc_j W
        for (int i = warp; i < rows; i += RPM_WARPS) {
            const float ri = rt[i];
            float acc = 0.f;
            for (int j = lane; j < N2; j += 32) {
                const float w = __expf(s.z[i * ld + j] - ri - ct[j]) * s.c[j];
                s.e[i * ld + j] += w;
                acc += w;
            }
            acc = ge::warp_sum(acc);
            if (lane == 0) s.r[i] -= acc;
        }
        __syncthreads();
        // adjoint of the row pass r_t = LSE_j(z - c_{t-1}): W = exp(z - c_{t-1} - r_t)
        //   E += gr_i W ;  gc'_j = -sum_i gr_i W   (cluster-wide column sum)
        float* pp = s.part + parity * 2 * N2p;
        for (int j = tid; j < N2; j += RPM_THREADS) {
            const float cp = cprev[j];
            float acc = 0.f;
            for (int i = 0; i < rows; ++i) {
                const float w = __expf(s.z[i * ld + j] - cp - rt[i]) * s.r[i];
                s.e[i * ld + j] += w;
                acc += w;
            }
            pp[j] = acc;
        }
        cl.sync();
        for (int j = tid; j < N2; j += RPM_THREADS) {
            float acc = 0.f;
            for (unsigned q = 0; q < cs; ++q) acc += cl.map_shared_rank(pp, q)[j];
            s.c[j] = -acc;
        }
        for (int i = tid; i < rows; i += RPM_THREADS) s.r[i] = 0.f;
        parity ^= 1;
        __syncthreads();
    }

    // instance-norm adjoint: dM = rstd * (dz - mean(dz) - z * mean(dz o z))
    if (apply_instnorm) {
        float a = 0.f, bsum = 0.f;
        for (int e = tid; e < rows * N2; e += RPM_THREADS) {
            const int i = e / N2, j = e - i * N2;
            const float dz = s.e[i * ld + j];
            a += dz;
            bsum += dz * s.z[i * ld + j];
        }
        float v4[4];
        v4[0] = ge::block_sum(a, s.scratch);
        v4[1] = ge::block_sum(bsum, s.scratch);
        int xp = 0;
        cluster_sum4(cl, s.xchg, xp, v4, 2);
        const float inv_n = 1.f / ((float)N1 * (float)N2);
        const float m1 = v4[0] * inv_n, m2 = v4[1] * inv_n;
        for (int e = tid; e < rows * N2; e += RPM_THREADS) {
            const int i = e / N2, j = e - i * N2;
            dM[(size_t)(row0 + i) * N2 + j] = rstd * (s.e[i * ld + j] - m1 - s.z[i * ld + j] * m2);
        }
    } else {
        for (int e = tid; e < rows * N2; e += RPM_THREADS) {
            const int i = e / N2, j = e - i * N2;
            dM[(size_t)(row0 + i) * N2 + j] = s.e[i * ld + j];
        }
    }
    cl.sync();
}

constexpr size_t kSmemCap = 220 * 1024;

int pick_cluster(int N1, int N2, bool bwd, int requested, int batch = 1) {
    const int ld = (N2 + 3) & ~3;
    const int cands[5] = {1, 2, 4, 8, 16};
    if (requested > 0) {
        for (int c : cands)
            if (c == requested) {
                const int R = ge::cdiv(N1, c);
                return rpm_smem_floats(R, ld, ld, bwd) * sizeof(float) <= kSmemCap ? c : -1;
            }
        return -1;
    }
    // Throughput heuristic for batches that fill the machine anyway: the smallest cluster that holds the matrix
    // (more problems in flight, fewer DSMEM exchanges per pass).
    if ((long long)batch * 8 >= 2LL * ge::sm_count()) {
        for (int c : cands) {
            const int R = ge::cdiv(N1, c);
            if (rpm_smem_floats(R, ld, ld, bwd) * sizeof(float) <= kSmemCap) return c;
        }
        return -1;
    }
    // Latency heuristic: aim for <= 48 rows per CTA, grow further only if capacity demands it.
    for (int c : cands) {
        const int R = ge::cdiv(N1, c);
        const bool fits = rpm_smem_floats(R, ld, ld, bwd) * sizeof(float) <= kSmemCap;
        if (fits && (R <= 48 || c == 8)) return c;
    }
    {
        const int R = ge::cdiv(N1, 16);
        if (rpm_smem_floats(R, ld, ld, bwd) * sizeof(float) <= kSmemCap) return 16;
    }
    return -1;
}

template <typename K, typename... Args>
int launch_cluster(K kernel, const char* name, int cs, int batch, size_t smem, cudaStream_t st, Args... args) {
    // attributes are set once per process and kernel (template instantiation), so that a stream
    // capture only ever sees the launch itself
    static size_t max_smem_set = 0;
    static bool nonportable_set = false;
    if (smem > max_smem_set) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
        max_smem_set = smem;
    }
    if (cs > 8 && !nonportable_set) {
        GE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), name);
        nonportable_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * batch));
    cfg.blockDim = dim3(RPM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GE_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...), name);
    ge_count_launches(1);
    return GE_OK;
}

}  // namespace

// 0 = automatic, 1 = log-domain (shared-memory) kernels only, 2 / 3 = register path with 4 / 8 rows per thread
static int g_rpm_path = 0;

extern "C" int ge_sinkhorn_rpm_set_path(int path) {
    GE_REQUIRE(path >= 0 && path <= 3, GE_ERR_ARG, "ge_sinkhorn_rpm_set_path: path must be 0..3 (got %d)", path);
    g_rpm_path = path;
    return GE_OK;
}

extern "C" int ge_sinkhorn_rpm_cluster_size(int N1, int N2, int backward) {
    if (N1 <= 0 || N2 <= 0) return -1;
    return pick_cluster(N1, N2, backward != 0, 0);
}

// cluster_size > 0 requests the log-domain kernel with that cluster size; 0 takes the register-resident
// exponent-domain kernel when the problem fits it (sinkhorn_rpm_reg.cu) and runs the log-domain kernel
// behind it, gated per problem on the flag the first kernel leaves in stats[2] (max z too large).
static bool use_reg_path(int N1, int N2, int n_iters, int cluster_size) {
    return cluster_size == 0 && g_rpm_path != 1 && ge::rpmreg::fits(N1, N2, n_iters);
}

extern "C" int ge_sinkhorn_rpm_fwd(const float* M, float* P, float* hist_r, float* hist_c, float* stats,
                                   int batch, int N1, int N2, int n_iters, int apply_instnorm,
                                   int cluster_size, ge_stream_t stream) {
    GE_REQUIRE(M && P && hist_r && hist_c && stats, GE_ERR_ARG, "ge_sinkhorn_rpm_fwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && n_iters >= 0, GE_ERR_ARG, "ge_sinkhorn_rpm_fwd: bad dimension");
    const int cs = pick_cluster(N1, N2, false, cluster_size, batch);
    GE_REQUIRE(cs > 0, GE_ERR_CAPACITY,
               "ge_sinkhorn_rpm_fwd: %dx%d does not fit the on-chip cluster layout (cluster_size=%d)", N1, N2, cluster_size);
    const int ld = (N2 + 3) & ~3;
    const size_t smem = rpm_smem_floats(ge::cdiv(N1, cs), ld, ld, false) * sizeof(float);
    const bool reg = use_reg_path(N1, N2, n_iters, cluster_size);
    if (reg) {
        // 8 rows per thread halve the cluster size (fewer exchanges per iteration): worth it once the 4-row CTAs
        // would fill the machine anyway (measured: 74 x 252^2 0.060 vs 0.107 ms; 8 x 252^2 0.054 vs 0.050 ms)
        const int cs4 = ge::cdiv(N1, 64), cs8 = ge::cdiv(N1, 128);
        const int tr = g_rpm_path == 2 ? 4 : g_rpm_path == 3 ? 8
                       : (cs8 < cs4 && (long long)batch * cs4 >= ge::sm_count() ? 8 : 4);
        const int rc = ge::rpmreg::fwd(M, P, hist_r, hist_c, stats, batch, N1, N2, n_iters, apply_instnorm, tr,
                                       (cudaStream_t)stream);
        if (rc != GE_OK) return rc;
    }
    return launch_cluster(sinkhorn_rpm_fwd_kernel, "ge_sinkhorn_rpm_fwd", cs, batch, smem, (cudaStream_t)stream,
                          M, P, hist_r, hist_c, stats, N1, N2, n_iters, apply_instnorm, reg ? 1 : 0);
}

extern "C" int ge_sinkhorn_rpm_bwd(const float* M, const float* G, const float* hist_r, const float* hist_c,
                                   const float* stats, float* dM, int batch, int N1, int N2, int n_iters,
                                   int apply_instnorm, int cluster_size, ge_stream_t stream) {
    GE_REQUIRE(M && G && hist_r && hist_c && stats && dM, GE_ERR_ARG, "ge_sinkhorn_rpm_bwd: null pointer");
    GE_REQUIRE(batch > 0 && N1 > 0 && N2 > 0 && n_iters >= 0, GE_ERR_ARG, "ge_sinkhorn_rpm_bwd: bad dimension");
    const int cs = pick_cluster(N1, N2, true, cluster_size, batch);
    GE_REQUIRE(cs > 0, GE_ERR_CAPACITY,
               "ge_sinkhorn_rpm_bwd: %dx%d does not fit the on-chip cluster layout (cluster_size=%d)", N1, N2, cluster_size);
    const int ld = (N2 + 3) & ~3;
    const size_t smem = rpm_smem_floats(ge::cdiv(N1, cs), ld, ld, true) * sizeof(float);
    const bool reg = use_reg_path(N1, N2, n_iters, cluster_size);
    if (reg) {
        const int rc = ge::rpmreg::bwd(M, G, hist_r, hist_c, stats, dM, batch, N1, N2, n_iters, apply_instnorm,
                                       (cudaStream_t)stream);
        if (rc != GE_OK) return rc;
    }
    return launch_cluster(sinkhorn_rpm_bwd_kernel, "ge_sinkhorn_rpm_bwd", cs, batch, smem, (cudaStream_t)stream,
                          M, G, hist_r, hist_c, stats, dM, N1, N2, n_iters, apply_instnorm, reg ? 1 : 0);
}
