// Internal (C++ linkage) launchers of the register-resident Sinkhorn kernels (sinkhorn_rpm_reg.cu);
// the C-ABI entry points that dispatch to them are ge_sinkhorn_rpm_fwd/bwd in sinkhorn_rpm.cu.
#pragma once
#include <cuda_runtime.h>

namespace ge { namespace rpmreg {

constexpr int kWarps = 16;       // warps per CTA
constexpr int kMaxCluster = 8;   // portable cluster size

// true when an [N1, N2] problem with n_iters saved iterations runs on the register path
// (forward and backward take the same decision: they share the history format)
bool fits(int N1, int N2, int n_iters);

// stats[prob*4 + 2] is written 0 (done here) or 1 (max z too large for the exponent domain: the caller
// must run the log-domain kernel for that problem).  hist_u/hist_v receive u_t, v_t.
int fwd(const float* M, float* P, float* hist_u, float* hist_v, float* stats, int batch, int N1, int N2,
        int n_iters, int apply_instnorm, int rows_per_thread, cudaStream_t st);

// problems whose stats[prob*4 + 2] != 0 are skipped
int bwd(const float* M, const float* G, const float* hist_u, const float* hist_v, const float* stats, float* dM,
        int batch, int N1, int N2, int n_iters, int apply_instnorm, cudaStream_t st);

} }  // namespace ge::rpmreg
