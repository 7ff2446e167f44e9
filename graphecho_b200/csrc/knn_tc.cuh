// Tensor-core (tcgen05 / TMEM / TMA) path of the k-NN graph build; see knn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace ge {
// true when the tcgen05 path covers this problem (no relative_pos, C % 32 == 0, k*dilation <= 32, N, M >= 128)
bool knn_tc_applicable(int B, int C, int N, int M, int K, bool has_rel);
size_t knn_tc_workspace_bytes(int B, int C, int N, int M);
// layout 0: x, y [B,C,N] fp32;  1: [B,N,C] fp32;  2: [B,N,C] bf16 (node-major / channels-last)
int knn_tc_run(const void* x, const void* y, int layout, long long* edge_index, void* workspace,
               int B, int C, int N, int M, int K, int dilation, cudaStream_t st);
}  // namespace ge
