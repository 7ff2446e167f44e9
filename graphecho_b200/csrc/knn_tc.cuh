// Tensor-core (tcgen05 / TMEM / TMA) path of the k-NN graph build; see knn_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace ge {
// true when the tcgen05 path covers this problem (no relative_pos, C % 32 == 0, k*dilation <= 32, N, M >= 128)
bool knn_tc_applicable(int B, int C, int N, int M, int K, bool has_rel);
size_t knn_tc_workspace_bytes(int B, int C, int N, int M);
int knn_tc_run(const float* x, const float* y, long long* edge_index, void* workspace,
               int B, int C, int N, int M, int K, int dilation, cudaStream_t st);
}  // namespace ge
