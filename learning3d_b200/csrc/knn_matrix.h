// Internal (C++ linkage) helpers shared between translation units of libl3d_b200.
#pragma once
#include <cuda_runtime.h>
namespace l3d {
// top-k (largest key first, lower index on ties) of every row of a key matrix keys[rows][N] -> idx[rows][k]
int knn_select_from_matrix(const float* keys, long rows, int N, int k, long long* idx, cudaStream_t stream);
// knn() on xyz clouds with uint16 indices (knn.cu): device half of the host-buffer entry point
int knn_expansion_u16(const float* x_dev, int B, int N, int k, unsigned short* idx_dev, cudaStream_t stream);
// value of the l3d_debug_force_slow_path() testing hook
int knn_force_slow_flag();
}  // namespace l3d
