// get_graph_feature() gather (utils/model_common_utils.py:132-155) and its backward.
//   out[b, c,     n, j] = x[b, c, idx[b,n,j]]      (neighbour half, channels 0..C-1)
//   out[b, C + c, n, j] = x[b, c, n]               (centre half)
// HBM-bound: reads B*N*k int64 indices once, writes 4*2C*k bytes per row.  One thread owns
// 4 consecutive j of one (b,n) row and streams all 2C channel planes with 128-bit stores;
// the gathered reads hit x[b,c,:] (4 KB per channel at N=1024) in L1/L2.
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

namespace l3d {

template <int VEC>
__global__ void __launch_bounds__(256) graph_feature_kernel(const float* __restrict__ x,
                                                            const long long* __restrict__ idx,
                                                            int B, int C, int N, int k,
                                                            float* __restrict__ out) {
  // flattened over (b, n, j/VEC)
  const long per_row = k / VEC;
  const long total = (long)B * N * per_row;
  const size_t plane = (size_t)N * k;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const long row = t / per_row;            // b*N + n
    const int jq = (int)(t - row * per_row);  // which VEC-group of the k neighbours
    const int b = (int)(row / N);
    const int n = (int)(row - (long)b * N);
    int nb[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) nb[v] = (int)idx[row * k + jq * VEC + v];
    const float* xb = x + (size_t)b * C * N;
    float* ob = out + (size_t)b * 2 * C * plane + (size_t)n * k + (size_t)jq * VEC;
    for (int c = 0; c < C; ++c) {
      const float* xc = xb + (size_t)c * N;
      float g[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) g[v] = __ldg(xc + nb[v]);
      const float ctr = __ldg(xc + n);
      float* o1 = ob + (size_t)c * plane;
      float* o2 = ob + (size_t)(C + c) * plane;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(o1) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(o2) = make_float4(ctr, ctr, ctr, ctr);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { o1[v] = g[v]; o2[v] = ctr; }
      }
    }
  }
}

// grad_x[b,c,m] += sum_{(n,j): idx[b,n,j]==m} go[b,c,n,j]  +  (m==n) sum_j go[b,C+c,n,j]
__global__ void __launch_bounds__(256) graph_feature_grad_kernel(const float* __restrict__ go,
                                                                 const long long* __restrict__ idx,
                                                                 int B, int C, int N, int k,
                                                                 float* __restrict__ gx) {
  const long total = (long)B * N * k;
  const size_t plane = (size_t)N * k;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const long row = t / k;
    const int b = (int)(row / N);
    const int n = (int)(row - (long)b * N);
    const int m = (int)idx[t];
    const float* gb = go + (size_t)b * 2 * C * plane + (size_t)(t - (long)b * N * k);
    float* gxb = gx + (size_t)b * C * N;
    for (int c = 0; c < C; ++c) {
      atomicAdd(gxb + (size_t)c * N + m, gb[(size_t)c * plane]);
      atomicAdd(gxb + (size_t)c * N + n, gb[(size_t)(C + c) * plane]);
    }
  }
}

}  // namespace l3d

extern "C" int l3d_graph_feature(const float* x_dev, const int64_t* idx_dev, int B, int C, int N,
                                 int k, float* out_dev, void* stream) {
  if (!x_dev || !idx_dev || !out_dev || B < 0 || C < 1 || N < 1 || k < 1) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  const bool vec4 = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_dev) & 15u) == 0);
  const long total = (long)B * N * (vec4 ? k / 4 : k);
  long grid = (total + 255) / 256;
  if (grid > 148L * 16) grid = 148L * 16;
  if (vec4)
    l3d::graph_feature_kernel<4><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        x_dev, (const long long*)idx_dev, B, C, N, k, out_dev);
  else
    l3d::graph_feature_kernel<1><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        x_dev, (const long long*)idx_dev, B, C, N, k, out_dev);
  l3d::count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" int l3d_graph_feature_grad(const float* grad_out_dev, const int64_t* idx_dev, int B,
                                      int C, int N, int k, float* grad_x_dev, void* stream) {
  if (!grad_out_dev || !idx_dev || !grad_x_dev || B < 0 || C < 1 || N < 1 || k < 1)
    return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  const long total = (long)B * N * k;
  long grid = (total + 255) / 256;
  if (grid > 148L * 16) grid = 148L * 16;
  l3d::graph_feature_grad_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      grad_out_dev, (const long long*)idx_dev, B, C, N, k, grad_x_dev);
  l3d::count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}
