// RPMNet's matching tail (models/rpmnet.py:157-254), sm_100a.
//
//   sinkhorn(log_alpha, n_iters, slack)          :157-218  log-domain Sinkhorn with a slack row and column
//   exp / row-sum / perm @ xyz_ref               :283-287  (RPMNet.spam)
//   compute_rigid_transform(a, b, weights)       :221-254  weighted Kabsch  -> kabsch.cu (l3d_weighted_rigid_transform)
// (match_features / square_distance on C-dimensional features, :130-154, is the tensor-core Gram pipeline of
// softcorr.cu: l3d_feature_square_distance.)
//
// Sinkhorn.  The reference rewrites the whole padded (J+1) x (K+1) matrix twice per iteration (logsumexp,
// subtract, torch.cat: ~6 passes of the matrix per iteration).  Here the matrix is READ-ONLY: after any number of
// row / column normalisations the padded matrix is  A_pad[j,k] - u[j] - v[k]  with u[J] = v[K] = 0 (the slack row
// is never row-normalised, the slack column never column-normalised), and a normalisation step is just
//     u[j] = logsumexp_{k <= K}(A_pad[j,k] - v[k])        v[k] = logsumexp_{j <= J}(A_pad[j,k] - u[j]).
// One row sweep and one column sweep of A per iteration (the same tile loop as the EMD sweeps, emd.cu), the
// potentials live in L2, the result is written once.
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

#include <math.h>

namespace l3d {

struct Lse {           // running (max, sum of exp(x - max))
  float m, s;
  __device__ __forceinline__ void add(float x) {
    if (x > m) { s = s * expf(m - x) + 1.0f; m = x; }
    else s += expf(x - m);
  }
  __device__ __forceinline__ void merge(float om, float os) {
    if (os == 0.0f) return;
    if (om > m) { s = s * expf(m - om) + os; m = om; }
    else s += os * expf(om - m);
  }
};

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;

// u[b,j] = logsumexp_k (A[b,j,k] - v[b,k])  (+ the slack column's exp(0)); warp per row, lanes along k (coalesced)
__global__ void __launch_bounds__(SK_THREADS) sinkhorn_row_kernel(const float* __restrict__ A, const float* __restrict__ v,
                                                                  float* __restrict__ u, int J, int K, int slack) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int j = blockIdx.x * SK_WARPS + (threadIdx.x >> 5);
  if (j >= J) return;
  const float* row = A + ((size_t)b * J + j) * K;
  const float* vb = v + (size_t)b * K;
  Lse acc{-INFINITY, 0.0f};
  for (int k = lane; k < K; k += 32) acc.add(row[k] - vb[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(L3D_FULL_MASK, acc.m, o), os = __shfl_xor_sync(L3D_FULL_MASK, acc.s, o);
    acc.merge(om, os);
  }
  if (slack) acc.merge(0.0f, 1.0f);          // A_pad[j,K] - v[K] = 0
  if (lane == 0) u[(size_t)b * J + j] = acc.m + logf(acc.s);
}

// v[b,k] = logsumexp_j (A[b,j,k] - u[b,j])  (+ the slack row): a CTA owns 32 columns (lanes, coalesced), its 8 warps
// split the rows
__global__ void __launch_bounds__(SK_THREADS) sinkhorn_col_kernel(const float* __restrict__ A, const float* __restrict__ u,
                                                                  float* __restrict__ v, int J, int K, int slack) {
  __shared__ float sm[SK_WARPS][32], ss[SK_WARPS][32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const float* Ab = A + (size_t)b * J * K;
  const float* ub = u + (size_t)b * J;
  // four independent accumulators per thread (rows j, j+8, j+16, j+24 of this warp's share): the expf chains of
  // consecutive rows overlap instead of serialising (51 -> see profiles/r02)
  Lse a4[4] = {{-INFINITY, 0.0f}, {-INFINITY, 0.0f}, {-INFINITY, 0.0f}, {-INFINITY, 0.0f}};
  if (k < K) {
    int j = warp;
    for (; j + 3 * SK_WARPS < J; j += 4 * SK_WARPS) {
      float xv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) xv[q] = Ab[(size_t)(j + q * SK_WARPS) * K + k] - ub[j + q * SK_WARPS];
#pragma unroll
      for (int q = 0; q < 4; ++q) a4[q].add(xv[q]);
    }
    for (; j < J; j += SK_WARPS) a4[0].add(Ab[(size_t)j * K + k] - ub[j]);
  }
  Lse acc = a4[0];
#pragma unroll
  for (int q = 1; q < 4; ++q) acc.merge(a4[q].m, a4[q].s);
  sm[warp][lane] = acc.m; ss[warp][lane] = acc.s;
  __syncthreads();
  if (warp == 0 && k < K) {
    for (int w = 1; w < SK_WARPS; ++w) acc.merge(sm[w][lane], ss[w][lane]);
    if (slack) acc.merge(0.0f, 1.0f);
    v[(size_t)b * K + k] = acc.m + logf(acc.s);
  }
}

// out[b,j,k] = A - u[j] - v[k]  (log of the normalised matrix, what sinkhorn() returns); optionally fused with
// RPMNet.spam's tail: perm = exp(out), rowsum[j] = sum_k perm, weighted[j] = perm[j,:] @ xyz_ref / (rowsum + eps)
__global__ void __launch_bounds__(SK_THREADS) sinkhorn_finish_kernel(const float* __restrict__ A, const float* __restrict__ u,
                                                                     const float* __restrict__ v, int J, int K,
                                                                     float* __restrict__ log_out, float* __restrict__ perm_out,
                                                                     const float* __restrict__ xyz_ref, float eps,
                                                                     float* __restrict__ weighted, float* __restrict__ rowsum) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int j = blockIdx.x * SK_WARPS + (threadIdx.x >> 5);
  if (j >= J) return;
  const size_t ro = ((size_t)b * J + j) * K;
  const float uj = u ? u[(size_t)b * J + j] : 0.0f;
  const float* vb = v ? v + (size_t)b * K : nullptr;
  float rs = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float lp = A[ro + k] - uj - (vb ? vb[k] : 0.0f);
    if (log_out) log_out[ro + k] = lp;
    if (perm_out || weighted) {
      const float pm = expf(lp);
      if (perm_out) perm_out[ro + k] = pm;
      if (weighted) {
        const float* q = xyz_ref + ((size_t)b * K + k) * 3;
        rs += pm;
        wx = fmaf(pm, q[0], wx); wy = fmaf(pm, q[1], wy); wz = fmaf(pm, q[2], wz);
      }
    }
  }
  if (weighted) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rs += __shfl_xor_sync(L3D_FULL_MASK, rs, o);
      wx += __shfl_xor_sync(L3D_FULL_MASK, wx, o);
      wy += __shfl_xor_sync(L3D_FULL_MASK, wy, o);
      wz += __shfl_xor_sync(L3D_FULL_MASK, wz, o);
    }
    if (lane == 0) {
      const float d = rs + eps;
      float* o = weighted + ((size_t)b * J + j) * 3;
      o[0] = wx / d; o[1] = wy / d; o[2] = wz / d;
      if (rowsum) rowsum[(size_t)b * J + j] = rs;
    }
  }
}

__global__ void sinkhorn_zero_kernel(float* p, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0f;
}

}  // namespace l3d

using namespace l3d;

extern "C" size_t l3d_sinkhorn_ws_bytes(int B, int J, int K) {
  if (B < 1 || J < 1 || K < 1) return 0;
  return sizeof(float) * (size_t)B * ((size_t)J + K);
}

// Shared driver: n_iters row+column normalisations of log_alpha, then one finishing pass.
static int sinkhorn_run(const float* A, int B, int J, int K, int n_iters, int slack, float* log_out, float* perm_out,
                        const float* xyz_ref, float eps, float* weighted, float* rowsum, void* ws_dev, cudaStream_t s) {
  if (B < 0 || J < 1 || K < 1 || n_iters < 0) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  if (!A || !ws_dev || B > 65535) return L3D_ERR_INVALID;
  float* u = reinterpret_cast<float*>(ws_dev);
  float* v = u + (size_t)B * J;
  const long nz = (long)B * (J + K);
  sinkhorn_zero_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, s>>>(u, nz);
  count_launch();
  L3D_LAUNCH_CHECK();
  const dim3 grow((J + SK_WARPS - 1) / SK_WARPS, B), gcol((K + 31) / 32, B);
  for (int it = 0; it < n_iters; ++it) {
    sinkhorn_row_kernel<<<grow, SK_THREADS, 0, s>>>(A, v, u, J, K, slack);
    count_launch();
    L3D_LAUNCH_CHECK();
    sinkhorn_col_kernel<<<gcol, SK_THREADS, 0, s>>>(A, u, v, J, K, slack);
    count_launch();
    L3D_LAUNCH_CHECK();
  }
  sinkhorn_finish_kernel<<<grow, SK_THREADS, 0, s>>>(A, u, v, J, K, log_out, perm_out, xyz_ref, eps, weighted, rowsum);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" int l3d_sinkhorn(const float* log_alpha_dev, int B, int J, int K, int n_iters, int slack, float* out_dev,
                            void* ws_dev, void* stream) {
  if (!out_dev) return L3D_ERR_INVALID;
  return sinkhorn_run(log_alpha_dev, B, J, K, n_iters, slack, out_dev, nullptr, nullptr, 0.f, nullptr, nullptr, ws_dev,
                      (cudaStream_t)stream);
}

extern "C" int l3d_rpm_match_tail(const float* affinity_dev, const float* xyz_ref_dev, int B, int J, int K, int n_iters,
                                  int slack, float eps, float* perm_out_dev, float* weighted_out_dev,
                                  float* rowsum_out_dev, void* ws_dev, void* stream) {
  if (!xyz_ref_dev || !weighted_out_dev) return L3D_ERR_INVALID;
  return sinkhorn_run(affinity_dev, B, J, K, n_iters, slack, nullptr, perm_out_dev, xyz_ref_dev, eps, weighted_out_dev,
                      rowsum_out_dev, ws_dev, (cudaStream_t)stream);
}
