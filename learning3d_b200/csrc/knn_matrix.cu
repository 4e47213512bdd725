// Top-k selection over the rows of a precomputed key matrix (sm_100a).
//
// Second half of knn() for C != 3 (utils/model_common_utils.py:8, `pairwise_distance.topk(k)`): the keys
// come from the tensor-core Gram kernel in softcorr.cu instead of being evaluated in registers, the
// selection is the one of knn.cu's hot path: a warp owns two rows, each lane holds 32 keys
// of a 1024-key tile, the k-th largest lane maximum is a proven threshold, survivors become 64-bit
// (order-preserving key bits, ~index) composites and go through the flip + half-cleaner network.
// Rows whose survivors overflow (exact ties: duplicated feature vectors after ReLU / max-pool are common)
// and k > 24 take the exact k-round arg-max scan.  Largest key first, lower index first on equal keys.
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "knn_matrix.h"
#include "launch_count.h"

#include <math.h>

namespace l3d {

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_TILE = 1024;
constexpr int KM_R = 2;
constexpr int KM_CAP = 64;

struct KnnMatParams {
  const float* keys;    // [rows, N]
  long long* out_idx;   // [rows, k]
  long rows;
  int N, k;
  int force_slow;
};

// exact O(k*N) selection: k rounds of "best pair strictly after the previous one"
__device__ __noinline__ void km_row_slow(const KnnMatParams& p, long row, int lane) {
  const float* kp = p.keys + row * (long)p.N;
  float pv = INFINITY;
  uint32_t pi = 0;
  bool first = true;
  for (int r = 0; r < p.k; ++r) {
    float bv = -INFINITY;
    uint32_t bi = 0xffffffffu;
    for (int j = lane; j < p.N; j += 32) {
      const float d = __fadd_rn(__ldg(kp + j), 0.0f);
      const bool after = first || better(pv, pi, d, (uint32_t)j);
      if (after && better(d, (uint32_t)j, bv, bi)) { bv = d; bi = (uint32_t)j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(L3D_FULL_MASK, bv, o);
      const uint32_t oi = __shfl_xor_sync(L3D_FULL_MASK, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) p.out_idx[row * p.k + r] = (long long)bi;
    pv = bv; pi = bi; first = false;
  }
}

template <int R>
__device__ __forceinline__ void km_rows(const KnnMatParams& p, unsigned long long* __restrict__ cbuf, long row0,
                                        int lane) {
  const int k = p.k, N = p.N;
  const int ntiles = (N + KM_TILE - 1) / KM_TILE;
  const float* kp[R];
  int base[R];
  float kth[R];
  bool ovf[R];
  unsigned long long best[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    kp[r] = p.keys + (row0 + r) * (long)N;
    base[r] = 0; kth[r] = -INFINITY; ovf[r] = (p.force_slow != 0); best[r] = 0ull;
  }
  for (int t = 0; t < ntiles; ++t) {
    float d[R][32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int j = t * KM_TILE + e * 32 + lane;
#pragma unroll
      for (int r = 0; r < R; ++r) d[r][e] = (j < N) ? __ldg(kp[r] + j) : -INFINITY;
    }
    float mx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float m = d[r][0];
#pragma unroll
      for (int e = 1; e < 32; ++e) m = fmaxf(m, d[r][e]);
      mx[r] = warp_sort32_keys_desc(m, lane);
    }
    uint32_t mask[R];
    int cnt[R], incl[R], total[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float thr = fmaxf(__shfl_sync(L3D_FULL_MASK, mx[r], k - 1), kth[r]);
      uint32_t mk = 0u;
#pragma unroll
      for (int e = 0; e < 32; ++e) mk |= (d[r][e] >= thr) ? (1u << e) : 0u;   // NaN keys never survive
      mask[r] = mk;
      cnt[r] = __popc(mk);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) incl[r] = warp_inclusive_scan(cnt[r], lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      total[r] = __shfl_sync(L3D_FULL_MASK, incl[r], 31);
      // (second clause: a first tile with NaN keys can leave fewer than k survivors -> exact path)
      if (base[r] + total[r] > KM_CAP || (t == 0 && total[r] < k)) ovf[r] = true;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (ovf[r]) continue;
      unsigned long long* cb = cbuf + r * KM_CAP;
      if (base[r] && lane < base[r]) cb[lane] = best[r];
      int off = base[r] + incl[r] - cnt[r];
      uint32_t mk = mask[r];
      while (mk) {
        const int e = __ffs(mk) - 1;
        mk &= mk - 1;
        const int j = t * KM_TILE + e * 32 + lane;
        cb[off++] = pack_pair(__ldg(kp[r] + j), (uint32_t)j);
      }
    }
    __syncwarp();
    unsigned long long a[R];
    int n_in[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      n_in[r] = base[r] + total[r];
      a[r] = (!ovf[r] && lane < n_in[r]) ? cbuf[r * KM_CAP + lane] : 0ull;
    }
    warp_sort32_desc_x<R>(a, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (n_in[r] > 32 && !ovf[r]) {
        unsigned long long b = (lane + 32 < n_in[r]) ? cbuf[r * KM_CAP + lane + 32] : 0ull;
        b = shfl_xor_u64(warp_sort32_desc(b, lane), 31);
        unsigned long long c = (b > a[r]) ? b : a[r];
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) c = cmpx_u64(c, j, (lane & j) == 0);
        a[r] = c;
      }
      best[r] = a[r];
      if (t + 1 < ntiles) {
        kth[r] = f32_unorder(__shfl_sync(L3D_FULL_MASK, (uint32_t)(best[r] >> 32), k - 1));
        base[r] = k;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (ovf[r]) km_row_slow(p, row0 + r, lane);
    else if (lane < k) p.out_idx[(row0 + r) * k + lane] = (long long)(~(uint32_t)best[r]);
  }
}

__global__ void __launch_bounds__(KM_THREADS) knn_matrix_kernel(const KnnMatParams p) {
  __shared__ unsigned long long cbuf_all[KM_WARPS][KM_R * KM_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* cbuf = cbuf_all[warp];
  const long stride = (long)gridDim.x * KM_WARPS * KM_R;
  for (long row = ((long)blockIdx.x * KM_WARPS + warp) * KM_R; row < p.rows; row += stride) {
    if (p.k > 24) {
      for (int r = 0; r < KM_R && row + r < p.rows; ++r) km_row_slow(p, row + r, lane);
    } else if (row + KM_R <= p.rows) {
      km_rows<KM_R>(p, cbuf, row, lane);
    } else {
      km_rows<1>(p, cbuf, row, lane);
    }
  }
}

int knn_select_from_matrix(const float* keys, long rows, int N, int k, long long* idx, cudaStream_t stream) {
  if (rows <= 0) return L3D_OK;
  if (!keys || !idx || k < 1 || k > N) return L3D_ERR_INVALID;
  KnnMatParams p;
  p.keys = keys; p.out_idx = idx; p.rows = rows; p.N = N; p.k = k; p.force_slow = knn_force_slow_flag();
  long grid = (rows + KM_WARPS * KM_R - 1) / (KM_WARPS * KM_R);
  if (grid > 148L * 8) grid = 148L * 8;
  knn_matrix_kernel<<<(unsigned)grid, KM_THREADS, 0, stream>>>(p);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

}  // namespace l3d

// Row-wise top-k of a key matrix that is already in HBM (largest key first, lower index on ties): the selection
// stage of l3d_knn_features as an entry point of its own.  Used for knn_point() on C-dimensional features
// (utils/model_common_utils.py:84-100 with C != 3): keys = -square_distance from l3d_feature_square_distance.
extern "C" int l3d_topk_rows(const float* keys_dev, long long rows, int N, int k, int64_t* idx_dev, void* stream) {
  if (rows < 0 || N < 1 || k < 1 || k > N) return L3D_ERR_INVALID;
  if (rows == 0) return L3D_OK;
  if (!keys_dev || !idx_dev) return L3D_ERR_INVALID;
  return l3d::knn_select_from_matrix(keys_dev, (long)rows, N, k, reinterpret_cast<long long*>(idx_dev),
                                     (cudaStream_t)stream);
}
