// Chamfer distance: nearest-neighbour search in both directions + deterministic backward.
//
// Replaces losses/cuda/chamfer_distance/chamfer_distance.cu (K1 :6-137, K2 :158-187) and the
// CPU path chamfer_distance.cpp:59-177 whose arithmetic it reproduces BIT-FOR-BIT:
//   d = (dx*dx + dy*dy) + dz*dz with every operation rounded (no fma), strict '<' so the lowest
//   index wins (chamfer_distance.cpp:68-80).
// Backward is a GATHER (no atomics): every output point accumulates its own term and the terms
// scattered to it in exactly the order of the reference's sequential CPU loops
// (chamfer_distance.cpp:141-176), so gradients are bit-identical and run-to-run deterministic
// (the reference CUDA kernel uses atomicAdd, chamfer_distance.cu:177-182).
//
// Fused loss path (ChamferDistanceLoss, losses/chamfer_distance.py:34-51): the forward kernel
// also produces 0.5*(mean sqrt d1 + mean sqrt d2) with a fixed-order two-level reduction (last
// CTA finishes), the backward kernel folds the sqrt/mean chain rule in — 2 launches per
// fwd+bwd instead of the reference's 2 + 2 memsets + 2 + ~10 elementwise launches.
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

namespace l3d {

constexpr int CH_THREADS = 256;
constexpr int CH_LPQ = 8;                       // lanes cooperating on one query
constexpr int CH_QPP = CH_THREADS / CH_LPQ;     // queries per pass (32)
constexpr int CH_CHUNK = 2048;                  // candidates staged per chunk (32 KB as float4)

struct ChamferFwdParams {
  const float* xyz[2];   // [B,n,3], [B,m,3]
  float* dist[2];        // [B,n], [B,m]
  int* idx[2];
  int cnt[2];            // n, m
  int B;
  int passes;            // query passes per CTA (tile = passes * 32 queries)
  // fused loss (optional)
  float* partial;        // [2 * gridDim.x * B] per-CTA sums of sqrt(d), or nullptr
  unsigned int* ticket;  // self-resetting arrival counter
  float* loss;           // scalar out
};

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(L3D_FULL_MASK, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < CH_THREADS / 32; ++w) s += red[w];
  }
  __syncthreads();
  return s;  // valid on thread 0
}

// d = x2*x2 + y2*y2 + z2*z2 with x2 = candidate - query, every operation rounded (.cpp:74-77)
__device__ __forceinline__ float chamfer_d2(const float4 cc, float qx, float qy, float qz) {
  const float dx = __fsub_rn(cc.x, qx), dy = __fsub_rn(cc.y, qy), dz = __fsub_rn(cc.z, qz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// grid = (query tiles, 2 directions, B).  A group of CH_LPQ lanes scans the candidates for Q queries at
// once: every candidate float4 is loaded from shared memory once per Q queries and the Q distance
// chains are independent (ILP) — Q = 4 for large batches, 1 when the grid would otherwise be too small.
template <int Q>
__global__ void __launch_bounds__(CH_THREADS) chamfer_fwd_kernel(const ChamferFwdParams p) {
  __shared__ float4 cand[CH_CHUNK];
  __shared__ float red[CH_THREADS / 32];
  __shared__ bool is_last;

  const int dir = blockIdx.y, b = blockIdx.z;
  const int nq = p.cnt[dir], nc = p.cnt[1 - dir];
  const float* q_xyz = p.xyz[dir] + (size_t)b * nq * 3;
  const float* c_xyz = p.xyz[1 - dir] + (size_t)b * nc * 3;
  const int tid = threadIdx.x;
  const int grp = tid / CH_LPQ, sub = tid % CH_LPQ;
  const int tile0 = blockIdx.x * p.passes * CH_QPP * Q;
  const int nchunks = (nc + CH_CHUNK - 1) / CH_CHUNK;

  float sqrt_sum = 0.f;
  if (tile0 < nq) {
    for (int pass = 0; pass < p.passes; ++pass) {
      int qi[Q];
      float qx[Q], qy[Q], qz[Q], best[Q];
      int besti[Q];
#pragma unroll
      for (int u = 0; u < Q; ++u) {
        qi[u] = tile0 + (pass * Q + u) * CH_QPP + grp;
        const int qc = min(qi[u], nq - 1);      // out-of-range slots recompute a valid query, never stored
        qx[u] = q_xyz[qc * 3]; qy[u] = q_xyz[qc * 3 + 1]; qz[u] = q_xyz[qc * 3 + 2];
        best[u] = INFINITY; besti[u] = 0;
      }
      bool have = false;
      for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * CH_CHUNK;
        const int cn = min(CH_CHUNK, nc - c0);
        if (nchunks > 1 || pass == 0) {
          __syncthreads();
          for (int j = tid; j < cn; j += CH_THREADS) {
            const float* s = c_xyz + (size_t)(c0 + j) * 3;
            cand[j] = make_float4(s[0], s[1], s[2], 0.f);
          }
          __syncthreads();
        }
        int j = sub;
        if (!have && j < cn) {   // `k == 0 ||` of the reference: the first candidate is taken as is
          const float4 cc = cand[j];
#pragma unroll
          for (int u = 0; u < Q; ++u) { best[u] = chamfer_d2(cc, qx[u], qy[u], qz[u]); besti[u] = c0 + j; }
          have = true;
          j += CH_LPQ;
        }
#pragma unroll 4
        for (; j < cn; j += CH_LPQ) {
          const float4 cc = cand[j];
#pragma unroll
          for (int u = 0; u < Q; ++u) {
            const float d = chamfer_d2(cc, qx[u], qy[u], qz[u]);
            if (d < best[u]) { best[u] = d; besti[u] = c0 + j; }   // strict '<' (.cpp:78)
          }
        }
      }
      // combine the 8 lanes of the group: smaller d, then lower index
#pragma unroll
      for (int u = 0; u < Q; ++u) {
        bool hv = have;   // a sub-lane beyond the cloud (nc < 8) holds nothing until it takes a partner's
#pragma unroll
        for (int o = CH_LPQ / 2; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(L3D_FULL_MASK, best[u], o);
          const int oi = __shfl_xor_sync(L3D_FULL_MASK, besti[u], o);
          const bool oh = __shfl_xor_sync(L3D_FULL_MASK, (int)hv, o) != 0;
          const bool take = oh && (!hv || od < best[u] || (od == best[u] && oi < besti[u]));
          if (take) { best[u] = od; besti[u] = oi; hv = true; }
        }
        if (qi[u] < nq && sub == 0) {
          p.dist[dir][(size_t)b * nq + qi[u]] = best[u];
          p.idx[dir][(size_t)b * nq + qi[u]] = besti[u];
          sqrt_sum += sqrtf(best[u]);
        }
      }
    }
  }

  if (p.partial) {
    // fixed-order two-level reduction of sum sqrt(d); the last CTA to arrive finishes the loss
    const float s = block_sum_256(sqrt_sum, red);
    const unsigned int ncta = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int lin = (blockIdx.y * gridDim.z + blockIdx.z) * gridDim.x + blockIdx.x;
    if (tid == 0) {
      p.partial[lin] = s;
      __threadfence();
      const unsigned int t = atomicAdd(p.ticket, 1u);
      is_last = (t == ncta - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const unsigned int per_dir = gridDim.x * gridDim.z;
      float tot[2];
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        float a = 0.f;
        for (unsigned int i = tid; i < per_dir; i += CH_THREADS)
          a += __ldcg(p.partial + d * per_dir + i);
        tot[d] = block_sum_256(a, red);
      }
      if (tid == 0) {
        // mean(sqrt(d1)), mean(sqrt(d2)), (a + b) / 2.0   (losses/chamfer_distance.py:38-40)
        const float c0 = tot[0] / (float)((long)p.B * p.cnt[0]);
        const float c1 = tot[1] / (float)((long)p.B * p.cnt[1]);
        *p.loss = (c0 + c1) / 2.0f;
        *p.ticket = 0u;   // self-reset for the next call
      }
    }
  }
}

struct ChamferBwdParams {
  const float* xyz[2];
  const int* idx[2];
  const float* graddist[2];  // [B,n], [B,m] upstream grads of the squared distances, or nullptr
  // fused-loss mode: graddist derived from dist + the scalar upstream gradient
  const float* dist[2];
  const float* grad_loss;    // device scalar
  float* grad[2];            // outputs [B,n,3], [B,m,3]
  int cnt[2];
  int B;
};

// g = graddist*2 as the reference's backward uses it (chamfer_distance.cpp:151,168)
__device__ __forceinline__ float chamfer_g(const ChamferBwdParams& p, int side, size_t off,
                                           float gl_half) {
  float gd;
  if (p.graddist[side]) {
    gd = p.graddist[side][off];
  } else {
    // d loss / d dist = (g/2) / numel / (2*sqrt(dist))   (autograd of div, mean, sqrt)
    const float u = gl_half / (float)((long)p.B * p.cnt[side]);
    gd = u / (2.0f * sqrtf(p.dist[side][off]));
  }
  return gd * 2.0f;
}

// grid = (tiles of 256 output points, 2 sides, B).  One lane per output point; a warp ballots 32
// arg-min indices of the OTHER cloud at a time and hands each hit to the lane that owns the target point,
// in ascending j — the reference's sequential accumulation order.  The other cloud's (x, y, z, g_j) and
// arg-min indices are staged in shared memory per CTA, so a hit costs two LDS (a first version chased
// dependent global loads: ~600 cycles per hit, 20 us at B=4 for 3 us of work).
constexpr int CHB_CHUNK = 2048;
__global__ void __launch_bounds__(CH_THREADS) chamfer_bwd_kernel2(const ChamferBwdParams p) {
  __shared__ float4 s_pt[CHB_CHUNK];
  __shared__ int s_idx[CHB_CHUNK];
  const int side = blockIdx.y, b = blockIdx.z;
  const int other = 1 - side;
  const int no = p.cnt[side], nx = p.cnt[other];
  const int tid = threadIdx.x, lane = tid & 31;
  const int i0 = (blockIdx.x * (CH_THREADS / 32) + (tid >> 5)) * 32;  // warp's first output point
  const int i = i0 + lane;
  const bool active = i < no;
  const float* my = p.xyz[side] + (size_t)b * no * 3;
  const float* ot = p.xyz[other] + (size_t)b * nx * 3;
  const int* oidx = p.idx[other] + (size_t)b * nx;
  const float gl_half = p.grad_loss ? (*p.grad_loss) * 0.5f : 0.f;

  float px = 0.f, py = 0.f, pz = 0.f;
  if (active) { px = my[i * 3]; py = my[i * 3 + 1]; pz = my[i * 3 + 2]; }
  // own term: grad[i] += g*(p_i - q_idx[i])        (.cpp:153-155 / :170-172)
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (active) {
    const int j2 = p.idx[side][(size_t)b * no + i];
    const float g = chamfer_g(p, side, (size_t)b * no + i, gl_half);
    ox = __fmul_rn(g, __fsub_rn(px, ot[j2 * 3]));
    oy = __fmul_rn(g, __fsub_rn(py, ot[j2 * 3 + 1]));
    oz = __fmul_rn(g, __fsub_rn(pz, ot[j2 * 3 + 2]));
  }
  float ax = 0.f, ay = 0.f, az = 0.f;
  // side 0 (xyz1): its own loop runs first, the scattered terms of loop 2 follow;
  // side 1 (xyz2): the scattered terms of loop 1 come first, then its own loop.
  if (side == 0) { ax = __fadd_rn(ax, ox); ay = __fadd_rn(ay, oy); az = __fadd_rn(az, oz); }

  for (int c0 = 0; c0 < nx; c0 += CHB_CHUNK) {
    const int cn = min(CHB_CHUNK, nx - c0);
    __syncthreads();
    for (int j = tid; j < cn; j += CH_THREADS) {
      const int jj = c0 + j;
      s_pt[j] = make_float4(ot[jj * 3], ot[jj * 3 + 1], ot[jj * 3 + 2],
                            chamfer_g(p, other, (size_t)b * nx + jj, gl_half));
      s_idx[j] = oidx[jj];
    }
    __syncthreads();
    if (i0 < no) {
      for (int j0 = 0; j0 < cn; j0 += 32) {
        const int j = j0 + lane;
        const int rel = (j < cn) ? (s_idx[j] - i0) : -1;
        unsigned hits = __ballot_sync(L3D_FULL_MASK, rel >= 0 && rel < 32);
        while (hits) {
          const int src = __ffs(hits) - 1;
          hits &= hits - 1;
          const int tgt = __shfl_sync(L3D_FULL_MASK, rel, src);
          if (lane == tgt) {
            // grad[idx[j]] -= g_j * (q_j - p_idx[j])      (.cpp:156-158 / :173-175)
            const float4 q = s_pt[j0 + src];
            ax = __fsub_rn(ax, __fmul_rn(q.w, __fsub_rn(q.x, px)));
            ay = __fsub_rn(ay, __fmul_rn(q.w, __fsub_rn(q.y, py)));
            az = __fsub_rn(az, __fmul_rn(q.w, __fsub_rn(q.z, pz)));
          }
        }
      }
    }
  }
  if (side == 1) { ax = __fadd_rn(ax, ox); ay = __fadd_rn(ay, oy); az = __fadd_rn(az, oz); }
  if (active) {
    float* o = p.grad[side] + ((size_t)b * no + i) * 3;
    o[0] = ax; o[1] = ay; o[2] = az;
  }
}

static int chamfer_fwd_launch(ChamferFwdParams p, cudaStream_t stream, unsigned* grid_x_out) {
  const int nmax = p.cnt[0] > p.cnt[1] ? p.cnt[0] : p.cnt[1];
  // queries per CTA = passes * Q * 32: as many as possible (Q-fold register blocking, staging amortised
  // over `passes`) while the grid still covers ~2 waves of the 148 SMs
  auto ctas = [&](int per_cta) { return (long)((nmax + per_cta - 1) / per_cta) * 2 * p.B; };
  int Q = 4;
  while (Q > 1 && ctas(Q * CH_QPP) < 2 * 148) Q >>= 1;
  int passes = 4;
  while (passes > 1 && ctas(passes * Q * CH_QPP) < 2 * 148) passes >>= 1;
  p.passes = passes;
  const int per_cta = passes * Q * CH_QPP;
  dim3 grid((nmax + per_cta - 1) / per_cta, 2, p.B);
  if (grid_x_out) *grid_x_out = grid.x;
  if (Q == 4) chamfer_fwd_kernel<4><<<grid, CH_THREADS, 0, stream>>>(p);
  else if (Q == 2) chamfer_fwd_kernel<2><<<grid, CH_THREADS, 0, stream>>>(p);
  else chamfer_fwd_kernel<1><<<grid, CH_THREADS, 0, stream>>>(p);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

static unsigned chamfer_grid_x_max(int n, int m) {
  const int nmax = n > m ? n : m;
  return (unsigned)((nmax + CH_QPP - 1) / CH_QPP);   // passes >= 1
}

}  // namespace l3d

using namespace l3d;

static bool chamfer_args_ok(const void* a, const void* b, int B, int n, int m) {
  return a && b && B >= 0 && n >= 1 && m >= 1 && B <= 65535;
}

extern "C" int l3d_chamfer_forward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                                   float* dist1_dev, float* dist2_dev, int32_t* idx1_dev,
                                   int32_t* idx2_dev, void* stream) {
  if (!chamfer_args_ok(xyz1_dev, xyz2_dev, B, n, m) || !dist1_dev || !dist2_dev || !idx1_dev || !idx2_dev)
    return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  ChamferFwdParams p{};
  p.xyz[0] = xyz1_dev; p.xyz[1] = xyz2_dev;
  p.dist[0] = dist1_dev; p.dist[1] = dist2_dev;
  p.idx[0] = idx1_dev; p.idx[1] = idx2_dev;
  p.cnt[0] = n; p.cnt[1] = m; p.B = B;
  return chamfer_fwd_launch(p, (cudaStream_t)stream, nullptr);
}

extern "C" int l3d_chamfer_backward(const float* xyz1_dev, const float* xyz2_dev, int B, int n, int m,
                                    const float* graddist1_dev, const float* graddist2_dev,
                                    const int32_t* idx1_dev, const int32_t* idx2_dev,
                                    float* gradxyz1_dev, float* gradxyz2_dev, void* stream) {
  if (!chamfer_args_ok(xyz1_dev, xyz2_dev, B, n, m) || !graddist1_dev || !graddist2_dev || !idx1_dev ||
      !idx2_dev || !gradxyz1_dev || !gradxyz2_dev)
    return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  ChamferBwdParams p{};
  p.xyz[0] = xyz1_dev; p.xyz[1] = xyz2_dev;
  p.idx[0] = idx1_dev; p.idx[1] = idx2_dev;
  p.graddist[0] = graddist1_dev; p.graddist[1] = graddist2_dev;
  p.grad[0] = gradxyz1_dev; p.grad[1] = gradxyz2_dev;
  p.cnt[0] = n; p.cnt[1] = m; p.B = B;
  const int nmax = n > m ? n : m;
  dim3 grid((nmax + CH_THREADS - 1) / CH_THREADS, 2, B);
  chamfer_bwd_kernel2<<<grid, CH_THREADS, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" size_t l3d_chamfer_ws_bytes(int B, int n, int m) {
  if (B < 0 || n < 1 || m < 1) return 0;
  return 16 + sizeof(float) * 2 * (size_t)chamfer_grid_x_max(n, m) * (size_t)(B > 0 ? B : 1);
}

extern "C" int l3d_chamfer_loss_forward(const float* xyz1_dev, const float* xyz2_dev, int B, int n,
                                        int m, float* dist1_dev, float* dist2_dev, int32_t* idx1_dev,
                                        int32_t* idx2_dev, float* loss_dev, void* ws_dev,
                                        void* stream) {
  if (!chamfer_args_ok(xyz1_dev, xyz2_dev, B, n, m) || !dist1_dev || !dist2_dev || !idx1_dev ||
      !idx2_dev || !loss_dev || !ws_dev || B < 1)
    return L3D_ERR_INVALID;
  ChamferFwdParams p{};
  p.xyz[0] = xyz1_dev; p.xyz[1] = xyz2_dev;
  p.dist[0] = dist1_dev; p.dist[1] = dist2_dev;
  p.idx[0] = idx1_dev; p.idx[1] = idx2_dev;
  p.cnt[0] = n; p.cnt[1] = m; p.B = B;
  p.ticket = reinterpret_cast<unsigned int*>(ws_dev);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws_dev) + 16);
  p.loss = loss_dev;
  return chamfer_fwd_launch(p, (cudaStream_t)stream, nullptr);
}

extern "C" int l3d_chamfer_loss_backward(const float* xyz1_dev, const float* xyz2_dev, int B, int n,
                                         int m, const float* dist1_dev, const float* dist2_dev,
                                         const int32_t* idx1_dev, const int32_t* idx2_dev,
                                         const float* grad_loss_dev, float* gradxyz1_dev,
                                         float* gradxyz2_dev, void* stream) {
  if (!chamfer_args_ok(xyz1_dev, xyz2_dev, B, n, m) || !dist1_dev || !dist2_dev || !idx1_dev ||
      !idx2_dev || !grad_loss_dev || !gradxyz1_dev || !gradxyz2_dev)
    return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  ChamferBwdParams p{};
  p.xyz[0] = xyz1_dev; p.xyz[1] = xyz2_dev;
  p.idx[0] = idx1_dev; p.idx[1] = idx2_dev;
  p.dist[0] = dist1_dev; p.dist[1] = dist2_dev;
  p.grad_loss = grad_loss_dev;
  p.grad[0] = gradxyz1_dev; p.grad[1] = gradxyz2_dev;
  p.cnt[0] = n; p.cnt[1] = m; p.B = B;
  const int nmax = n > m ? n : m;
  dim3 grid((nmax + CH_THREADS - 1) / CH_THREADS, 2, B);
  chamfer_bwd_kernel2<<<grid, CH_THREADS, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}
