/*
 * l3d_oracle_group.c — CPU restatement of the grouping family.  TEST INFRASTRUCTURE ONLY
 * (see l3d_oracle.c for the rules).  Two sources are restated:
 *   (a) the pointnet2 CUDA kernels under utils/lib/src/ — CUDA-only in the reference, so these
 *       functions follow the kernels statement by statement, including nvcc's fma contraction of
 *       `dx*dx + dy*dy + dz*dz` (FMUL dy,dy; FFMA dx,dx; FFMA dz,dz — read off the SASS of the
 *       reference files, oracle/README.md) and the block-level tree reductions;
 *   (b) the pure-torch helpers of utils/model_common_utils.py, pointconv_util.py, ppfnet_util.py —
 *       pinned against the real reference on CPU (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float pn2_d2(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}
static inline float dot3_gemm(float a0, float a1, float a2, float b0, float b1, float b2) {
  float acc = a0 * b0;
  acc = fmaf(a1, b1, acc);
  acc = fmaf(a2, b2, acc);
  return acc;
}
static inline float sumsq3(float x, float y, float z) {
  float s = x * x;
  s = s + y * y;
  s = s + z * z;
  return s;
}
static inline float sqdist_exp(const float* s, const float* d) {
  float v = -2.0f * dot3_gemm(s[0], s[1], s[2], d[0], d[1], d[2]);
  v = v + sumsq3(s[0], s[1], s[2]);
  v = v + sumsq3(d[0], d[1], d[2]);
  return v;
}

/* ball_query_kernel_fast: utils/lib/src/ball_query_gpu.cu:9-45 (idx pre-zeroed by
 * utils/lib/pointnet2_utils.py:245). */
void l3d_oracle_pn2_ball_query(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                               const float* xyz, int32_t* idx) {
  memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
  const float radius2 = radius * radius;
  for (int bs = 0; bs < b; ++bs)
    for (int pt = 0; pt < m; ++pt) {
      const float* q = new_xyz + ((size_t)bs * m + pt) * 3;
      const float* X = xyz + (size_t)bs * n * 3;
      int32_t* o = idx + ((size_t)bs * m + pt) * nsample;
      int cnt = 0;
      for (int k = 0; k < n; ++k) {
        const float d2 = pn2_d2(q[0] - X[k * 3], q[1] - X[k * 3 + 1], q[2] - X[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
          if (cnt >= nsample) break;
        }
      }
    }
}

/* query_ball_point: utils/model_common_utils.py:102-130, pointconv_util.py:85-105,
 * ppfnet_util.py:96-131.  group_idx = arange(N); [itself -> N]; group_idx[sqrdists > r2] = N;
 * sort; [:nsample]; entries == N replaced by the first entry (or by itself_indices). */
void l3d_oracle_query_ball_point(const float* xyz, const float* new_xyz, int B, int N, int S,
                                 float radius2, int nsample, const int64_t* itself, int64_t* out,
                                 int64_t* cnt) {
  for (long r = 0; r < (long)B * S; ++r) {
    const int b = (int)(r / S);
    int64_t* o = out + (size_t)r * nsample;
    int w = 0;
    int64_t total = 0;
    for (int k = 0; k < N; ++k) {
      const float d = sqdist_exp(new_xyz + (size_t)r * 3, xyz + ((size_t)b * N + k) * 3);
      const int masked = (d > radius2) || (itself && itself[r] == k);
      if (!masked) {
        if (w < nsample) o[w++] = k;
        ++total;
      }
    }
    const int64_t first = itself ? itself[r] : (w > 0 ? o[0] : (int64_t)N);
    for (int s = w; s < nsample; ++s) o[s] = first;
    if (cnt) cnt[r] = total;
  }
}

/* group_points_kernel_fast / gather_points_kernel_fast: group_points_gpu.cu:47-66,
 * sampling_gpu.cu:8-24.  out[b,c,pos] = points[b,c,idx[b,pos]], P = npoints*nsample. */
void l3d_oracle_pn2_group_points(int b, int c, int n, long P, const float* points,
                                 const int32_t* idx, float* out) {
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (long p = 0; p < P; ++p)
        out[((size_t)bi * c + ci) * P + p] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * P + p]];
}
/* group_points_grad / gather_points_grad: group_points_gpu.cu:8-25, sampling_gpu.cu:46-63
 * (atomicAdd there; sequential here). */
void l3d_oracle_pn2_group_points_grad(int b, int c, int n, long P, const float* grad_out,
                                      const int32_t* idx, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (long p = 0; p < P; ++p)
        grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * P + p]] +=
            grad_out[((size_t)bi * c + ci) * P + p];
}

/* three_interpolate_kernel_fast: interpolate_gpu.cu:149-169, contracted as
 * fma(w2,p2, fma(w0,p0, w1*p1)) (SASS of the reference file). */
void l3d_oracle_pn2_three_interpolate(int b, int c, int m, int n, const float* points,
                                      const int32_t* idx, const float* weight, float* out) {
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (int p = 0; p < n; ++p) {
        const int32_t* ip = idx + ((size_t)bi * n + p) * 3;
        const float* w = weight + ((size_t)bi * n + p) * 3;
        const float* pp = points + ((size_t)bi * c + ci) * m;
        out[((size_t)bi * c + ci) * n + p] = fmaf(w[2], pp[ip[2]], fmaf(w[0], pp[ip[0]], w[1] * pp[ip[1]]));
      }
}
/* three_interpolate_grad_kernel_fast: interpolate_gpu.cu:192-214. */
void l3d_oracle_pn2_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out,
                                           const int32_t* idx, const float* weight,
                                           float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
  for (int bi = 0; bi < b; ++bi)
    for (int ci = 0; ci < c; ++ci)
      for (int p = 0; p < n; ++p) {
        const int32_t* ip = idx + ((size_t)bi * n + p) * 3;
        const float* w = weight + ((size_t)bi * n + p) * 3;
        const float g = grad_out[((size_t)bi * c + ci) * n + p];
        float* gp = grad_points + ((size_t)bi * c + ci) * m;
        gp[ip[0]] += g * w[0];
        gp[ip[1]] += g * w[1];
        gp[ip[2]] += g * w[2];
      }
}

/* furthest_point_sampling_kernel<block_size>: sampling_gpu.cu:93-209, simulated literally:
 * block_size threads striding over the points, per-thread strict '>' (:136-137), shared-memory
 * tree with __update keeping the left entry on ties (:86-91), block size from opt_n_threads
 * (cuda_utils.h:10-14).  temp [b,n] holds 1e10 on entry. */
void l3d_oracle_pn2_fps(int b, int n, int m, const float* dataset_all, float* temp_all,
                        int32_t* idxs_all) {
  if (m <= 0) return;
  const int pow_2 = (int)(log((double)n) / log(2.0));
  int bs = 1 << pow_2;
  if (bs > 1024) bs = 1024;
  if (bs < 1) bs = 1;
  float* dists = (float*)malloc(sizeof(float) * (size_t)bs);
  int* dists_i = (int*)malloc(sizeof(int) * (size_t)bs);
  for (int bi = 0; bi < b; ++bi) {
    const float* dataset = dataset_all + (size_t)bi * n * 3;
    float* temp = temp_all + (size_t)bi * n;
    int32_t* idxs = idxs_all + (size_t)bi * m;
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1;
        for (int k = tid; k < n; k += bs) {
          const float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          const float d = pn2_d2(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = d < temp[k] ? d : temp[k]; /* min(d, temp[k]) */
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int s = bs >> 1; s >= 1; s >>= 1)
        for (int tid = 0; tid < s; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + s];
          const int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      old = dists_i[0];
      idxs[j] = old;
    }
  }
  free(dists);
  free(dists_i);
}

/* farthest_point_sample: model_common_utils.py:58-82 (start optional), pointconv_util.py:60-83
 * (start 0), ppfnet_util.py:71-93 (random start): dist = sum((xyz - centroid)**2, -1);
 * distance = min(distance, dist) through the mask; farthest = first arg-max. */
void l3d_oracle_farthest_point_sample(const float* xyz_all, int B, int N, int npoint,
                                      const int64_t* start, int64_t* centroids) {
  float* distance = (float*)malloc(sizeof(float) * (size_t)N);
  for (int b = 0; b < B; ++b) {
    const float* xyz = xyz_all + (size_t)b * N * 3;
    for (int k = 0; k < N; ++k) distance[k] = 1e10f;
    int64_t farthest = start ? start[b] : 0;
    for (int i = 0; i < npoint; ++i) {
      centroids[(size_t)b * npoint + i] = farthest;
      const float cx = xyz[farthest * 3], cy = xyz[farthest * 3 + 1], cz = xyz[farthest * 3 + 2];
      int64_t arg = 0;
      float best = -INFINITY;
      for (int k = 0; k < N; ++k) {
        const float dx = xyz[k * 3] - cx, dy = xyz[k * 3 + 1] - cy, dz = xyz[k * 3 + 2] - cz;
        float dist = dx * dx;
        dist = dist + dy * dy;
        dist = dist + dz * dz;
        if (dist < distance[k]) distance[k] = dist;
        if (distance[k] > best) { best = distance[k]; arg = k; }
      }
      farthest = arg;
    }
  }
  free(distance);
}

/* index_points: model_common_utils.py:40-56.  points [B,N,C], idx [B,R] -> out [B,R,C]. */
void l3d_oracle_index_points(const float* points, const int64_t* idx, int B, int N, long R, int C,
                             float* out) {
  for (int b = 0; b < B; ++b)
    for (long r = 0; r < R; ++r)
      memcpy(out + ((size_t)b * R + r) * C, points + ((size_t)b * N + idx[(size_t)b * R + r]) * C,
             sizeof(float) * (size_t)C);
}

/* compute_density: pointconv_util.py:199-209, row mean accumulated in double (a checker for a
 * fp32 result compared with a tolerance). */
void l3d_oracle_compute_density(const float* xyz, int B, int N, float two_bw2, float norm,
                                float* out) {
  for (long r = 0; r < (long)B * N; ++r) {
    const int b = (int)(r / N);
    double acc = 0;
    for (int j = 0; j < N; ++j) {
      const float d = sqdist_exp(xyz + (size_t)r * 3, xyz + ((size_t)b * N + j) * 3);
      acc += (double)(expf(-d / two_bw2) / norm);
    }
    out[r] = (float)(acc / N);
  }
}
