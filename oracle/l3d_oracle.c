/*
 * l3d_oracle.c — CPU restatement of learning3d's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (learning3d_b200/) never does.  Every function cites the reference
 * file:line it restates (paths relative to the learning3d checkout) and spells out the fp32
 * rounding sequence explicitly: compile with -ffp-contract=off so that the ONLY fused
 * operations are the fmaf() calls written below.
 *
 * Pinning (see oracle/README.md, tests/golden/make_golden.py): every function here was
 * checked against the real reference imported from /root/reference in the build container
 * (pure-torch functions, and the Chamfer CPU extension) and the outputs are committed as
 * fixtures under tests/golden/.  Functions restating CUDA-only reference kernels (pointnet2,
 * EMD) say so in their comment.
 *
 * Selection order everywhere: better key first, ties -> lower index first.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int l3d_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- helpers ---------------------------------------------------------------------- */

/* dot product with the accumulation order of a K=3 sgemm (MKL and cuBLAS):
 * acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc).  Verified bit-for-bit against
 * torch.matmul on CPU (oracle/README.md). */
static inline float dot3_gemm(float a0, float a1, float a2, float b0, float b1, float b2) {
  float acc = a0 * b0;
  acc = fmaf(a1, b1, acc);
  acc = fmaf(a2, b2, acc);
  return acc;
}

/* torch.sum(x**2, dim) over three elements: ((x*x + y*y) + z*z), each op rounded. */
static inline float sumsq3(float x, float y, float z) {
  float s = x * x;
  s = s + y * y;
  s = s + z * z;
  return s;
}

/* insert (key, j) into a list sorted by (key desc, idx asc); candidates arrive in ascending j,
 * so an equal key goes AFTER the entries already there. */
static inline void topk_insert(float* keys, int64_t* ids, int* cnt, int k, float key, int64_t j) {
  int n = *cnt;
  if (n == k && !(key > keys[k - 1])) return;
  int pos = n < k ? n : k - 1;
  while (pos > 0 && key > keys[pos - 1]) {
    keys[pos] = keys[pos - 1];
    ids[pos] = ids[pos - 1];
    --pos;
  }
  keys[pos] = key;
  ids[pos] = j;
  if (n < k) *cnt = n + 1;
}

/* ---- kNN family ------------------------------------------------------------------- */

/* knn(x, k): utils/model_common_utils.py:3-9.
 *   inner = -2*matmul(x^T, x)                          (:5)
 *   xx = sum(x**2, dim=1, keepdim)                     (:6)
 *   pd = -xx - inner - xx^T  == ((-xx[j]) - inner[i][j]) - xx[i]   (:7, App. A of SURVEY.md)
 *   idx = pd.topk(k)[1]  (largest first)               (:8)
 * x [B,3,N]; idx [B,N,k] int64; val (optional) [B,N,k] = pd of the selected entries. */
void l3d_oracle_knn_expansion(const float* x, int B, int N, int k, int64_t* idx, float* val) {
#pragma omp parallel
  {
    float* keys = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    float* xx = (float*)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      const float* X = x + (size_t)b * 3 * N;
      const float* Y = X + N;
      const float* Z = Y + N;
      for (int j = 0; j < N; ++j) xx[j] = sumsq3(X[j], Y[j], Z[j]);
      for (int i = 0; i < N; ++i) {
        int cnt = 0;
        for (int j = 0; j < N; ++j) {
          const float dot = dot3_gemm(X[i], Y[i], Z[i], X[j], Y[j], Z[j]);
          const float inner = -2.0f * dot;
          float pd = (-xx[j]) - inner;
          pd = pd - xx[i];
          topk_insert(keys, ids, &cnt, k, pd, j);
        }
        const size_t o = ((size_t)b * N + i) * k;
        for (int r = 0; r < k; ++r) {
          idx[o + r] = ids[r];
          if (val) val[o + r] = keys[r];
        }
      }
    }
    free(keys); free(ids); free(xx);
  }
}

/* Same as above but parallel over rows (b, i) — used only for timing (cpu_baseline) where
 * B may be smaller than the core count. */
void l3d_oracle_knn_expansion_mt(const float* x, int B, int N, int k, int64_t* idx) {
  float* xx = (float*)malloc(sizeof(float) * (size_t)B * N);
  for (int b = 0; b < B; ++b) {
    const float* X = x + (size_t)b * 3 * N;
    for (int j = 0; j < N; ++j) xx[(size_t)b * N + j] = sumsq3(X[j], X[N + j], X[2 * N + j]);
  }
#pragma omp parallel
  {
    float* keys = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(static)
    for (long row = 0; row < (long)B * N; ++row) {
      const int b = (int)(row / N), i = (int)(row % N);
      const float* X = x + (size_t)b * 3 * N;
      const float* Y = X + N;
      const float* Z = Y + N;
      const float* xxb = xx + (size_t)b * N;
      int cnt = 0;
      for (int j = 0; j < N; ++j) {
        const float dot = dot3_gemm(X[i], Y[i], Z[i], X[j], Y[j], Z[j]);
        float pd = (-xxb[j]) - (-2.0f * dot);
        pd = pd - xxb[i];
        topk_insert(keys, ids, &cnt, k, pd, j);
      }
      for (int r = 0; r < k; ++r) idx[(size_t)row * k + r] = ids[r];
    }
    free(keys); free(ids);
  }
  free(xx);
}

/* get_graph_feature gather: utils/model_common_utils.py:149-154.
 * out[b, c, n, j] = x[b, c, idx[b,n,j]];  out[b, C+c, n, j] = x[b, c, n]. */
void l3d_oracle_graph_feature(const float* x, const int64_t* idx, int B, int C, int N, int k,
                              float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int n = 0; n < N; ++n)
        for (int j = 0; j < k; ++j) {
          const int64_t m = idx[((size_t)b * N + n) * k + j];
          out[(((size_t)b * 2 * C + c) * N + n) * k + j] = x[((size_t)b * C + c) * N + m];
          out[(((size_t)b * 2 * C + C + c) * N + n) * k + j] = x[((size_t)b * C + c) * N + n];
        }
}

/* backward of the gather (autograd of :149-154): sequential accumulation. */
void l3d_oracle_graph_feature_grad(const float* go, const int64_t* idx, int B, int C, int N, int k,
                                   float* gx) {
  memset(gx, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int n = 0; n < N; ++n)
        for (int j = 0; j < k; ++j) {
          const int64_t m = idx[((size_t)b * N + n) * k + j];
          gx[((size_t)b * C + c) * N + m] += go[(((size_t)b * 2 * C + c) * N + n) * k + j];
          gx[((size_t)b * C + c) * N + n] += go[(((size_t)b * 2 * C + C + c) * N + n) * k + j];
        }
}

/* square_distance(src, dst): utils/model_common_utils.py:19-38 (copies: pointconv_util.py:18-39,
 * ppfnet_util.py:29-48).  dist = -2*matmul(src, dst^T); dist += |src|^2; dist += |dst|^2.
 * src [B,N,3], dst [B,M,3] -> out [B,N,M]. */
static inline float sqdist_exp(const float* s, const float* d) {
  const float dot = dot3_gemm(s[0], s[1], s[2], d[0], d[1], d[2]);
  float v = -2.0f * dot;
  v = v + sumsq3(s[0], s[1], s[2]);
  v = v + sumsq3(d[0], d[1], d[2]);
  return v;
}
void l3d_oracle_square_distance(const float* src, const float* dst, int B, int N, int M,
                                float* out) {
#pragma omp parallel for schedule(static)
  for (long r = 0; r < (long)B * N; ++r) {
    const int b = (int)(r / N);
    for (int m = 0; m < M; ++m)
      out[(size_t)r * M + m] = sqdist_exp(src + (size_t)r * 3, dst + ((size_t)b * M + m) * 3);
  }
}

/* pointconv_util.knn_point(nsample, xyz, new_xyz): utils/pointconv_util.py:107-118.
 * topk(square_distance(new_xyz, xyz), nsample, largest=False, sorted=False).  The reference
 * leaves the order unspecified; this oracle returns ascending distance, ties -> lower index.
 * xyz [B,N,3], new_xyz [B,S,3] -> idx [B,S,nsample] int64. */
void l3d_oracle_knn_sqdist(const float* xyz, const float* new_xyz, int B, int N, int S,
                           int nsample, int64_t* idx) {
#pragma omp parallel
  {
    float* keys = (float*)malloc(sizeof(float) * (size_t)nsample);
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)nsample);
#pragma omp for schedule(static)
    for (long r = 0; r < (long)B * S; ++r) {
      const int b = (int)(r / S);
      int cnt = 0;
      for (int j = 0; j < N; ++j) {
        const float v = sqdist_exp(new_xyz + (size_t)r * 3, xyz + ((size_t)b * N + j) * 3);
        topk_insert(keys, ids, &cnt, nsample, -v, j);
      }
      for (int t = 0; t < nsample; ++t) idx[(size_t)r * nsample + t] = ids[t];
    }
    free(keys); free(ids);
  }
}

/* knn_point(k, pos1, pos2): utils/model_common_utils.py:84-100.
 *   dist = sum(-(pos1-pos2)**2, -1); val, idx = dist.topk(k); return sqrt(-val), idx
 * pos1 = data [B,N,3], pos2 = query [B,M,3] -> val [B,M,k], idx [B,M,k] int64. */
void l3d_oracle_knn_point(const float* data, const float* query, int B, int N, int M, int k,
                          float* val, int64_t* idx) {
#pragma omp parallel
  {
    float* keys = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(static)
    for (long r = 0; r < (long)B * M; ++r) {
      const int b = (int)(r / M);
      const float* q = query + (size_t)r * 3;
      int cnt = 0;
      for (int j = 0; j < N; ++j) {
        const float* p = data + ((size_t)b * N + j) * 3;
        const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
        float s = -(dx * dx);
        s = s + (-(dy * dy));
        s = s + (-(dz * dz));
        topk_insert(keys, ids, &cnt, k, s, j);
      }
      for (int t = 0; t < k; ++t) {
        idx[(size_t)r * k + t] = ids[t];
        if (val) val[(size_t)r * k + t] = sqrtf(-keys[t]);
      }
    }
    free(keys); free(ids);
  }
}

/* Squared distance as the pointnet2 CUDA kernels evaluate it.  The source expression is
 * dx*dx + dy*dy + dz*dz (ball_query_gpu.cu:33, interpolate_gpu.cu:38,104, sampling_gpu.cu:131);
 * nvcc 12.9 -O2 for sm_100 contracts it to FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.) — read off
 * the SASS of the reference files themselves (oracle/README.md). */
static inline float pn2_d2(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* pointnet2 knn_kernel_fast / three_nn_kernel_fast: utils/lib/src/interpolate_gpu.cu:9-57,
 * 81-124 (CUDA-only reference; restated).  d = (ux-x)*(ux-x) + (uy-y)*(uy-y) + (uz-z)*(uz-z)
 * as nvcc contracts it (pn2_d2 above); insertion with strict '<' keeps the
 * earlier index first among equal distances; output ascending d2 + int32 idx.
 * unknown [b,n,3] queries, known [b,m,3] data. */
void l3d_oracle_pn2_knn(int b, int n, int m, int k, const float* unknown, const float* known,
                        float* dist2, int32_t* idx) {
#pragma omp parallel
  {
    float* keys = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(static)
    for (long r = 0; r < (long)b * n; ++r) {
      const int bi = (int)(r / n);
      const float* u = unknown + (size_t)r * 3;
      int cnt = 0;
      for (int j = 0; j < m; ++j) {
        const float* p = known + ((size_t)bi * m + j) * 3;
        const float dx = u[0] - p[0], dy = u[1] - p[1], dz = u[2] - p[2];
        const float d = pn2_d2(dx, dy, dz);
        topk_insert(keys, ids, &cnt, k, -d, j);
      }
      for (int t = 0; t < k; ++t) {
        idx[(size_t)r * k + t] = (int32_t)ids[t];
        dist2[(size_t)r * k + t] = -keys[t];
      }
    }
    free(keys); free(ids);
  }
}

/* ---- Chamfer ---------------------------------------------------------------------- */

/* nnsearch: losses/cuda/chamfer_distance/chamfer_distance.cpp:59-87.  For every point of
 * xyz1 [b,n,3] the squared distance to and index of its nearest point of xyz2 [b,m,3];
 * products and sums in float, strict '<' (lowest index wins). */
static void nnsearch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist,
                     int32_t* idx) {
#pragma omp parallel for schedule(static)
  for (long r = 0; r < (long)b * n; ++r) {
    const int i = (int)(r / n);
    const float x1 = xyz1[r * 3 + 0], y1 = xyz1[r * 3 + 1], z1 = xyz1[r * 3 + 2];
    double best = 0;
    int besti = 0;
    for (int k = 0; k < m; ++k) {
      const float x2 = xyz2[((size_t)i * m + k) * 3 + 0] - x1;
      const float y2 = xyz2[((size_t)i * m + k) * 3 + 1] - y1;
      const float z2 = xyz2[((size_t)i * m + k) * 3 + 2] - z1;
      const float df = x2 * x2 + y2 * y2 + z2 * z2; /* float expression, then widened (:77) */
      const double d = df;
      if (k == 0 || d < best) { best = d; besti = k; }
    }
    dist[r] = (float)best;
    idx[r] = besti;
  }
}

/* chamfer_distance_forward: chamfer_distance.cpp:90-111. */
void l3d_oracle_chamfer_forward(const float* xyz1, const float* xyz2, int b, int n, int m,
                                float* dist1, float* dist2, int32_t* idx1, int32_t* idx2) {
  nnsearch(b, n, m, xyz1, xyz2, dist1, idx1);
  nnsearch(b, m, n, xyz2, xyz1, dist2, idx2);
}

/* chamfer_distance_backward: chamfer_distance.cpp:114-177 (sequential, same statement order). */
void l3d_oracle_chamfer_backward(const float* xyz1, const float* xyz2, int b, int n, int m,
                                 const float* graddist1, const float* graddist2,
                                 const int32_t* idx1, const int32_t* idx2, float* gradxyz1,
                                 float* gradxyz2) {
  for (long i = 0; i < (long)b * n * 3; ++i) gradxyz1[i] = 0;
  for (long i = 0; i < (long)b * m * 3; ++i) gradxyz2[i] = 0;
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      const float x1 = xyz1[((size_t)i * n + j) * 3 + 0];
      const float y1 = xyz1[((size_t)i * n + j) * 3 + 1];
      const float z1 = xyz1[((size_t)i * n + j) * 3 + 2];
      const int j2 = idx1[(size_t)i * n + j];
      const float x2 = xyz2[((size_t)i * m + j2) * 3 + 0];
      const float y2 = xyz2[((size_t)i * m + j2) * 3 + 1];
      const float z2 = xyz2[((size_t)i * m + j2) * 3 + 2];
      const float g = graddist1[(size_t)i * n + j] * 2;
      gradxyz1[((size_t)i * n + j) * 3 + 0] += g * (x1 - x2);
      gradxyz1[((size_t)i * n + j) * 3 + 1] += g * (y1 - y2);
      gradxyz1[((size_t)i * n + j) * 3 + 2] += g * (z1 - z2);
      gradxyz2[((size_t)i * m + j2) * 3 + 0] -= (g * (x1 - x2));
      gradxyz2[((size_t)i * m + j2) * 3 + 1] -= (g * (y1 - y2));
      gradxyz2[((size_t)i * m + j2) * 3 + 2] -= (g * (z1 - z2));
    }
    for (int j = 0; j < m; ++j) {
      const float x1 = xyz2[((size_t)i * m + j) * 3 + 0];
      const float y1 = xyz2[((size_t)i * m + j) * 3 + 1];
      const float z1 = xyz2[((size_t)i * m + j) * 3 + 2];
      const int j2 = idx2[(size_t)i * m + j];
      const float x2 = xyz1[((size_t)i * n + j2) * 3 + 0];
      const float y2 = xyz1[((size_t)i * n + j2) * 3 + 1];
      const float z2 = xyz1[((size_t)i * n + j2) * 3 + 2];
      const float g = graddist2[(size_t)i * m + j] * 2;
      gradxyz2[((size_t)i * m + j) * 3 + 0] += g * (x1 - x2);
      gradxyz2[((size_t)i * m + j) * 3 + 1] += g * (y1 - y2);
      gradxyz2[((size_t)i * m + j) * 3 + 2] += g * (z1 - z2);
      gradxyz1[((size_t)i * n + j2) * 3 + 0] -= (g * (x1 - x2));
      gradxyz1[((size_t)i * n + j2) * 3 + 1] -= (g * (y1 - y2));
      gradxyz1[((size_t)i * n + j2) * 3 + 2] -= (g * (z1 - z2));
    }
  }
}

/* chamfer_distance(): losses/chamfer_distance.py:34-40 on top of the native forward:
 * (mean(sqrt(d1)) + mean(sqrt(d2))) / 2, accumulated in double (a checker, not a bit model
 * of torch.mean's summation tree). */
double l3d_oracle_chamfer_loss(const float* xyz1, const float* xyz2, int b, int n, int m) {
  float* d1 = (float*)malloc(sizeof(float) * (size_t)b * n);
  float* d2 = (float*)malloc(sizeof(float) * (size_t)b * m);
  int32_t* i1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)b * n);
  int32_t* i2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)b * m);
  l3d_oracle_chamfer_forward(xyz1, xyz2, b, n, m, d1, d2, i1, i2);
  double s1 = 0, s2 = 0;
  for (long i = 0; i < (long)b * n; ++i) s1 += sqrt((double)d1[i]);
  for (long i = 0; i < (long)b * m; ++i) s2 += sqrt((double)d2[i]);
  free(d1); free(d2); free(i1); free(i2);
  return 0.5 * (s1 / ((double)b * n) + s2 / ((double)b * m));
}
