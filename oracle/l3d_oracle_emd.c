/*
 * l3d_oracle_emd.c — CPU restatement of the approximate EMD (approxmatch / matchcost / grads).
 * TEST INFRASTRUCTURE ONLY (rules in l3d_oracle.c).
 *
 * The reference is CUDA-only, so this file follows losses/cuda/emd_torch/pkg/include/cuda/emd.cuh
 * statement by statement, sequentially, in fp32, with libm expf standing in for the kernel's __expf
 * (ex2.approx): parity with a GPU implementation is to a tolerance (1e-5 relative on cost / match mass /
 * gradients), not bit-exact.  Pinning: the reference extension's host glue no longer compiles (AT_CHECK /
 * tensor.type()), but its KERNELS do behind oracle/ref_shims/emd_shim.cu (oracle/_ref/libemd_ref.so); on the
 * GPU box tests/test_gpu_emd_svd.py::test_emd_against_reference_cuda_kernels runs them next to the product
 * kernels, and this restatement agrees with both to the same tolerance.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* approxmatch<float>: emd.cuh:6-185.  xyz1 [b,n,3], xyz2 [b,m,3] -> match [b, m*n] with the
 * reference's index l*n + k (l over xyz2, k over xyz1; emd.cuh:158). */
void l3d_oracle_emd_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2,
                                float* match) {
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = (float)(n / m); }       /* integer division, :10-16 */
  else { multiL = (float)(m / n); multiR = 1; }
#pragma omp parallel for schedule(dynamic)                  /* batch items are independent (:19) */
  for (int i = 0; i < b; ++i) {
    float* remainL = (float*)malloc(sizeof(float) * (size_t)n);
    float* remainR = (float*)malloc(sizeof(float) * (size_t)m);
    float* ratioL = (float*)malloc(sizeof(float) * (size_t)n);
    float* ratioR = (float*)malloc(sizeof(float) * (size_t)m);
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    float* mt = match + (size_t)i * n * m;
    memset(mt, 0, sizeof(float) * (size_t)n * m);
    for (int j = 0; j < n; ++j) remainL[j] = multiL;
    for (int j = 0; j < m; ++j) remainR[j] = multiR;
    for (int j = 7; j >= -2; --j) {
      float level = -powf(4.0f, (float)j);
      if (j == -2) level = 0;
      for (int k = 0; k < n; ++k) {                          /* :32-60 */
        const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) {
          const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
          const float d = level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1));
          suml += expf(d) * remainR[l];
        }
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {                          /* :80-116 */
        const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
        float sumr = 0;
        for (int k = 0; k < n; ++k) {
          const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
          sumr += expf(level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1))) * ratioL[k];
        }
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {                          /* :135-168 */
        const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        float suml = 0;
        const float rl = ratioL[k];
        for (int l = 0; l < m; ++l) {
          const float x2 = p2[l * 3], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
          const float w = expf(level * ((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1))) * rl * ratioR[l];
          mt[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    free(remainL); free(remainR); free(ratioL); free(ratioR);
  }
}

/* matchcost<float>: emd.cuh:201-244 (double accumulation of the block sum: a checker). */
void l3d_oracle_emd_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2,
                              const float* match, float* cost) {
  for (int i = 0; i < b; ++i) {
    double s = 0;
    for (int k = 0; k < n; ++k) {
      const float* p = xyz1 + ((size_t)i * n + k) * 3;
      for (int l = 0; l < m; ++l) {
        const float* q = xyz2 + ((size_t)i * m + l) * 3;
        const float d = sqrtf((q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]));
        s += (double)(d * match[(size_t)i * n * m + (size_t)l * n + k]);
      }
    }
    cost[i] = (float)s;
  }
}

/* matchcostgrad1 / matchcostgrad2: emd.cuh:301-323, 258-299 (rsqrtf(fmaxf(d2, 1e-20f))). */
void l3d_oracle_emd_grads(int b, int n, int m, const float* xyz1, const float* xyz2,
                          const float* match, float* grad1, float* grad2) {
  for (int i = 0; i < b; ++i) {
    for (int l = 0; l < n; ++l) {
      const float* p = xyz1 + ((size_t)i * n + l) * 3;
      double dx = 0, dy = 0, dz = 0;
      for (int k = 0; k < m; ++k) {
        const float* q = xyz2 + ((size_t)i * m + k) * 3;
        const float ex = p[0] - q[0], ey = p[1] - q[1], ez = p[2] - q[2];
        const float d = match[(size_t)i * n * m + (size_t)k * n + l] / sqrtf(fmaxf(ex * ex + ey * ey + ez * ez, 1e-20f));
        dx += (double)(ex * d); dy += (double)(ey * d); dz += (double)(ez * d);
      }
      grad1[((size_t)i * n + l) * 3] = (float)dx;
      grad1[((size_t)i * n + l) * 3 + 1] = (float)dy;
      grad1[((size_t)i * n + l) * 3 + 2] = (float)dz;
    }
    for (int k = 0; k < m; ++k) {
      const float* q = xyz2 + ((size_t)i * m + k) * 3;
      double sx = 0, sy = 0, sz = 0;
      for (int j = 0; j < n; ++j) {
        const float* p = xyz1 + ((size_t)i * n + j) * 3;
        const float ex = q[0] - p[0], ey = q[1] - p[1], ez = q[2] - p[2];
        const float d = match[(size_t)i * n * m + (size_t)k * n + j] / sqrtf(fmaxf(ex * ex + ey * ey + ez * ez, 1e-20f));
        sx += (double)(ex * d); sy += (double)(ey * d); sz += (double)(ez * d);
      }
      grad2[((size_t)i * m + k) * 3] = (float)sx;
      grad2[((size_t)i * m + k) * 3 + 1] = (float)sy;
      grad2[((size_t)i * m + k) * 3 + 2] = (float)sz;
    }
  }
}
