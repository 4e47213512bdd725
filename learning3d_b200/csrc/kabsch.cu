// Batched 3x3 Kabsch / SVD-head tail (sm_100a).
//
// Replaces the tail of SVDHead.forward (utils/svd.py:29-58): centring, H = src_c * corr_c^T,
// the Python `for i in range(B)` loop of torch.svd + torch.det + a host-synchronising
// `if r_det < 0` branch (:38-51), and t = -R*mean(src) + mean(src_corr) (:58).
// One CTA per batch item reduces the means and H over N (fp64 accumulation), thread 0 then runs a
// one-sided Jacobi SVD of the 3x3 in fp64 (a handful of rotations), fixes the determinant by
// negating the right singular vector of the SMALLEST singular value (v * diag(1,1,-1), :44-45)
// and writes R, t.  No host round trip, one launch for the whole batch.  Latency bound by design
// (36 B in / 48 B out per item once H is known).
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

namespace l3d {

// `fact` (optional, 22 doubles): U (9, row-major), V (9), sigma (3, descending), d = +-1 of the
// determinant fix — what the backward pass needs.
__device__ void kabsch_from_H(const double Hin[9], const float mu_s[3], const float mu_c[3],
                              float* __restrict__ R_out, float* __restrict__ t_out,
                              double* __restrict__ fact = nullptr) {
  double A[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { A[i][j] = Hin[i * 3 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }

  // one-sided (Hestenes) Jacobi: rotate column pairs of A (and V) until mutually orthogonal
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pair = 0; pair < 3; ++pair) {
      const int p = (pair == 2) ? 1 : 0;
      const int q = (pair == 0) ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        alpha += A[i][p] * A[i][p];
        beta += A[i][q] * A[i][q];
        gamma += A[i][p] * A[i][q];
      }
      if (fabs(gamma) > 1e-15 * sqrt(alpha * beta) && gamma != 0.0) {
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq;
          A[i][q] = s * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
        rotated = true;
      }
    }
    if (!rotated) break;
  }
  // singular values = column norms; sort descending (torch.svd order) by swapping columns
  double sig[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const int a = (pass == 1) ? 1 : 0, b = (pass == 1) ? 2 : 1;   // (0,1), (1,2), (0,1)
    if (sig[a] < sig[b]) {
      const double ts = sig[a]; sig[a] = sig[b]; sig[b] = ts;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double tv = A[i][a]; A[i][a] = A[i][b]; A[i][b] = tv;
        tv = V[i][a]; V[i][a] = V[i][b]; V[i][b] = tv;
      }
    }
  }
  // U = A * Sigma^-1; a vanishing singular value leaves its direction free: complete the basis
  double U[3][3];
  const double tiny = 1e-300;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double inv = sig[j] > tiny ? 1.0 / sig[j] : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][j] = A[i][j] * inv;
  }
  if (sig[2] > 1e-14 * sig[0] && sig[2] > tiny) {
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][2] = A[i][2] / sig[2];
  } else {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
  // r = v u^T; if det(r) < 0: v <- v * diag(1,1,-1)   (svd.py:40-45)
  double R[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
  const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) -
                     R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                     R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (det < 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] -= 2.0 * V[i][2] * U[j][2];
  }
  if (fact) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { fact[i * 3 + j] = U[i][j]; fact[9 + i * 3 + j] = V[i][j]; }
    fact[18] = sig[0]; fact[19] = sig[1]; fact[20] = sig[2];
    fact[21] = det < 0 ? -1.0 : 1.0;
  }
  float Rf[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Rf[i * 3 + j] = (float)R[i][j];
      if (R_out) R_out[i * 3 + j] = Rf[i * 3 + j];
    }
  if (!t_out) return;
  // t = matmul(-R, mean(src)) + mean(src_corr)   (svd.py:58), fp32 like the reference
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float acc = fmaf(-Rf[i * 3 + 2], mu_s[2], fmaf(-Rf[i * 3 + 1], mu_s[1], (-Rf[i * 3]) * mu_s[0]));
    t_out[i] = acc + mu_c[i];
  }
}

__global__ void __launch_bounds__(32) kabsch_kernel(const float* __restrict__ H, const float* __restrict__ mu_s,
                                                    const float* __restrict__ mu_c, int B,
                                                    float* __restrict__ R, float* __restrict__ t) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double h[9];
  float ms[3], mc[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) h[i] = (double)H[(size_t)b * 9 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { ms[i] = mu_s[b * 3 + i]; mc[i] = mu_c[b * 3 + i]; }
  kabsch_from_H(h, ms, mc, R + (size_t)b * 9, t + (size_t)b * 3);
}

// src, corr: [B,3,N].  One CTA per batch item.
constexpr int SVDH_THREADS = 256;
__global__ void __launch_bounds__(SVDH_THREADS) svd_head_tail_kernel(const float* __restrict__ src,
                                                                     const float* __restrict__ corr, int N,
                                                                     float* __restrict__ R,
                                                                     float* __restrict__ t) {
  __shared__ double red[SVDH_THREADS / 32][15];
  __shared__ double tot[15];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* s = src + (size_t)b * 3 * N;
  const float* c = corr + (size_t)b * 3 * N;
  // sums: 0-2 sum src, 3-5 sum corr, 6-14 sum src_i*corr_j ; H = sum s c^T - N mu_s mu_c^T
  double acc[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.0;
  for (int n = tid; n < N; n += SVDH_THREADS) {
    const double sv[3] = {(double)s[n], (double)s[N + n], (double)s[2 * N + n]};
    const double cv[3] = {(double)c[n], (double)c[N + n], (double)c[2 * N + n]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { acc[i] += sv[i]; acc[3 + i] += cv[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[6 + i * 3 + j] += sv[i] * cv[j];
  }
#pragma unroll
  for (int i = 0; i < 15; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(L3D_FULL_MASK, acc[i], o);
    if (lane == 0) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (tid < 15) {
    double v = 0.0;
    for (int w = 0; w < SVDH_THREADS / 32; ++w) v += red[w][tid];
    tot[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    const double inv = 1.0 / (double)N;
    double h[9];
    float ms[3], mc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ms[i] = (float)(tot[i] * inv); mc[i] = (float)(tot[3 + i] * inv); }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) h[i * 3 + j] = tot[6 + i * 3 + j] - tot[i] * tot[3 + j] * inv;
    kabsch_from_H(h, ms, mc, R + (size_t)b * 9, t + (size_t)b * 3);
  }
}

// ---- backward of the SVD-head tail ---------------------------------------------------------------
// R = V D U^T (D = diag(1,1,d)), t = -R mu_s + mu_c, H = sum_n (s_n - mu_s)(c_n - mu_c)^T = U S V^T.
// With G = dL/dR - g_t mu_s^T and M = V^T G U, the differential of the SVD gives dL/dH = U Q V^T,
//   Q_ij = [ M_ij (D_jj s_i - D_ii s_j) - M_ji (D_ii s_i - D_jj s_j) ] / (s_j^2 - s_i^2)   (i != j), Q_ii = 0
// (for D = I this is -(M_ij - M_ji)/(s_i + s_j): the polar-factor derivative, no s_i - s_j pole).
// Then dL/ds_n = (dL/dH)(c_n - mu_c) - R^T g_t / N,  dL/dc_n = (dL/dH)^T (s_n - mu_s) + g_t / N.
// Same numbers as torch autograd through torch.svd on the reference formulation (tests).
__global__ void __launch_bounds__(SVDH_THREADS) svd_head_tail_bwd_kernel(
    const float* __restrict__ src, const float* __restrict__ corr, const float* __restrict__ gR,
    const float* __restrict__ gt, int N, float* __restrict__ g_src, float* __restrict__ g_corr) {
  __shared__ double red[SVDH_THREADS / 32][15];
  __shared__ double tot[15];
  __shared__ float s_gh[9], s_mu[6], s_ds[3], s_dc[3];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* s = src + (size_t)b * 3 * N;
  const float* c = corr + (size_t)b * 3 * N;
  double acc[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.0;
  for (int n = tid; n < N; n += SVDH_THREADS) {
    const double sv[3] = {(double)s[n], (double)s[N + n], (double)s[2 * N + n]};
    const double cv[3] = {(double)c[n], (double)c[N + n], (double)c[2 * N + n]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { acc[i] += sv[i]; acc[3 + i] += cv[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[6 + i * 3 + j] += sv[i] * cv[j];
  }
#pragma unroll
  for (int i = 0; i < 15; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(L3D_FULL_MASK, acc[i], o);
    if (lane == 0) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (tid < 15) {
    double v = 0.0;
    for (int w = 0; w < SVDH_THREADS / 32; ++w) v += red[w][tid];
    tot[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    const double inv = 1.0 / (double)N;
    double h[9], fact[22], mu_s[3], mu_c[3];
    float ms[3], mc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      mu_s[i] = tot[i] * inv; mu_c[i] = tot[3 + i] * inv;
      ms[i] = (float)mu_s[i]; mc[i] = (float)mu_c[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) h[i * 3 + j] = tot[6 + i * 3 + j] - tot[i] * tot[3 + j] * inv;
    kabsch_from_H(h, ms, mc, nullptr, nullptr, fact);
    const double* U = fact; const double* V = fact + 9; const double* sg = fact + 18;
    const double D[3] = {1.0, 1.0, fact[21]};
    double G[9], g3[3], R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) g3[i] = (double)gt[(size_t)b * 3 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        G[i * 3 + j] = (double)gR[(size_t)b * 9 + i * 3 + j] - g3[i] * mu_s[j];
        R[i * 3 + j] = V[i * 3 + 0] * D[0] * U[j * 3 + 0] + V[i * 3 + 1] * D[1] * U[j * 3 + 1] + V[i * 3 + 2] * D[2] * U[j * 3 + 2];
      }
    double M[9], Q[9], T[9], GH[9];
    // M = V^T G U
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) T[i * 3 + j] = V[0 * 3 + i] * G[0 * 3 + j] + V[1 * 3 + i] * G[1 * 3 + j] + V[2 * 3 + i] * G[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) M[i * 3 + j] = T[i * 3 + 0] * U[0 * 3 + j] + T[i * 3 + 1] * U[1 * 3 + j] + T[i * 3 + 2] * U[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (i == j) { Q[i * 3 + j] = 0.0; continue; }
        const double num = M[i * 3 + j] * (D[j] * sg[i] - D[i] * sg[j]) - M[j * 3 + i] * (D[i] * sg[i] - D[j] * sg[j]);
        const double den = sg[j] * sg[j] - sg[i] * sg[i];
        double q;
        if (D[i] == D[j]) q = -(M[i * 3 + j] - M[j * 3 + i]) / fmax(sg[i] + sg[j], 1e-300);   // pole-free form
        else q = num / (den != 0.0 ? den : 1e-300);
        Q[i * 3 + j] = q;
      }
    // GH = U Q V^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) T[i * 3 + j] = U[i * 3 + 0] * Q[0 * 3 + j] + U[i * 3 + 1] * Q[1 * 3 + j] + U[i * 3 + 2] * Q[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) GH[i * 3 + j] = T[i * 3 + 0] * V[j * 3 + 0] + T[i * 3 + 1] * V[j * 3 + 1] + T[i * 3 + 2] * V[j * 3 + 2];
#pragma unroll
    for (int i = 0; i < 9; ++i) s_gh[i] = (float)GH[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      s_mu[i] = ms[i]; s_mu[3 + i] = mc[i];
      // d mu_s = -R^T g_t, d mu_c = g_t, each spread over the N points
      s_ds[i] = (float)(-(R[0 * 3 + i] * g3[0] + R[1 * 3 + i] * g3[1] + R[2 * 3 + i] * g3[2]) * inv);
      s_dc[i] = (float)(g3[i] * inv);
    }
  }
  __syncthreads();
  float* gs = g_src + (size_t)b * 3 * N;
  float* gc = g_corr + (size_t)b * 3 * N;
  for (int n = tid; n < N; n += SVDH_THREADS) {
    const float sx = s[n] - s_mu[0], sy = s[N + n] - s_mu[1], sz = s[2 * N + n] - s_mu[2];
    const float cx = c[n] - s_mu[3], cy = c[N + n] - s_mu[4], cz = c[2 * N + n] - s_mu[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gs[(size_t)i * N + n] = s_gh[i * 3] * cx + s_gh[i * 3 + 1] * cy + s_gh[i * 3 + 2] * cz + s_ds[i];
      gc[(size_t)i * N + n] = s_gh[i] * sx + s_gh[3 + i] * sy + s_gh[6 + i] * sz + s_dc[i];
    }
  }
}

// compute_rigid_transform (rpmnet.py:221-254): one CTA per item, fp64 sums.
//   w~ = w / (sum w + eps);  ca = sum w~ a;  cb = sum w~ b;  cov = sum w~ (a - ca)(b - cb)^T
//      = sum w~ a b^T - (2 - W) ca cb^T,  W = sum w~
//   R = V U^T of cov = U S V^T (third column of V negated when det < 0), t = -R ca + cb
constexpr int RT_THREADS = 256;
__global__ void __launch_bounds__(RT_THREADS) weighted_rigid_kernel(const float* __restrict__ a, const float* __restrict__ bp,
                                                                    const float* __restrict__ w, int M, float eps,
                                                                    float* __restrict__ T) {
  __shared__ double red[RT_THREADS / 32][16];
  __shared__ double tot[16];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* A = a + (size_t)b * M * 3;
  const float* Bq = bp + (size_t)b * M * 3;
  const float* W = w + (size_t)b * M;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;
  for (int n = tid; n < M; n += RT_THREADS) {
    const double wn = (double)W[n];
    const double av[3] = {(double)A[n * 3], (double)A[n * 3 + 1], (double)A[n * 3 + 2]};
    const double bv[3] = {(double)Bq[n * 3], (double)Bq[n * 3 + 1], (double)Bq[n * 3 + 2]};
    acc[15] += wn;
#pragma unroll
    for (int i = 0; i < 3; ++i) { acc[i] += wn * av[i]; acc[3 + i] += wn * bv[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[6 + i * 3 + j] += wn * av[i] * bv[j];
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(L3D_FULL_MASK, acc[i], o);
    if (lane == 0) red[warp][i] = acc[i];
  }
  __syncthreads();
  if (tid < 16) {
    double v = 0.0;
    for (int q = 0; q < RT_THREADS / 32; ++q) v += red[q][tid];
    tot[tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    const double inv = 1.0 / (tot[15] + (double)eps);      // weights_normalized = w / (sum w + eps)
    const double Wn = tot[15] * inv;
    double ca[3], cb[3], h[9];
    float caf[3], cbf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ca[i] = tot[i] * inv; cb[i] = tot[3 + i] * inv; caf[i] = (float)ca[i]; cbf[i] = (float)cb[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) h[i * 3 + j] = tot[6 + i * 3 + j] * inv - (2.0 - Wn) * ca[i] * cb[j];
    float R[9], t[3];
    kabsch_from_H(h, caf, cbf, R, t, nullptr);
    float* o = T + (size_t)b * 12;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      o[i * 4 + 0] = R[i * 3 + 0]; o[i * 4 + 1] = R[i * 3 + 1]; o[i * 4 + 2] = R[i * 3 + 2]; o[i * 4 + 3] = t[i];
    }
  }
}

}  // namespace l3d

using namespace l3d;

extern "C" int l3d_kabsch3x3_batched(const float* H_dev, const float* src_mean_dev,
                                     const float* corr_mean_dev, int B, float* R_dev, float* t_dev,
                                     void* stream) {
  if (!H_dev || !src_mean_dev || !corr_mean_dev || !R_dev || !t_dev || B < 0) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  kabsch_kernel<<<(B + 31) / 32, 32, 0, (cudaStream_t)stream>>>(H_dev, src_mean_dev, corr_mean_dev, B,
                                                                R_dev, t_dev);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" int l3d_svd_head_tail(const float* src_dev, const float* src_corr_dev, int B, int N,
                                 float* R_dev, float* t_dev, void* stream) {
  if (!src_dev || !src_corr_dev || !R_dev || !t_dev || B < 0 || N < 1) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  svd_head_tail_kernel<<<B, SVDH_THREADS, 0, (cudaStream_t)stream>>>(src_dev, src_corr_dev, N, R_dev, t_dev);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" int l3d_svd_head_tail_backward(const float* src_dev, const float* src_corr_dev,
                                          const float* grad_R_dev, const float* grad_t_dev, int B, int N,
                                          float* grad_src_dev, float* grad_src_corr_dev, void* stream) {
  if (!src_dev || !src_corr_dev || !grad_R_dev || !grad_t_dev || !grad_src_dev || !grad_src_corr_dev ||
      B < 0 || N < 1)
    return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  svd_head_tail_bwd_kernel<<<B, SVDH_THREADS, 0, (cudaStream_t)stream>>>(
      src_dev, src_corr_dev, grad_R_dev, grad_t_dev, N, grad_src_dev, grad_src_corr_dev);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}

extern "C" int l3d_weighted_rigid_transform(const float* a_dev, const float* b_dev, const float* w_dev, int B, int M,
                                            float eps, float* T_dev, void* stream) {
  if (B < 0 || M < 1) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  if (!a_dev || !b_dev || !w_dev || !T_dev) return L3D_ERR_INVALID;
  weighted_rigid_kernel<<<B, RT_THREADS, 0, (cudaStream_t)stream>>>(a_dev, b_dev, w_dev, M, eps, T_dev);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}
