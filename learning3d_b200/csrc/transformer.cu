// Pieces of DCP's "pointer" transformer (utils/transformer.py) that are not GEMMs, for activations kept
// channel-major ([B, d_model, N], the layout DGCNN produces and SVDHead consumes — the reference transposes to
// [B, N, d_model] at the transformer's entry and back at its exit, transformer.py:257-262).
//
// LayerNorm (transformer.py:128-137, the "annotated transformer" flavour): per position, over the d_model channels:
//     y = a_2 * (x - mean) / (std + eps) + b_2,   std = UNBIASED standard deviation (torch.std default), eps OUTSIDE
// The GEMMs (nn.Linear, q k^T, p v) run on the tcgen05 pipelines of edgeconv.cu / softcorr.cu.
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

namespace l3d {

constexpr int LN_POS = 32;       // positions per CTA (one 128-byte line per channel row)
constexpr int LN_SLICES = 8;     // channel slices per position
// VPT = channels per thread held in REGISTERS (D <= 8 * VPT): x is read once and written once (2 passes of HBM
// instead of 4).  VPT = 0: any D, three reads.
template <int VPT>
__global__ void __launch_bounds__(LN_POS * LN_SLICES) layernorm_cm_kernel(const float* __restrict__ x,
                                                                          const float* __restrict__ a2,
                                                                          const float* __restrict__ b2, float eps, int D,
                                                                          int N, float* __restrict__ out) {
  __shared__ float red[LN_SLICES][LN_POS + 1];
  __shared__ float s_mean[LN_POS], s_inv[LN_POS];
  const int b = blockIdx.y, ln = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int n = blockIdx.x * LN_POS + ln;
  const bool ok = n < N;
  const float* xp = x + (size_t)b * D * N + (ok ? n : 0);
  float* op = out + (size_t)b * D * N + (ok ? n : 0);
  const int c0 = (int)((long)D * sl / LN_SLICES), c1 = (int)((long)D * (sl + 1) / LN_SLICES);
  float v[VPT > 0 ? VPT : 1];
  float acc = 0.f;
  if (VPT > 0) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      v[i] = (ok && c0 + i < c1) ? __ldg(xp + (size_t)(c0 + i) * N) : 0.f;
      acc += v[i];
    }
  } else if (ok) {
    for (int c = c0; c < c1; ++c) acc += __ldg(xp + (size_t)c * N);
  }
  red[sl][ln] = acc;
  __syncthreads();
  if (sl == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_SLICES; ++k) s += red[k][ln];
    s_mean[ln] = s / (float)D;
  }
  __syncthreads();
  const float mean = s_mean[ln];
  acc = 0.f;
  if (VPT > 0) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const float d = (c0 + i < c1) ? v[i] - mean : 0.f;
      acc = fmaf(d, d, acc);
    }
  } else if (ok) {
    for (int c = c0; c < c1; ++c) { const float d = __ldg(xp + (size_t)c * N) - mean; acc = fmaf(d, d, acc); }
  }
  __syncthreads();
  red[sl][ln] = acc;
  __syncthreads();
  if (sl == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_SLICES; ++k) s += red[k][ln];
    s_inv[ln] = 1.0f / (sqrtf(s / (float)(D > 1 ? D - 1 : 1)) + eps);      // unbiased std, eps added to the std
  }
  __syncthreads();
  const float inv = s_inv[ln];
  if (!ok) return;
  if (VPT > 0) {
#pragma unroll
    for (int i = 0; i < VPT; ++i)
      if (c0 + i < c1) op[(size_t)(c0 + i) * N] = fmaf(__ldg(a2 + c0 + i) * (v[i] - mean), inv, __ldg(b2 + c0 + i));
  } else {
    for (int c = c0; c < c1; ++c)
      op[(size_t)c * N] = fmaf(__ldg(a2 + c) * (__ldg(xp + (size_t)c * N) - mean), inv, __ldg(b2 + c));
  }
}

}  // namespace l3d

using namespace l3d;

extern "C" int l3d_layernorm_cm(const float* x_dev, const float* a2_dev, const float* b2_dev, float eps, int B, int D,
                                int N, float* out_dev, void* stream) {
  if (B < 0 || D < 1 || N < 0) return L3D_ERR_INVALID;
  if (B == 0 || N == 0) return L3D_OK;
  if (!x_dev || !a2_dev || !b2_dev || !out_dev || B > 65535) return L3D_ERR_INVALID;
  const dim3 grid((N + LN_POS - 1) / LN_POS, B);
  const int per = (D + LN_SLICES - 1) / LN_SLICES;           // widest channel slice of a thread
  if (per <= 64)
    layernorm_cm_kernel<64><<<grid, LN_POS * LN_SLICES, 0, (cudaStream_t)stream>>>(x_dev, a2_dev, b2_dev, eps, D, N, out_dev);
  else
    layernorm_cm_kernel<0><<<grid, LN_POS * LN_SLICES, 0, (cudaStream_t)stream>>>(x_dev, a2_dev, b2_dev, eps, D, N, out_dev);
  count_launch();
  L3D_LAUNCH_CHECK();
  return L3D_OK;
}
