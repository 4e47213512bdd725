// Library-level entry points of libl3d_b200.so: version, error strings, launch counter and the
// HOST-buffer convenience calls (H2D -> kernel -> D2H on an internal stream).
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace l3d {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Grow-only device scratch for the *_host entry points (one per process, mutex-guarded).
struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;
  void* in = nullptr;  size_t in_cap = 0;
  void* out = nullptr; size_t out_cap = 0;
  int ensure(size_t in_bytes, size_t out_bytes) {
    cudaError_t e;
    if (!stream) { e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking); if (e) return (int)e; }
    if (in_bytes > in_cap) {
      if (in) cudaFree(in);
      in = nullptr; in_cap = 0;
      e = cudaMalloc(&in, in_bytes); if (e) return (int)e;
      in_cap = in_bytes;
    }
    if (out_bytes > out_cap) {
      if (out) cudaFree(out);
      out = nullptr; out_cap = 0;
      e = cudaMalloc(&out, out_bytes); if (e) return (int)e;
      out_cap = out_bytes;
    }
    return L3D_OK;
  }
};
static HostCtx g_host;
}  // namespace l3d

extern "C" int l3d_abi_version(void) { return 1; }

extern "C" const char* l3d_error_string(int code) {
  if (code == L3D_OK) return "ok";
  if (code == L3D_ERR_INVALID) return "l3d: invalid argument";
  if (code == L3D_ERR_UNSUPPORTED) return "l3d: shape not supported by this build";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "l3d: unknown error";
}

extern "C" uint64_t l3d_launch_count(void) {
  return l3d::g_launches.load(std::memory_order_relaxed);
}

extern "C" int l3d_knn_expansion_host(const float* x_host, int B, int N, int k, int64_t* idx_host) {
  if (!x_host || !idx_host || B < 0 || N < 1 || k < 1 || k > N) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  std::lock_guard<std::mutex> lock(l3d::g_host.mu);
  const size_t in_bytes = (size_t)B * 3 * N * sizeof(float);
  const size_t out_bytes = (size_t)B * N * k * sizeof(int64_t);
  int rc = l3d::g_host.ensure(in_bytes, out_bytes);
  if (rc) return rc;
  // Zero-copy output (L3D_KNN_HOST_ZEROCOPY=1, experiment): when idx_host is pinned, the kernel stores the
  // indices straight into host memory over PCIe, so the write-back overlaps the whole kernel instead of
  // following it.  Measured on the B200 box: see DESIGN.md §7 before enabling by default.
  static int zero_copy = -1;
  if (zero_copy < 0) {
    const char* ev = getenv("L3D_KNN_HOST_ZEROCOPY");
    zero_copy = (ev && ev[0] == '1') ? 1 : 0;
  }
  if (zero_copy) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, idx_host) == cudaSuccess && at.type == cudaMemoryTypeHost &&
        at.devicePointer) {
      cudaStream_t s = l3d::g_host.stream;
      cudaError_t ze = cudaMemcpyAsync(l3d::g_host.in, x_host, in_bytes, cudaMemcpyHostToDevice, s);
      if (ze) return (int)ze;
      rc = l3d_knn_expansion((const float*)l3d::g_host.in, B, N, k, (int64_t*)at.devicePointer, nullptr, s);
      if (rc) return rc;
      return (int)cudaStreamSynchronize(s);
    }
    cudaGetLastError();   // not a pinned buffer: fall through to the copy pipeline
  }
  // The call is PCIe-bound: 8*k bytes of int64 indices return per 12 bytes of input.  The batch is
  // cut into slices that ping-pong over two streams so the device-to-host copy of slice i overlaps
  // the kernel of slice i+1 (clouds are independent; each slice is a whole number of clouds).
  cudaError_t e;
  if (!l3d::g_host.stream2) {
    e = cudaStreamCreateWithFlags(&l3d::g_host.stream2, cudaStreamNonBlocking);
    if (e) return (int)e;
  }
  cudaStream_t st[2] = {l3d::g_host.stream, l3d::g_host.stream2};
  const int nslice = B >= 8 ? 4 : (B >= 2 ? 2 : 1);
  const float* din = (const float*)l3d::g_host.in;
  int64_t* dout = (int64_t*)l3d::g_host.out;
  int b0 = 0;
  for (int i = 0; i < nslice; ++i) {
    const int b1 = (int)((long)B * (i + 1) / nslice);
    const int nb = b1 - b0;
    if (nb > 0) {
      cudaStream_t s = st[i & 1];
      const size_t io = (size_t)b0 * 3 * N, oo = (size_t)b0 * N * k;
      e = cudaMemcpyAsync((void*)(din + io), x_host + io, (size_t)nb * 3 * N * sizeof(float),
                          cudaMemcpyHostToDevice, s);
      if (e) return (int)e;
      rc = l3d_knn_expansion(din + io, nb, N, k, dout + oo, nullptr, s);
      if (rc) return rc;
      e = cudaMemcpyAsync(idx_host + oo, dout + oo, (size_t)nb * N * k * sizeof(int64_t),
                          cudaMemcpyDeviceToHost, s);
      if (e) return (int)e;
    }
    b0 = b1;
  }
  e = cudaStreamSynchronize(st[0]);
  if (e) return (int)e;
  e = cudaStreamSynchronize(st[1]);
  return (int)e;
}

// Chamfer loss + both gradients from HOST buffers in one call (the "Chamfer fwd+bwd" half of the headline metric
// for callers that are not PyTorch): H2D of the two clouds, the fused loss forward and backward launches with
// dL/dloss = 1, D2H of the scalar and the two gradient clouds, one synchronisation.
extern "C" int l3d_chamfer_loss_fwd_bwd_host(const float* xyz1_host, const float* xyz2_host, int B, int n, int m,
                                             float* loss_host, float* grad1_host, float* grad2_host) {
  if (!xyz1_host || !xyz2_host || !loss_host || B < 1 || n < 1 || m < 1) return L3D_ERR_INVALID;
  std::lock_guard<std::mutex> lock(l3d::g_host.mu);
  const size_t b1 = (size_t)B * n * 3 * sizeof(float), b2 = (size_t)B * m * 3 * sizeof(float);
  const size_t ws = l3d_chamfer_ws_bytes(B, n, m);
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // out region: dist1 | dist2 | idx1 | idx2 | loss,one | ws | grad1 | grad2
  const size_t o_d1 = 0, o_d2 = o_d1 + up((size_t)B * n * 4), o_i1 = o_d2 + up((size_t)B * m * 4),
               o_i2 = o_i1 + up((size_t)B * n * 4), o_l = o_i2 + up((size_t)B * m * 4), o_ws = o_l + 256,
               o_g1 = o_ws + up(ws), o_g2 = o_g1 + up(b1), o_end = o_g2 + up(b2);
  int rc = l3d::g_host.ensure(up(b1) + up(b2), o_end);
  if (rc) return rc;
  cudaStream_t s = l3d::g_host.stream;
  unsigned char* in = (unsigned char*)l3d::g_host.in;
  unsigned char* out = (unsigned char*)l3d::g_host.out;
  float* x1 = (float*)in;
  float* x2 = (float*)(in + up(b1));
  cudaError_t e = cudaMemcpyAsync(x1, xyz1_host, b1, cudaMemcpyHostToDevice, s);
  if (e) return (int)e;
  e = cudaMemcpyAsync(x2, xyz2_host, b2, cudaMemcpyHostToDevice, s);
  if (e) return (int)e;
  float* loss = (float*)(out + o_l);
  static const float one = 1.0f;
  e = cudaMemcpyAsync(loss + 1, &one, sizeof(float), cudaMemcpyHostToDevice, s);
  if (e) return (int)e;
  rc = l3d_chamfer_loss_forward(x1, x2, B, n, m, (float*)(out + o_d1), (float*)(out + o_d2), (int32_t*)(out + o_i1),
                                (int32_t*)(out + o_i2), loss, out + o_ws, s);
  if (rc) return rc;
  if (grad1_host && grad2_host) {
    rc = l3d_chamfer_loss_backward(x1, x2, B, n, m, (const float*)(out + o_d1), (const float*)(out + o_d2),
                                   (const int32_t*)(out + o_i1), (const int32_t*)(out + o_i2), loss + 1,
                                   (float*)(out + o_g1), (float*)(out + o_g2), s);
    if (rc) return rc;
    e = cudaMemcpyAsync(grad1_host, out + o_g1, b1, cudaMemcpyDeviceToHost, s);
    if (e) return (int)e;
    e = cudaMemcpyAsync(grad2_host, out + o_g2, b2, cudaMemcpyDeviceToHost, s);
    if (e) return (int)e;
  }
  e = cudaMemcpyAsync(loss_host, loss, sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e) return (int)e;
  return (int)cudaStreamSynchronize(s);
}
