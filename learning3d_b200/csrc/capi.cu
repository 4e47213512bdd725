// Library-level entry points of libl3d_b200.so: version, error strings, launch counter and the
// HOST-buffer convenience calls (H2D -> kernel -> D2H on an internal stream).
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"

#include <atomic>
#include <mutex>

namespace l3d {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Grow-only device scratch for the *_host entry points (one per process, mutex-guarded).
struct HostCtx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
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
  cudaStream_t s = l3d::g_host.stream;
  cudaError_t e = cudaMemcpyAsync(l3d::g_host.in, x_host, in_bytes, cudaMemcpyHostToDevice, s);
  if (e) return (int)e;
  rc = l3d_knn_expansion((const float*)l3d::g_host.in, B, N, k, (int64_t*)l3d::g_host.out, nullptr, s);
  if (rc) return rc;
  e = cudaMemcpyAsync(idx_host, l3d::g_host.out, out_bytes, cudaMemcpyDeviceToHost, s);
  if (e) return (int)e;
  e = cudaStreamSynchronize(s);
  return (int)e;
}
