// Library-level entry points of libl3d_b200.so: version, error strings, launch counter and the
// HOST-buffer convenience calls (H2D -> kernel -> D2H on an internal stream).
#include "common.cuh"
#include "../../include/l3d_b200.h"
#include "launch_count.h"
#include "knn_matrix.h"

#include <atomic>
#include <omp.h>
#include <thread>
#include <cstdlib>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <climits>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

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

// Widening workers of l3d_knn_expansion_host.  An OpenMP parallel region per chunk costs a fork / join (~10 us) every
// time, which caps the useful thread count at 4 and leaves the call bound by the 10.5 MB of int64 the host has to write
// (profiles/r02/knn_host_path_sweep.txt).  These threads sleep on a condition variable between calls, are woken when a
// call starts (the wake-up overlaps the H2D copy and the kernel), spin on the `ready` counter while the copies land and
// each widen a fixed share of every chunk.  The calling thread is worker 0.
struct WidenPool {
  std::mutex mu;
  std::condition_variable cv;
  uint64_t gen = 0;                 // bumped once per call, under mu
  int nworkers = 1;                 // including the caller
  std::vector<std::thread> threads;
  // the job of the current call (written before gen is bumped)
  const unsigned short* src = nullptr;
  int64_t* dst = nullptr;
  long cuts[65] = {0};
  int ncuts = 0;
  std::atomic<int> ready{0};        // chunks whose copy has landed
  std::atomic<int> done{0};         // helper threads that have finished the current job
  std::atomic<int> abort{0};

  static void relax() {
#if defined(__x86_64__)
    _mm_pause();
#endif
  }
  void widen_share(int c, int w) {
    const long i0 = cuts[c], i1 = cuts[c + 1], n = i1 - i0;
    const long a = i0 + n * w / nworkers, z = i0 + n * (w + 1) / nworkers;
    const unsigned short* s = src;
    int64_t* d = dst;
    for (long t = a; t < z; ++t) d[t] = (int64_t)s[t];
  }
  void run_job(int w) {
    for (int c = 0; c < ncuts; ++c) {
      while (ready.load(std::memory_order_acquire) <= c) relax();
      if (abort.load(std::memory_order_relaxed)) break;
      widen_share(c, w);
    }
  }
  void helper(int w) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return gen != seen; });
        seen = gen;
      }
      run_job(w);
      done.fetch_add(1, std::memory_order_release);
    }
  }
  void start(int n) {               // once, under g_host.mu
    nworkers = n < 1 ? 1 : n;
    for (int w = 1; w < nworkers; ++w) {
      threads.emplace_back([this, w] { helper(w); });
      threads.back().detach();      // they live as long as the process; the pool itself is never destroyed
    }
  }
  void begin() {                    // job fields are set: wake the helpers
    ready.store(0, std::memory_order_relaxed);
    done.store(0, std::memory_order_relaxed);
    abort.store(0, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu);
      ++gen;
    }
    cv.notify_all();
  }
  void finish(bool failed) {        // caller: release everything on failure, then wait for the helpers
    if (failed) {
      abort.store(1, std::memory_order_relaxed);
      ready.store(INT_MAX, std::memory_order_release);
    }
    while (done.load(std::memory_order_acquire) < nworkers - 1) relax();
  }
};
static WidenPool* g_pool = nullptr;
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

// knn() from HOST buffers.  The call is PCIe-bound on its OUTPUT: 8*k bytes of int64 indices return per 12 bytes of
// input (5.2 MB per C2 batch, ~100 us of D2H at the ~50 GB/s a pinned copy reaches), three times the kernel.  Every
// index is < N (N <= 65536 on this path; larger clouds return int64 directly), so the device writes uint16, 2 bytes per index cross the bus (1.3 MB) into an internal pinned
// staging buffer, and the host widens them to the caller's int64 array with a few OpenMP threads while the next
// slice is still in flight.  The batch is cut into (by default two) slices on two streams (clouds are independent);
// more slices overlap better on paper but every slice costs four driver calls, and those ~4 us each are what the
// call is bound by once the bytes are narrow (C2: 171 us with int64 over the bus and 4 slices -> 104 us).
extern "C" int l3d_knn_expansion_host(const float* x_host, int B, int N, int k, int64_t* idx_host) {
  if (!x_host || !idx_host || B < 0 || N < 1 || k < 1 || k > N) return L3D_ERR_INVALID;
  if (B == 0) return L3D_OK;
  std::lock_guard<std::mutex> lock(l3d::g_host.mu);
  const size_t in_bytes = (size_t)B * 3 * N * sizeof(float);
  const size_t n_idx = (size_t)B * N * k;
  cudaError_t e;
  if (N > 65536) {
    // an index no longer fits the 16-bit wire format: int64 indices straight into the caller's buffer, one launch
    int rc64 = l3d::g_host.ensure(in_bytes, n_idx * sizeof(int64_t));
    if (rc64) return rc64;
    cudaStream_t s = l3d::g_host.stream;
    e = cudaMemcpyAsync(l3d::g_host.in, x_host, in_bytes, cudaMemcpyHostToDevice, s);
    if (e) return (int)e;
    rc64 = l3d_knn_expansion((const float*)l3d::g_host.in, B, N, k, (int64_t*)l3d::g_host.out, nullptr, (void*)s);
    if (rc64) return rc64;
    e = cudaMemcpyAsync(idx_host, l3d::g_host.out, n_idx * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e) return (int)e;
    return (int)cudaStreamSynchronize(s);
  }
  int rc = l3d::g_host.ensure(in_bytes, n_idx * sizeof(unsigned short));
  if (rc) return rc;
  // grow-only pinned staging buffer for the narrow indices
  static unsigned short* stage = nullptr;
  static size_t stage_cap = 0;
  if (n_idx > stage_cap) {
    if (stage) cudaFreeHost(stage);
    stage = nullptr; stage_cap = 0;
    e = cudaHostAlloc((void**)&stage, n_idx * sizeof(unsigned short), cudaHostAllocDefault);
    if (e) return (int)e;
    stage_cap = n_idx;
  }
  if (!l3d::g_host.stream2) {
    e = cudaStreamCreateWithFlags(&l3d::g_host.stream2, cudaStreamNonBlocking);
    if (e) return (int)e;
  }
  // Work plan: `nslice` kernel slices (clouds are independent) on two streams, each slice's indices returned in
  // `nchunk` device-to-host copies with an event after every copy, so the host widens chunk c while later chunks are
  // still on the bus or still being computed (profiles/r02/knn_host_path_sweep.txt has the sweeps of both knobs).
  constexpr int MAX_SLICES = 8, MAX_CHUNKS = 8;
  static cudaEvent_t ev[MAX_SLICES * MAX_CHUNKS] = {};
  cudaStream_t st[2] = {l3d::g_host.stream, l3d::g_host.stream2};
  static int slice_cap = 0, chunk_cfg = 0, nthreads = 0, use_pool = 1;
  if (slice_cap == 0) {
    const char* ev_s = getenv("L3D_HOST_SLICES");
    slice_cap = ev_s ? atoi(ev_s) : 2;
    if (slice_cap < 1) slice_cap = 1;
    if (slice_cap > MAX_SLICES) slice_cap = MAX_SLICES;
    const char* ev_c = getenv("L3D_HOST_D2H_CHUNKS");
    chunk_cfg = ev_c ? atoi(ev_c) : 1;
    if (chunk_cfg < 1) chunk_cfg = 1;
    if (chunk_cfg > MAX_CHUNKS) chunk_cfg = MAX_CHUNKS;
    const char* ev_p = getenv("L3D_HOST_POOL");
    use_pool = ev_p ? atoi(ev_p) : 1;
    const char* ev_t = getenv("L3D_HOST_THREADS");
    // measured at C2 (profiles/r02/knn_host_path_sweep.txt): OpenMP 4 threads 108.7 us; pool 4 / 8 / 16 threads
    // 109.4 / 94.6 / 89.1 us.  The helpers spin while a call is in flight, so the default leaves half of the machine
    // to the other ranks of a node (torchrun exports LOCAL_WORLD_SIZE; OMP_NUM_THREADS = 1 there is NOT a clamp here).
    const int hw = (int)std::thread::hardware_concurrency();
    const char* ev_w = getenv("LOCAL_WORLD_SIZE");
    const int lws = (ev_w && atoi(ev_w) > 0) ? atoi(ev_w) : 1;
    int dflt = use_pool ? 16 : 4;
    if (use_pool && hw > 0 && dflt > hw / (2 * lws)) dflt = hw / (2 * lws) < 4 ? 4 : hw / (2 * lws);
    nthreads = ev_t ? atoi(ev_t) : dflt;
    if (nthreads < 1) nthreads = 1;
    if (hw > 0 && nthreads > hw) nthreads = hw;
    if (nthreads > 32) nthreads = 32;
    if (use_pool) {
      l3d::g_pool = new l3d::WidenPool();
      l3d::g_pool->start(nthreads);
    }
  }
  int nslice = B >= 16 ? 8 : (B >= 8 ? 4 : (B >= 2 ? 2 : 1));
  if (nslice > slice_cap) nslice = slice_cap;
  const float* din = (const float*)l3d::g_host.in;
  unsigned short* dout = (unsigned short*)l3d::g_host.out;
  int bounds[MAX_SLICES + 1];
  for (int i = 0; i <= nslice; ++i) bounds[i] = (int)((long)B * i / nslice);
  // copy plan first (it depends on sizes only), so that the widening workers can be woken before any CUDA call
  long cuts[MAX_SLICES * MAX_CHUNKS + 1];
  int chunks_of[MAX_SLICES];
  int ncuts = 0;
  cuts[0] = 0;
  for (int i = 0; i < nslice; ++i) {
    const long cnt = (long)(bounds[i + 1] - bounds[i]) * N * k, oo = (long)bounds[i] * N * k;
    chunks_of[i] = (cnt >= (long)chunk_cfg * 65536) ? chunk_cfg : 1;      // small outputs: one copy
    for (int c = 0; c < chunks_of[i]; ++c) cuts[++ncuts] = oo + cnt * (c + 1) / chunks_of[i];
  }
  l3d::WidenPool* pool = use_pool ? l3d::g_pool : nullptr;
  if (pool) {
    pool->src = stage; pool->dst = idx_host; pool->ncuts = ncuts;
    for (int i = 0; i <= ncuts; ++i) pool->cuts[i] = cuts[i];
    pool->begin();
  }
  int fail = 0;
  int cut = 0;
  for (int i = 0; i < nslice && !fail; ++i) {
    const int b0 = bounds[i], nb = bounds[i + 1] - b0;
    cudaStream_t s = st[i & 1];
    const size_t io = (size_t)b0 * 3 * N;
    if (nb > 0) {
      e = cudaMemcpyAsync((void*)(din + io), x_host + io, (size_t)nb * 3 * N * sizeof(float), cudaMemcpyHostToDevice, s);
      if (e) { fail = (int)e; break; }
      rc = l3d::knn_expansion_u16(din + io, nb, N, k, dout + (size_t)b0 * N * k, s);
      if (rc) { fail = rc; break; }
    }
    for (int c = 0; c < chunks_of[i]; ++c, ++cut) {
      const long c0 = cuts[cut], c1 = cuts[cut + 1];
      if (c1 > c0) {
        e = cudaMemcpyAsync(stage + c0, dout + c0, (size_t)(c1 - c0) * sizeof(unsigned short), cudaMemcpyDeviceToHost, s);
        if (e) { fail = (int)e; break; }
      }
      if (!ev[cut]) { e = cudaEventCreateWithFlags(&ev[cut], cudaEventDisableTiming); if (e) { fail = (int)e; break; } }
      e = cudaEventRecord(ev[cut], s);
      if (e) { fail = (int)e; break; }
    }
  }
  // widen chunk by chunk as the copies land
  for (int i = 0; i < ncuts && !fail; ++i) {
    e = cudaEventSynchronize(ev[i]);
    if (e) { fail = (int)e; break; }
    if (pool) {
      pool->ready.store(i + 1, std::memory_order_release);
      pool->widen_share(i, 0);
    } else {
      const long i0 = cuts[i], i1 = cuts[i + 1];
      const unsigned short* src = stage;
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (long c = i0 / 4096; c < (i1 + 4095) / 4096; ++c) {
        const long a = c * 4096 < i0 ? i0 : c * 4096, z = (c + 1) * 4096 > i1 ? i1 : (c + 1) * 4096;
        for (long t = a; t < z; ++t) idx_host[t] = (int64_t)src[t];
      }
    }
  }
  if (pool) pool->finish(fail != 0);
  if (fail) {
    cudaStreamSynchronize(st[0]);
    cudaStreamSynchronize(st[1]);
    return fail;
  }
  return L3D_OK;
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
