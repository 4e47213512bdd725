// Micro-benchmarks behind the kNN thread-per-row design: broadcast LDS.128 / LDS.64 throughput per
// SM and FFMA2 vs FFMA issue rate per scheduler on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>   // 0: LDS.128 broadcast, 1: LDS.64 broadcast x2, 2: LDS.128 lane-distinct (conflict-free)
__global__ void lds_kernel(unsigned long long* out, long long* cyc, int iters) {
  __shared__ __align__(16) unsigned long long buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i * 0x9e3779b97f4a7c15ull;
  __syncthreads();
  unsigned long long acc0 = 0, acc1 = 0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int base = (it & 15) * 64;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&buf[base + 2 * j]);
        acc0 ^= v.x; acc1 += v.y;
      } else if (MODE == 1) {
        acc0 ^= buf[base + 2 * j]; acc1 += buf[base + 2 * j + 1];
      } else {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&buf[(base + 2 * j + 2 * lane) & 4095]);
        acc0 ^= v.x; acc1 += v.y;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>   // 0: FFMA2 x 8 chains, 1: FFMA x 16 chains, 2: FFMA2 x 4 chains + FMNMX x 4 chains
__global__ void fma_kernel(float* out, long long* cyc, int iters, float a, float b) {
  const int lane = threadIdx.x & 31;
  long long t0, t1;
  float r = 0.f;
  if (MODE == 0) {
    unsigned long long x[8];
    const unsigned long long aa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int i = 0; i < 8; ++i) x[i] = aa + i + lane;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = f2_fma(x[i], aa, bb);
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) r += __uint_as_float((uint32_t)x[i]) + __uint_as_float((uint32_t)(x[i] >> 32));
  } else {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = a + i + lane;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
    t1 = clock64();
    for (int i = 0; i < 16; ++i) r += x[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

int main() {
  unsigned long long* out; long long* cyc; float* fout;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096); cudaMalloc(&fout, 1 << 20);
  long long h[64];
  const int iters = 256;
  const int warps_list[] = {1, 2, 4, 7, 8, 16};
  for (int mode = 0; mode < 3; ++mode) {
    for (int w : warps_list) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) lds_kernel<0><<<1, 32 * w>>>(out, cyc, iters);
        if (mode == 1) lds_kernel<1><<<1, 32 * w>>>(out, cyc, iters);
        if (mode == 2) lds_kernel<2><<<1, 32 * w>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(long long) * w, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx;
      const double n = (mode == 1 ? 2.0 : 1.0) * iters * 16.0 * w;
      printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"cycles\": %lld, \"sm_cycles_per_warp_instr\": %.3f}\n",
             mode == 0 ? "LDS.128 broadcast" : mode == 1 ? "LDS.64 broadcast" : "LDS.128 lane-distinct", w, mx, mx / n);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    for (int w : {1, 4, 5, 8, 16}) {     // 4 = one warp per scheduler, 5 = two on scheduler 0, 8 = two on each
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) fma_kernel<0><<<1, 32 * w>>>(fout, cyc, 1024, 1.0001f, 0.5f);
        else fma_kernel<1><<<1, 32 * w>>>(fout, cyc, 1024, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(long long) * w, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per_warp = 1024.0 * (mode == 0 ? 8 : 16);
      printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"cycles\": %lld, \"cycles_per_instr_of_one_warp\": %.3f, \"fma_lanes_per_clk_per_sm\": %.1f}\n",
             mode == 0 ? "FFMA2 (8 chains)" : "FFMA (16 chains)", w, mx, mx / per_warp,
             per_warp * w * 32 * (mode == 0 ? 2 : 1) / mx);
    }
  }
  return 0;
}
