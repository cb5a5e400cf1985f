// Micro-benchmark: tcgen05.ld throughput (TMEM -> registers) per SM as a function of the number of reading warps,
// and MUFU.EX2 / FMA-polynomial exp2 throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
// Output: cycles per 32x32b.x32 load (= 4 KB per warp-instruction) and the implied bytes / clk / SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: one ld then wait (latency-bound chain); mode 1: four lds in flight then one wait
__global__ void tmem_ld_kernel(int iters, int mode, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  uint32_t r0[32], r1[32], r2[32], r3[32];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      tmem_ld32(base, r0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r0[q];
    } else {
      tmem_ld32(base, r0);
      tmem_ld32(base + 32, r1);
      tmem_ld32(base + 64, r2);
      tmem_ld32(base + 96, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r0[q] ^ r1[q] ^ r2[q] ^ r3[q];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2p(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0551716648f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// poly_every: 0 = all MUFU, k = every k-th element on the FMA polynomial
template <int POLY_EVERY>
__global__ void exp_kernel(int iters, long long* cycles, float* sink) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * float(threadIdx.x + i);
  float sum = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = fmaf(v[i], 1.0001f, -0.25f);
      const float e = (POLY_EVERY > 0 && (i % POLY_EVERY) == POLY_EVERY - 1) ? ex2p(x) : ex2a(x);
      sum += e;
      v[i] = x * 0.5f;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (sum == 123.456f) sink[0] = sum;
}

int main() {
  long long* d_cycles; uint32_t* d_sink; float* d_fsink;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4); cudaMalloc(&d_fsink, 4);
  long long h[148];
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      tmem_ld_kernel<<<148, warps * 32>>>(iters, mode, d_cycles, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("tmem_ld mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_iter = double(h[0]) / iters;
      const int lds = mode == 0 ? 1 : 4;
      printf("tcgen05.ld.32x32b.x32 %s, %2d warps/SM: %.1f cycles per warp-iteration (%d ld) -> %.0f B/clk/SM\n",
             mode == 0 ? "ld+wait      " : "4 x ld + wait", warps, per_iter, lds, lds * 4096.0 * warps / per_iter);
    }
  for (int warps : {4, 8, 16}) {
    exp_kernel<0><<<148, warps * 32>>>(2000, d_cycles, d_fsink); cudaDeviceSynchronize();
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    const double a = double(h[0]) / 2000 / 16;
    exp_kernel<4><<<148, warps * 32>>>(2000, d_cycles, d_fsink); cudaDeviceSynchronize();
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    const double b = double(h[0]) / 2000 / 16;
    exp_kernel<2><<<148, warps * 32>>>(2000, d_cycles, d_fsink); cudaDeviceSynchronize();
    cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
    const double c = double(h[0]) / 2000 / 16;
    printf("exp2 per element-step, %2d warps/SM: all MUFU %.2f clk | 1 of 4 poly %.2f clk | 1 of 2 poly %.2f clk  (per warp; x warps/%d SMSP-sharing)\n",
           warps, a, b, c, 4);
  }
  return 0;
}
