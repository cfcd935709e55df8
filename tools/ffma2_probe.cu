// Probe: fp32 FMA issue rate per SM, scalar FFMA vs packed fma.rn.f32x2 (FFMA2), with 1..4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 65536)[0] = t1 - t0;
}

__global__ void ffma2_kernel(float* out, int iters, float a, float b) {
  unsigned long long acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
    acc[i] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
  }
  const unsigned long long aa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
  const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)(acc[i] & 0xffffffffu)) + __uint_as_float((unsigned)(acc[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 65536)[0] = t1 - t0;
}

int main() {
  float* out;
  CK(cudaMalloc(&out, (65536 + 16) * sizeof(float)));
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int which = 0; which < 2; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        if (which == 0) ffma_kernel<<<148, warps * 32>>>(out, iters, 0.999f, 0.001f);
        else ffma2_kernel<<<148, warps * 32>>>(out, iters, 0.999f, 0.001f);
        CK(cudaDeviceSynchronize());
      }
      long long cyc;
      CK(cudaMemcpy(&cyc, out + 65536, sizeof cyc, cudaMemcpyDeviceToHost));
      const double inst = (double)iters * 64;                  // per warp
      const double fma_per_clk_sm = inst * warps * 32 * (which ? 2 : 1) / (double)cyc;
      printf("%s warps/SM=%2d: %.2f cycles per warp-instruction per scheduler-warp, %.1f FMA/clk/SM\n",
             which ? "FFMA2" : "FFMA ", warps, (double)cyc / inst, fma_per_clk_sm);
    }
  }
  return 0;
}
