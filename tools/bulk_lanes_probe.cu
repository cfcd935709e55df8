// Is the ~431-cycle cost of a cp.async.bulk (tools/r2_probe.cu) per issuing THREAD or per SM?  `lanes` warps of a CTA each
// run their own ring of `depth` slots (lane 0 issues); the CTA's total copies per cycle should scale with the lanes if the
// cost is the issuing thread's.  Also times the issuing thread around the two instructions of a request.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_lanes_probe tools/bulk_lanes_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* src, size_t region, int copy_bytes, int depth, int lanes, int iters,
                                                long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bars[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < lanes) {
    const uint8_t* base = src + ((size_t)(blockIdx.x % 16) * 4 + w) * region;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    long long t_exp = 0, t_cp = 0, t_wait = 0;
    const long long t0 = clock64();
    size_t off = 0;
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      const uint32_t bar = s32(&bars[w * 8 + s]);
      if (i >= depth) {
        const uint32_t par = (uint32_t)((i / depth - 1) & 1);
        uint32_t ok = 0;
        const long long w0 = clock64();
        if (mode == 0) { while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); }
        else { while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory"); }
        t_wait += clock64() - w0;
      }
      const long long a = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)copy_bytes) : "memory");
      const long long b = clock64();
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                   ::"r"(s32(sm + ((size_t)w * depth + s) * copy_bytes)), "l"(base + off), "r"(copy_bytes), "r"(bar), "l"(pol) : "memory");
      const long long c = clock64();
      t_exp += b - a; t_cp += c - b;
      off += copy_bytes;
      if (off + copy_bytes > region) off = 0;
    }
    for (int i = iters; i < iters + depth; ++i) {
      const int s = i % depth;
      const uint32_t bar = s32(&bars[w * 8 + s]);
      const uint32_t par = (uint32_t)((i / depth - 1) & 1);
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    }
    out[(blockIdx.x * 4 + w) * 3 + 0] = clock64() - t0;
    out[(blockIdx.x * 4 + w) * 3 + 1] = t_exp + (t_wait << 32);
    out[(blockIdx.x * 4 + w) * 3 + 2] = t_cp;
  }
}
int main() {
  const size_t region = 2304 * 1024;
  uint8_t* src;
  CK(cudaMalloc(&src, region * 64));
  CK(cudaMemset(src, 0, region * 64));
  long long* out;
  CK(cudaMalloc(&out, 148 * 4 * 3 * sizeof(long long)));
  std::vector<long long> h(148 * 4 * 3);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int g = 112;
  for (int mode : {0, 1}) for (int sz : {4096, 16384}) for (int depth : {2, 4, 8}) for (int lanes : {1, 2}) {
    if ((size_t)sz * depth * lanes > 200 * 1024) continue;
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) { probe<<<g, 128, 200 * 1024>>>(src, region, sz, depth, lanes, iters, out, mode); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; double e = 0, c = 0, wt = 0;
    for (int b = 0; b < g; ++b) for (int w = 0; w < lanes; ++w) { mx = std::max(mx, h[(b * 4 + w) * 3]); e += h[(b * 4 + w) * 3 + 1] & 0xffffffffll; wt += h[(b * 4 + w) * 3 + 1] >> 32; c += h[(b * 4 + w) * 3 + 2]; }
    printf("%s copy=%5d B depth=%d lanes=%d: %.0f cycles per copy per lane, %.1f B/clk/SM; issuing thread: wait %.0f, expect_tx %.0f, cp.async.bulk %.0f cycles\n",
           mode ? "test_wait" : "try_wait ", sz, depth, lanes, (double)mx / iters, (double)iters * sz * lanes / mx, wt / (g * lanes * iters), e / (g * lanes * iters), c / (g * lanes * iters));
  }
  return 0;
}
