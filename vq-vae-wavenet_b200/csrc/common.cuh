// Shared device-side helpers for the vqwn kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqwn {

// ---------------------------------------------------------------------------------------
// Loads/stores for data that other CTAs of the same (persistent) launch produce.  L1 is not
// coherent across SMs, so everything exchanged between CTAs is read at L2 (.cg).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_cg(float* p, float v) { __stcg(p, v); }

// ---------------------------------------------------------------------------------------
// Grid-wide barrier for cooperative (co-resident) launches.  A monotonically increasing
// 64-bit arrival counter: barrier number n (1-based) is passed when counter >= n * gridDim.x.
// Same structure as cooperative_groups::grid_group::sync (bar.sync, fence, atomic, spin,
// bar.sync) without the per-launch bookkeeping.
// ---------------------------------------------------------------------------------------
struct GridBarrier {
  unsigned long long* counter;   // counter[0]: arrivals (monotonic); counter[16]: released epoch (own 128 B line)
  unsigned long long epoch;      // barriers passed so far by this CTA (uniform across the grid)
  int* err;                      // bounded wait: a broken launch reports an error instead of hanging the GPU
  // split barrier: arrive() publishes this CTA's stores and counts it in; wait() blocks until every CTA
  // has arrived.  Work that does not depend on other CTAs' data can be placed between the two.
  __device__ __forceinline__ void arrive() {
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
      const unsigned long long target = epoch * (unsigned long long)gridDim.x;
      // release: this CTA's stores (ordered before by bar.sync) become visible before the arrival is counted
      unsigned long long prev;
      asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(prev) : "l"(counter) : "memory");
      if (prev + 1 == target) {
        // last arriver releases everybody; pollers never touch the arrival counter's line
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(counter + 16), "l"(epoch) : "memory");
      }
    }
  }
  __device__ __forceinline__ void wait() {
    if (threadIdx.x == 0) {
      unsigned long long* flag = counter + 16;
      unsigned long long v;
      long long spins = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= epoch) break;
        if (++spins > (1LL << 27)) {
          if (err) atomicExch(err, 3);
          break;
        }
        if ((spins & 0xFFFF) == 0 && err && *reinterpret_cast<volatile int*>(err) != 0) break;
      }
    }
    __syncthreads();
  }
  __device__ __forceinline__ void sync() {
    arrive();
    wait();
  }
};

// counter-based uniform generator used when the caller supplies no uniforms (sample mode):
// splitmix64 over (seed, step, stream) -> double in [0,1) with 53 random bits.
__device__ __forceinline__ double counter_uniform(uint64_t seed, uint64_t t, uint64_t b) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (t * 0x100000001B3ULL + b + 1ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// mu_law_encode float path (mu_law_ops.py:6-8) for arbitrary inputs (step / teacher-forced
// API).  The generation loop itself uses the host-built 257-entry LUT.
__device__ __forceinline__ float mu_law_encode_dev(float x, float mu, float inv_log1p_mu_is_unused) {
  (void)inv_log1p_mu_is_unused;
  x = fminf(fmaxf(x, -1.0f), 1.0f);
  const float s = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);
  // sign(x) * log1p(mu*|x|) / log1p(mu), evaluated left to right in float32
  return __fdiv_rn(__fmul_rn(s, log1pf(__fmul_rn(mu, fabsf(x)))), log1pf(mu));
}

}  // namespace vqwn
