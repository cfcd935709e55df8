// Round-2 hardware probes for the tensor-core generation kernel (sm_100a).  Timing only (operands are zeros).
//   1. sustained cp.async.bulk L2 -> shared memory throughput per SM, by number of SMs pulling, copy size and depth
//   2. tcgen05.mma kind::f16 cost per instruction: M = 64 / 128, N = 32 / 64 / 128, no-swizzle planes vs SWIZZLE_128B
//   3. all-gather hand-off inside a 16-CTA cluster: per-thread st.async vs bulk shared::cta -> shared::cluster copies
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/r2_probe tools/r2_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ 1. bulk copy throughput
__global__ void __launch_bounds__(128, 1) probe_bulk(const uint8_t* src, size_t region_bytes, int nregions, int copy_bytes,
                                                     int depth, int iters, int keep, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (size_t)(blockIdx.x % nregions) * region_bytes;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const long long t0 = clock64();
    size_t off = 0;
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      if (i >= depth) mbar_wait(&bars[s], (uint32_t)((i / depth - 1) & 1));
      mbar_expect(&bars[s], (uint32_t)copy_bytes);
      if (keep)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(sm + (size_t)s * copy_bytes)), "l"(base + off), "r"(copy_bytes), "r"(smem_u32(&bars[s])), "l"(pol) : "memory");
      else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm + (size_t)s * copy_bytes)), "l"(base + off), "r"(copy_bytes), "r"(smem_u32(&bars[s])) : "memory");
      off += copy_bytes;
      if (off + copy_bytes > region_bytes) off = 0;
    }
    for (int i = iters; i < iters + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) mbar_wait(&bars[s], (uint32_t)((i / depth - 1) & 1));
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

// ------------------------------------------------------------------------------------------------ 2. MMA cost
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ uint64_t desc_swz128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B
  d |= 1ull << 46;
  d |= 2ull << 61;                        // SWIZZLE_128B
  return d;
}
// M rows x (nmma * 16) K, N columns.  layout 0: K-major planes (plane = 8 k of every row), 1: SWIZZLE_128B K-blocks of 64
__global__ void __launch_bounds__(128, 1) probe_mma(int M, int N, int nmma, int layout, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
  const uint32_t a_base = smem_u32(sm), b_base = smem_u32(sm) + 128 * 1024;
  long long best = 1ll << 60, best_issue = 0;
  uint32_t ph = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    long long t1 = 0;
    if (tid == 0) {
      // descriptors advance by a constant per K chunk: the issue loop is two 64-bit adds + the MMA (a dependent chain of
      // integer instructions in ONE thread costs ~4.5 cycles each and would otherwise hide the pipe cost)
      uint64_t da0, db0, sa, sb;
      if (layout == 0) {
        da0 = desc_noswz(a_base, (uint32_t)M * 16u, 128); sa = (2u * (uint32_t)M * 16u) >> 4;
        db0 = desc_noswz(b_base, (uint32_t)N * 16u, 128); sb = (2u * (uint32_t)N * 16u) >> 4;
      } else if (layout == 2) {
        da0 = desc_noswz(a_base, 128, 32 * 128); sa = 256 >> 4;
        db0 = desc_noswz(b_base, 128, 32 * 128); sb = 256 >> 4;
      } else {
        da0 = desc_swz128(a_base); sa = 32 >> 4;      // within one 64-element K block (timing only)
        db0 = desc_swz128(b_base); sb = 32 >> 4;
      }
      uint64_t da = da0, db = db0;
#pragma unroll 4
      for (int ks = 0; ks < nmma; ++ks) {
        const uint32_t acc = ks > 0 ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        da += sa; db += sb;
        if ((ks & 3) == 3 && layout == 1) { da = da0; db = db0; }
        if ((ks & 15) == 15) { da = da0; db = db0; }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, ph);
    ph ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t2 = clock64();
    if (tid == 0 && t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
  }
  if (tid == 0) { out[0] = best; out[1] = best_issue; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ------------------------------------------------------------------------------------------------ 3. cluster all-gather
// every CTA pushes a slice of `slice` bytes to the same offset (rank * slice) of all CS CTAs, waits for its own CS
// slices, repeats.  method 0: 128 threads, st.async v4 (thread = 16-byte chunk, loops over peers); method 1: lanes of
// warp 0 issue one bulk shared::cta -> shared::cluster copy per peer; method 2: as 1 but the 16 copies come from 16
// different warps' lane 0 (512 threads)
__global__ void probe_gather(int CS, int slice, int method, int iters, long long* out, uint8_t* gbuf) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  uint8_t* rx = sm;                 // CS * slice
  uint8_t* tx = sm + 32 * 1024;     // slice
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = tid; i < 40 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = (uint32_t)i;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect(&bar, (uint32_t)(CS * slice));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cl_sync();
  const uint32_t rx_u = smem_u32(rx) + rank * (uint32_t)slice, bar_u = smem_u32(&bar), tx_u = smem_u32(tx);
  uint32_t ph = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (method == 0) {
      for (int c = tid; c < slice / 16; c += blockDim.x) {
        const float4 x = *reinterpret_cast<const float4*>(tx + c * 16);
        for (int k = 0; k < CS; ++k) {
          const uint32_t pr = (rank + 1 + k) % CS;
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                       ::"r"(mapa(rx_u + c * 16, pr)), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w), "r"(mapa(bar_u, pr)) : "memory");
        }
      }
    } else if (method == 1) {
      if (warp == 0 && lane < CS) {
        const uint32_t pr = (rank + 1 + lane) % CS;
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa(rx_u, pr)), "r"(tx_u), "r"(slice), "r"(mapa(bar_u, pr)) : "memory");
      }
    } else if (method == 3) {
      // through L2: slice -> global, then ONE multicast bulk copy global -> the same offset of every CTA of the cluster
      uint8_t* g = gbuf + ((size_t)blockIdx.x * 2 + (it & 1)) * 4096;
      for (int c = tid; c < slice / 16; c += blockDim.x)
        *reinterpret_cast<float4*>(g + c * 16) = *reinterpret_cast<const float4*>(tx + c * 16);
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        const uint16_t mask = (uint16_t)((1u << CS) - 1u);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(rx_u), "l"(g), "r"(slice), "r"(bar_u), "h"(mask) : "memory");
      }
    } else {
      if (lane == 0 && warp < CS) {
        const uint32_t pr = (rank + 1 + warp) % CS;
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(mapa(rx_u, pr)), "r"(tx_u), "r"(slice), "r"(mapa(bar_u, pr)) : "memory");
      }
    }
    // receive: one thread waits and re-arms, everybody follows through the CTA barrier
    if (tid == 0) {
      mbar_wait(&bar, ph);
      mbar_expect(&bar, (uint32_t)(CS * slice));
    }
    ph ^= 1u;
    __syncthreads();
  }
  const long long t1 = clock64();
  cl_sync();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------------ 2b. who pays the ~64 cycles?
// `nw` warps (one elected thread each) issue nmma / nw MMAs each into their own accumulator (M = 128, N = 32, planes layout)
// style 0: descriptor advanced by 64-bit adds; style 1: the 16-bit address field patched into a constant high word
__global__ void __launch_bounds__(128, 1) probe_mma_multi(int nw, int nmma, int style, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, (uint32_t)nw);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int M = 128, N = 32;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (((uint32_t)M >> 4) << 24);
  const uint32_t a_base = smem_u32(sm), b_base = smem_u32(sm) + 128 * 1024;
  long long best = 1ll << 60;
  uint32_t ph = 0;
  for (int r = 0; r < 20; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nw && lane == 0) {
      const int per = nmma / nw;
      uint64_t da = desc_noswz(a_base + (uint32_t)warp * 16384u, (uint32_t)M * 16u, 128);
      uint64_t db = desc_noswz(b_base + (uint32_t)warp * 4096u, (uint32_t)N * 16u, 128);
      const uint32_t dcol = tmem + 32u * (uint32_t)warp;
      if (style == 0) {
        for (int ks = 0; ks < per; ++ks) {
          const uint32_t acc = ks > 0 ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          da += (2u * M * 16u) >> 4; db += (2u * N * 16u) >> 4;
        }
      } else {
        // fully unrolled, descriptors are compile-time offsets from two base registers
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          if (ks < per) {
            const uint64_t da_k = da + (uint64_t)(ks * ((2 * 128 * 16) >> 4));
            const uint64_t db_k = db + (uint64_t)(ks * ((2 * 32 * 16) >> 4));
            if (ks == 0)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 1;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                           ::"r"(dcol), "l"(da_k), "l"(db_k), "r"(idesc) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 1;\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                           ::"r"(dcol), "l"(da_k), "l"(db_k), "r"(idesc) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, ph);
    ph ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t2 = clock64();
    if (tid == 0 && t2 - t0 < best) best = t2 - t0;
  }
  if (tid == 0) out[0] = best;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ------------------------------------------------------------------------------------------------ 2c. MMA chain under load
// 64 MMAs (M = 128, N = 64, block layout LBO 128 / SBO 256) in groups of 4 with a commit per group (cgrp) or one commit at
// the end, while warp 3 streams `bg` concurrent 16 KB bulk copies from L2 into other shared memory (0 = none)
__global__ void __launch_bounds__(128, 1) probe_mma_load(const uint8_t* src, int cgrp, int bg, int N, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar, gbar[4], lbar[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&gbar[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&lbar[i], 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 3 && lane == 0 && bg > 0) {
    // background stream: bg copies in flight, 16 KB each, into smem 128 KB..
    long long n = 0;
    size_t off = 0;
    while (!stop) {
      const int s = (int)(n % bg);
      if (n >= bg) mbar_wait(&lbar[s], (uint32_t)((n / bg - 1) & 1));
      mbar_expect(&lbar[s], 16384);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(sm + 128 * 1024 + s * 16384)), "l"(src + off), "r"(16384), "r"(smem_u32(&lbar[s])) : "memory");
      off = (off + 16384) % (4u << 20);
      n += 1;
    }
    for (long long i = n; i < n + bg; ++i) { const int s = (int)(i % bg); if (i >= bg) mbar_wait(&lbar[s], (uint32_t)((i / bg - 1) & 1)); }
    out[2] = n;
  }
  if (warp == 0) {
    long long best = 1ll << 60;
    uint32_t ph = 0;
    for (int r = 0; r < 10; ++r) {
      const long long t0 = clock64();
      if (lane == 0) {
        for (int g = 0; g < 16; ++g) {
          uint64_t da = 0, db = 0;
          da |= (uint64_t)(((smem_u32(sm) + (g & 3) * 16384) & 0x3FFFF) >> 4); da |= (uint64_t)(128 >> 4) << 16; da |= (uint64_t)(256 >> 4) << 32; da |= 1ull << 46;
          db |= (uint64_t)(((smem_u32(sm) + 65536 + (g & 3) * 8192) & 0x3FFFF) >> 4); db |= (uint64_t)(128 >> 4) << 16; db |= (uint64_t)(256 >> 4) << 32; db |= 1ull << 46;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t acc = (g | j) ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            da += 4096 >> 4; db += 2048 >> 4;
          }
          if (cgrp) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&gbar[g & 3])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long t2 = clock64();
      if (t2 - t0 < best) best = t2 - t0;
    }
    if (lane == 0) { out[0] = best; stop = 1; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  long long* out;
  CK(cudaMalloc(&out, 1024 * sizeof(long long)));
  std::vector<long long> h(1024);

  // ---- 1
  {
    const size_t region = 4608 * 1024;      // 4.5 MB per cluster CTA and step
    const int nregions = 16;
    uint8_t* src;
    CK(cudaMalloc(&src, region * nregions));
    CK(cudaMemset(src, 0, region * nregions));
    CK(cudaFuncSetAttribute(probe_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int grids[] = {};
    const int sizes[] = {32768};
    const int depths[] = {2, 4};
    for (int g : grids) for (int sz : sizes) for (int d : depths) {
      if ((size_t)sz * d > 200 * 1024) continue;
      const int iters = (int)(2 * region / sz);
      for (int rep = 0; rep < 2; ++rep) {
        probe_bulk<<<g, 128, 200 * 1024>>>(src, region, nregions, sz, d, iters, 1, out);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), out, g * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0, mn = 1ll << 60;
      for (int i = 0; i < g; ++i) { mx = std::max(mx, h[i]); mn = std::min(mn, h[i]); }
      printf("bulk  SMs=%3d copy=%5d B depth=%d: %.1f .. %.1f B/clk/SM (slowest .. fastest CTA)\n", g, sz, d,
             (double)iters * sz / mx, (double)iters * sz / mn);
    }
    CK(cudaFree(src));
  }
  // ---- 2
  {
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int Ms[] = {128};
    const int Ns[] = {32};
    const int nm[] = {16, 48};
    for (int layout = 0; layout < 0; ++layout) for (int M : Ms) for (int N : Ns) for (int n : nm) {
      probe_mma<<<1, 128, 200 * 1024>>>(M, N, n, layout, 20, out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      printf("mma   %s M=%3d N=%3d nmma=%2d: total %5lld (issue %5lld) -> %.1f cyc/mma\n", layout == 0 ? "planes" : (layout == 1 ? "swz128" : "rowgrp"), M, N, n,
             h[0], h[1], (double)h[0] / n);
    }
  }
  // ---- 2b
  {
    CK(cudaFuncSetAttribute(probe_mma_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int style = 0; style < 2; ++style) for (int nw : {1, 2, 4}) for (int n : {16}) {
      probe_mma_multi<<<1, 128, 200 * 1024>>>(nw, n, style, out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, sizeof(long long), cudaMemcpyDeviceToHost));
      printf("mma-multi style=%d issuing warps=%d total mma=%d: %lld cycles -> %.1f cyc/mma\n", style, nw, n, h[0], (double)h[0] / n);
    }
  }
  // ---- 2c
  {
    uint8_t* src2;
    CK(cudaMalloc(&src2, 8u << 20));
    CK(cudaMemset(src2, 0, 8u << 20));
    CK(cudaFuncSetAttribute(probe_mma_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int N : {32, 64}) for (int cgrp = 0; cgrp < 2; ++cgrp) for (int bg : {0, 2, 4}) {
      probe_mma_load<<<1, 128, 200 * 1024>>>(src2, cgrp, bg, N, out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, 3 * sizeof(long long), cudaMemcpyDeviceToHost));
      printf("mma-load N=%d commit-per-4=%d background copies in flight=%d: 64 mma in %lld cycles -> %.1f cyc/mma\n", N, cgrp, bg, h[0], (double)h[0] / 64);
    }
  }
  // ---- 3
  {
    CK(cudaFuncSetAttribute(probe_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(probe_gather, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    uint8_t* gbuf;
    CK(cudaMalloc(&gbuf, 1 << 20));
    const int CSs[] = {};
    const int slices[] = {1024};
    for (int CS : CSs) for (int slice : slices) for (int method = 0; method < 4; ++method) for (int ncl : {1, 4}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(CS * ncl);
      cfg.blockDim = dim3(method == 2 ? 512 : 128);
      cfg.dynamicSmemBytes = 200 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      const int iters = 2000;
      CK(cudaLaunchKernelEx(&cfg, probe_gather, CS, slice, method, iters, out, gbuf));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), out, CS * ncl * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < CS * ncl; ++i) mx = std::max(mx, h[i]);
      printf("gather CS=%2d slice=%4d B method=%d (%s) clusters=%d: %.0f cycles per all-gather round\n", CS, slice, method,
             method == 0 ? "st.async" : (method == 1 ? "bulk, 1 warp" : (method == 2 ? "bulk, 16 warps" : "L2 + multicast")), ncl, (double)mx / iters);
    }
  }
  printf("probe done\n");
  return 0;
}
