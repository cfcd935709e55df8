// Hardware probe for the building blocks of the bf16 tensor-core generation kernel (sm_100a):
//   1. tcgen05.mma kind::f16, A and B from shared memory (K-major, no swizzle), D in TMEM
//   2. tcgen05.mma with A from TMEM (written with tcgen05.st as packed bf16 pairs)
//   3. cluster of 4: DSMEM stores + remote mbarrier arrive round trip
//   4. L2 flag hand-off latency between two CTAs (release store / acquire poll)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  return d;          // layout type 0 = no swizzle
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_LOOP;\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// element (r, k) of a K-major no-swizzle operand tile: 8x8 core matrices of 128 contiguous bytes,
// k-chunks (8 elements) LBO apart, 8-row groups SBO apart
__device__ __forceinline__ uint32_t tile_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (r >> 3) * sbo + (k >> 3) * lbo + (r & 7) * 16 + (k & 7) * 2;
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe_mma(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int K,
                                                    int a_in_tmem, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t LBO = 128, SBO = (K / 8) * 128;
  uint8_t* sA = smem;
  uint8_t* sB = a_in_tmem ? smem : smem + 128 * K * 2;
  if (!a_in_tmem) for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + tile_off(r, k, LBO, SBO)) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + tile_off(r, k, LBO, SBO)) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;          // columns [0, N)
  const uint32_t tmem_a = tmem + 128;    // columns [128, 128 + K/2)
  if (a_in_tmem) {
    // thread = row; pack (k even -> low half)
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K);
    for (int c = 0; c < K / 2; c += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = arow[c + j];
      const uint32_t addr = tmem_a + ((uint32_t)(warp * 32) << 16) + c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  long long t0 = clock64();
  if (tid == 0) {
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      if (a_in_tmem) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "r"(tmem_a + ks * 8), "l"(db), "r"(idesc), "r"(acc));
      } else {
        const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  // epilogue: thread = row
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    const uint32_t addr = tmem_d + ((uint32_t)(warp * 32) << 16) + c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

// ---------------------------------------------------------------- cluster DSMEM + remote mbarrier
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1) probe_dsmem(int iters, int payload_words, long long* cycles, int* errors) {
  __shared__ uint64_t bar;
  __shared__ __align__(16) uint32_t inbox[4][1024];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 4 * 128);   // every thread of every CTA arrives once per round
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  int err = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (uint32_t peer = 0; peer < 4; ++peer) {
      uint32_t remote_inbox, remote_bar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_inbox) : "r"(smem_u32(&inbox[rank][0])), "r"(peer));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote_bar) : "r"(smem_u32(&bar)), "r"(peer));
      for (int w = tid; w < payload_words; w += 128)
        asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(remote_inbox + 4 * w), "r"((uint32_t)(it * 131 + rank * 7 + w)) : "memory");
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote_bar) : "memory");
    }
    // wait for all 4 producers (acquire at cluster scope)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP2:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP2;\n\t}\n" :: "r"(smem_u32(&bar)), "r"((uint32_t)(it & 1)) : "memory");
    for (uint32_t src = 0; src < 4; ++src)
      for (int w = tid; w < payload_words; w += 128)
        if (inbox[src][w] != (uint32_t)(it * 131 + src * 7 + w)) err++;
    // all CTAs must have consumed before the next round overwrites
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = (t1 - t0) / iters;
  if (err) atomicAdd(errors, err);
}

// ---------------------------------------------------------------- L2 flag hand-off between CTAs
__global__ void __launch_bounds__(128, 1) probe_l2(int iters, int payload_floats, float* buf0, float* buf1,
                                                   unsigned* flag0, unsigned* flag1, long long* cycles, int* errors) {
  const int tid = threadIdx.x;
  const int me = blockIdx.x;   // 0 or 1 (other CTAs idle)
  if (me > 1) return;
  float* mine = me == 0 ? buf0 : buf1;
  float* theirs = me == 0 ? buf1 : buf0;
  unsigned* my_flag = me == 0 ? flag0 : flag1;
  unsigned* their_flag = me == 0 ? flag1 : flag0;
  int err = 0;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (me == 0) {
      for (int w = tid; w < payload_floats; w += 128) __stcg(mine + w, (float)(it + w));
      __syncthreads();
      if (tid == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(my_flag), "r"((unsigned)it) : "memory"); }
    }
    // wait for the other side's message `it`
    if (tid == 0) {
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(their_flag) : "memory"); } while (v < (unsigned)it);
    }
    __syncthreads();
    for (int w = tid; w < payload_floats; w += 128)
      if (__ldcg(theirs + w) != (float)(it + w)) err++;
    if (me == 1) {
      for (int w = tid; w < payload_floats; w += 128) __stcg(mine + w, (float)(it + w));
      __syncthreads();
      if (tid == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(my_flag), "r"((unsigned)it) : "memory"); }
    }
  }
  long long t1 = clock64();
  if (tid == 0) cycles[me] = (t1 - t0) / iters;   // one round trip = 2 hops
  if (err) atomicAdd(errors, err);
}


// ---------------------------------------------------------------- MMA latency structure
// nmma MMAs (M=128, K=16 each) round-robin over nacc independent accumulators (N columns each)
template <int N>
__global__ void __launch_bounds__(128, 1) probe_chain(int nmma, int nacc, int a_in_tmem, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t LBO = 128, SBO = 256;   // one K=16 slice, re-read by every MMA
  for (int i = tid; i < (128 + N) * 16 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (a_in_tmem) {
    uint32_t v = 0x3c003c00u;
    for (int c = 0; c < 8; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(tmem + ((uint32_t)(warp * 32) << 16) + 504 + c), "r"(v));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = make_desc(smem_u32(smem), LBO, SBO);
  const uint64_t db = make_desc(smem_u32(smem) + 128 * 16 * 2, LBO, SBO);
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    long long t0 = clock64(), t_issue = 0;
    if (tid == 0) {
      for (int i = 0; i < nmma; ++i) {
        const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
        const uint32_t acc = (i >= nacc) ? 1u : 0u;
        if (a_in_tmem)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                       :: "r"(d), "r"(tmem + 504), "l"(db), "r"(idesc), "r"(acc));
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
      }
      t_issue = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, rep & 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t1 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = t_issue - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

template <int N>
void run_chain(int nmma, int nacc, int ts) {
  long long* dC; CK(cudaMalloc(&dC, 64));
  size_t smem = (128 + N) * 16 * 2;
  probe_chain<N><<<1, 128, smem>>>(nmma, nacc, ts, dC);
  CK(cudaDeviceSynchronize());
  long long c[2]; CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
  printf("chain %s N=%3d nmma=%3d nacc=%2d: total=%5lld cycles (issue loop %4lld) -> %.1f cyc/mma\n", ts ? "TS" : "SS", N, nmma, nacc, c[0], c[1], (double)c[0] / nmma);
  cudaFree(dC);
}

// clean issue path: warp-uniform branch, elect.sync predicate, compile-time accumulator rotation
template <int N, int NACC, int NMMA>
__global__ void __launch_bounds__(128, 1) probe_chain2(int a_in_tmem, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * 16 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = make_desc(smem_u32(smem), 128, 256);
  const uint64_t db = make_desc(smem_u32(smem) + 128 * 16 * 2, 128, 256);
  for (int rep = 0; rep < 3; ++rep) {
    __syncthreads();
    long long t0 = clock64(), t_issue = 0;
    if (warp == 0) {
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(elected));
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        const uint32_t d = tmem + (uint32_t)((i % NACC) * N);
        const uint32_t acc = (i >= NACC) ? 1u : 0u;
        if (a_in_tmem)
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                       :: "r"(d), "r"(tmem + 504), "l"(db), "r"(idesc), "r"(acc), "r"(elected));
        else
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(elected));
      }
      t_issue = clock64();
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                   "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(&bar)), "r"(elected) : "memory");
    }
    mbar_wait(&bar, rep & 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t1 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = t_issue - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}

template <int N, int NACC, int NMMA>
void run_chain2(int ts) {
  long long* dC; CK(cudaMalloc(&dC, 64));
  size_t smem = (128 + N) * 16 * 2;
  probe_chain2<N, NACC, NMMA><<<1, 128, smem>>>(ts, dC);
  CK(cudaDeviceSynchronize());
  long long c[2]; CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
  printf("chain2 %s N=%3d nmma=%3d nacc=%2d: total=%5lld cycles (issue loop %4lld) -> %.1f cyc/mma\n", ts ? "TS" : "SS", N, NMMA, NACC, c[0], c[1], (double)c[0] / NMMA);
  cudaFree(dC);
}

// ---------------------------------------------------------------- TMEM load latency / throughput
template <int X>
__device__ __forceinline__ void ldtm(uint32_t addr, uint32_t* v);
template <>
__device__ __forceinline__ void ldtm<32>(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
}
template <>
__device__ __forceinline__ void ldtm<8>(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}

template <int X, int DEPTH>
__global__ void __launch_bounds__(128, 1) probe_ldtm(int iters, int nwarps, long long* cycles, unsigned* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      uint32_t v[DEPTH][X];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) ldtm<X>(tmem + (uint32_t)(((it * DEPTH + d) * X) & 511 & ~(X - 1)), v[d]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= v[d][j];
    }
  }
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  sink[tid] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s));
}

template <int X, int DEPTH>
void run_ldtm(int nwarps) {
  long long* dC; unsigned* dS; CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dS, 1024));
  const int iters = 256;
  probe_ldtm<X, DEPTH><<<1, 128>>>(iters, nwarps, dC, dS);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
  double per = (double)c / (iters * DEPTH);
  printf("ldtm 32x32b.x%d depth=%d warps=%d: %.1f cycles per load (%d B/warp) -> %.1f B/clk/SM\n", X, DEPTH, nwarps, per, X * 128,
         nwarps * X * 128.0 / per);
  cudaFree(dC); cudaFree(dS);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int N>
int run_mma(int K, int a_in_tmem) {
  std::vector<__nv_bfloat16> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K), ref(128 * N), out(128 * N);
  srand(1 + K + N);
  for (int i = 0; i < 128 * K; ++i) { float v = bf((rand() % 2001 - 1000) / 1000.0f); fA[i] = v; hA[i] = __float2bfloat16(v); }
  for (int i = 0; i < N * K; ++i) { float v = bf((rand() % 2001 - 1000) / 1000.0f); fB[i] = v; hB[i] = __float2bfloat16(v); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __nv_bfloat16 *dA, *dB; float* dD; long long* dC;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, out.size() * 4)); CK(cudaMalloc(&dC, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, out.size() * 4));
  size_t smem = (size_t)((a_in_tmem ? 0 : 128) + N) * K * 2;
  CK(cudaFuncSetAttribute(probe_mma<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long cyc = 0;
  for (int rep = 0; rep < 3; ++rep) {
    probe_mma<N><<<1, 128, smem>>>(dA, dB, dD, K, a_in_tmem, dC);
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < out.size(); ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  int ok = maxerr <= 1e-3 * maxref;
  printf("mma %s N=%d K=%d: max_err=%.3e (max_ref %.3f) cycles(issue..commit-wait)=%lld %s\n", a_in_tmem ? "TS" : "SS", N, K, maxerr,
         maxref, cyc, ok ? "PASS" : "FAIL");
  if (!ok) {
    printf("  sample out/ref: ");
    for (int i = 0; i < 8; ++i) printf("%.3f/%.3f ", out[i * N + i % N], ref[i * N + i % N]);
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return ok;
}

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  if (which == 0 || which == 1) {
    run_mma<16>(64, 0);
    run_mma<64>(64, 0);
    run_mma<16>(768, 0);
    run_mma<64>(512, 0);
    run_mma<128>(256, 0);
  }
  if (which == 0 || which == 2) {
    run_mma<16>(64, 1);
    run_mma<16>(768, 1);
    run_mma<64>(768, 1);
  }
  if (which == 0 || which == 5) {
    for (int n : {1, 2, 4, 8, 16, 48}) run_chain<16>(n, 1, 0);
    for (int a : {2, 4, 8, 16}) run_chain<16>(48, a, 0);
    run_chain<64>(48, 1, 0); run_chain<64>(48, 4, 0); run_chain<64>(48, 8, 0);
    run_chain<256>(16, 1, 0); run_chain<256>(16, 2, 0);
    run_chain<16>(48, 1, 1); run_chain<16>(48, 8, 1); run_chain<16>(48, 16, 1); run_chain<64>(48, 4, 1);
    run_chain<32>(48, 1, 0); run_chain<32>(48, 8, 0);
  }
  if (which == 0 || which == 6) {
    run_chain2<16, 1, 1>(0); run_chain2<16, 1, 4>(0); run_chain2<16, 1, 16>(0); run_chain2<16, 1, 48>(0);
    run_chain2<16, 4, 48>(0); run_chain2<16, 8, 48>(0); run_chain2<16, 16, 48>(0);
    run_chain2<64, 1, 48>(0); run_chain2<64, 4, 48>(0);
    run_chain2<256, 1, 16>(0); run_chain2<256, 2, 16>(0);
    run_chain2<16, 1, 48>(1); run_chain2<16, 8, 48>(1); run_chain2<64, 4, 48>(1);
  }
  if (which == 0 || which == 7) {
    run_ldtm<32, 1>(1); run_ldtm<32, 1>(4); run_ldtm<32, 2>(4); run_ldtm<32, 4>(4);
    run_ldtm<8, 1>(1); run_ldtm<8, 4>(4); run_ldtm<8, 8>(4);
  }
  if (which == 0 || which == 3) {
    long long* dC; int* dE;
    CK(cudaMalloc(&dC, 64 * 8)); CK(cudaMalloc(&dE, 4)); CK(cudaMemset(dE, 0, 4));
    for (int words : {32, 512, 1024}) {
      CK(cudaMemset(dE, 0, 4));
      probe_dsmem<<<4, 128>>>(200, words, dC, dE);
      CK(cudaDeviceSynchronize());
      long long c[4]; int e = 0;
      CK(cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&e, dE, 4, cudaMemcpyDeviceToHost));
      printf("dsmem all-to-all (4 CTAs, %d B per peer) + 2 cluster barriers: %lld cycles/round, errors=%d\n", words * 4, c[0], e);
    }
    int maxc = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(prop.multiProcessorCount / 4 * 4); cfg.blockDim = dim3(128);
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e4 = cudaOccupancyMaxActiveClusters(&maxc, probe_dsmem, &cfg);
    printf("max active clusters of 4 (128 thr, ~16KB smem): %d (%s)\n", maxc, cudaGetErrorString(e4));
  }
  if (which == 0 || which == 4) {
    float *b0, *b1; unsigned *f; long long* dC; int* dE;
    CK(cudaMalloc(&b0, 1 << 20)); CK(cudaMalloc(&b1, 1 << 20)); CK(cudaMalloc(&f, 1024)); CK(cudaMalloc(&dC, 64)); CK(cudaMalloc(&dE, 4));
    for (int pf : {32, 4096, 16384}) {
      CK(cudaMemset(f, 0, 1024)); CK(cudaMemset(dE, 0, 4));
      probe_l2<<<2, 128>>>(500, pf, b0, b1, f, f + 64, dC, dE);
      CK(cudaDeviceSynchronize());
      long long c[2]; int e = 0;
      CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&e, dE, 4, cudaMemcpyDeviceToHost));
      printf("L2 hand-off (%d B payload): %lld cycles per round trip (2 hops), errors=%d\n", pf * 4, c[0], e);
    }
  }
  printf("probe done\n");
  return 0;
}
