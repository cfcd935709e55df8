// Probe: can 8 clusters of 16 CTAs (one CTA per SM, ~200 KB shared memory each) be co-resident on this GPU,
// and what does a "push 512 B to each of the 16 CTAs + cluster barrier" hand-off cost?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(unsigned addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// each CTA owns `slice` floats of a [csize*slice] vector; every iteration it pushes its slice (values depend on
// the previous iteration's full vector) to all CTAs of the cluster, then a cluster barrier
template <int CS>
__global__ void handoff_kernel(int iters, int slice_floats, float* out, long long* cycles, int* smid) {
  extern __shared__ __align__(16) float sm[];
  float* buf0 = sm;                         // [CS*slice]
  float* buf1 = sm + CS * slice_floats;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * CS * slice_floats; i += blockDim.x) sm[i] = 1.0f;
  if (tid == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smid[blockIdx.x] = (int)s; }
  __syncthreads();
  cluster_arrive(); cluster_wait();
  const int nvec = slice_floats / 4;        // float4 per slice
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float* src = (it & 1) ? buf1 : buf0;
    float* dst = (it & 1) ? buf0 : buf1;
    // thread (peer p, vec v): p = tid / nvec ... loop
    for (int w = tid; w < CS * nvec; w += blockDim.x) {
      const int p = w / nvec, v = w - p * nvec;
      float4 x = reinterpret_cast<const float4*>(src)[((rank + 1) % CS) * nvec + v];   // depends on a peer's slice
      x.x = x.x * 0.999f + 0.001f;
      const unsigned a = mapa(smem_u32(dst + rank * slice_floats + 4 * v), (unsigned)p);
      st_cluster_v4(a, x);
    }
    cluster_arrive();
    cluster_wait();
  }
  long long t1 = clock64();
  if (tid == 0) {
    cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x] = buf0[0] + buf1[1];
  }
}

template <int CS>
int run(int smem_bytes, int iters, int slice_floats) {
  auto kern = handoff_kernel<CS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int nclusters = 1; nclusters <= 148 / CS; ++nclusters) {
    cfg.gridDim = dim3(nclusters * CS);
    int maxc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
    if (nclusters == 1) printf("cluster size %d, smem %d B: cudaOccupancyMaxActiveClusters = %d (%s)\n", CS, smem_bytes, maxc, cudaGetErrorString(e));
  }
  const int ncl = 128 / CS;
  cfg.gridDim = dim3(ncl * CS);
  float* out; long long* cyc; int* smid;
  CK(cudaMalloc(&out, 4 * 256)); CK(cudaMalloc(&cyc, 8 * 256)); CK(cudaMalloc(&smid, 4 * 256));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, iters, slice_floats, out, cyc, smid));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long hc[256]; int hs[256];
  CK(cudaMemcpy(hc, cyc, 8 * ncl * CS, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs, smid, 4 * ncl * CS, cudaMemcpyDeviceToHost));
  long long mx = 0; for (int i = 0; i < ncl * CS; ++i) mx = hc[i] > mx ? hc[i] : mx;
  printf("  %d clusters x %d CTAs, slice %d B pushed to %d peers + barrier: %.0f cycles per hand-off (%.3f ms total, %d iters)\n",
         ncl, CS, slice_floats * 4, CS, (double)mx / iters, ms, iters);
  printf("  smid of cluster 0:");
  for (int i = 0; i < CS; ++i) printf(" %d", hs[i]);
  printf("\n");
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs, smem optin %zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  const int smem = 220 * 1024;
  if (run<16>(smem, 2000, 128)) return 1;     // 512 B slices (8 streams x 16 channels)
  if (run<16>(smem, 2000, 256)) return 1;     // 1 KB slices
  if (run<8>(smem, 2000, 128)) return 1;
  if (run<8>(smem, 2000, 512)) return 1;
  if (run<4>(smem, 2000, 1024)) return 1;
  return 0;
}
