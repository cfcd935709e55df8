// VQ bottleneck kernels (reference: model.py:57-74, model.py:19-27, decoder_ops.py:39-43).
//
// vq_direct_kernel: exact float32 direct-form squared distance (z - e)^2 summed over D in
// index order, lowest-index argmin, fused gather + straight-through z_q = z + (e_k - z) and
// (optionally) the speaker-row concat into the [B,F,D+spk] condition buffer.
//
// Data layout: z [N,D] row-major (N = B*F encoder frames), codebook [K,D] row-major.
// Each thread owns ONE codebook row in registers for the whole launch (K <= 512 threads per
// CTA), the CTA walks vectors VB at a time; z rows are staged in shared memory and read as
// broadcast float4.  Algorithmic bytes per vector: 4*D in + 4*D (or 4*(D+spk)) out + 8 idx.
#pragma once
#include "common.cuh"

namespace vqwn {

template <int D, int VB>
__global__ void __launch_bounds__(512, 1)
vq_direct_kernel(const float* __restrict__ z, const float* __restrict__ E, int K, long long N,
                 long long* __restrict__ idx_out, float* __restrict__ zq_out, int out_stride,
                 const float* __restrict__ spk_table, const int* __restrict__ spk_idx,
                 int spk_dim, int F, int out_code, int expanded = 0) {
  __shared__ __align__(16) float zs[VB][D];
  __shared__ float wmin_d[VB][16];
  __shared__ int wmin_k[VB][16];
  __shared__ int best_k[VB];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const bool has_code = tid < K;

  float e[D];
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    float4 v = has_code ? __ldg(reinterpret_cast<const float4*>(E + (size_t)tid * D + d))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    e[d] = v.x; e[d + 1] = v.y; e[d + 2] = v.z; e[d + 3] = v.w;
  }

  const long long nblocks = (N + VB - 1) / VB;
  for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const long long v0 = blk * VB;
    // stage VB vectors
    if (tid < VB * D / 4) {
      const int v = tid / (D / 4), q = tid % (D / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + v < N) val = __ldg(reinterpret_cast<const float4*>(z + (size_t)(v0 + v) * D) + q);
      reinterpret_cast<float4*>(&zs[v][0])[q] = val;
    }
    __syncthreads();

    float dist[VB];
#pragma unroll
    for (int v = 0; v < VB; ++v) dist[v] = 0.f;
    if (!expanded) {
#pragma unroll
      for (int d = 0; d < D; d += 4) {
#pragma unroll
        for (int v = 0; v < VB; ++v) {
          const float4 zv = *reinterpret_cast<const float4*>(&zs[v][d]);
          float t;
          t = __fsub_rn(zv.x, e[d]);     dist[v] = __fmaf_rn(t, t, dist[v]);
          t = __fsub_rn(zv.y, e[d + 1]); dist[v] = __fmaf_rn(t, t, dist[v]);
          t = __fsub_rn(zv.z, e[d + 2]); dist[v] = __fmaf_rn(t, t, dist[v]);
          t = __fsub_rn(zv.w, e[d + 3]); dist[v] = __fmaf_rn(t, t, dist[v]);
        }
      }
    } else {
      // Magenta/sonnet.py:91-95: ||z||^2 - 2 z.w + ||w||^2 (the three terms formed separately, then combined in that order)
      float ee = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) ee = __fmaf_rn(e[d], e[d], ee);
#pragma unroll
      for (int v = 0; v < VB; ++v) {
        float zz = 0.f, dot = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float zv = zs[v][d];
          zz = __fmaf_rn(zv, zv, zz);
          dot = __fmaf_rn(zv, e[d], dot);
        }
        dist[v] = __fadd_rn(__fsub_rn(zz, __fmul_rn(2.0f, dot)), ee);
      }
    }
    // warp argmin, lowest index on ties (tf.argmin de-facto order)
#pragma unroll
    for (int v = 0; v < VB; ++v) {
      float bd = has_code ? dist[v] : INFINITY;
      int bk = has_code ? tid : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, off);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
        if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
      }
      if (lane == 0) { wmin_d[v][warp] = bd; wmin_k[v][warp] = bk; }
    }
    __syncthreads();
    if (warp < VB) {
      float bd = (lane < nwarps) ? wmin_d[warp][lane] : INFINITY;
      int bk = (lane < nwarps) ? wmin_k[warp][lane] : 0x7fffffff;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, off);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
        if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
      }
      if (lane == 0) {
        best_k[warp] = bk;
        if (idx_out != nullptr && v0 + warp < N) idx_out[v0 + warp] = (long long)bk;
      }
    }
    __syncthreads();
    // fused gather + straight-through + speaker concat
    if (zq_out != nullptr) {
      const int width = D + spk_dim;
      for (int i = tid; i < VB * width; i += blockDim.x) {
        const int v = i / width, c = i % width;
        const long long gv = v0 + v;
        if (gv < N) {
          float o;
          if (c < D) {
            const float zz = zs[v][c];
            const float ek = __ldg(E + (size_t)best_k[v] * D + c);
            o = out_code ? ek : __fadd_rn(zz, __fsub_rn(ek, zz));      // Magenta/config.py:242 (e_k) : model.py:73
          } else {
            const int b = (int)(gv / F);
            o = __ldg(spk_table + (size_t)spk_idx[b] * spk_dim + (c - D));
          }
          zq_out[(size_t)gv * out_stride + c] = o;
        }
      }
    }
    __syncthreads();
  }
}

// speaker lookup + concat only (vqwn_build_condition; also the use_vq = 0 identity path)
__global__ void build_condition_kernel(const float* __restrict__ zq, const float* __restrict__ spk_table,
                                       const int* __restrict__ spk_idx, int D, int spk_dim, int F,
                                       long long N, float* __restrict__ cond) {
  const int width = D + spk_dim;
  const long long total = N * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / width;
    const int c = (int)(i % width);
    float o;
    if (c < D) o = zq[v * D + c];
    else o = __ldg(spk_table + (size_t)spk_idx[(int)(v / F)] * spk_dim + (c - D));
    cond[i] = o;
  }
}

}  // namespace vqwn
