// utils.decode on the device (reference: utils.py:13-46, mu_law_ops.py:26-31).
// One warp per stream.  greedy: first argmax of probs.  sample: sequential float32 running
// sum of probs (np.cumsum order), count of cdf entries < u compared in float64
// (ndarray.searchsorted 'left'); the count can reach q -> LUT has q+1 entries.
#pragma once
#include "common.cuh"

namespace vqwn {

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ probs, int B, int Q, int mode,
                                                     const double* __restrict__ uniforms,
                                                     const float* __restrict__ dec_lut, int* __restrict__ idx_out,
                                                     float* __restrict__ audio_out) {
  extern __shared__ float ps[];   // [8][Q]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  int k;
  if (mode == 0) {
    float bv = -INFINITY; int bi = 0;
    for (int i = lane; i < Q; i += 32) {
      const float v = probs[(long long)b * Q + i];
      if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    k = bi;
  } else {
    float* pw = ps + warp * Q;
    for (int i = lane; i < Q; i += 32) pw[i] = probs[(long long)b * Q + i];
    __syncwarp();
    int cnt = 0;
    if (lane == 0) {
      const double u = uniforms[b];
      float c = 0.f;
      for (int i = 0; i < Q; ++i) {
        c = __fadd_rn(c, pw[i]);
        cnt += ((double)c < u) ? 1 : 0;
      }
    }
    k = __shfl_sync(0xffffffffu, cnt, 0);
  }
  if (lane == 0) {
    if (idx_out) idx_out[b] = k;
    if (audio_out) audio_out[b] = dec_lut[k];
  }
}

}  // namespace vqwn
