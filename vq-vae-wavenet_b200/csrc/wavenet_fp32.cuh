// WaveNet fast generation, float32 CUDA-core path: ONE persistent cooperative kernel that
// runs the whole sample loop of generate.py:103-113 on the device.
//
// Reference semantics (file:line relative to the reference root):
//   wavenet.py:103-172        one-step graph (preprocess FIR, skip start, 30 residual
//                             stacks, postprocess1/2, softmax)
//   wavenet_ops.py:163-195    fast_conv1d: current tap kernel[k-1], queue i holds the layer
//                             input delayed by i*d, zero pre-filled
//   wavenet_ops.py:198-267    condition add, tanh*sigmoid gate, skip / residual 1x1
//   utils.py:13-46            greedy argmax(probs) / float32 cumsum + searchsorted draw
//   mu_law_ops.py:5-31        mu-law encode / decode (257-entry LUTs built on the host)
//
// Decomposition: each stage is a [streams x K] x [K x N] contraction cut into 16-stream x
// 16-channel tiles; a tile is computed by one 256-thread CTA (8 warps split K, partials are
// reduced through shared memory).  Stages are separated by a grid barrier.  Per-layer
// dilation queues are ring buffers of depth 2d in HBM: slot t mod 2d holds the layer input
// of step t-2d (read as the oldest tap, then overwritten with the step-t input), slot
// (t-d) mod 2d holds the middle tap.
//
// Layouts (all float32, stream-major like the reference's [B,C] tensors):
//   cur [Bp,R]  g [Bp,G]  skip [Bp,S]  n1 [Bp,S]  logits [Bp,Q]  u_hist [Bp,PK]
//   ring_l [2d, Bp, R]
//   w1_l [3R+C, 2G] rows = [gated/kernel[2] ; kernel[1] ; kernel[0] ; local_condition/kernel[0]]
//   w2_l [G, R+S]   cols = [residual/kernel[0] | skip/kernel[0]],  b2 = [residual/bias | skip/bias]
//   post1_w [S+C, S] rows = [postprocess1/kernel[0] ; postprocess1/local_condition/kernel[0]]
#pragma once
#include "common.cuh"

namespace vqwn {

constexpr int FP32_TB = 16;        // streams per tile
constexpr int FP32_TN = 16;        // output channels per tile
constexpr int FP32_THREADS = 256;  // 8 warps
constexpr int FP32_WARPS = 8;

enum GenMode { GEN_GREEDY = 0, GEN_SAMPLE = 1, GEN_STEP = 2, GEN_TEACHER = 3 };

struct LayerDev {
  const float* w1;
  const float* b1;
  const float* w2;
  const float* b2;
  float* ring;
  int d;
  int pad_;
};

struct GenParams {
  int L, R, G, S, Q, C, PK;
  int B, Bp;
  int lda;  // shared-memory row stride of the activation tile (floats)
  const float *pre_k, *pre_b, *skip0_w, *skip0_b, *post1_w, *post1_b, *post2_w, *post2_b;
  const LayerDev* layers;
  const float *enc_lut, *dec_lut;
  float *u_hist, *cur, *g, *skip, *n1, *logits;
  long long t0, T;
  int mode;
  const float* cond;        // stream b, step t -> cond + b*cond_bstride + ((t-t0)/ratio)*C
  long long cond_bstride;
  int ratio;                // 0: one frame only (step API)
  const float* ext_audio;   // GEN_STEP: [B]; GEN_TEACHER: x [B,T]
  const double* uniforms;   // [T,B] or null
  unsigned long long seed;
  float* audio_out;         // [B,T]
  int* idx_out;             // [B,T] or null
  float* logits_out;        // GEN_STEP: [B,Q]; GEN_TEACHER: [B,T,Q]; else null
  float* probs_out;         // GEN_STEP: [B,Q] or null
  unsigned long long* barrier;
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// 16 rows x len floats from global (written by other CTAs -> .cg) into the activation tile
__device__ __forceinline__ void load_rows(float* act_s, int lda, int koff, const float* src,
                                          long long row_stride, int len, bool relu) {
  const int q4 = len >> 2;
  for (int idx = threadIdx.x; idx < FP32_TB * q4; idx += FP32_THREADS) {
    const int i = idx / q4, q = idx - i * q4;
    float4 v = ld_cg4(src + (long long)i * row_stride + 4 * q);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(act_s + i * lda + koff + 4 * q) = v;
  }
}

__device__ __forceinline__ void load_cond_rows(float* act_s, int lda, int koff, const GenParams& p,
                                               int sb, long long t) {
  const int q4 = p.C >> 2;
  const long long frame = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
  for (int idx = threadIdx.x; idx < FP32_TB * q4; idx += FP32_THREADS) {
    const int i = idx / q4, q = idx - i * q4;
    int b = sb * FP32_TB + i;
    if (b >= p.B) b = p.B - 1;   // padded streams reuse the last real stream's condition
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.cond + (long long)b * p.cond_bstride + frame * p.C) + q);
    *reinterpret_cast<float4*>(act_s + i * lda + koff + 4 * q) = v;
  }
}

// out(stream i = tid/16, column c = tid%16) = sum_k act_s[i][k] * W[k][col(c)]
// `col` is the global column of W this thread's lane (c = lane & 15) accumulates.
__device__ __forceinline__ float tile_gemm(const float* __restrict__ W, int ldw, int col, int Ktot,
                                           const float* act_s, int lda, float* red_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, c = lane & 15;
  const int Kw = Ktot / FP32_WARPS;
  const int k0 = warp * Kw, k1 = k0 + Kw;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float* wp = W + col;
  const float* ap = act_s + (half * 8) * lda;
#pragma unroll 2
  for (int k = k0; k < k1; k += 4) {
    const float w0 = __ldg(wp + (long long)(k + 0) * ldw);
    const float w1 = __ldg(wp + (long long)(k + 1) * ldw);
    const float w2 = __ldg(wp + (long long)(k + 2) * ldw);
    const float w3 = __ldg(wp + (long long)(k + 3) * ldw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(ap + j * lda + k);
      acc[j] = fmaf(a.x, w0, acc[j]);
      acc[j] = fmaf(a.y, w1, acc[j]);
      acc[j] = fmaf(a.z, w2, acc[j]);
      acc[j] = fmaf(a.w, w3, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red_s[warp * 256 + (half * 8 + j) * 16 + c] = acc[j];
  __syncthreads();
  float out = 0.f;
#pragma unroll
  for (int w = 0; w < FP32_WARPS; ++w) out += red_s[w * 256 + threadIdx.x];
  return out;
}

__global__ void __launch_bounds__(FP32_THREADS, 1) wavenet_fp32_persistent(const GenParams p) {
  extern __shared__ __align__(16) float smem[];
  float* act_s = smem;                              // [16][lda]
  float* red_s = act_s + FP32_TB * p.lda;           // [8][256]
  float* u_s = red_s + FP32_WARPS * 256;            // [16][PK]
  float* ps = u_s + FP32_TB * p.PK;                 // [8][Q]

  GridBarrier bar{p.barrier, 0ULL};
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tc = tid & 15;           // epilogue identity: stream-in-tile, column-in-tile
  const int lane = tid & 31, warp = tid >> 5;
  const int nsb = p.Bp / FP32_TB;
  const int lda = p.lda;
  const float mu = (float)(p.Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {
    // ------------------------------------------------------------------ stage 0: preprocess FIR + skip start
    {
      const int ncb = p.S / FP32_TN;
      for (int tile = blockIdx.x; tile < ncb * nsb; tile += gridDim.x) {
        const int cb = tile / nsb, sb = tile - cb * nsb;
        for (int idx = tid; idx < FP32_TB * p.PK; idx += FP32_THREADS) {
          const int i = idx / p.PK, j = idx - i * p.PK;
          const int b = sb * FP32_TB + i;
          float u;
          if (j == 0 && ext) {
            float x = 0.f;
            if (b < p.B) {
              if (p.mode == GEN_STEP) x = p.ext_audio[b];
              else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
            }
            u = mu_law_encode_dev(x, mu, 0.f);
            if (cb == 0) st_cg(p.u_hist + (long long)b * p.PK + (int)(t % p.PK), u);
          } else {
            long long s = (t - j) % p.PK;
            if (s < 0) s += p.PK;
            u = ld_cg(p.u_hist + (long long)b * p.PK + s);
          }
          u_s[i * p.PK + j] = u;
        }
        __syncthreads();
        // h0[i][n] = (u0*K[PK-1] + bias) + u1*K[PK-2] + ...   (wavenet_ops.py:178,193)
        for (int idx = tid; idx < FP32_TB * p.R; idx += FP32_THREADS) {
          const int i = idx / p.R, n = idx - i * p.R;
          float acc = fmaf(u_s[i * p.PK], __ldg(p.pre_k + (p.PK - 1) * p.R + n), __ldg(p.pre_b + n));
          for (int j = 1; j < p.PK; ++j)
            acc = fmaf(u_s[i * p.PK + j], __ldg(p.pre_k + (p.PK - 1 - j) * p.R + n), acc);
          act_s[i * lda + n] = acc;
        }
        __syncthreads();
        if (cb < p.R / FP32_TN) {
          const int n = cb * FP32_TN + tc;
          st_cg(p.cur + (long long)(sb * FP32_TB + ti) * p.R + n, act_s[ti * lda + n]);
        }
        const int col = cb * FP32_TN + tc;
        const float acc = tile_gemm(p.skip0_w, p.S, col, p.R, act_s, lda, red_s);
        st_cg(p.skip + (long long)(sb * FP32_TB + ti) * p.S + col, acc + __ldg(p.skip0_b + col));
        __syncthreads();
      }
    }
    bar.sync();

    // ------------------------------------------------------------------ residual stacks
    for (int l = 0; l < p.L; ++l) {
      const LayerDev ly = p.layers[l];
      const int d2 = 2 * ly.d;
      const int slot_old = (int)(t % d2);                 // holds input of step t-2d; overwritten below
      const int slot_mid = (int)((t + ly.d) % d2);        // holds input of step t-d
      const long long ring_slot = (long long)p.Bp * p.R;
      // ---- S1: gated conv + condition + tanh*sigmoid  -> g
      {
        const int ncb = p.G / 8;
        const int K1 = 3 * p.R + p.C;
        for (int tile = blockIdx.x; tile < ncb * nsb; tile += gridDim.x) {
          const int cb = tile / nsb, sb = tile - cb * nsb;
          const long long row0 = (long long)sb * FP32_TB;
          load_rows(act_s, lda, 0, p.cur + row0 * p.R, p.R, p.R, false);
          load_rows(act_s, lda, p.R, ly.ring + slot_mid * ring_slot + row0 * p.R, p.R, p.R, false);
          load_rows(act_s, lda, 2 * p.R, ly.ring + slot_old * ring_slot + row0 * p.R, p.R, p.R, false);
          load_cond_rows(act_s, lda, 3 * p.R, p, sb, t);
          __syncthreads();
          const int col = (tc < 8) ? (cb * 8 + tc) : (p.G + cb * 8 + (tc - 8));
          float v = tile_gemm(ly.w1, 2 * p.G, col, K1, act_s, lda, red_s) + __ldg(ly.b1 + col);
          const float partner = __shfl_down_sync(0xffffffffu, v, 8);
          if (tc < 8) st_cg(p.g + (row0 + ti) * p.G + cb * 8 + tc, tanhf(v) * sigmoid_f(partner));
          __syncthreads();
        }
      }
      bar.sync();
      // ---- S2: residual (+ queue push) and skip accumulation
      {
        const int nres = p.R / FP32_TN;
        const int ncb = (p.R + p.S) / FP32_TN;
        const bool last = (l == p.L - 1);
        for (int tile = blockIdx.x; tile < ncb * nsb; tile += gridDim.x) {
          const int cb = tile / nsb, sb = tile - cb * nsb;
          const long long row = (long long)sb * FP32_TB + ti;
          const int col = cb * FP32_TN + tc;
          if (last && cb < nres) {
            // last layer: residual output is dead (wavenet.py:145 result unused), queue push stays
            const float old = ld_cg(p.cur + row * p.R + col);
            st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, old);
            continue;
          }
          load_rows(act_s, lda, 0, p.g + (long long)sb * FP32_TB * p.G, p.G, p.G, false);
          __syncthreads();
          const float v = tile_gemm(ly.w2, p.R + p.S, col, p.G, act_s, lda, red_s) + __ldg(ly.b2 + col);
          if (col < p.R) {
            const float old = ld_cg(p.cur + row * p.R + col);
            st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, old);   // push_ops
            st_cg(p.cur + row * p.R + col, old + v);
          } else {
            float* sp = p.skip + row * p.S + (col - p.R);
            st_cg(sp, ld_cg(sp) + v);
          }
          __syncthreads();
        }
      }
      bar.sync();
    }

    // ------------------------------------------------------------------ postprocess1 (+ condition)
    {
      const int ncb = p.S / FP32_TN;
      for (int tile = blockIdx.x; tile < ncb * nsb; tile += gridDim.x) {
        const int cb = tile / nsb, sb = tile - cb * nsb;
        load_rows(act_s, lda, 0, p.skip + (long long)sb * FP32_TB * p.S, p.S, p.S, true);
        load_cond_rows(act_s, lda, p.S, p, sb, t);
        __syncthreads();
        const int col = cb * FP32_TN + tc;
        const float v = tile_gemm(p.post1_w, p.S, col, p.S + p.C, act_s, lda, red_s) + __ldg(p.post1_b + col);
        st_cg(p.n1 + ((long long)sb * FP32_TB + ti) * p.S + col, v);
        __syncthreads();
      }
    }
    bar.sync();
    // ------------------------------------------------------------------ postprocess2 -> logits
    {
      const int ncb = p.Q / FP32_TN;
      for (int tile = blockIdx.x; tile < ncb * nsb; tile += gridDim.x) {
        const int cb = tile / nsb, sb = tile - cb * nsb;
        load_rows(act_s, lda, 0, p.n1 + (long long)sb * FP32_TB * p.S, p.S, p.S, true);
        __syncthreads();
        const int col = cb * FP32_TN + tc;
        const float v = tile_gemm(p.post2_w, p.Q, col, p.S, act_s, lda, red_s) + __ldg(p.post2_b + col);
        st_cg(p.logits + ((long long)sb * FP32_TB + ti) * p.Q + col, v);
        __syncthreads();
      }
    }
    bar.sync();
    // ------------------------------------------------------------------ softmax + draw + mu-law decode
    {
      const int NQ = p.Q / 32;   // <= 8
      for (int b = blockIdx.x * FP32_WARPS + warp; b < p.B; b += gridDim.x * FP32_WARPS) {
        float lg[8], pr[8];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          lg[i] = (i < NQ) ? ld_cg(p.logits + (long long)b * p.Q + lane + 32 * i) : -INFINITY;
          m = fmaxf(m, lg[i]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { pr[i] = (i < NQ) ? expf(lg[i] - m) : 0.f; s += pr[i]; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = __fdiv_rn(pr[i], s);

        if (p.mode == GEN_STEP) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < NQ && p.logits_out) p.logits_out[(long long)b * p.Q + lane + 32 * i] = lg[i];
            if (i < NQ && p.probs_out) p.probs_out[(long long)b * p.Q + lane + 32 * i] = pr[i];
          }
          continue;
        }
        if (p.mode == GEN_TEACHER) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < NQ) p.logits_out[((long long)b * p.T + (t - p.t0)) * p.Q + lane + 32 * i] = lg[i];
          continue;
        }
        int k;
        if (p.mode == GEN_GREEDY) {
          // np.argmax(probs): first maximum (utils.py:43)
          float bv = -1.f; int bi = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < NQ && pr[i] > bv) { bv = pr[i]; bi = lane + 32 * i; }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          k = bi;
        } else {
          // utils.py:20-25: sequential float32 cumsum, float64 compare, searchsorted 'left'
          float* pw = ps + warp * p.Q;
#pragma unroll
          for (int i = 0; i < 8; ++i) if (i < NQ) pw[lane + 32 * i] = pr[i];
          __syncwarp();
          int cnt = 0;
          if (lane == 0) {
            const double u = p.uniforms ? p.uniforms[(t - p.t0) * p.B + b]
                                        : counter_uniform(p.seed, (unsigned long long)t, (unsigned long long)b);
            float c = 0.f;
            for (int i = 0; i < p.Q; ++i) {
              c = __fadd_rn(c, pw[i]);
              cnt += ((double)c < u) ? 1 : 0;
            }
          }
          k = __shfl_sync(0xffffffffu, cnt, 0);
          __syncwarp();
        }
        if (lane == 0) {
          p.audio_out[(long long)b * p.T + (t - p.t0)] = __ldg(p.dec_lut + k);
          if (p.idx_out) p.idx_out[(long long)b * p.T + (t - p.t0)] = k;
          st_cg(p.u_hist + (long long)b * p.PK + (int)((t + 1) % p.PK), __ldg(p.enc_lut + k));
        }
      }
    }
    bar.sync();
  }
}

}  // namespace vqwn
