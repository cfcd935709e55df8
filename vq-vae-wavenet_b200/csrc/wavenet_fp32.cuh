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
// Decomposition: a step is a chain of stages (preprocess+skip start, 2 per layer, post1,
// post2, draw) separated by a grid barrier.  Each stage is a [streams x K] x [K x N]
// contraction cut into (16 streams x NC channels) tiles, one tile per CTA.  Inside a tile the
// 256 threads form K-groups; a thread owns an 8-stream x 4-channel register tile (128 FMA per
// 12 shared-memory vector loads), partial sums are reduced through shared memory.
// The weight tile of the NEXT stage is prefetched with cp.async while the current stage
// computes and while the CTA waits at the barrier (weights do not depend on the data);
// activations (written by other CTAs) are fetched with cp.async.cg after the barrier.
// Per-layer dilation queues are ring buffers of depth 2d in HBM: slot t mod 2d holds the layer
// input of step t-2d (read as the oldest tap, then overwritten with the step-t input), slot
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
constexpr int FP32_THREADS = 256;  // 8 warps
constexpr int FP32_WARPS = 8;
constexpr int FP32_RED_FLOATS = 8192;   // K-groups x tile outputs (32 x 256 or 16 x 512)

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
  int lda;      // shared-memory row stride of the activation tile (floats)
  int wfloats;  // floats per weight-tile buffer
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
  long long* prof;          // optional [8] cycle counters of CTA 0 (VQWN_PROFILE=1), else null
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one tile of one stage: which weight columns it needs and which streams it covers
struct TileInfo {
  const float* W;   // null: no contraction for this tile (last layer's dead residual)
  int ldw, K, NC, col0, col1, cb, sb;
};

// stage ids inside a step: 0 preprocess+skip start | 1+2l gated (S1) | 2+2l residual/skip (S2)
//                          | 2L+1 postprocess1 | 2L+2 postprocess2 | 2L+3 draw (no tiles)
__device__ __forceinline__ int stage_tiles(const GenParams& p, int s) {
  const int nsb = p.Bp / FP32_TB;
  if (s == 0) return (p.S / 16) * nsb;
  if (s <= 2 * p.L) return ((s & 1) ? (p.G / 8) : ((p.R + p.S) / 32)) * nsb;
  if (s == 2 * p.L + 1) return (p.S / 16) * nsb;
  if (s == 2 * p.L + 2) return (p.Q / 16) * nsb;
  return 0;
}

__device__ __forceinline__ void stage_tile(const GenParams& p, int s, int tile, TileInfo& ti) {
  const int nsb = p.Bp / FP32_TB;
  ti.cb = tile / nsb;
  ti.sb = tile - ti.cb * nsb;
  ti.col1 = -1;
  if (s == 0) {
    ti.W = p.skip0_w; ti.ldw = p.S; ti.K = p.R; ti.NC = 16; ti.col0 = ti.cb * 16;
  } else if (s <= 2 * p.L) {
    const int l = (s - 1) >> 1;
    if (s & 1) {
      ti.W = p.layers[l].w1; ti.ldw = 2 * p.G; ti.K = 3 * p.R + p.C; ti.NC = 16;
      ti.col0 = ti.cb * 8; ti.col1 = p.G + ti.cb * 8;      // 8 tanh channels + their 8 sigmoid partners
    } else {
      ti.W = p.layers[l].w2; ti.ldw = p.R + p.S; ti.K = p.G; ti.NC = 32; ti.col0 = ti.cb * 32;
      if (l == p.L - 1 && ti.col0 < p.R) ti.W = nullptr;   // wavenet.py:145: last residual is dead
    }
  } else if (s == 2 * p.L + 1) {
    ti.W = p.post1_w; ti.ldw = p.S; ti.K = p.S + p.C; ti.NC = 16; ti.col0 = ti.cb * 16;
  } else {
    ti.W = p.post2_w; ti.ldw = p.Q; ti.K = p.S; ti.NC = 16; ti.col0 = ti.cb * 16;
  }
}

// weight tile [K][NC] -> shared memory (16-byte async copies)
__device__ __forceinline__ void issue_w_tile(float* wbuf, const TileInfo& ti) {
  if (ti.W == nullptr) return;
  const int sh = (ti.NC == 16) ? 2 : 3;          // log2(16-byte chunks per row)
  const int c4 = threadIdx.x & ((1 << sh) - 1);
  int col = ti.col0 + 4 * c4;
  if (ti.col1 >= 0 && c4 >= 2) col = ti.col1 + 4 * (c4 - 2);
  const float* src = ti.W + col;
  float* dst = wbuf + 4 * c4;
  const int kstep = FP32_THREADS >> sh;
  for (int k = threadIdx.x >> sh; k < ti.K; k += kstep)
    cp_async16(dst + k * ti.NC, src + (long long)k * ti.ldw);
}

__device__ __forceinline__ void issue_w_prefetch(const GenParams& p, float* wbuf, int s) {
  if ((int)blockIdx.x < stage_tiles(p, s)) {
    TileInfo ti;
    stage_tile(p, s, blockIdx.x, ti);
    issue_w_tile(wbuf, ti);
  }
}

// 16 rows x len floats produced by other CTAs (global, L2-coherent .cg) -> activation tile
__device__ __forceinline__ void issue_rows(float* act_s, int lda, int koff, const float* src, long long row_stride, int len) {
  const int q4 = len >> 2;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x >> 5; i < FP32_TB; i += FP32_WARPS)        // warp w: rows w, w+8
    for (int q = lane; q < q4; q += 32)
      cp_async16(act_s + i * lda + koff + 4 * q, src + (long long)i * row_stride + 4 * q);
}

__device__ __forceinline__ void issue_cond_rows(float* act_s, int lda, int koff, const GenParams& p, int sb, long long t) {
  const int q4 = p.C >> 2;
  const long long frame = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x >> 5; i < FP32_TB; i += FP32_WARPS) {
    int b = sb * FP32_TB + i;
    if (b >= p.B) b = p.B - 1;   // padded streams reuse the last real stream's condition
    const float* src = p.cond + (long long)b * p.cond_bstride + frame * p.C;
    for (int q = lane; q < q4; q += 32) cp_async16(act_s + i * lda + koff + 4 * q, src + 4 * q);
  }
}

__device__ __forceinline__ void relu_rows(float* act_s, int lda, int len) {
  const int q4 = len >> 2;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x >> 5; i < FP32_TB; i += FP32_WARPS)
    for (int q = lane; q < q4; q += 32) {
      float4* a = reinterpret_cast<float4*>(act_s + i * lda + 4 * q);
      float4 v = *a;
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      *a = v;
    }
}

// partial sums of the (16 streams x NC channels) tile into red_s[kgroup][stream*NC + channel]
template <int NC>
__device__ __forceinline__ void tile_compute(const float* wbuf, const float* act_s, int lda, int K, float* red_s) {
  constexpr int CQ = NC / 4;          // channel quads
  constexpr int TPG = 2 * CQ;         // threads per K-group (2 stream halves x channel quads)
  constexpr int KG = FP32_THREADS / TPG;
  const int tid = threadIdx.x;
  const int kg = tid / TPG, r = tid - kg * TPG;
  const int sh = r / CQ, q = r - sh * CQ;
  const int klen = K / KG;
  const int k0 = kg * klen;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
  const float* wp = wbuf + q * 4;
  const float* ap = act_s + sh * lda;
  for (int k = k0; k < k0 + klen; k += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(wp + (k + 0) * NC);
    const float4 w1 = *reinterpret_cast<const float4*>(wp + (k + 1) * NC);
    const float4 w2 = *reinterpret_cast<const float4*>(wp + (k + 2) * NC);
    const float4 w3 = *reinterpret_cast<const float4*>(wp + (k + 3) * NC);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(ap + (2 * j) * lda + k);   // stream 2j+sh
      acc[j][0] = fmaf(a.x, w0.x, acc[j][0]); acc[j][1] = fmaf(a.x, w0.y, acc[j][1]);
      acc[j][2] = fmaf(a.x, w0.z, acc[j][2]); acc[j][3] = fmaf(a.x, w0.w, acc[j][3]);
      acc[j][0] = fmaf(a.y, w1.x, acc[j][0]); acc[j][1] = fmaf(a.y, w1.y, acc[j][1]);
      acc[j][2] = fmaf(a.y, w1.z, acc[j][2]); acc[j][3] = fmaf(a.y, w1.w, acc[j][3]);
      acc[j][0] = fmaf(a.z, w2.x, acc[j][0]); acc[j][1] = fmaf(a.z, w2.y, acc[j][1]);
      acc[j][2] = fmaf(a.z, w2.z, acc[j][2]); acc[j][3] = fmaf(a.z, w2.w, acc[j][3]);
      acc[j][0] = fmaf(a.w, w3.x, acc[j][0]); acc[j][1] = fmaf(a.w, w3.y, acc[j][1]);
      acc[j][2] = fmaf(a.w, w3.z, acc[j][2]); acc[j][3] = fmaf(a.w, w3.w, acc[j][3]);
    }
  }
  float* rp = red_s + kg * (FP32_TB * NC) + q * 4;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(rp + (2 * j + sh) * NC) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}

template <int NC>
__device__ __forceinline__ float tile_reduce(const float* red_s, int o) {
  constexpr int KG = FP32_THREADS / (NC / 2);
  float s = 0.f;
#pragma unroll 8
  for (int kg = 0; kg < KG; ++kg) s += red_s[kg * (FP32_TB * NC) + o];
  return s;
}

__global__ void __launch_bounds__(FP32_THREADS, 1) wavenet_fp32_persistent(const GenParams p) {
  extern __shared__ __align__(16) float smem[];
  float* wbuf0 = smem;                              // 2 weight-tile buffers
  float* act_s = smem + 2 * p.wfloats;              // [16][lda]
  float* red_s = act_s + FP32_TB * p.lda;           // [kgroups][tile outputs]
  float* u_s = red_s + FP32_RED_FLOATS;             // [16][PK]
  float* ps = u_s + FP32_TB * p.PK;                 // [8][Q]

  GridBarrier bar{p.barrier, 0ULL};
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int lda = p.lda;
  const float mu = (float)(p.Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  const int NS = 2 * p.L + 4;
  const int S_P1 = 2 * p.L + 1, S_P2 = 2 * p.L + 2, S_DRAW = 2 * p.L + 3;
  const long long ring_slot = (long long)p.Bp * p.R;

  long long pf_bar = 0, pf_wait = 0, pf_comp = 0, pf_epi = 0, pf_draw = 0, pf_issue = 0, pf_t = 0;
  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
#define PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define PF_ADD(x) do { if (prof) { long long n_ = clock64(); x += n_ - pf_t; pf_t = n_; } } while (0)
  int wb = 0;   // buffer holding the weight tile of the current stage's first tile
  issue_w_prefetch(p, wbuf0, 0);
  cp_async_commit();

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {
    for (int s = 0; s < NS; ++s) {
      if (s == S_DRAW) {
        // -------------------------------------------------------------- softmax + draw + mu-law decode
        const int NQ = p.Q / 32;   // <= 8
        PF_START();
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
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { pr[i] = (i < NQ) ? expf(lg[i] - m) : 0.f; sum += pr[i]; }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll
          for (int i = 0; i < 8; ++i) pr[i] = __fdiv_rn(pr[i], sum);

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
        PF_ADD(pf_draw);
        bar.sync();
        PF_ADD(pf_bar);
        continue;
      }

      // ---------------------------------------------------------------- contraction stages
      const int ntiles = stage_tiles(p, s);
      const int s_next = (s + 1 == S_DRAW) ? 0 : s + 1;   // next stage that owns weight tiles
      const int l = (s - 1) >> 1;                          // layer of S1/S2 stages
      bool first = true;
      if ((int)blockIdx.x >= ntiles) {
        // no tile here: still stream the next stage's weights
        issue_w_prefetch(p, wbuf0 + (wb ^ 1) * p.wfloats, s_next);
        cp_async_commit();
      }
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TileInfo ti;
        stage_tile(p, s, tile, ti);
        float* wcur = wbuf0 + wb * p.wfloats;
        const long long row0 = (long long)ti.sb * FP32_TB;
        if (!first) issue_w_tile(wcur, ti);   // later rounds: this tile's weights were not prefetched
        PF_START();

        // ---- activations
        if (s == 0) {
          for (int idx = tid; idx < FP32_TB * p.PK; idx += FP32_THREADS) {
            const int i = idx / p.PK, j = idx - i * p.PK;
            const int b = ti.sb * FP32_TB + i;
            float u;
            if (j == 0 && ext) {
              float x = 0.f;
              if (b < p.B) {
                if (p.mode == GEN_STEP) x = p.ext_audio[b];
                else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
              }
              u = mu_law_encode_dev(x, mu, 0.f);
              if (ti.cb == 0) st_cg(p.u_hist + (long long)b * p.PK + (int)(t % p.PK), u);
            } else {
              long long sl = (t - j) % p.PK;
              if (sl < 0) sl += p.PK;
              u = ld_cg(p.u_hist + (long long)b * p.PK + sl);
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
        } else if (s <= 2 * p.L) {
          const LayerDev ly = p.layers[l];
          if (s & 1) {
            const int d2 = 2 * ly.d;
            const int slot_old = (int)(t % d2);
            const int slot_mid = (int)((t + ly.d) % d2);
            issue_rows(act_s, lda, 0, p.cur + row0 * p.R, p.R, p.R);
            issue_rows(act_s, lda, p.R, ly.ring + slot_mid * ring_slot + row0 * p.R, p.R, p.R);
            issue_rows(act_s, lda, 2 * p.R, ly.ring + slot_old * ring_slot + row0 * p.R, p.R, p.R);
            issue_cond_rows(act_s, lda, 3 * p.R, p, ti.sb, t);
          } else if (ti.W != nullptr) {
            issue_rows(act_s, lda, 0, p.g + row0 * p.G, p.G, p.G);
          }
        } else if (s == S_P1) {
          issue_rows(act_s, lda, 0, p.skip + row0 * p.S, p.S, p.S);
          issue_cond_rows(act_s, lda, p.S, p, ti.sb, t);
        } else {
          issue_rows(act_s, lda, 0, p.n1 + row0 * p.S, p.S, p.S);
        }
        cp_async_commit();
        if (first) {
          issue_w_prefetch(p, wbuf0 + (wb ^ 1) * p.wfloats, s_next);
          cp_async_commit();
          first = false;
        } else {
          cp_async_commit();   // keep the group count uniform
        }
        PF_ADD(pf_issue);
        cp_async_wait<1>();      // everything but the newest group (next stage's weights) has landed
        __syncthreads();
        PF_ADD(pf_wait);
        if (s >= S_P1) {         // relu on the contraction input (wavenet.py:153,163)
          relu_rows(act_s, lda, p.S);
          __syncthreads();
        }

        // ---- contraction + epilogue
        if (ti.NC == 16) {
          const int i = tid >> 4, c = tid & 15;
          const int col = (ti.col1 >= 0 && c >= 8) ? (ti.col1 + c - 8) : (ti.col0 + c);
          const float* bias_p = (s == 0) ? p.skip0_b : (s <= 2 * p.L) ? p.layers[l].b1 : (s == S_P1) ? p.post1_b : p.post2_b;
          const float bias = __ldg(bias_p + col);            // in flight during the contraction
          if (s == 0 && ti.cb < p.R / 16) {
            const int n = ti.cb * 16 + c;
            st_cg(p.cur + (row0 + i) * p.R + n, act_s[i * lda + n]);
          }
          tile_compute<16>(wcur, act_s, lda, ti.K, red_s);
          __syncthreads();
          PF_ADD(pf_comp);
          float v = tile_reduce<16>(red_s, tid) + bias;
          const long long row = row0 + i;
          if (s == 0) {
            st_cg(p.skip + row * p.S + col, v);
          } else if (s <= 2 * p.L) {
            const float partner = __shfl_down_sync(0xffffffffu, v, 8);
            if (c < 8) st_cg(p.g + row * p.G + col, tanhf(v) * sigmoid_f(partner));
          } else if (s == S_P1) {
            st_cg(p.n1 + row * p.S + col, v);
          } else {
            st_cg(p.logits + row * p.Q + col, v);
          }
        } else {
          // S2: residual (+ queue push) and skip accumulation; 2 outputs per thread
          const LayerDev ly = p.layers[l];
          const int slot_old = (int)(t % (2 * ly.d));
          const bool is_res = ti.col0 < p.R;
          // old residual / skip values and biases: loads in flight during the contraction
          float oldv[2], bias[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o = tid + h * FP32_THREADS;
            const long long row = row0 + (o >> 5);
            const int col = ti.col0 + (o & 31);
            oldv[h] = is_res ? ld_cg(p.cur + row * p.R + col) : ld_cg(p.skip + row * p.S + (col - p.R));
            bias[h] = __ldg(ly.b2 + col);
          }
          if (ti.W != nullptr) {
            tile_compute<32>(wcur, act_s, lda, ti.K, red_s);
            __syncthreads();
            PF_ADD(pf_comp);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o = tid + h * FP32_THREADS;
            const long long row = row0 + (o >> 5);
            const int col = ti.col0 + (o & 31);
            if (ti.W == nullptr) {
              // last layer: residual output is dead, the queue push stays
              st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, oldv[h]);
            } else {
              const float v = tile_reduce<32>(red_s, o) + bias[h];
              if (is_res) {
                st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, oldv[h]);   // push_ops
                st_cg(p.cur + row * p.R + col, oldv[h] + v);
              } else {
                st_cg(p.skip + row * p.S + (col - p.R), oldv[h] + v);
              }
            }
          }
        }
        __syncthreads();
        PF_ADD(pf_epi);
      }
      wb ^= 1;
      PF_START();
      bar.sync();
      PF_ADD(pf_bar);
    }
  }
  cp_async_wait<0>();
  if (prof) {
    p.prof[0] = pf_bar; p.prof[1] = pf_wait; p.prof[2] = pf_comp; p.prof[3] = pf_epi; p.prof[4] = pf_draw; p.prof[5] = pf_issue;
  }
}

}  // namespace vqwn
