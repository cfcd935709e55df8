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
//
// Data movement is bulk-asynchronous (cp.async.bulk + mbarrier complete_tx, the TMA copy
// engine): weights are pre-packed tile-major so a stage's weight tile is ONE contiguous block;
// it and the rows that do not depend on the previous stage (older dilation taps, condition)
// are requested while the previous stage still computes / waits at the barrier; only the rows
// produced by the previous stage are fetched after the barrier.  Stages alternate between two
// shared-memory buffer classes (A: gated conv / post1, B: skip start / residual+skip / post2),
// so the prefetch of stage s+1 never touches the buffers stage s computes from.
// Per-layer dilation queues are ring buffers of depth 2d in HBM: slot t mod 2d holds the layer
// input of step t-2d (read as the oldest tap, then overwritten with the step-t input), slot
// (t-d) mod 2d holds the middle tap.
//
// Layouts (all float32, stream-major like the reference's [B,C] tensors):
//   cur [Bp,R]  g [Bp,G]  skip [Bp,S]  n1 [Bp,S]  logits [Bp,Q]  u_hist [Bp,PK]
//   ring_l [2d, Bp, R]
//   weight tiles (tile-major, [tile][K][NC]):
//     w1t_l [G/8][3R+C][16]   K rows = gated/kernel[2] ; kernel[1] ; kernel[0] ; local_condition,
//                             columns = 8 tanh channels | their 8 sigmoid partners
//     w2t_l [(R+S)/32][G][32] columns of [residual/kernel[0] | skip/kernel[0]]
//     skip0t [S/16][R][16]   post1t [S/16][S+C][16]   post2t [Q/16][S][16]
#pragma once
#include "common.cuh"

namespace vqwn {

constexpr int FP32_TB = 16;        // streams per tile
constexpr int FP32_THREADS = 256;  // 8 warps
constexpr int FP32_WARPS = 8;
constexpr int FIR_TAPS = 32;       // preprocess kernel size with a register-resident fast path
constexpr int FP32_RED_FLOATS = 8192;   // K-groups x tile outputs (32 x 256 or 16 x 512)

enum GenMode { GEN_GREEDY = 0, GEN_SAMPLE = 1, GEN_STEP = 2, GEN_TEACHER = 3 };

struct LayerDev {
  const float* w1t;
  const float* b1;
  const float* w2t;
  const float* b2;
  float* ring;
  int d;
  int pad_;
};

struct GenParams {
  int L, R, G, S, Q, C, PK;
  int B, Bp;
  int actA_floats, actB_floats;   // activation buffers: segment-major [segment][16 streams][segment length]
  int wfloatsA, wfloatsB;
  const float *pre_k, *pre_b, *skip0t, *skip0_b, *post1t, *post1_b, *post2t, *post2_b;
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
  int b_offset;             // global index of stream 0 (vqwn_set_stream_offset): keys the seeded generator
  float* audio_out;         // [B,T]
  int* idx_out;             // [B,T] or null
  float* logits_out;        // GEN_STEP: [B,Q]; GEN_TEACHER: [B,T,Q]; else null
  float* probs_out;         // GEN_STEP: [B,Q] or null
  unsigned long long* barrier;
  long long* prof;          // optional cycle counters of CTA 0 (VQWN_PROFILE=1), else null
  int* err;                 // set when a bounded wait expires
};

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// One warp, one stream: softmax over the Q <= 256 logits held as lg[i] = logit[lane + 32 i] (i < Q/32, else -inf), then
// what the mode asks for (utils.py:13-46).  GEN_STEP / GEN_TEACHER only report (logits, probs) and return -1; greedy
// returns np.argmax(probs) (first maximum, utils.py:43); sample reproduces utils.py:20-25 - sequential float32 cumsum,
// float64 compare, searchsorted 'left' (the result can be Q).  The drawn class is also written to audio_out (mu-law
// decode LUT) and idx_out.  `scratch`: Q floats of shared memory private to the warp (sample mode).
// Shared by every generation kernel (P = GenParams / ClParams / BcParams).
template <typename P>
__device__ __forceinline__ int warp_softmax_draw(const P& p, int Q, const float (&lg)[8], int b, long long t, int lane,
                                                 float* scratch) {
  const int NQ = Q >> 5;
  float pr[8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) m = fmaxf(m, lg[i]);
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
      if (i < NQ && p.logits_out) p.logits_out[(long long)b * Q + lane + 32 * i] = lg[i];
      if (i < NQ && p.probs_out) p.probs_out[(long long)b * Q + lane + 32 * i] = pr[i];
    }
    return -1;
  }
  if (p.mode == GEN_TEACHER) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < NQ) p.logits_out[((long long)b * p.T + (t - p.t0)) * Q + lane + 32 * i] = lg[i];
    return -1;
  }
  int k;
  if (p.mode == GEN_GREEDY) {
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
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < NQ) scratch[lane + 32 * i] = pr[i];
    __syncwarp();
    int cnt = 0;
    if (lane == 0) {
      const double u = p.uniforms ? p.uniforms[(t - p.t0) * p.B + b]
                                  : counter_uniform(p.seed, (unsigned long long)t, (unsigned long long)(b + p.b_offset));
      float c = 0.f;
      for (int i = 0; i < Q; ++i) {
        c = __fadd_rn(c, scratch[i]);
        cnt += ((double)c < u) ? 1 : 0;
      }
    }
    k = __shfl_sync(0xffffffffu, cnt, 0);
    __syncwarp();
  }
  if (lane == 0) {
    p.audio_out[(long long)b * p.T + (t - p.t0)] = __ldg(p.dec_lut + k);
    if (p.idx_out) p.idx_out[(long long)b * p.T + (t - p.t0)] = k;
  }
  return k;
}

// ---------------------------------------------------------------------------------------
// bulk async copies + mbarrier
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f32_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(f32_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(f32_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f32_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long* bar, unsigned parity, int* err) {
  const unsigned addr = f32_smem_u32(bar);
#pragma unroll 1
  for (int spin = 0; spin < (1 << 26); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
  }
  atomicExch(err, 2);
  return false;
}

// one tile of one stage
struct TileInfo {
  const float* W;   // contiguous [K][NC] weight tile; null: no contraction (last layer's dead residual)
  int K, NC, col0, col1, cb, sb, cls;
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
// buffer class: A (1) for gated conv / post1, B (0) for the rest
__device__ __forceinline__ int stage_class(const GenParams& p, int s) { return (s >= 1 && s <= 2 * p.L + 1) ? (s & 1) : 0; }

__device__ __forceinline__ void stage_tile(const GenParams& p, int s, int tile, TileInfo& ti) {
  const int nsb = p.Bp / FP32_TB;
  ti.cb = tile / nsb;
  ti.sb = tile - ti.cb * nsb;
  ti.col1 = -1;
  ti.cls = stage_class(p, s);
  if (s == 0) {
    ti.K = p.R; ti.NC = 16; ti.col0 = ti.cb * 16;
    ti.W = p.skip0t + (long long)ti.cb * ti.K * 16;
  } else if (s <= 2 * p.L) {
    const int l = (s - 1) >> 1;
    if (s & 1) {
      ti.K = 3 * p.R + p.C; ti.NC = 16;
      ti.col0 = ti.cb * 8; ti.col1 = p.G + ti.cb * 8;      // 8 tanh channels + their 8 sigmoid partners
      ti.W = p.layers[l].w1t + (long long)ti.cb * ti.K * 16;
    } else {
      ti.K = p.G; ti.NC = 32; ti.col0 = ti.cb * 32;
      ti.W = p.layers[l].w2t + (long long)ti.cb * ti.K * 32;
      if (l == p.L - 1 && ti.col0 < p.R) ti.W = nullptr;   // wavenet.py:145: last residual is dead
    }
  } else if (s == 2 * p.L + 1) {
    ti.K = p.S + p.C; ti.NC = 16; ti.col0 = ti.cb * 16;
    ti.W = p.post1t + (long long)ti.cb * ti.K * 16;
  } else {
    ti.K = p.S; ti.NC = 16; ti.col0 = ti.cb * 16;
    ti.W = p.post2t + (long long)ti.cb * ti.K * 16;
  }
}

// partial sums of the (16 streams x NC channels) tile into red_s[kgroup][stream*NC + channel].
// Activations are segment-major: k < Kmain lives in act_s[(k >> sl_log)][stream][k & (SL-1)] (SL = 1 << sl_log
// floats per row, 16 rows per segment); k >= Kmain is the condition tile cond_s[stream][k - Kmain] (C per row).
template <int NC>
__device__ __forceinline__ void tile_compute(const float* wbuf, const float* act_s, int sl_log, int Kmain,
                                             const float* cond_s, int C, int K, float* red_s) {
  constexpr int CQ = NC / 4;          // channel quads
  constexpr int TPG = 2 * CQ;         // threads per K-group (2 stream halves x channel quads)
  constexpr int KG = FP32_THREADS / TPG;
  const int tid = threadIdx.x;
  const int kg = tid / TPG, r = tid - kg * TPG;
  const int sh = r / CQ, q = r - sh * CQ;
  const int klen = K / KG;
  const int k0 = kg * klen;
  const int SL = 1 << sl_log;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
  const float* wp = wbuf + q * 4;
  for (int k = k0; k < k0 + klen; k += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(wp + (k + 0) * NC);
    const float4 w1 = *reinterpret_cast<const float4*>(wp + (k + 1) * NC);
    const float4 w2 = *reinterpret_cast<const float4*>(wp + (k + 2) * NC);
    const float4 w3 = *reinterpret_cast<const float4*>(wp + (k + 3) * NC);
    const float* ap;
    int astr;
    if (k < Kmain) { ap = act_s + ((k >> sl_log) << (sl_log + 4)) + (k & (SL - 1)) + sh * SL; astr = 2 * SL; }
    else { ap = cond_s + (k - Kmain) + sh * C; astr = 2 * C; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(ap + j * astr);   // stream 2j+sh
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

__device__ __forceinline__ void relu_block(float* act_s, int nfloats) {
  for (int q = threadIdx.x; q < (nfloats >> 2); q += FP32_THREADS) {
    float4* a = reinterpret_cast<float4*>(act_s) + q;
    float4 v = *a;
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    *a = v;
  }
}

// ---------------------------------------------------------------------------------------
// load plans.  Every helper is executed by all 256 threads with uniform arguments; thread 0
// posts the expected byte count, threads 0..(n-1) issue one bulk copy each.
// ---------------------------------------------------------------------------------------
constexpr unsigned FP32_WCHUNK = 32768;   // bytes per weight bulk copy (at most 2 per tile)

__device__ __forceinline__ void issue_weights(float* wbuf, const TileInfo& ti, unsigned long long* bar) {
  if (ti.W == nullptr) return;
  const unsigned bytes = (unsigned)ti.K * ti.NC * 4u;
  const int nchunk = (int)((bytes + FP32_WCHUNK - 1) / FP32_WCHUNK);
  if (threadIdx.x == 0) mbar_expect(bar, bytes);
  if ((int)threadIdx.x < nchunk) {
    const unsigned off = threadIdx.x * FP32_WCHUNK;
    const unsigned n = (bytes - off < FP32_WCHUNK) ? (bytes - off) : FP32_WCHUNK;
    bulk_g2s(wbuf + off / 4, ti.W + off / 4, n, bar);
  }
}

// rows that do NOT depend on the previous stage: the two older dilation taps of a gated conv (S1).
// 16 consecutive streams of one ring slot are contiguous in HBM -> one 16 KB copy per tap.
__device__ __forceinline__ bool stage_has_pre(const GenParams& p, int s) { return s >= 1 && s <= 2 * p.L && (s & 1); }

__device__ __forceinline__ void issue_pre_rows(const GenParams& p, int s, const TileInfo& ti, long long t, float* act,
                                               unsigned long long* bar) {
  const int tid = threadIdx.x;
  if (tid < 2) {
    const int which = tid;                                                  // 0: middle tap, 1: oldest tap
    const LayerDev ly = p.layers[(s - 1) >> 1];
    const int d2 = 2 * ly.d;
    const long long ring_slot = (long long)p.Bp * p.R;
    const unsigned bytes = (unsigned)(FP32_TB * p.R * 4);
    if (which == 0) mbar_expect(bar, 2 * bytes);
    const long long slot = (which == 0) ? ((t + ly.d) % d2) : (t % d2);
    bulk_g2s(act + (1 + which) * FP32_TB * p.R, ly.ring + slot * ring_slot + (long long)ti.sb * FP32_TB * p.R, bytes, bar);
  }
}

// rows produced by the previous stage (fetched after the grid barrier): one contiguous [16][len] block
__device__ __forceinline__ bool issue_post_rows(const GenParams& p, int s, const TileInfo& ti, float* act,
                                                unsigned long long* bar) {
  const float* src;
  int len;
  if (s == 0) return false;                                   // computed locally (preprocess FIR)
  if (s <= 2 * p.L) {
    if (s & 1) { src = p.cur + (long long)ti.sb * FP32_TB * p.R; len = p.R; }
    else {
      if (ti.W == nullptr) return false;
      src = p.g + (long long)ti.sb * FP32_TB * p.G; len = p.G;
    }
  } else if (s == 2 * p.L + 1) { src = p.skip + (long long)ti.sb * FP32_TB * p.S; len = p.S; }
  else { src = p.n1 + (long long)ti.sb * FP32_TB * p.S; len = p.S; }
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(FP32_TB * len * 4);
    mbar_expect(bar, bytes);
    bulk_g2s(act, src, bytes, bar);
  }
  return true;
}

// condition tile [16 streams][C] of stream block sb at `frame`; reloaded only when the frame changes
__device__ __forceinline__ void issue_cond_tile(const GenParams& p, int sb, long long frame, float* cond_s,
                                                unsigned long long* bar) {
  const int tid = threadIdx.x;
  if (tid == 0) mbar_expect(bar, (unsigned)(FP32_TB * p.C * 4));
  if (tid < FP32_TB) {
    int b = sb * FP32_TB + tid;
    if (b >= p.B) b = p.B - 1;   // padded streams reuse the last real stream's condition
    bulk_g2s(cond_s + tid * p.C, p.cond + (long long)b * p.cond_bstride + frame * p.C, (unsigned)p.C * 4u, bar);
  }
}

__global__ void __launch_bounds__(FP32_THREADS, 1) wavenet_fp32_persistent(const GenParams p_in) {
  extern __shared__ __align__(128) float smem[];
  // per-layer table (pointers, dilation) cached in shared memory: with ~223 KB of dynamic shared memory the L1 is
  // a few KB, and every table lookup from global memory would cost an L2 round trip on the critical path
  __shared__ LayerDev layers_s[64];
  for (int i = threadIdx.x; i < p_in.L; i += FP32_THREADS) layers_s[i] = p_in.layers[i];
  GenParams p = p_in;
  p.layers = layers_s;
  __syncthreads();
  float* const wA = smem;                            // class A weights
  float* const wB = wA + p.wfloatsA;                 // class B weights
  float* const actA = wB + p.wfloatsB;               // class A activations (segment-major)
  float* const actB = actA + p.actA_floats;          // class B activations
  float* const cond_s = actB + p.actB_floats;        // [16][C] condition tile of this CTA's stream block
  float* red_s = cond_s + FP32_TB * p.C;             // [kgroups][tile outputs]
  float* u_s = red_s + FP32_RED_FLOATS;             // [16][PK]
  float* ps = u_s + FP32_TB * p.PK;                 // [8][Q]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ps + FP32_WARPS * p.Q);
  unsigned long long* wbar = bars;        // [2] weight tile landed (per class)
  unsigned long long* prebar = bars + 2;  // [1] class-A pre rows landed
  unsigned long long* postbar = bars + 3; // [2] post rows landed (per class)
  unsigned long long* condbar = bars + 5; // [1] condition tile landed
  unsigned wphA = 0u, wphB = 0u, preph = 0u, postphA = 0u, postphB = 0u;

  GridBarrier bar{p.barrier, 0ULL, p.err};
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float mu = (float)(p.Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  const int NS = 2 * p.L + 4;
  const int S_P1 = 2 * p.L + 1, S_DRAW = 2 * p.L + 3;
  const long long ring_slot = (long long)p.Bp * p.R;

  if (tid == 0) {
    for (int i = 0; i < 6; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f32_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  long long pf_bar = 0, pf_wait = 0, pf_comp = 0, pf_epi = 0, pf_draw = 0, pf_issue = 0, pf_arrive = 0, pf_pref = 0, pf_t = 0;
  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && tid == 0;
#define PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define PF_ADD(x) do { if (prof) { long long n_ = clock64(); x += n_ - pf_t; pf_t = n_; } } while (0)

  // the first tile of stage `sn` at step `tn` for this CTA: weights + independent rows
  auto prefetch_stage = [&](int sn, long long tn) {
    if ((int)blockIdx.x < stage_tiles(p, sn)) {
      TileInfo tn_i;
      stage_tile(p, sn, blockIdx.x, tn_i);
      issue_weights(tn_i.cls ? wA : wB, tn_i, &wbar[tn_i.cls]);
      if (stage_has_pre(p, sn)) issue_pre_rows(p, sn, tn_i, tn, actA, prebar);
    }
  };

  // preprocess FIR taps of channel `tid`, newest sample first (fast path for the default geometry: one channel
  // per thread, 32 taps); the generic path reads them through the L2 every step, which costs ~50k cycles per step
  const bool fir_regs = (p.R == FP32_THREADS) && (p.PK == FIR_TAPS);
  float fir_k[FIR_TAPS];
  float fir_b = 0.f;
#pragma unroll
  for (int j = 0; j < FIR_TAPS; ++j) fir_k[j] = fir_regs ? __ldg(p.pre_k + (FIR_TAPS - 1 - j) * p.R + tid) : 0.f;
  if (fir_regs) fir_b = __ldg(p.pre_b + tid);

  bool alive = true;
  prefetch_stage(0, p.t0);
  // condition tile: stages that use it (gated conv, post1) map tile -> stream block identically, so one tile
  // per CTA serves a whole frame (ratio steps x 31 stages)
  long long cond_frame = -1;
  int cond_sb = -1;
  unsigned condph = 0u;
  const int sl_R = 31 - __clz(p.R), sl_S = 31 - __clz(p.S), sl_G = 31 - __clz(p.G);

  for (long long t = p.t0; t < p.t0 + p.T && alive; ++t) {
    const long long frame_t = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    for (int s = 0; s < NS && alive; ++s) {
      if (s == S_DRAW) {
        // weights of the next step's first stage (same buffer class as post2, free by now)
        if (t + 1 < p.t0 + p.T) prefetch_stage(0, t + 1);
        // -------------------------------------------------------------- softmax + draw + mu-law decode
        const int NQ = p.Q / 32;   // <= 8
        PF_START();
        for (int b = blockIdx.x * FP32_WARPS + warp; b < p.B; b += gridDim.x * FP32_WARPS) {
          float lg[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) lg[i] = (i < NQ) ? ld_cg(p.logits + (long long)b * p.Q + lane + 32 * i) : -INFINITY;
          const int k = warp_softmax_draw(p, p.Q, lg, b, t, lane, ps + warp * p.Q);
          if (k >= 0 && lane == 0)
            st_cg(p.u_hist + (long long)b * p.PK + (int)((t + 1) % p.PK), __ldg(p.enc_lut + k));   // input of step t+1
        }
        PF_ADD(pf_draw);
        bar.sync();
        PF_ADD(pf_bar);
        continue;
      }

      // ---------------------------------------------------------------- contraction stages
      const int ntiles = stage_tiles(p, s);
      const int l = (s - 1) >> 1;                          // layer of S1/S2 stages
      const int cls = stage_class(p, s);
      float* act_s = cls ? actA : actB;
      bool first = true;
      PF_START();
      for (int tile = blockIdx.x; tile < ntiles && alive; tile += gridDim.x) {
        TileInfo ti;
        stage_tile(p, s, tile, ti);
        float* wcur = cls ? wA : wB;
        const long long row0 = (long long)ti.sb * FP32_TB;
        if (!first) {   // later rounds (more than gridDim tiles): nothing was prefetched for this tile
          issue_weights(wcur, ti, &wbar[cls]);
          if (stage_has_pre(p, s)) issue_pre_rows(p, s, ti, t, act_s, prebar);
        }
        const bool has_post = issue_post_rows(p, s, ti, act_s, &postbar[cls]);
        const bool uses_cond = cls == 1;
        if (uses_cond) {
          const long long frame = frame_t;
          if (frame != cond_frame || ti.sb != cond_sb) {          // uniform across the CTA
            __syncthreads();                                      // nobody still reads the old tile
            issue_cond_tile(p, ti.sb, frame, cond_s, condbar);
            alive = alive && mbar_wait_bounded(condbar, condph, p.err);
            condph ^= 1u;
            cond_frame = frame; cond_sb = ti.sb;
          }
        }

        if (s == 0) {
          // ---- preprocess FIR computed in place (wavenet_ops.py:178,193): h0 = (u0*K[PK-1] + b) + u1*K[PK-2] + ...
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
          if (fir_regs) {
            // thread = channel, taps in registers, the 16 streams' histories broadcast from shared memory
#pragma unroll 4
            for (int i = 0; i < FP32_TB; ++i) {
              const float4* up = reinterpret_cast<const float4*>(u_s + i * FIR_TAPS);
              float acc = fir_b;
#pragma unroll
              for (int j4 = 0; j4 < FIR_TAPS / 4; ++j4) {
                const float4 u4 = up[j4];
                acc = fmaf(u4.x, fir_k[4 * j4 + 0], acc); acc = fmaf(u4.y, fir_k[4 * j4 + 1], acc);
                acc = fmaf(u4.z, fir_k[4 * j4 + 2], acc); acc = fmaf(u4.w, fir_k[4 * j4 + 3], acc);
              }
              act_s[i * p.R + tid] = acc;
            }
          } else {
            for (int idx = tid; idx < FP32_TB * p.R; idx += FP32_THREADS) {
              const int i = idx / p.R, n = idx - i * p.R;
              float acc = fmaf(u_s[i * p.PK], __ldg(p.pre_k + (p.PK - 1) * p.R + n), __ldg(p.pre_b + n));
              for (int j = 1; j < p.PK; ++j)
                acc = fmaf(u_s[i * p.PK + j], __ldg(p.pre_k + (p.PK - 1 - j) * p.R + n), acc);
              act_s[i * p.R + n] = acc;
            }
          }
        }
        first = false;
        PF_ADD(pf_issue);
        // ---- wait for this tile's operands
        if (ti.W != nullptr) {
          alive = alive && mbar_wait_bounded(&wbar[cls], cls ? wphA : wphB, p.err);
          if (cls) wphA ^= 1u; else wphB ^= 1u;
        }
        if (stage_has_pre(p, s)) { alive = alive && mbar_wait_bounded(prebar, preph, p.err); preph ^= 1u; }
        if (has_post) {
          alive = alive && mbar_wait_bounded(&postbar[cls], cls ? postphA : postphB, p.err);
          if (cls) postphA ^= 1u; else postphB ^= 1u;
        }
        __syncthreads();
        PF_ADD(pf_wait);
        if (s >= S_P1) {         // relu on the contraction input (wavenet.py:153,163)
          relu_block(act_s, FP32_TB * p.S);
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
            st_cg(p.cur + (row0 + i) * p.R + n, act_s[i * p.R + n]);
          }
          {
            const int slg = (s == 0 || (s <= 2 * p.L)) ? sl_R : sl_S;
            const int kmain = (s == 0) ? p.R : (s <= 2 * p.L) ? 3 * p.R : p.S;
            tile_compute<16>(wcur, act_s, slg, kmain, cond_s, p.C, ti.K, red_s);
          }
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
            tile_compute<32>(wcur, act_s, sl_G, p.G, cond_s, p.C, ti.K, red_s);
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
      // publish this stage's results, then request the NEXT stage's weights and independent rows while the
      // barrier completes (they do not depend on other CTAs' data and this stage's buffers of that class are idle)
      PF_START();
      bar.arrive();
      PF_ADD(pf_arrive);
      if (s + 1 != S_DRAW) prefetch_stage(s + 1, t);
      PF_ADD(pf_pref);
      bar.wait();
      PF_ADD(pf_bar);
    }
  }
  if (prof) {
    p.prof[0] = pf_bar; p.prof[1] = pf_wait; p.prof[2] = pf_comp; p.prof[3] = pf_epi; p.prof[4] = pf_draw; p.prof[5] = pf_issue; p.prof[6] = pf_arrive; p.prof[7] = pf_pref;
  }
}

// [K][ldw] row-major weights -> tile-major [tile][K][NC]; tile cb takes columns col0(cb)..+NC
// (paired: 8 columns from cb*8 and 8 from G + cb*8)
__global__ void pack_tiles_kernel(const float* __restrict__ src, int ldw, int K, int NC, int ntiles, int paired_G,
                                  float* __restrict__ dst) {
  const long long total = (long long)ntiles * K * NC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % NC);
    const long long r = i / NC;
    const int k = (int)(r % K);
    const int cb = (int)(r / K);
    int col;
    if (paired_G > 0) col = (c < 8) ? (cb * 8 + c) : (paired_G + cb * 8 + (c - 8));
    else col = cb * NC + c;
    dst[i] = src[(long long)k * ldw + col];
  }
}

}  // namespace vqwn
