// WaveNet fast generation, float32, DATAFLOW variant of the persistent kernel (wavenet_fp32.cuh).
//
// Same arithmetic, same tiles, same weight/tap streaming as wavenet_fp32_persistent, but there is
// NO grid barrier: every value one CTA hands to another travels as an 8-byte packet {float, tag}
// written with one 64-bit store and polled with 64-bit L2 loads by the consumers (tag = step+1).
// A stage's tile starts as soon as the packets it needs carry the current tag, so the fence +
// atomic + flag round trips of a barrier (about 3.5k cycles per stage) disappear from the
// dependency chain of a step.  Per time step the chain is the same 63 stages:
//   preprocess FIR + skip start -> 30 x (gated conv -> residual/skip) -> post1 -> post2 -> draw
// Requirements: every stage's tile count <= gridDim (B <= 64 streams on 148 SMs), so tile index ==
// blockIdx for every stage and a CTA owns the same channels of the same streams in every layer:
//   * the residual stream slice and the skip accumulator slice of a CTA stay in REGISTERS across
//     the 30 layers (the barrier kernel round-trips them through L2 every layer);
//   * the preprocess FIR history of a CTA's streams stays in shared memory across steps.
// Buffers are per layer, so a packet is overwritten one full step after it was consumed (every
// consumer of step t has finished before the draw of step t, which precedes all writes of t+1).
// Reference semantics: see wavenet_fp32.cuh.
#pragma once
#include "wavenet_fp32.cuh"

namespace vqwn {

typedef unsigned long long ll_packet;   // low 32 bits: float value, high 32 bits: tag

struct DfParams {
  GenParams g;
  ll_packet* cur_ll;    // [L][Bp][R]   input of layer l
  ll_packet* g_ll;      // [L][Bp][G]   gated output of layer l
  ll_packet* skip0_ll;  // [Bp][S]      skip start (preprocess -> first layer's skip owners)
  ll_packet* skip_ll;   // [Bp][S]      skip sum (last layer -> post1)
  ll_packet* n1_ll;     // [Bp][S]
  ll_packet* logit_ll;  // [Bp][Q]
  ll_packet* u_ll;      // [Bp]         next network input (draw -> preprocess of the next step)
};

__device__ __forceinline__ ll_packet ll_pack(float v, unsigned tag) {
  return ((ll_packet)tag << 32) | (ll_packet)__float_as_uint(v);
}
__device__ __forceinline__ void ll_store(ll_packet* p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(ll_pack(v, tag)) : "memory");
}
__device__ __forceinline__ ll_packet ll_load(const ll_packet* p) {
  ll_packet v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// one packet, spin until it carries `tag` (bounded by wall-clock cycles; a dead producer becomes an error)
__device__ __forceinline__ float ll_wait_one(const ll_packet* p, unsigned tag, int* err, bool& alive) {
  ll_packet v = ll_load(p);
  if ((unsigned)(v >> 32) != tag) {
    const long long t0 = clock64();
    do {
      v = ll_load(p);
      if ((unsigned)(v >> 32) == tag) break;
      if (clock64() - t0 > 2000000000LL || *reinterpret_cast<volatile int*>(err) != 0) {
        atomicExch(err, 4);
        alive = false;
        break;
      }
    } while (true);
  }
  return __uint_as_float((unsigned)v);
}

// block-cooperative: n packets (multiple of 256) -> n floats in shared memory
template <int PER_THREAD>
__device__ __forceinline__ void ll_read_block(const ll_packet* src, unsigned tag, float* dst, int* err, bool& alive) {
  ll_packet v[PER_THREAD];
#pragma unroll
  for (int j = 0; j < PER_THREAD; ++j) v[j] = ll_load(src + threadIdx.x + j * FP32_THREADS);
#pragma unroll
  for (int j = 0; j < PER_THREAD; ++j) {
    float f;
    if ((unsigned)(v[j] >> 32) == tag) f = __uint_as_float((unsigned)v[j]);
    else f = ll_wait_one(src + threadIdx.x + j * FP32_THREADS, tag, err, alive);
    dst[threadIdx.x + j * FP32_THREADS] = f;
  }
}
__device__ __forceinline__ void ll_read_block_n(const ll_packet* src, int n, unsigned tag, float* dst, int* err, bool& alive) {
  // n in {4096, 8192}
  for (int base = 0; base < n; base += 8 * FP32_THREADS) ll_read_block<8>(src + base, tag, dst + base, err, alive);
}

__global__ void __launch_bounds__(FP32_THREADS, 1) wavenet_fp32_dataflow(const DfParams dp) {
  extern __shared__ __align__(128) float smem[];
  __shared__ LayerDev layers_s[64];
  for (int i = threadIdx.x; i < dp.g.L; i += FP32_THREADS) layers_s[i] = dp.g.layers[i];
  GenParams p = dp.g;
  p.layers = layers_s;
  __syncthreads();

  float* const wA = smem;
  float* const wB = wA + p.wfloatsA;
  float* const actA = wB + p.wfloatsB;
  float* const actB = actA + p.actA_floats;
  float* const cond_s = actB + p.actB_floats;
  float* red_s = cond_s + FP32_TB * p.C;
  float* u_s = red_s + FP32_RED_FLOATS;             // [16][PK] FIR history of this CTA's streams (persistent)
  float* ps = u_s + FP32_TB * p.PK;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(ps + FP32_WARPS * p.Q);
  unsigned long long* wbar = bars;        // [2]
  unsigned long long* prebar = bars + 2;
  unsigned long long* condbar = bars + 5;
  unsigned wphA = 0u, wphB = 0u, preph = 0u, condph = 0u;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float mu = (float)(p.Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  const int NS = 2 * p.L + 4;
  const int S_P1 = 2 * p.L + 1, S_P2 = 2 * p.L + 2, S_DRAW = 2 * p.L + 3;
  const long long ring_slot = (long long)p.Bp * p.R;
  const int nsb = p.Bp / FP32_TB;
  const int tile = blockIdx.x;             // this CTA's tile index in EVERY stage
  const int sl_R = 31 - __clz(p.R), sl_S = 31 - __clz(p.S), sl_G = 31 - __clz(p.G);

  if (tid == 0) {
    for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f32_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // FIR history of the streams of this CTA's preprocess tile (u_hist holds it between launches)
  const bool has_st0 = tile < stage_tiles(p, 0);
  const int st0_sb = tile % nsb, st0_cb = tile / nsb;
  if (has_st0)
    for (int idx = tid; idx < FP32_TB * p.PK; idx += FP32_THREADS)
      u_s[idx] = p.u_hist[(long long)(st0_sb * FP32_TB) * p.PK + idx];
  __syncthreads();

  bool alive = true;
  // has this CTA a tile in stage s?
  auto has_tile = [&](int s) { return s != S_DRAW && tile < stage_tiles(p, s); };
  // next stage (same or later step) in which this CTA owns a contraction tile
  auto next_stage = [&](int s, long long t, int& sn, long long& tn) {
    sn = s; tn = t;
    for (int i = 0; i < NS; ++i) {
      ++sn;
      if (sn == NS) { sn = 0; ++tn; }
      if (has_tile(sn)) return true;
    }
    return false;
  };
  auto prefetch = [&](int sn, long long tn) {
    if (tn >= p.t0 + p.T) return;
    TileInfo tn_i;
    stage_tile(p, sn, tile, tn_i);
    issue_weights(tn_i.cls ? wA : wB, tn_i, &wbar[tn_i.cls]);
    if (stage_has_pre(p, sn)) issue_pre_rows(p, sn, tn_i, tn, actA, prebar);
  };
  {
    int s0 = -1; long long tt = p.t0;
    int sn; long long tn;
    if (next_stage(s0, tt, sn, tn)) prefetch(sn, tn);
  }

  long long cond_frame = -1;
  float own[2] = {0.f, 0.f};     // residual-stream slice or skip-sum slice owned by this CTA (2 values per thread)

  for (long long t = p.t0; t < p.t0 + p.T && alive; ++t) {
    const unsigned tag = (unsigned)(t + 1);
    const long long frame = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    for (int s = 0; s < NS && alive; ++s) {
      if (s == S_DRAW) {
        // -------------------------------------------------------------- softmax + draw + mu-law decode
        const int NQ = p.Q / 32;
        for (int b = blockIdx.x * FP32_WARPS + warp; b < p.B; b += gridDim.x * FP32_WARPS) {
          float lg[8], pr[8];
          float m = -INFINITY;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            lg[i] = (i < NQ) ? ll_wait_one(dp.logit_ll + (long long)b * p.Q + lane + 32 * i, tag, p.err, alive) : -INFINITY;
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
            const float un = __ldg(p.enc_lut + k);
            ll_store(dp.u_ll + b, un, tag);                       // consumed by the preprocess stage of step t+1
            st_cg(p.u_hist + (long long)b * p.PK + (int)((t + 1) % p.PK), un);   // and kept for a later launch
          }
        }
        continue;
      }
      if (!has_tile(s)) continue;

      // ---------------------------------------------------------------- contraction stage, this CTA's tile
      TileInfo ti;
      stage_tile(p, s, tile, ti);
      const int l = (s - 1) >> 1;
      const int cls = ti.cls;
      float* act_s = cls ? actA : actB;
      float* wcur = cls ? wA : wB;
      const long long row0 = (long long)ti.sb * FP32_TB;
      int sn; long long tn;
      const bool have_next = next_stage(s, t, sn, tn);
      const bool next_same_class = have_next && stage_class(p, sn) == cls;

      // ---- inputs produced by other CTAs (packets)
      if (s == 0) {
        // newest network input of this tile's 16 streams -> FIR history column t % PK
        if (tid < FP32_TB) {
          const int b = ti.sb * FP32_TB + tid;
          float u = 0.f;
          if (ext) {
            float x = 0.f;
            if (b < p.B) {
              if (p.mode == GEN_STEP) x = p.ext_audio[b];
              else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
            }
            u = mu_law_encode_dev(x, mu, 0.f);
          } else if (t > p.t0 && b < p.B) {
            u = ll_wait_one(dp.u_ll + b, (unsigned)t, p.err, alive);   // tag of step t-1
          } else if (t == p.t0 && t > 0 && b < p.B) {
            u = p.u_hist[(long long)b * p.PK + (int)(t % p.PK)];       // continued generation: stored by the last launch
          }
          u_s[tid * p.PK + (int)(t % p.PK)] = u;
        }
        __syncthreads();
        // h0[i][n] = (u[t]*K[PK-1] + b) + u[t-1]*K[PK-2] + ...   (wavenet_ops.py:178,193)
        for (int idx = tid; idx < FP32_TB * p.R; idx += FP32_THREADS) {
          const int i = idx / p.R, n = idx - i * p.R;
          float acc = fmaf(u_s[i * p.PK + (int)(t % p.PK)], __ldg(p.pre_k + (p.PK - 1) * p.R + n), __ldg(p.pre_b + n));
          for (int j = 1; j < p.PK; ++j) {
            int sl = (int)((t - j) % p.PK);
            if (sl < 0) sl += p.PK;
            acc = fmaf(u_s[i * p.PK + sl], __ldg(p.pre_k + (p.PK - 1 - j) * p.R + n), acc);
          }
          act_s[i * p.R + n] = acc;
        }
      } else if (s <= 2 * p.L) {
        if (s & 1) ll_read_block_n(dp.cur_ll + ((long long)l * p.Bp + row0) * p.R, FP32_TB * p.R, tag, act_s, p.err, alive);
        else if (ti.W != nullptr) ll_read_block_n(dp.g_ll + ((long long)l * p.Bp + row0) * p.G, FP32_TB * p.G, tag, act_s, p.err, alive);
      } else if (s == S_P1) {
        ll_read_block_n(dp.skip_ll + row0 * p.S, FP32_TB * p.S, tag, act_s, p.err, alive);
      } else {
        ll_read_block_n(dp.n1_ll + row0 * p.S, FP32_TB * p.S, tag, act_s, p.err, alive);
      }
      // ---- condition tile (gated conv / post1), reloaded when the frame changes
      if (cls == 1 && frame != cond_frame) {
        __syncthreads();
        issue_cond_tile(p, ti.sb, frame, cond_s, condbar);
        alive = alive && mbar_wait_bounded(condbar, condph, p.err);
        condph ^= 1u;
        cond_frame = frame;
      }
      // ---- operands streamed by the copy engine
      if (ti.W != nullptr) {
        alive = alive && mbar_wait_bounded(&wbar[cls], cls ? wphA : wphB, p.err);
        if (cls) wphA ^= 1u; else wphB ^= 1u;
      }
      if (stage_has_pre(p, s)) { alive = alive && mbar_wait_bounded(prebar, preph, p.err); preph ^= 1u; }
      __syncthreads();
      if (have_next && !next_same_class) prefetch(sn, tn);     // other buffer class: safe to stream during the math
      if (s >= S_P1) {
        relu_block(act_s, FP32_TB * p.S);
        __syncthreads();
      }

      // ---- contraction + epilogue
      if (ti.NC == 16) {
        const int i = tid >> 4, c = tid & 15;
        const int col = (ti.col1 >= 0 && c >= 8) ? (ti.col1 + c - 8) : (ti.col0 + c);
        const float* bias_p = (s == 0) ? p.skip0_b : (s <= 2 * p.L) ? p.layers[l].b1 : (s == S_P1) ? p.post1_b : p.post2_b;
        const float bias = __ldg(bias_p + col);
        const long long row = row0 + i;
        if (s == 0 && ti.cb < p.R / 16) {
          const int n = ti.cb * 16 + c;
          ll_store(dp.cur_ll + row * p.R + n, act_s[i * p.R + n], tag);          // input of layer 0
        }
        {
          const int slg = (s <= 2 * p.L) ? sl_R : sl_S;
          const int kmain = (s == 0) ? p.R : (s <= 2 * p.L) ? 3 * p.R : p.S;
          tile_compute<16>(wcur, act_s, slg, kmain, cond_s, p.C, ti.K, red_s);
        }
        __syncthreads();
        if (have_next && next_same_class) prefetch(sn, tn);    // same class: only after the math has read it
        float v = tile_reduce<16>(red_s, tid) + bias;
        if (s == 0) {
          ll_store(dp.skip0_ll + row * p.S + col, v, tag);
        } else if (s <= 2 * p.L) {
          const float partner = __shfl_down_sync(0xffffffffu, v, 8);
          if (c < 8) ll_store(dp.g_ll + ((long long)l * p.Bp + row) * p.G + col, tanhf(v) * sigmoid_f(partner), tag);
        } else if (s == S_P1) {
          ll_store(dp.n1_ll + row * p.S + col, v, tag);
        } else {
          ll_store(dp.logit_ll + row * p.Q + col, v, tag);
        }
      } else {
        // residual / skip owners: 2 values per thread, kept in registers from layer to layer
        const LayerDev ly = p.layers[l];
        const int slot_old = (int)(t % (2 * ly.d));
        const bool is_res = ti.col0 < p.R;
        float bias[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o = tid + h * FP32_THREADS;
          const long long row = row0 + (o >> 5);
          const int col = ti.col0 + (o & 31);
          bias[h] = __ldg(ly.b2 + col);
          if (l == 0) {
            own[h] = is_res ? ll_wait_one(dp.cur_ll + row * p.R + col, tag, p.err, alive)
                            : ll_wait_one(dp.skip0_ll + row * p.S + (col - p.R), tag, p.err, alive);
          }
        }
        if (ti.W != nullptr) {
          tile_compute<32>(wcur, act_s, sl_G, p.G, cond_s, p.C, ti.K, red_s);
          __syncthreads();
        }
        if (have_next && next_same_class) prefetch(sn, tn);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o = tid + h * FP32_THREADS;
          const long long row = row0 + (o >> 5);
          const int col = ti.col0 + (o & 31);
          if (is_res) {
            st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, own[h]);      // push_ops (layer input of step t)
            if (ti.W != nullptr) {
              own[h] = own[h] + (tile_reduce<32>(red_s, o) + bias[h]);
              ll_store(dp.cur_ll + ((long long)(l + 1) * p.Bp + row) * p.R + col, own[h], tag);
            }
          } else {
            own[h] = own[h] + (tile_reduce<32>(red_s, o) + bias[h]);
            if (l == p.L - 1) ll_store(dp.skip_ll + row * p.S + (col - p.R), own[h], tag);
          }
        }
      }
      __syncthreads();
    }
  }
  // FIR history back to global for a later launch (one writer per stream block)
  if (has_st0 && st0_cb == 0) {
    const int skip_col = (int)((p.t0 + p.T) % p.PK);     // written by the draw stage (input of the next step)
    for (int idx = tid; idx < FP32_TB * p.PK; idx += FP32_THREADS)
      if (idx % p.PK != skip_col) p.u_hist[(long long)(st0_sb * FP32_TB) * p.PK + idx] = u_s[idx];
  }
}

}  // namespace vqwn
