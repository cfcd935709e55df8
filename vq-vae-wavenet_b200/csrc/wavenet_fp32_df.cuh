// WaveNet fast generation, float32, DATAFLOW variant of the persistent kernel (wavenet_fp32.cuh).
//
// Same arithmetic, same tiles, same weight/tap streaming as wavenet_fp32_persistent, but NO grid
// barrier: dependencies are tracked per (stage, 16-stream block) with monotonic counters.  A tile
// that has stored its outputs does bar.sync + one red.release.gpu on the counter of its stage and
// stream block; a consumer tile polls (ld.acquire.gpu, one thread) until all producers of the stage
// it depends on have signalled, then that same thread issues the bulk copy of the rows it needs.
// Compared with the barrier kernel a hand-off loses the "last arriver -> release flag -> everybody"
// round trip and only waits for the 8-32 producers it really depends on instead of all 148 CTAs.
// Requirements: every stage's tile count <= gridDim (B <= 64 streams on 148 SMs), so tile index ==
// blockIdx in every stage and a CTA owns the same channels of the same streams in every layer:
//   * the residual-stream slice and the skip-sum slice of a CTA stay in REGISTERS across the 30
//     layers (the barrier kernel round-trips them through L2 every layer);
//   * the preprocess FIR history of a CTA's streams stays in shared memory across steps.
// cur and g are per-layer buffers, so a buffer is overwritten one full step after its last reader
// (every consumer of step t has finished before the draw of step t, which precedes step t+1).
// Reference semantics: see wavenet_fp32.cuh.
#pragma once
#include "wavenet_fp32.cuh"

namespace vqwn {

struct DfParams {
  GenParams g;
  float* cur_l;      // [L][Bp][R]  input of layer l
  float* g_l;        // [L][Bp][G]  gated output of layer l
  float* skip0;      // [Bp][S]     skip start (preprocess -> first layer's skip owners)
  unsigned* cnt;     // [2L+4][nsb] arrivals per (stage, stream block), zeroed at launch
  unsigned* cnt_res; // [L][nsb]    arrivals of the residual tiles of S2_l only
};

__device__ __forceinline__ void df_signal(unsigned* c) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
}
// one thread: wait until *c >= target (bounded by cycles; a dead producer becomes an error)
__device__ __forceinline__ bool df_wait(const unsigned* c, unsigned target, int* err) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
  if (v >= target) return true;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
    if (v >= target) return true;
    if (clock64() - t0 > 2000000000LL || *reinterpret_cast<volatile int*>(err) != 0) {
      atomicExch(err, 4);
      return false;
    }
  }
}

__global__ void __launch_bounds__(FP32_THREADS, 1) wavenet_fp32_dataflow(const DfParams dp) {
  extern __shared__ __align__(128) float smem[];
  __shared__ LayerDev layers_s[64];
  __shared__ int alive_s;
  for (int i = threadIdx.x; i < dp.g.L; i += FP32_THREADS) layers_s[i] = dp.g.layers[i];
  GenParams p = dp.g;
  p.layers = layers_s;
  if (threadIdx.x == 0) alive_s = 1;
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
  unsigned long long* postbar = bars + 3; // [2]
  unsigned long long* condbar = bars + 5;
  unsigned wphA = 0u, wphB = 0u, preph = 0u, postphA = 0u, postphB = 0u, condph = 0u;

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

  const bool pf = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
  long long pc[6] = {0, 0, 0, 0, 0, 0};
  long long pt = clock64();
#define DF_PF(i) do { if (pf) { long long n_ = clock64(); pc[i] += n_ - pt; pt = n_; } } while (0)
  bool alive = true;
  auto has_tile = [&](int s) { return s != S_DRAW && tile < stage_tiles(p, s); };
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
    const unsigned trel = (unsigned)(t - p.t0 + 1);       // counters are zeroed at launch
    const long long frame = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    for (int s = 0; s < NS && alive; ++s) {
      if (s == S_DRAW) {
        // -------------------------------------------------------------- softmax + draw + mu-law decode
        const int NQ = p.Q / 32;
        for (int b = blockIdx.x * FP32_WARPS + warp; b < p.B; b += gridDim.x * FP32_WARPS) {
          // all post2 tiles of this stream's block have published their logits
          if (lane == 0) alive = df_wait(dp.cnt + (size_t)S_P2 * nsb + b / FP32_TB, (unsigned)(p.Q / 16) * trel, p.err) && alive;
          alive = __shfl_sync(0xffffffffu, alive ? 1 : 0, 0) != 0;
          float lg[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) lg[i] = (i < NQ) ? ld_cg(p.logits + (long long)b * p.Q + lane + 32 * i) : -INFINITY;
          const int k = warp_softmax_draw(p, p.Q, lg, b, t, lane, ps + warp * p.Q);
          if (k >= 0 && lane == 0) {
            st_cg(p.u_hist + (long long)b * p.PK + (int)((t + 1) % p.PK), __ldg(p.enc_lut + k));   // input of step t+1
            df_signal(dp.cnt + (size_t)S_DRAW * nsb + b / FP32_TB);
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

      // ---- wait for the producers this tile depends on, then request the rows they produced (thread 0)
      bool has_post = false;
      DF_PF(5);
      if (s == 0) {
        // newest network input of this tile's 16 streams -> FIR history column t % PK
        if (!ext && t > p.t0) {
          if (tid == 0) alive_s = df_wait(dp.cnt + (size_t)S_DRAW * nsb + ti.sb, (unsigned)min(FP32_TB, p.B - ti.sb * FP32_TB) * (trel - 1), p.err) ? 1 : 0;
          __syncthreads();
          alive = alive_s != 0;
        }
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
          } else if (b < p.B && t > 0) {
            u = ld_cg(p.u_hist + (long long)b * p.PK + (int)(t % p.PK));   // written by the draw of step t-1
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
      } else {
        const float* src = nullptr;
        int len = 0;
        const unsigned* c = nullptr;
        unsigned target = 0;
        if (s <= 2 * p.L) {
          if (s & 1) {
            src = dp.cur_l + ((long long)l * p.Bp + row0) * p.R; len = p.R;
            if (l == 0) { c = dp.cnt + ti.sb; target = (unsigned)(p.S / 16) * trel; }      // all preprocess tiles
            else { c = dp.cnt_res + (size_t)(l - 1) * nsb + ti.sb; target = (unsigned)(p.R / 32) * trel; }
          } else {
            c = dp.cnt + (size_t)(s - 1) * nsb + ti.sb; target = (unsigned)(p.G / 8) * trel;   // all gated tiles of this layer
            if (ti.W != nullptr) { src = dp.g_l + ((long long)l * p.Bp + row0) * p.G; len = p.G; }
          }
        } else if (s == S_P1) {
          src = p.skip + row0 * p.S; len = p.S;
          c = dp.cnt + (size_t)(2 * p.L) * nsb + ti.sb; target = (unsigned)(p.S / 32) * trel;     // skip tiles of the last layer
        } else {
          src = p.n1 + row0 * p.S; len = p.S;
          c = dp.cnt + (size_t)S_P1 * nsb + ti.sb; target = (unsigned)(p.S / 16) * trel;
        }
        if (tid == 0) {
          const bool ok = df_wait(c, target, p.err);
          alive_s = ok ? 1 : 0;
          if (ok && src != nullptr) {
            const unsigned bytes = (unsigned)(FP32_TB * len * 4);
            mbar_expect(&postbar[cls], bytes);
            bulk_g2s(act_s, src, bytes, &postbar[cls]);
          }
        }
        has_post = src != nullptr;
      }
      DF_PF(0);
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
      __syncthreads();                                   // alive_s and (for stage 0) the FIR tile are visible
      alive = alive && alive_s != 0;
      if (has_post && alive) {
        alive = mbar_wait_bounded(&postbar[cls], cls ? postphA : postphB, p.err);
        if (cls) postphA ^= 1u; else postphB ^= 1u;
      }
      DF_PF(1);
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
          st_cg(dp.cur_l + row * p.R + n, act_s[i * p.R + n]);          // input of layer 0
        }
        {
          const int slg = (s <= 2 * p.L) ? sl_R : sl_S;
          const int kmain = (s == 0) ? p.R : (s <= 2 * p.L) ? 3 * p.R : p.S;
          tile_compute<16>(wcur, act_s, slg, kmain, cond_s, p.C, ti.K, red_s);
        }
        __syncthreads();
        DF_PF(2);
        float v = tile_reduce<16>(red_s, tid) + bias;
        if (s == 0) {
          st_cg(dp.skip0 + row * p.S + col, v);
        } else if (s <= 2 * p.L) {
          const float partner = __shfl_down_sync(0xffffffffu, v, 8);
          if (c < 8) st_cg(dp.g_l + ((long long)l * p.Bp + row) * p.G + col, tanhf(v) * sigmoid_f(partner));
        } else if (s == S_P1) {
          st_cg(p.n1 + row * p.S + col, v);
        } else {
          st_cg(p.logits + row * p.Q + col, v);
        }
        __syncthreads();                                       // every thread's stores precede the signal
        if (tid == 0) df_signal(dp.cnt + (size_t)s * nsb + ti.sb);
        DF_PF(3);
        if (have_next) prefetch(sn, tn);   // weights + older taps of this CTA's next tile: overlaps the wait for its producers
        DF_PF(4);
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
          if (l == 0) own[h] = is_res ? ld_cg(dp.cur_l + row * p.R + col) : ld_cg(dp.skip0 + row * p.S + (col - p.R));
        }
        if (ti.W != nullptr) {
          tile_compute<32>(wcur, act_s, sl_G, p.G, cond_s, p.C, ti.K, red_s);
          __syncthreads();
        }
        DF_PF(2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o = tid + h * FP32_THREADS;
          const long long row = row0 + (o >> 5);
          const int col = ti.col0 + (o & 31);
          if (is_res) {
            st_cg(ly.ring + slot_old * ring_slot + row * p.R + col, own[h]);      // push_ops (layer input of step t)
            if (ti.W != nullptr) {
              own[h] = own[h] + (tile_reduce<32>(red_s, o) + bias[h]);
              st_cg(dp.cur_l + ((long long)(l + 1) * p.Bp + row) * p.R + col, own[h]);
            }
          } else {
            own[h] = own[h] + (tile_reduce<32>(red_s, o) + bias[h]);
            if (l == p.L - 1) st_cg(p.skip + row * p.S + (col - p.R), own[h]);
          }
        }
        __syncthreads();
        if (tid == 0) {
          if (is_res) df_signal(dp.cnt_res + (size_t)l * nsb + ti.sb);
          else df_signal(dp.cnt + (size_t)s * nsb + ti.sb);
        }
        DF_PF(3);
        if (have_next) prefetch(sn, tn);
        DF_PF(4);
      }
    }
  }
  if (pf) for (int i = 0; i < 6; ++i) p.prof[8 + i] = pc[i];
  // FIR history back to global for a later launch (one writer per stream block)
  if (has_st0 && st0_cb == 0) {
    const int skip_col = (int)((p.t0 + p.T) % p.PK);     // written by the draw stage (input of the next step)
    for (int idx = tid; idx < FP32_TB * p.PK; idx += FP32_THREADS)
      if (idx % p.PK != skip_col) p.u_hist[(long long)(st0_sb * FP32_TB) * p.PK + idx] = u_s[idx];
  }
}

}  // namespace vqwn
