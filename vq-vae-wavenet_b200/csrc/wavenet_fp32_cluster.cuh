// WaveNet fast generation, float32, thread-block-cluster form: the synchronisation domain of a
// time step is ONE CLUSTER of 16 CTAs instead of the whole grid.
//
// Reference semantics (file:line relative to the reference root) are those of wavenet_fp32.cuh:
//   wavenet.py:103-172 (one-step graph), wavenet_ops.py:163-267 (fast_conv1d queues, condition,
//   gate, skip / residual 1x1), utils.py:13-46 (draw), mu_law_ops.py:5-31 (mu-law).
//
// Why: a step is a chain of 2L+3 dependent contractions.  With the whole GPU working on every
// link of the chain (wavenet_fp32.cuh) each hand-off is a grid barrier through the L2 plus a
// reload of the activations (~5k cycles, more than the contraction itself).  Here a cluster of
// 16 CTAs owns MS streams (MS <= 10) for the whole run and splits every stage by OUTPUT
// CHANNELS: CTA r computes channels [r*N/16, (r+1)*N/16) of all MS streams.  The activations
// never leave shared memory: a CTA pushes its slice of the stage output into the shared memory
// of all 16 CTAs with st.async.shared::cluster ... mbarrier::complete_tx (SASS STAS): every store
// also counts its bytes on the RECEIVER's mbarrier, so a receiver waits on its own barrier for the
// bytes of all 16 senders - no cluster-wide barrier per hand-off and no release stall on the sender
// (two hardware cluster barriers per step remain: logits -> CTA 0, new sample -> all; they also
// order the HBM ring stores against later bulk-copy reads).  Clusters are independent of each
// other: no grid-wide synchronisation exists.  The price is that every cluster streams all float32
// weights (78.6 MB) once per step, 4.9 MB per CTA, through cp.async.bulk with an L2 evict-last
// hint, one stage ahead (S1 buffer / S2 buffer); they do not stay in L2 (85 MB of DRAM reads per
// step, profiles/r1_cluster_full_summary.txt).
//
// Stage -> per-CTA tile (MS streams x NC columns, K), default geometry:
//   skip start   MS x 32 skip channels,                 K = R          (input: local FIR, no hand-off)
//   S1 (gated)   MS x (16 tanh + 16 sigmoid partners),  K = 3R + C     -> g slice pushed to all
//   S2           MS x (16 residual + 32 skip),          K = G          -> new layer input pushed to all,
//                                                                         skip slice accumulates locally,
//                                                                         queue push to the HBM ring
//   post1        MS x 32,                               K = S + C      -> relu slice pushed to all
//   post2        MS x 16 logits,                        K = S          -> pushed to CTA 0 of the cluster
//   draw         CTA 0: softmax + greedy/sample draw, mu-law; new sample pushed to all
// Dilation queues: the same HBM rings as the barrier kernel ([2d][Bp][R] per layer), so the two
// kernels can continue each other's state (step API).
#pragma once
#include "wavenet_fp32.cuh"

namespace vqwn {

constexpr int CL_CS = 16;            // CTAs per cluster (non-portable size)
constexpr int CL_THREADS = 256;
constexpr int CL_MAX_MS = 10;        // streams per cluster

struct ClLayerDev {
  const float* w1c;   // [CS][3R+C][2G/CS]   columns: G/CS tanh channels | their sigmoid partners
  const float* b1;    // [2G]
  const float* w2c;   // [CS][G][R/CS + S/CS] columns: residual slice | skip slice
  const float* b2;    // [R+S]
  float* ring;        // [2d][Bp][R]
  int d;
  int pad_;
};

struct ClParams {
  int L, R, G, S, Q, C, PK;
  int B, Bp;
  int cluster0;             // first cluster of this launch (batches above the co-resident capacity run as several launches)
  int kg_s0, kg_s1, kg_s2, kg_p1, kg_p2;   // K-groups per stage (host-chosen: kg * NC/4 <= 256, K % (4 kg) == 0)
  int w1_floats, w2_floats;                // shared-memory weight buffers
  const float *pre_k, *pre_b, *skip0c, *skip0_b, *post1c, *post1_b, *post2c, *post2_b;
  const ClLayerDev* layers;
  const float *enc_lut, *dec_lut;
  float* u_hist;            // [Bp][PK] persistent input history (shared with the barrier kernel)
  long long t0, T;
  int mode;
  const float* cond;
  long long cond_bstride;
  int ratio;
  const float* ext_audio;
  const double* uniforms;
  unsigned long long seed;
  int b_offset;
  float* audio_out;
  int* idx_out;
  float* logits_out;
  float* probs_out;
  long long* prof;
  int* err;
};

__device__ __forceinline__ unsigned cl_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_v4(unsigned addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cl_st_f32(unsigned addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// hand-off: every CTA's pushes (and global ring stores) before the barrier are visible to every CTA after it
__device__ __forceinline__ void cl_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// remote store that also counts its 16 bytes on the DESTINATION CTA's mbarrier (complete_tx): the receiver waits on
// its own barrier for the bytes of all 16 senders, no cluster-wide barrier and no release stall on the sender
__device__ __forceinline__ void cl_st_async_v4(unsigned addr, float4 v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar) : "memory");
}
// weight tiles are re-read by every cluster every step: keep them in L2 ahead of the streaming queue traffic
__device__ __forceinline__ void cl_bulk_g2s_keep(float* smem_dst, const float* gsrc, unsigned bytes, unsigned long long* bar) {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(f32_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(f32_smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// push geometry of a thread for slices of `nvec` float4 with `vec_per_row` float4 per stream row: thread groups of nvec
// threads take every ngroups-th destination CTA.  Packed once per kernel (no integer divisions in the step loop):
// j | i << 8 | grp << 16 | ngroups << 24; grp = 255: thread idle.
__device__ __forceinline__ unsigned cl_push_geo(int nvec, int vec_per_row) {
  const int tid = threadIdx.x;
  const int ngroups = CL_THREADS / nvec;          // nvec <= 256 (host-checked)
  const int grp = tid / nvec, v = tid - grp * nvec;
  const int i = v / vec_per_row, j = v - i * vec_per_row;
  return (unsigned)j | ((unsigned)i << 8) | ((unsigned)(grp < ngroups ? grp : 255) << 16) | ((unsigned)ngroups << 24);
}

// push a [rows][4*vec_per_row] slice (row stride src_stride floats in the local staging buffer) to the same place in
// `nranks` CTAs of the cluster: element (i, j) goes to dst_local + i*row_stride + col0 + 4j, each 16-byte store also
// counting on the destination's receive barrier (rx_bar) when given.
__device__ __forceinline__ void cl_push_all(unsigned geo, const float* stage, int src_stride, float* dst_local,
                                            int row_stride, int col0, unsigned first_rank, unsigned nranks,
                                            unsigned long long* rx_bar = nullptr) {
  const unsigned grp = (geo >> 16) & 255u, ngroups = geo >> 24;
  if (grp == 255u) return;
  const int i = (int)((geo >> 8) & 255u), j = (int)(geo & 255u);
  const float4 x = *reinterpret_cast<const float4*>(stage + i * src_stride + 4 * j);
  const unsigned a = f32_smem_u32(dst_local) + (unsigned)(i * row_stride + col0 + 4 * j) * 4u;
  if (rx_bar != nullptr) {
    const unsigned mb = f32_smem_u32(rx_bar);
    for (unsigned pr = grp; pr < nranks; pr += ngroups)
      cl_st_async_v4(cl_mapa(a, first_rank + pr), x, cl_mapa(mb, first_rank + pr));
  } else {
    for (unsigned pr = grp; pr < nranks; pr += ngroups) cl_st_v4(cl_mapa(a, first_rank + pr), x);
  }
}

// register tile of one thread: MS streams x 4 columns over its K range.  Weights w[K][NC]; activations are
// segment-major (k < Kmain: act[(k >> sl_log)][stream][k & (SL-1)], SL floats per row, MS rows per segment),
// k >= Kmain: cond[stream][k - Kmain].  Threads >= KG * NC/4 idle.
template <int MS>
__device__ __forceinline__ void cl_contract(const float* __restrict__ w, int NC, int K, int KG, const float* act, int sl_log,
                                            int Kmain, const float* cond, int C, float (&acc)[MS][4]) {
  const int CQ = NC >> 2;
  const int tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < MS; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
  if (tid >= KG * CQ) return;
  const int kg = tid / CQ, q = tid - kg * CQ;
  const int klen = K / KG;
  const int k0 = kg * klen;
  const int SL = 1 << sl_log;
  const float* wp = w + q * 4;
  for (int k = k0; k < k0 + klen; k += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(wp + (k + 0) * NC);
    const float4 w1 = *reinterpret_cast<const float4*>(wp + (k + 1) * NC);
    const float4 w2 = *reinterpret_cast<const float4*>(wp + (k + 2) * NC);
    const float4 w3 = *reinterpret_cast<const float4*>(wp + (k + 3) * NC);
    const float* ap;
    int astr;
    if (k < Kmain) { ap = act + (k >> sl_log) * (MS << sl_log) + (k & (SL - 1)); astr = SL; }
    else { ap = cond + (k - Kmain); astr = C; }
#pragma unroll
    for (int j = 0; j < MS; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(ap + j * astr);
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
}

// partial sums -> red[kg][stream][NC] (red aliases the weight buffer: the caller has synchronised)
template <int MS>
__device__ __forceinline__ void cl_store_partials(float* red, int NC, int KG, const float (&acc)[MS][4]) {
  const int CQ = NC >> 2;
  const int tid = threadIdx.x;
  if (tid >= KG * CQ) return;
  const int kg = tid / CQ, q = tid - kg * CQ;
  float* rp = red + (size_t)kg * MS * NC + q * 4;
#pragma unroll
  for (int j = 0; j < MS; ++j)
    *reinterpret_cast<float4*>(rp + j * NC) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
}

template <int MS>
__global__ void __launch_bounds__(CL_THREADS, 1) wavenet_fp32_cluster(const ClParams p_in) {
  extern __shared__ __align__(128) float smem[];
  __shared__ ClLayerDev layers_s[64];     // per-layer table: no L2 round trips on the critical path
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p_in.L; i += CL_THREADS) layers_s[i] = p_in.layers[i];
  ClParams p = p_in;
  p.layers = layers_s;
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int cluster = p_in.cluster0 + (int)blockIdx.x / CL_CS;
  const int b0 = cluster * MS;                                  // first stream of this cluster
  const int nvalid = min(MS, p.B - b0);                         // >= 1 (host launches only clusters with work)

  const int R = p.R, G = p.G, S = p.S, Q = p.Q, C = p.C, PK = p.PK, L = p.L;
  const int GP = G / CL_CS;            // gate pairs per CTA
  const int NC1 = 2 * GP;
  const int NR = R / CL_CS, NSK = S / CL_CS, NC2 = NR + NSK;
  const int NQ = Q / CL_CS;
  const int K1 = 3 * R + C;
  const int sl_R = 31 - __clz(R), sl_S = 31 - __clz(S), sl_G = 31 - __clz(G);

  // ---- shared memory carve-up (floats)
  float* const wS1 = smem;                              // S1 / post1 weights; partial sums after the contraction
  float* const wS2 = wS1 + p.w1_floats;                 // skip start / S2 / post2 weights; partial sums
  float* const gfull = wS2 + p.w2_floats;               // [MS][G] gate outputs of all channels
  float* const seg0 = gfull + MS * G;                   // [MS][R] layer input   (gfull..seg0: post1 output [MS][S])
  float* const seg1 = seg0 + MS * R;                    // [MS][R] tap t-d       (seg1..seg2: relu(skip) [MS][S])
  float* const seg2 = seg1 + MS * R;                    // [MS][R] tap t-2d
  float* const cond_s = seg2 + MS * R;                  // [MS][C]
  float* const logits_s = cond_s + MS * C;              // [MS][Q] (used on CTA 0)
  float* const hist = logits_s + MS * Q;                // [MS][PK] input history ring
  float* const u_s = hist + MS * PK;                    // [MS][PK] history in tap order
  float* const skip_acc = u_s + MS * PK;                // [MS][NSK] this CTA's skip slice
  float* const stage = skip_acc + MS * NSK;             // [MS][max slice] staging for pushes
  int stage_cols = NC2 > NC1 ? NC2 : NC1;          // a full tile of any stage (host sizes the buffer the same way)
  if (NSK > stage_cols) stage_cols = NSK;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(stage + MS * stage_cols + ((MS * stage_cols) & 1));
  unsigned long long* wbar1 = bars;
  unsigned long long* wbar2 = bars + 1;
  unsigned long long* tapbar = bars + 2;
  unsigned long long* condbar = bars + 3;
  unsigned long long* gbar = bars + 4;    // gfull received from the 16 CTAs (S1 -> S2)
  unsigned long long* cbar = bars + 5;    // seg0 (next layer input) received (S2 -> S1)
  unsigned long long* skbar = bars + 6;   // relu(skip) received (last S2 -> post1)
  unsigned long long* n1bar = bars + 7;   // post1 output received (post1 -> post2)
  float* const skip_full = seg1;
  float* const n1_full = gfull;

  {
    // zero everything once: rows of streams past the end are never filled by the copies
    const int total = (int)(reinterpret_cast<float*>(bars) - smem);
    for (int i = tid; i < total; i += CL_THREADS) smem[i] = 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 8; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f32_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // receive barriers: armed one phase ahead with the bytes all 16 senders will deliver
    mbar_expect(gbar, (unsigned)(CL_CS * MS * (p.G / CL_CS) * 4));
    mbar_expect(cbar, (unsigned)(CL_CS * MS * (p.R / CL_CS) * 4));
    mbar_expect(skbar, (unsigned)(CL_CS * MS * (p.S / CL_CS) * 4));
    mbar_expect(n1bar, (unsigned)(CL_CS * MS * (p.S / CL_CS) * 4));
  }
  // input history of this cluster's streams
  for (int i = tid; i < nvalid * PK; i += CL_THREADS) hist[i] = ld_cg(p.u_hist + (long long)b0 * PK + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cl_barrier();      // every CTA of the cluster is running before the first remote store

  const float mu = (float)(Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  const long long ring_slot = (long long)p.Bp * R;
  unsigned ph1 = 0u, ph2 = 0u, phtap = 0u, phcond = 0u;
  bool alive = true;

  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && p.cluster0 == 0 && tid == 0;
  long long pf[24];
  for (int i = 0; i < 24; ++i) pf[i] = 0;
  long long pf_t = 0;
  int pf_cls = 2;     // 0: S1, 1: S2, 2: other stages
#define CL_PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define CL_PF_ADD(i) do { if (prof) { long long n_ = clock64(); pf[(i) + 8 * pf_cls] += n_ - pf_t; pf_t = n_; } } while (0)

  // ---- bulk-copy issue helpers (thread 0 posts the byte count; a few threads issue the copies)
  auto issue_w = [&](float* dst, const float* src, unsigned bytes, unsigned long long* bar) {
    // one bulk copy per weight tile (issuing a copy costs hundreds of cycles; the size only has to fit the
    // mbarrier's 2^20-1 transaction-byte range).  The buffer was last written by generic stores (partial sums,
    // ordered by the preceding __syncthreads): the issuing thread orders them before its async-proxy writes.
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect(bar, bytes);
      cl_bulk_g2s_keep(dst, src, bytes, bar);
    }
  };
  // older dilation taps of layer l at step t: issued by warp 1 so that they do not queue behind the weight copy
  auto issue_taps = [&](int l, long long t) {
    if (tid == 32 || tid == 33) {
      const int which = tid - 32;
      const ClLayerDev ly = p.layers[l];
      const int d2 = 2 * ly.d;
      const unsigned bytes = (unsigned)(nvalid * R * 4);
      if (which == 0) mbar_expect(tapbar, 2 * bytes);
      const long long slot = (which == 0) ? ((t + ly.d) % d2) : (t % d2);
      bulk_g2s(which == 0 ? seg1 : seg2, ly.ring + slot * ring_slot + (long long)b0 * R, bytes, tapbar);
    }
  };
  auto issue_cond = [&](long long frame) {
    if (tid == 0) mbar_expect(condbar, (unsigned)(nvalid * C * 4));
    if (tid < nvalid)
      bulk_g2s(cond_s + tid * C, p.cond + (long long)(b0 + tid) * p.cond_bstride + frame * C, (unsigned)C * 4u, condbar);
  };
  auto wait_bar = [&](unsigned long long* bar, unsigned& ph) {
    alive = alive && mbar_wait_bounded(bar, ph, p.err);
    ph ^= 1u;
  };

  // wait for a pushed activation block, then arm the barrier for its next use (the next block cannot be sent before
  // every CTA has consumed this one: each sender first needs this CTA's output of the stage that reads it)
  auto recv_wait = [&](unsigned long long* bar, unsigned& ph, unsigned bytes) {
    wait_bar(bar, ph);
    if (tid == 0) mbar_expect(bar, bytes);
  };
  unsigned phg = 0u, phc = 0u, phsk = 0u, phn1 = 0u;
  const unsigned rx_g = (unsigned)(CL_CS * MS * GP * 4), rx_c = (unsigned)(CL_CS * MS * NR * 4),
                 rx_s = (unsigned)(CL_CS * MS * NSK * 4);

  // per-thread index decompositions, once per kernel
  const unsigned geo_g = cl_push_geo(MS * GP / 4, GP / 4), geo_r = cl_push_geo(MS * NR / 4, NR / 4),
                 geo_s = cl_push_geo(MS * NSK / 4, NSK / 4), geo_q = cl_push_geo(MS * NQ / 4, NQ / 4);
  const int gate_i = tid / GP, gate_j = tid - gate_i * GP;                       // MS*GP <= 256 (host-checked)
  const int s2_i0 = tid / NC2, s2_c0 = tid - s2_i0 * NC2;                        // outputs tid and tid + 256 of an S2 tile
  const int s2_i1 = (tid + CL_THREADS) / NC2, s2_c1 = (tid + CL_THREADS) - s2_i1 * NC2;

  // one contraction stage: bias prefetch -> operand wait -> register-tile contraction -> partial sums through shared
  // memory (aliasing the weight buffer) -> stage[stream*NC + c] = sum over K + bias.  Column c of CTA `rank` is
  // global column rank*n0 + c (c < n0) or base1 + rank*n1 + (c - n0).
  float acc[MS][4];
  auto run_stage = [&](float* wbuf, int NC, int K, int KG, const float* act, int sl_log, int Kmain, const float* bias,
                       int n0, int base1, int n1, unsigned long long* barA, unsigned& phA, unsigned long long* barB,
                       unsigned* phB, int c0, int c1) {
    const int nout = MS * NC;                     // <= 512
    float bz[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = tid + h * CL_THREADS;
      if (o < nout) {
        const int c = h ? c1 : c0;                // o % NC, supplied by the caller (no runtime modulo here)
        bz[h] = __ldg(bias + ((c < n0) ? (rank * n0 + c) : (base1 + rank * n1 + (c - n0))));
      }
    }
    wait_bar(barA, phA);
    if (barB != nullptr) wait_bar(barB, *phB);
    CL_PF_ADD(1);
    cl_contract<MS>(wbuf, NC, K, KG, act, sl_log, Kmain, cond_s, C, acc);
    __syncthreads();
    cl_store_partials<MS>(wbuf, NC, KG, acc);
    __syncthreads();
    CL_PF_ADD(2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = tid + h * CL_THREADS;
      if (o < nout) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int kg = 0;
        for (; kg + 4 <= KG; kg += 4) {
          s0 += wbuf[(kg + 0) * nout + o]; s1 += wbuf[(kg + 1) * nout + o];
          s2 += wbuf[(kg + 2) * nout + o]; s3 += wbuf[(kg + 3) * nout + o];
        }
        for (; kg < KG; ++kg) s0 += wbuf[kg * nout + o];
        stage[o] = ((s0 + s1) + (s2 + s3)) + bz[h];
      }
    }
    __syncthreads();
    CL_PF_ADD(7);
  };

  // preprocess FIR taps of channel `tid` (fast path: one channel per thread, 32 taps)
  const bool fir_regs = (R == CL_THREADS) && (PK == FIR_TAPS);
  float fir_k[FIR_TAPS];
  float fir_b = 0.f;
#pragma unroll
  for (int j = 0; j < FIR_TAPS; ++j) fir_k[j] = fir_regs ? __ldg(p.pre_k + (FIR_TAPS - 1 - j) * R + tid) : 0.f;
  if (fir_regs) fir_b = __ldg(p.pre_b + tid);

  // ---- prologue: weights of the first two contractions, taps of layer 0, condition of frame 0
  issue_w(wS2, p.skip0c + (size_t)rank * R * NSK, (unsigned)(R * NSK * 4), wbar2);
  issue_w(wS1, p.layers[0].w1c + (size_t)rank * K1 * NC1, (unsigned)(K1 * NC1 * 4), wbar1);
  issue_taps(0, p.t0);
  long long cond_frame = -1;

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {   // a timed-out wait keeps running (all CTAs must reach every cluster barrier)
    const long long frame_t = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    CL_PF_START();
    if (frame_t != cond_frame) {       // cond_s was last read by post1 of the previous step (program order)
      issue_cond(frame_t);
      wait_bar(condbar, phcond);
      cond_frame = frame_t;
    }
    // ================================================================== stage 0: input history -> FIR -> skip start
    {
      const int slot_t = (int)(t % PK);
      if (ext) {
        if (tid < MS) {
          const int b = b0 + tid;
          float x = 0.f;
          if (b < p.B) {
            if (p.mode == GEN_STEP) x = p.ext_audio[b];
            else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
          }
          const float u = mu_law_encode_dev(x, mu, 0.f);
          hist[tid * PK + slot_t] = u;
          if (rank == 0 && b < p.B) st_cg(p.u_hist + (long long)b * PK + slot_t, u);
        }
        __syncthreads();
      }
      for (int idx = tid; idx < MS * PK; idx += CL_THREADS) {
        const int i = idx / PK, j = idx - i * PK;
        int sl = (int)((t - j) % PK);
        if (sl < 0) sl += PK;
        u_s[idx] = hist[i * PK + sl];
      }
      __syncthreads();
      // h0 = (u0*K[PK-1] + b) + u1*K[PK-2] + ...   (wavenet_ops.py:178,193)
      if (fir_regs) {
#pragma unroll 2
        for (int i = 0; i < MS; ++i) {
          const float4* up = reinterpret_cast<const float4*>(u_s + i * FIR_TAPS);
          float a = fir_b;
#pragma unroll
          for (int j4 = 0; j4 < FIR_TAPS / 4; ++j4) {
            const float4 u4 = up[j4];
            a = fmaf(u4.x, fir_k[4 * j4 + 0], a); a = fmaf(u4.y, fir_k[4 * j4 + 1], a);
            a = fmaf(u4.z, fir_k[4 * j4 + 2], a); a = fmaf(u4.w, fir_k[4 * j4 + 3], a);
          }
          seg0[i * R + tid] = a;
        }
      } else {
        for (int idx = tid; idx < MS * R; idx += CL_THREADS) {
          const int i = idx / R, n = idx - i * R;
          float a = fmaf(u_s[i * PK], __ldg(p.pre_k + (PK - 1) * R + n), __ldg(p.pre_b + n));
          for (int j = 1; j < PK; ++j) a = fmaf(u_s[i * PK + j], __ldg(p.pre_k + (PK - 1 - j) * R + n), a);
          seg0[idx] = a;
        }
      }
      __syncthreads();
      CL_PF_ADD(0);
      // skip start: skip = h0 . W_skip + b   (wavenet.py:117-121)
      run_stage(wS2, NSK, R, p.kg_s0, seg0, sl_R, R, p.skip0_b, NSK, 0, 0, wbar2, ph2, nullptr, nullptr, tid & (NSK - 1), (tid + CL_THREADS) & (NSK - 1));
      for (int o = tid; o < MS * NSK; o += CL_THREADS) skip_acc[o] = stage[o];
      __syncthreads();
      issue_w(wS2, p.layers[0].w2c + (size_t)rank * G * NC2, (unsigned)(G * NC2 * 4), wbar2);
      CL_PF_ADD(3);
    }

    // ================================================================== residual stacks
    for (int l = 0; l < L; ++l) {
      const ClLayerDev ly = p.layers[l];
      const bool last = (l == L - 1);
      // ---------------------------------------------------------------- S1: dilated conv + condition + gate
      pf_cls = 0;
      if (l > 0) { recv_wait(cbar, phc, rx_c); CL_PF_ADD(4); }
      run_stage(wS1, NC1, K1, p.kg_s1, seg0, sl_R, 3 * R, ly.b1, GP, G, GP, wbar1, ph1, tapbar, &phtap, tid & (NC1 - 1), (tid + CL_THREADS) & (NC1 - 1));
      if (tid < MS * GP)
        stage[gate_i * NC1 + gate_j] =
            tanhf(stage[gate_i * NC1 + gate_j]) * sigmoid_f(stage[gate_i * NC1 + GP + gate_j]);   // wavenet_ops.py:236-240
      __syncthreads();
      CL_PF_ADD(5);
      cl_push_all(geo_g, stage, NC1, gfull, G, rank * GP, 0u, CL_CS, gbar);
      CL_PF_ADD(3);
      // next S1-class weights (+ the next layer's older taps) stream in during the hand-off and S2
      if (!last) {
        issue_w(wS1, p.layers[l + 1].w1c + (size_t)rank * K1 * NC1, (unsigned)(K1 * NC1 * 4), wbar1);
        issue_taps(l + 1, t);
      } else {
        issue_w(wS1, p.post1c + (size_t)rank * (S + C) * NSK, (unsigned)((S + C) * NSK * 4), wbar1);
      }
      CL_PF_ADD(6);

      // ---------------------------------------------------------------- S2: residual + skip 1x1
      pf_cls = 1;
      recv_wait(gbar, phg, rx_g);
      CL_PF_ADD(4);
      run_stage(wS2, NC2, G, p.kg_s2, gfull, sl_G, G, ly.b2, NR, R, NSK, wbar2, ph2, nullptr, nullptr, s2_c0, s2_c1);
      {
        const int slot_old = (int)(t % (2 * ly.d));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int o = tid + h * CL_THREADS;
          if (o >= MS * NC2) break;
          const int i = h ? s2_i1 : s2_i0, c = h ? s2_c1 : s2_c0;
          const float v = stage[o];
          if (c < NR) {
            const float oldv = seg0[i * R + rank * NR + c];
            if (i < nvalid) st_cg(ly.ring + slot_old * ring_slot + (long long)(b0 + i) * R + rank * NR + c, oldv);  // push_ops
            stage[o] = oldv + v;
          } else {
            const float sk = skip_acc[i * NSK + (c - NR)] + v;
            skip_acc[i * NSK + (c - NR)] = sk;
            stage[o] = fmaxf(sk, 0.f);     // only read after the last layer (wavenet.py:153)
          }
        }
        // the ring stores above are generic-proxy writes that a later step reads with bulk copies (async proxy)
        asm volatile("fence.proxy.async.global;" ::: "memory");
      }
      __syncthreads();
      CL_PF_ADD(5);
      if (!last) cl_push_all(geo_r, stage, NC2, seg0, R, rank * NR, 0u, CL_CS, cbar);
      else cl_push_all(geo_s, stage + NR, NC2, skip_full, S, rank * NSK, 0u, CL_CS, skbar);   // wavenet.py:145: last residual is dead
      CL_PF_ADD(3);
      if (!last) issue_w(wS2, p.layers[l + 1].w2c + (size_t)rank * G * NC2, (unsigned)(G * NC2 * 4), wbar2);
      else issue_w(wS2, p.post2c + (size_t)rank * S * NQ, (unsigned)(S * NQ * 4), wbar2);
      CL_PF_ADD(6);
    }

    // ================================================================== postprocess1 (+ condition), relu
    pf_cls = 2;
    recv_wait(skbar, phsk, rx_s);
    CL_PF_ADD(4);
    run_stage(wS1, NSK, S + C, p.kg_p1, skip_full, sl_S, S, p.post1_b, NSK, 0, 0, wbar1, ph1, nullptr, nullptr, tid & (NSK - 1), (tid + CL_THREADS) & (NSK - 1));
    for (int o = tid; o < MS * NSK; o += CL_THREADS) stage[o] = fmaxf(stage[o], 0.f);     // wavenet.py:163
    __syncthreads();
    cl_push_all(geo_s, stage, NSK, n1_full, S, rank * NSK, 0u, CL_CS, n1bar);
    CL_PF_ADD(3);
    if (t + 1 < p.t0 + p.T) {
      issue_w(wS1, p.layers[0].w1c + (size_t)rank * K1 * NC1, (unsigned)(K1 * NC1 * 4), wbar1);
      issue_taps(0, t + 1);       // seg1/seg2 held relu(skip): this CTA's post1 contraction is done
    }
    CL_PF_ADD(6);

    // ================================================================== postprocess2 -> logits on CTA 0
    recv_wait(n1bar, phn1, rx_s);
    CL_PF_ADD(4);
    run_stage(wS2, NQ, S, p.kg_p2, n1_full, sl_S, S, p.post2_b, NQ, 0, 0, wbar2, ph2, nullptr, nullptr, tid & (NQ - 1), (tid + CL_THREADS) & (NQ - 1));
    cl_push_all(geo_q, stage, NQ, logits_s, Q, rank * NQ, 0u, 1u);
    CL_PF_ADD(3);
    cl_arrive();
    if (t + 1 < p.t0 + p.T) issue_w(wS2, p.skip0c + (size_t)rank * R * NSK, (unsigned)(R * NSK * 4), wbar2);
    CL_PF_ADD(6);
    cl_wait();
    CL_PF_ADD(4);


    // ================================================================== softmax + draw + mu-law decode (CTA 0)
    if (rank == 0) {
      const int NQW = Q / 32;   // <= 8
      for (int i = warp; i < nvalid; i += CL_THREADS / 32) {
        const int b = b0 + i;
        float lg[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) lg[q] = (q < NQW) ? logits_s[i * Q + lane + 32 * q] : -INFINITY;
        // (sample mode reuses the stream's logits row as scratch: the logits are in registers by then)
        const int k = warp_softmax_draw(p, Q, lg, b, t, lane, logits_s + i * Q);
        if (k >= 0 && lane == 0) {
          const float un = __ldg(p.enc_lut + k);
          const int slot_n = (int)((t + 1) % PK);
          st_cg(p.u_hist + (long long)b * PK + slot_n, un);
          const unsigned a = f32_smem_u32(hist + i * PK + slot_n);
          for (unsigned r = 0; r < (unsigned)CL_CS; ++r) cl_st_f32(cl_mapa(a, r), un);
        }
      }
    }
    CL_PF_ADD(5);
    cl_barrier();
    CL_PF_ADD(4);
  }
  // nobody leaves while a peer may still push into its shared memory
  cl_barrier();
  if (prof) for (int i = 0; i < 24; ++i) p.prof[i] = pf[i];
#undef CL_PF_START
#undef CL_PF_ADD
}

// [K][ldw] row-major -> [tiles][K][n0+n1]: tile cb takes columns cb*n0 .. +n0, then base1 + cb*n1 .. +n1
__global__ void pack_cluster_kernel(const float* __restrict__ src, int ldw, int K, int n0, int base1, int n1, int ntiles,
                                    float* __restrict__ dst) {
  const int NC = n0 + n1;
  const long long total = (long long)ntiles * K * NC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % NC);
    const long long r = i / NC;
    const int k = (int)(r % K);
    const int cb = (int)(r / K);
    const int col = (c < n0) ? (cb * n0 + c) : (base1 + cb * n1 + (c - n0));
    dst[i] = src[(long long)k * ldw + col];
  }
}

}  // namespace vqwn
