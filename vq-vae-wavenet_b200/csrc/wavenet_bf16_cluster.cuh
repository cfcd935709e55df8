// WaveNet fast generation on the 5th-generation tensor cores: bf16 operands, fp32 accumulation in TMEM
// (VQWN_PREC_BF16).  Same reference semantics as the float32 kernels (wavenet.py:103-172,
// wavenet_ops.py:163-267, utils.py:13-46, mu_law_ops.py:5-31); weights and the activations that feed a
// contraction are rounded to bfloat16, everything else (biases, residual and skip chains, gate, softmax, draw)
// stays float32.
//
// Structure = the float32 cluster kernel (wavenet_fp32_cluster.cuh) with the CUDA-core contraction replaced by
// tcgen05.mma: a cluster of 8 CTAs owns 16 streams for the whole run and splits every stage by output channels;
// the channels are the M dimension (weights = A operand, 128 rows of which 32/64/96 are real, the rest alias the
// following shared memory and produce ignored accumulator rows), the 16 streams are N, and a stage is a chain of
// K/16 MMAs issued by one elected thread.  Per CTA:
//   skip start 64 ch x K 256   S1 64 rows (32 tanh/sigmoid pairs, interleaved) x K 896   S2 32 res + 64 skip x K 256
//   post1 64 ch x K 640        post2 32 logits x K 512
// Operands sit in shared memory as K-major no-swizzle "planes": plane p holds k = 8p..8p+7 of every row
// (16 bytes per row, rows back to back), so LBO = plane stride and SBO = 128 in the matrix descriptors.  Weight
// tiles arrive pre-packed in that layout with one cp.async.bulk one stage ahead; activations are pushed between
// the CTAs of a cluster as 16-byte plane chunks with st.async + mbarrier complete_tx; the dilation queues are
// HBM rings in the same plane layout ([2d][cluster][32 planes][16 streams][8] bf16) so a tap is one 8 KB bulk copy.
// The accumulator (128 lanes x 16 columns of TMEM) is read with tcgen05.ld 32x32b.x16: thread = output channel.
// The geometry is fixed to the reference's default (R = G = 256, S = 512, Q = 256, C = 128, 32-tap preprocess).
#pragma once
#include <cuda_bf16.h>
#include "wavenet_fp32_cluster.cuh"

namespace vqwn {

constexpr int BC_CS = 8;             // CTAs per cluster
constexpr int BC_THREADS = 256;
constexpr int BC_NS = 16;            // streams per cluster = MMA N
constexpr int BC_R = 256, BC_G = 256, BC_S = 512, BC_Q = 256, BC_C = 128, BC_PK = 32;
constexpr int BC_K1 = 3 * BC_R + BC_C;                 // 896
constexpr int BC_ROWS_S1 = 2 * BC_G / BC_CS;           // 64: tanh / sigmoid rows interleaved
constexpr int BC_NR = BC_R / BC_CS, BC_NSK = BC_S / BC_CS, BC_NQ = BC_Q / BC_CS;   // 32, 64, 32
constexpr int BC_ROWS_S2 = BC_NR + BC_NSK;             // 96
constexpr int BC_PLANE_B = BC_NS * 16;                 // bytes of one activation plane (16 rows x 16 B)
// shared memory map (bytes)
constexpr int BC_W1_BYTES = BC_ROWS_S1 * BC_K1 * 2;    // 114688 (post1: 64 x 640 x 2 fits)
constexpr int BC_W2_BYTES = BC_ROWS_S2 * BC_G * 2;     // 49152  (skip start 64 x 256, post2 32 x 512 fit)
constexpr int BC_OFF_W1 = 0;
constexpr int BC_OFF_W2 = BC_OFF_W1 + BC_W1_BYTES;
constexpr int BC_OFF_G = BC_OFF_W2 + BC_W2_BYTES;                  // gate outputs: 32 planes
constexpr int BC_OFF_X = BC_OFF_G + 32 * BC_PLANE_B;               // cur 32 | tap t-d 32 | tap t-2d 32 | cond 16 planes
constexpr int BC_OFF_CUR32 = BC_OFF_X + 112 * BC_PLANE_B;          // [32 ch][16] fp32 residual chain of this CTA's channels
constexpr int BC_OFF_SKIP32 = BC_OFF_CUR32 + BC_NR * BC_NS * 4;    // [64 ch][16] fp32 skip accumulators
constexpr int BC_STAGE_PLANE = BC_PLANE_B + 16;                    // staging planes are padded by 16 B (bank spread of the 2-byte stores)
constexpr int BC_OFF_STAGE = BC_OFF_SKIP32 + BC_NSK * BC_NS * 4;   // 8 + 4 padded planes: slice to push | queue slice
constexpr int BC_OFF_PRE = BC_OFF_STAGE + 12 * BC_STAGE_PLANE;     // [16 streams][64 rows] fp32 pre-activations of the gate
constexpr int BC_OFF_HIST = BC_OFF_PRE + BC_NS * BC_ROWS_S1 * 4;   // [16][32] fp32 input history ring
constexpr int BC_OFF_US = BC_OFF_HIST + BC_NS * BC_PK * 4;         // [16][32] history in tap order
constexpr int BC_OFF_BARS = BC_OFF_US + BC_NS * BC_PK * 4;
constexpr int BC_SMEM = BC_OFF_BARS + 128;
// aliases: relu(skip) for post1 = tap planes (64 planes, followed by the 16 cond planes: K = 640 contiguous);
// post1 output for post2 = gate planes + cur planes (64 planes contiguous); logits (CTA 0) = tap planes again
constexpr int BC_OFF_SKF = BC_OFF_X + 32 * BC_PLANE_B;
constexpr int BC_OFF_N1F = BC_OFF_G;
constexpr int BC_OFF_LOGITS = BC_OFF_SKF;                          // [16][256] fp32 = 16 KB = 64 planes

struct BcLayerDev {
  const __nv_bfloat16* w1;   // [8 CTAs][112 planes][64 rows][8]   row 2j = tanh channel 32r+j, row 2j+1 = its sigmoid partner
  const float* b1;           // [2G]
  const __nv_bfloat16* w2;   // [8][32 planes][96 rows][8]         rows 0-31 residual 32r+i, rows 32-95 skip 64r+i
  const float* b2;           // [R+S]
  __nv_bfloat16* ring;       // [2d][clusters][32 planes][16][8]
  int d;
  int pad_;
};

struct BcParams {
  int L, B, nclusters;      // nclusters: clusters of the whole batch (ring layout)
  int cluster0;             // first cluster of this launch
  const float *pre_k, *pre_b;
  const __nv_bfloat16 *skip0, *post1, *post2;     // [8][32 planes][64][8], [8][80][64][8], [8][64][32][8]
  const float *skip0_b, *post1_b, *post2_b;
  const BcLayerDev* layers;
  const float *enc_lut, *dec_lut;
  float* u_hist;
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

__device__ __forceinline__ uint64_t bc_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;     // distance between core matrices adjacent in K (= plane stride)
  d |= (uint64_t)(128 >> 4) << 32;           // distance between 8-row groups inside a plane
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void bc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void bc_st_bf16(uint8_t* planes, int n, int k, float x) {
  *reinterpret_cast<__nv_bfloat16*>(planes + (k >> 3) * BC_PLANE_B + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(x);
}
// same element in the padded staging planes
__device__ __forceinline__ void bc_st_stage(uint8_t* stage, int n, int k, float x) {
  *reinterpret_cast<__nv_bfloat16*>(stage + (k >> 3) * BC_STAGE_PLANE + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(x);
}
// gate of the bf16 path: hardware tanh (MUFU.TANH, rel. error ~2^-11, far below the bf16 rounding of the result)
__device__ __forceinline__ float bc_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float bc_gate(float a, float b) { return bc_tanh(a) * fmaf(0.5f, bc_tanh(0.5f * b), 0.5f); }
// push `nchunks` 16-byte chunks (padded planes of 16 chunks in the local staging buffer, contiguous at dst_off in every
// destination CTA)
// executed by the first `nthr` threads of the CTA
__device__ __forceinline__ void bc_push(const uint8_t* stage, uint8_t* smem_base, int dst_off, int nchunks, unsigned nranks,
                                        unsigned long long* rx_bar, int nthr = BC_THREADS) {
  const unsigned dst = f32_smem_u32(smem_base + dst_off);
  const unsigned mb = rx_bar ? f32_smem_u32(rx_bar) : 0u;
  const int t0 = threadIdx.x, nt = nthr;
  if (t0 >= nthr) return;
  for (int w = t0; w < nchunks * (int)nranks; w += nt) {
    const int c = w % nchunks;
    const unsigned pr = (unsigned)(w / nchunks);
    const float4 x = *reinterpret_cast<const float4*>(stage + c * 16 + (c >> 4) * 16);
    if (rx_bar) cl_st_async_v4(cl_mapa(dst + c * 16, pr), x, cl_mapa(mb, pr));
    else cl_st_v4(cl_mapa(dst + c * 16, pr), x);
  }
}

__global__ void __launch_bounds__(BC_THREADS, 1) wavenet_bf16_cluster(const BcParams p_in) {
  extern __shared__ __align__(1024) uint8_t bsm[];
  __shared__ BcLayerDev layers_s[64];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p_in.L; i += BC_THREADS) layers_s[i] = p_in.layers[i];
  BcParams p = p_in;
  p.layers = layers_s;
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int cluster = p_in.cluster0 + (int)blockIdx.x / BC_CS;
  const int b0 = cluster * BC_NS;
  const int nvalid = min(BC_NS, p.B - b0);
  const int L = p.L;

  uint8_t* const wS1 = bsm + BC_OFF_W1;
  uint8_t* const wS2 = bsm + BC_OFF_W2;
  uint8_t* const gpl = bsm + BC_OFF_G;
  uint8_t* const xpl = bsm + BC_OFF_X;
  float* const cur32 = reinterpret_cast<float*>(bsm + BC_OFF_CUR32);
  float* const skip32 = reinterpret_cast<float*>(bsm + BC_OFF_SKIP32);
  uint8_t* const stage = bsm + BC_OFF_STAGE;
  float* const pre = reinterpret_cast<float*>(bsm + BC_OFF_PRE);
  float* const hist = reinterpret_cast<float*>(bsm + BC_OFF_HIST);
  float* const u_s = reinterpret_cast<float*>(bsm + BC_OFF_US);
  float* const logits_s = reinterpret_cast<float*>(bsm + BC_OFF_LOGITS);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(bsm + BC_OFF_BARS);
  unsigned long long* wbar1 = bars;       // S1-class weights landed
  unsigned long long* wbar2 = bars + 1;   // S2-class weights landed
  unsigned long long* tapbar = bars + 2;  // older taps landed
  unsigned long long* accbar = bars + 3;  // MMA chain complete (tcgen05.commit)
  unsigned long long* gbar = bars + 4;    // gate planes received from the 8 CTAs
  unsigned long long* cbar = bars + 5;    // next layer input planes received
  unsigned long long* skbar = bars + 6;   // relu(skip) planes received
  unsigned long long* n1bar = bars + 7;   // post1 output planes received

  for (int i = tid; i < (BC_OFF_BARS - BC_OFF_G) / 4; i += BC_THREADS) reinterpret_cast<uint32_t*>(bsm + BC_OFF_G)[i] = 0u;
  __syncthreads();
  constexpr unsigned RX_G = BC_CS * 4 * BC_PLANE_B, RX_C = BC_CS * 4 * BC_PLANE_B, RX_S = BC_CS * 8 * BC_PLANE_B;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f32_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect(gbar, RX_G);
    mbar_expect(cbar, RX_C);
    mbar_expect(skbar, RX_S);
    mbar_expect(n1bar, RX_S);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(f32_smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < nvalid * BC_PK; i += BC_THREADS) hist[i] = ld_cg(p.u_hist + (long long)b0 * BC_PK + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  cl_barrier();

  const float mu = (float)(BC_Q - 1);
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  unsigned ph1 = 0u, ph2 = 0u, phtap = 0u, phacc = 0u, phg = 0u, phc = 0u, phsk = 0u, phn1 = 0u;
  bool alive = true;
  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && p.cluster0 == 0 && tid == 0;
  long long pf[24];
  for (int i = 0; i < 24; ++i) pf[i] = 0;
  long long pf_t = 0;
  int pf_cls = 2;     // 0: S1, 1: S2, 2: other stages
#define BC_PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define BC_PF_ADD(i) do { if (prof) { long long n_ = clock64(); pf[(i) + 8 * pf_cls] += n_ - pf_t; pf_t = n_; } } while (0)
  uint32_t elected = 0;
  if (warp == 4) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(elected));
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BC_NS >> 3) << 17) | ((128u >> 4) << 24);

  // housekeeping runs on warps 5-6 so that the epilogue warps (0-3) and the MMA warp (4) never wait for it:
  // weight copies on warp 5, tap copies on warp 6
  auto issue_w = [&](uint8_t* dst, const __nv_bfloat16* src, unsigned bytes, unsigned long long* bar) {
    if (tid == 160) {
      mbar_expect(bar, bytes);
      cl_bulk_g2s_keep(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(src), bytes, bar);
    }
  };
  const long long ring_slot_elems = (long long)p.nclusters * BC_R * BC_NS;     // bf16 elements per ring slot
  auto issue_taps = [&](int l, long long t) {
    if (tid == 192 || tid == 193) {
      const int which = tid - 192;
      const BcLayerDev ly = p.layers[l];
      const int d2 = 2 * ly.d;
      const unsigned bytes = 32 * BC_PLANE_B;
      if (which == 0) mbar_expect(tapbar, 2 * bytes);
      const long long slot = (which == 0) ? ((t + ly.d) % d2) : (t % d2);
      bulk_g2s(reinterpret_cast<float*>(xpl + (32 + 32 * which) * BC_PLANE_B),
               reinterpret_cast<const float*>(ly.ring + slot * ring_slot_elems + (long long)cluster * BC_R * BC_NS), bytes, tapbar);
    }
  };
  // every stage starts here (all threads): operands written through the generic proxy by this CTA (FIR output,
  // condition planes) become visible to the tensor-core (async) proxy, and the TMEM reads of earlier epilogues are
  // ordered before the MMAs that overwrite those accumulators
  auto stage_sync = [&]() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // stages after the first: only the epilogue warps (0-3) and the MMA warp (4) meet; warps 5-7 (copy issue) are
  // paced by the accumulator barrier alone, so a slow bulk-copy issue never holds a stage up
  auto stage_sync5 = [&]() {
    if (warp < 5) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 3, 160;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  };
  auto ep_sync = [&]() { if (warp < 4) asm volatile("bar.sync 2, 128;" ::: "memory"); };
  // ---- warp 4 only: operand waits, MMA issue, commit.  Only this warp needs the operands; everybody else waits for
  // the accumulator.
  auto w4_wait = [&](unsigned long long* bar, unsigned& ph) {
    alive = alive && mbar_wait_bounded(bar, ph, p.err);
    ph ^= 1u;
  };
  auto w4_recv = [&](unsigned long long* bar, unsigned& ph, unsigned bytes) {
    // pushed activation block landed; re-arm the barrier for its next use (the next block cannot be sent before
    // every CTA has consumed this one: each sender first needs this CTA's output of the stage that reads it)
    w4_wait(bar, ph);
    if (lane == 0) mbar_expect(bar, bytes);
  };
  // K steps [ks0, ks1) of D[128 x 16] (TMEM column d_col) (+)= A . B^T; A planes a_lbo bytes apart, B planes 256 B
  auto mma_issue = [&](uint32_t d_col, uint8_t* a_base, uint32_t a_lbo, uint8_t* b_base, int ks0, int ks1, bool fresh) {
    // remote st.async / bulk-copy data observed through the mbarriers above -> tensor-core proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a0 = f32_smem_u32(a_base), bb = f32_smem_u32(b_base);
    for (int ks = ks0; ks < ks1; ++ks) {
      const uint64_t da = bc_desc(a0 + (uint32_t)ks * 2u * a_lbo, a_lbo);
      const uint64_t db = bc_desc(bb + (uint32_t)ks * 2u * BC_PLANE_B, BC_PLANE_B);
      const uint32_t accf = (fresh && ks == ks0) ? 0u : 1u;
      asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                   "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tmem + d_col), "l"(da), "l"(db), "r"(idesc), "r"(accf), "r"(elected) : "memory");
    }
  };
  auto mma_commit = [&]() {     // accbar fires when every MMA issued so far has completed
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                 ::"r"(f32_smem_u32(accbar)), "r"(elected) : "memory");
  };
  auto acc_wait = [&]() {
    alive = alive && mbar_wait_bounded(accbar, phacc, p.err);
    phacc ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  constexpr uint32_t D1 = 0, D2 = 16;     // TMEM accumulator columns: S1-class stages | S2-class stages
  constexpr int KS_CUR = BC_R / 16;       // K steps of the current-input part of S1 (the rest: older taps + condition)
  const uint32_t my_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int row = (warp & 3) * 32 + lane;          // accumulator row of an epilogue thread (warps 0-3)

  // preprocess FIR taps of channel `tid`
  float fir_k[BC_PK];
#pragma unroll
  for (int j = 0; j < BC_PK; ++j) fir_k[j] = __ldg(p.pre_k + (BC_PK - 1 - j) * BC_R + tid);
  const float fir_b = __ldg(p.pre_b + tid);

  // ---- prologue
  issue_w(wS2, p.skip0 + (size_t)rank * BC_NSK * BC_R, BC_NSK * BC_R * 2, wbar2);
  issue_w(wS1, p.layers[0].w1 + (size_t)rank * BC_ROWS_S1 * BC_K1, BC_ROWS_S1 * BC_K1 * 2, wbar1);
  long long cond_frame = -1;
  float v[16];

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {
    const long long frame_t = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    pf_cls = 2;
    BC_PF_START();
    issue_taps(0, t);          // tap planes held the logits of the previous step until its draw finished
    if (frame_t != cond_frame) {
      // condition rows -> bf16 planes (16 planes behind the taps): thread = (stream, 8 channels)
      {
        const int n = tid >> 4, c8 = tid & 15;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = 0.f;
        if (n < nvalid) {
          const float4* src = reinterpret_cast<const float4*>(p.cond + (long long)(b0 + n) * p.cond_bstride + frame_t * BC_C + c8 * 8);
          const float4 lo = __ldg(src), hi = __ldg(src + 1);
          x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w; x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) bc_st_bf16(xpl + 96 * BC_PLANE_B, n, c8 * 8 + e, x[e]);
      }
      cond_frame = frame_t;
    }
    // ================================================================== stage 0: history -> FIR -> skip start
    {
      const int slot_t = (int)(t % BC_PK);
      if (ext) {
        if (tid < BC_NS) {
          const int b = b0 + tid;
          float x = 0.f;
          if (b < p.B) {
            if (p.mode == GEN_STEP) x = p.ext_audio[b];
            else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
          }
          const float u = mu_law_encode_dev(x, mu, 0.f);
          hist[tid * BC_PK + slot_t] = u;
          if (rank == 0 && b < p.B) st_cg(p.u_hist + (long long)b * BC_PK + slot_t, u);
        }
        __syncthreads();
      }
      for (int idx = tid; idx < BC_NS * BC_PK; idx += BC_THREADS) {
        const int i = idx / BC_PK, j = idx - i * BC_PK;
        int sl = (int)((t - j) % BC_PK);
        if (sl < 0) sl += BC_PK;
        u_s[idx] = hist[i * BC_PK + sl];
      }
      __syncthreads();
      // h0 = (u0*K[PK-1] + b) + u1*K[PK-2] + ...   (wavenet_ops.py:178,193); thread = channel
#pragma unroll 2
      for (int i = 0; i < BC_NS; ++i) {
        const float4* up = reinterpret_cast<const float4*>(u_s + i * BC_PK);
        float a = fir_b;
#pragma unroll
        for (int j4 = 0; j4 < BC_PK / 4; ++j4) {
          const float4 u4 = up[j4];
          a = fmaf(u4.x, fir_k[4 * j4 + 0], a); a = fmaf(u4.y, fir_k[4 * j4 + 1], a);
          a = fmaf(u4.z, fir_k[4 * j4 + 2], a); a = fmaf(u4.w, fir_k[4 * j4 + 3], a);
        }
        bc_st_bf16(xpl, i, tid, a);
        if ((tid >> 5) == rank) cur32[i * BC_NR + (tid & 31)] = a;     // float32 residual chain of this CTA's channels
      }
      // skip start (wavenet.py:117-121): 64 skip channels of this CTA
      BC_PF_ADD(0);
      const float bias_s0 = (warp < 2) ? __ldg(p.skip0_b + rank * BC_NSK + row) : 0.f;
      stage_sync();
      if (warp == 4) {
        w4_wait(wbar2, ph2);
        mma_issue(D2, wS2, BC_NSK * 16, xpl, 0, BC_R / 16, true);
        mma_commit();
        // early part of layer 0's gated conv: older taps + condition (they do not depend on this step's chain)
        w4_wait(wbar1, ph1);
        w4_wait(tapbar, phtap);
        mma_issue(D1, wS1, BC_ROWS_S1 * 16, xpl, KS_CUR, BC_K1 / 16, true);
      }
      acc_wait();
      BC_PF_ADD(2);
      issue_w(wS2, p.layers[0].w2 + (size_t)rank * BC_ROWS_S2 * BC_G, BC_ROWS_S2 * BC_G * 2, wbar2);
      if (warp < 4) {
        bc_ld16(my_taddr + D2, v);
        if (row < BC_NSK) {
#pragma unroll
          for (int n = 0; n < 16; ++n) skip32[n * BC_NSK + row] = v[n] + bias_s0;
        }
      }
      BC_PF_ADD(3);
    }

    // ================================================================== residual stacks
    for (int l = 0; l < L; ++l) {
      const BcLayerDev ly = p.layers[l];
      const bool last = (l == L - 1);
      // ---------------------------------------------------------------- S1: dilated conv + condition + gate
      pf_cls = 0;
      // biases of this thread's accumulator rows: requested before the MMA chain so the L2 latency hides behind it
      const float bias_s1 = (warp < 2) ? __ldg(ly.b1 + ((row & 1) ? BC_G : 0) + rank * 32 + (row >> 1)) : 0.f;
      const float bias_s2 = (warp < 4 && row < BC_ROWS_S2)
                                ? __ldg(ly.b2 + (row < BC_NR ? rank * BC_NR + row : BC_R + rank * BC_NSK + (row - BC_NR))) : 0.f;
      stage_sync5();
      if (warp == 4) {
        if (l > 0) w4_recv(cbar, phc, RX_C);
        mma_issue(D1, wS1, BC_ROWS_S1 * 16, xpl, 0, KS_CUR, false);     // onto the early part
        mma_commit();
      }
      acc_wait();
      BC_PF_ADD(2);
      if (!last) {
        issue_w(wS1, p.layers[l + 1].w1 + (size_t)rank * BC_ROWS_S1 * BC_K1, BC_ROWS_S1 * BC_K1 * 2, wbar1);
        issue_taps(l + 1, t);
      } else {
        issue_w(wS1, p.post1 + (size_t)rank * BC_NSK * (BC_S + BC_C), BC_NSK * (BC_S + BC_C) * 2, wbar1);
      }
      if (warp < 2) {
        // rows 0-63: pre-activation + bias -> pre[stream][row] (even row tanh input, odd row its sigmoid partner)
        bc_ld16(my_taddr + D1, v);
#pragma unroll
        for (int n = 0; n < 16; ++n) pre[n * BC_ROWS_S1 + row] = v[n] + bias_s1;
      }
      ep_sync();
      // 512 gates over the 128 epilogue threads (wavenet_ops.py:236-240)
      if (warp < 4) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int idx = tid + h * 128;
          const int n = idx >> 5, j = idx & 31;
          const float2 ab = *reinterpret_cast<const float2*>(pre + n * BC_ROWS_S1 + 2 * j);
          bc_st_stage(stage, n, j, bc_gate(ab.x, ab.y));
        }
      }
      ep_sync();
      BC_PF_ADD(3);
      bc_push(stage, bsm, BC_OFF_G + rank * 4 * BC_PLANE_B, 4 * BC_PLANE_B / 16, BC_CS, gbar, 128);
      BC_PF_ADD(5);

      // ---------------------------------------------------------------- S2: residual + skip 1x1
      pf_cls = 1;
      stage_sync5();
      if (warp == 4) {
        w4_recv(gbar, phg, RX_G);
        w4_wait(wbar2, ph2);
        mma_issue(D2, wS2, BC_ROWS_S2 * 16, gpl, 0, BC_G / 16, true);
        mma_commit();
        if (!last) {
          // early part of the next layer's gated conv runs on the tensor pipe behind this chain, while the epilogue,
          // the pushes and the wait for the next layer input proceed
          w4_wait(wbar1, ph1);
          w4_wait(tapbar, phtap);
          mma_issue(D1, wS1, BC_ROWS_S1 * 16, xpl, KS_CUR, BC_K1 / 16, true);
        }
      }
      acc_wait();
      BC_PF_ADD(2);
      if (!last) issue_w(wS2, p.layers[l + 1].w2 + (size_t)rank * BC_ROWS_S2 * BC_G, BC_ROWS_S2 * BC_G * 2, wbar2);
      else issue_w(wS2, p.post2 + (size_t)rank * BC_NQ * BC_S, BC_NQ * BC_S * 2, wbar2);
      if (warp < 4) {
        bc_ld16(my_taddr + D2, v);
        BC_PF_ADD(7);
        // (all loads first: the byte stores into the staging planes may alias anything as far as the compiler knows)
        if (row < BC_NR) {
          float oldv[16];
#pragma unroll
          for (int n = 0; n < 16; ++n) oldv[n] = cur32[n * BC_NR + row];
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            const float nv = oldv[n] + (v[n] + bias_s2);
            cur32[n * BC_NR + row] = nv;
            if (!last) bc_st_stage(stage, n, row, nv);                       // next layer input slice (4 planes); dead after the last layer
            bc_st_stage(stage + 8 * BC_STAGE_PLANE, n, row, oldv[n]);        // push_ops: this step's layer input goes to the queue
          }
        } else if (row < BC_ROWS_S2) {
          const int c = row - BC_NR;
          float sk[16];
#pragma unroll
          for (int n = 0; n < 16; ++n) sk[n] = skip32[n * BC_NSK + c];
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            sk[n] += v[n] + bias_s2;
            skip32[n * BC_NSK + c] = sk[n];
            if (last) bc_st_stage(stage, n, c, fmaxf(sk[n], 0.f));           // wavenet.py:153 (the last residual is dead, :145)
          }
        }
      }
      BC_PF_ADD(6);
      // (the MMA warp is still issuing the next layer's early part: only the epilogue warps meet here)
      ep_sync();
      {
        // queue push: this CTA's 4 planes of ring slot t mod 2d
        const int slot_old = (t < 0x7fffffffLL) ? (int)((unsigned)t % (unsigned)(2 * ly.d)) : (int)(t % (2 * ly.d));
        if (tid < 4 * BC_PLANE_B / 16) {
          const float4 x = *reinterpret_cast<const float4*>(stage + 8 * BC_STAGE_PLANE + tid * 16 + (tid >> 4) * 16);
          __nv_bfloat16* dst = ly.ring + slot_old * ring_slot_elems + (long long)cluster * BC_R * BC_NS + (rank * 4) * (BC_PLANE_B / 2);
          *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(dst) + tid * 16) = x;
          // generic-proxy ring store, read by a later step's bulk copy (async proxy)
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
      }
      BC_PF_ADD(3);
      if (!last) bc_push(stage, bsm, BC_OFF_X + rank * 4 * BC_PLANE_B, 4 * BC_PLANE_B / 16, BC_CS, cbar, 128);
      else bc_push(stage, bsm, BC_OFF_SKF + rank * 8 * BC_PLANE_B, 8 * BC_PLANE_B / 16, BC_CS, skbar, 128);
      BC_PF_ADD(5);
    }
    pf_cls = 2;

    // ================================================================== postprocess1 (+ condition), relu
    const float bias_p1 = (warp < 2) ? __ldg(p.post1_b + rank * BC_NSK + row) : 0.f;
    const float bias_p2 = (warp < 1) ? __ldg(p.post2_b + rank * BC_NQ + row) : 0.f;
    stage_sync5();
    if (warp == 4) {
      w4_recv(skbar, phsk, RX_S);
      w4_wait(wbar1, ph1);
      mma_issue(D1, wS1, BC_NSK * 16, bsm + BC_OFF_SKF, 0, (BC_S + BC_C) / 16, true);
      mma_commit();
    }
    acc_wait();
    if (t + 1 < p.t0 + p.T) issue_w(wS1, p.layers[0].w1 + (size_t)rank * BC_ROWS_S1 * BC_K1, BC_ROWS_S1 * BC_K1 * 2, wbar1);
    if (warp < 4) {
      bc_ld16(my_taddr + D1, v);
      if (row < BC_NSK) {
#pragma unroll
        for (int n = 0; n < 16; ++n) bc_st_stage(stage, n, row, fmaxf(v[n] + bias_p1, 0.f));     // wavenet.py:163
      }
    }
    ep_sync();
    bc_push(stage, bsm, BC_OFF_N1F + rank * 8 * BC_PLANE_B, 8 * BC_PLANE_B / 16, BC_CS, n1bar, 128);

    // ================================================================== postprocess2 -> logits on CTA 0
    stage_sync5();
    if (warp == 4) {
      w4_recv(n1bar, phn1, RX_S);
      w4_wait(wbar2, ph2);
      mma_issue(D2, wS2, BC_NQ * 16, bsm + BC_OFF_N1F, 0, BC_S / 16, true);
      mma_commit();
    }
    acc_wait();
    if (t + 1 < p.t0 + p.T) issue_w(wS2, p.skip0 + (size_t)rank * BC_NSK * BC_R, BC_NSK * BC_R * 2, wbar2);
    if (warp < 4) {
      bc_ld16(my_taddr + D2, v);
      if (row < BC_NQ) {
        float* st = pre;                                      // [16 streams][32 logits] fp32
#pragma unroll
        for (int n = 0; n < 16; ++n) st[n * BC_NQ + row] = v[n] + bias_p2;
      }
    }
    __syncthreads();
    {
      // logits slice -> CTA 0: stream n, columns 32*rank .. +32 (8 chunks of 16 bytes per stream)
      const unsigned base = f32_smem_u32(logits_s);
      if (tid < BC_NS * 8) {
        const int n = tid >> 3, c = tid & 7;
        const float4 x = *reinterpret_cast<const float4*>(pre + n * BC_NQ + c * 4);
        cl_st_v4(cl_mapa(base + (unsigned)(n * BC_Q + rank * BC_NQ + c * 4) * 4u, 0u), x);
      }
    }
    cl_barrier();

    // ================================================================== softmax + draw + mu-law decode (CTA 0)
    if (rank == 0) {
      for (int i = warp; i < nvalid; i += BC_THREADS / 32) {
        const int b = b0 + i;
        float lg[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) lg[q] = logits_s[i * BC_Q + lane + 32 * q];
        const int k = warp_softmax_draw(p, BC_Q, lg, b, t, lane, logits_s + i * BC_Q);
        if (k >= 0 && lane == 0) {
          const float un = __ldg(p.enc_lut + k);
          const int slot_n = (int)((t + 1) % BC_PK);
          st_cg(p.u_hist + (long long)b * BC_PK + slot_n, un);
          const unsigned a = f32_smem_u32(hist + i * BC_PK + slot_n);
          for (unsigned r = 0; r < (unsigned)BC_CS; ++r) cl_st_f32(cl_mapa(a, r), un);
        }
      }
    }
    cl_barrier();
  }
  cl_barrier();
  if (prof) for (int i = 0; i < 24; ++i) p.prof[i] = pf[i];
#undef BC_PF_START
#undef BC_PF_ADD
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

// float32 [K][ldw] row-major -> bf16 K-major planes per cluster CTA: dst[cb][plane][row][8], row -> source column
//   mode 0: column = cb*rows + row                                      (skip start, post1, post2)
//   mode 1: row 2j -> cb*32 + j, row 2j+1 -> base1 + cb*32 + j          (S1: tanh / sigmoid partner interleaved)
//   mode 2: row < n0 -> cb*n0 + row, else base1 + cb*(rows-n0) + row-n0 (S2: residual | skip)
__global__ void pack_bf16_planes_kernel(const float* __restrict__ src, int ldw, int K, int rows, int mode, int n0, int base1,
                                        int ntiles, __nv_bfloat16* __restrict__ dst) {
  const long long total = (long long)ntiles * K * rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % 8);
    long long r = i / 8;
    const int row = (int)(r % rows); r /= rows;
    const int plane = (int)(r % (K / 8));
    const int cb = (int)(r / (K / 8));
    const int k = plane * 8 + e;
    int col;
    if (mode == 0) col = cb * rows + row;
    else if (mode == 1) col = (row & 1) ? (base1 + cb * (rows / 2) + (row >> 1)) : (cb * (rows / 2) + (row >> 1));
    else col = (row < n0) ? (cb * n0 + row) : (base1 + cb * (rows - n0) + (row - n0));
    dst[i] = __float2bfloat16_rn(src[(long long)k * ldw + col]);
  }
}

}  // namespace vqwn
