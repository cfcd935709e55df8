// WaveNet fast generation on the 5th-generation tensor cores at float32-grade accuracy (VQWN_PREC_TC).
// Same reference semantics as the float32 kernels (wavenet.py:103-172, wavenet_ops.py:163-267, utils.py:13-46,
// mu_law_ops.py:5-31).
//
// Arithmetic.  Every contraction operand x is split into two bfloat16 numbers x = hi + lo (hi = rn(x), lo = rn(x - hi),
// |x - hi - lo| <= 2^-17 |x|).  The weight tile stacks hi rows and lo rows along M, the activation tile stacks the hi
// and lo copies of the streams along N, so ONE tcgen05.mma (kind::f16, fp32 accumulation in TMEM) produces all four
// partial products W_hi.a_hi, W_hi.a_lo, W_lo.a_hi, W_lo.a_lo in four quadrants of the accumulator; the epilogue adds
// them.  Biases, residual chain, skip sum, gate (tanhf / expf), softmax and the draw are float32 as in the reference.
//
// Structure.  A cluster of 16 CTAs owns up to 16 streams for the whole run; every stage is split by output channels
// (CTA r: gate channels 16r..16r+15 as tanh/sigmoid pairs, residual 16r.., skip 32r.., post1 32r.., logits 16r..).
//   * weights: bf16 hi/lo tiles, K-major no-swizzle planes (plane p = k 8p..8p+7 of every row, 16 B per row), streamed
//     from L2 with one cp.async.bulk per tile into four single buffers (current tap | tap t-d | tap t-2d | residual+skip),
//     each re-filled as soon as the MMAs that read it have completed (tcgen05.commit);
//   * activations: a CTA publishes its 16-channel slice (one K chunk of 16 = 1 KB of hi/lo bf16) through L2: one warp
//     copies the staged slice to global memory and issues ONE multicast bulk copy (cp.async.bulk ... .multicast::cluster)
//     that delivers it to the same offset of all 16 CTAs and counts its bytes on each receiver's mbarrier - measured 920
//     cycles per 16-CTA all-gather round against 1650 for per-thread st.async and 1410 for 16 shared::cta ->
//     shared::cluster bulk copies (tools/r2_probe.cu).  The layer-input slices are published straight from the dilation
//     ring slot of the step, so the queue push (push_ops) and the hand-off are the same 1 KB store;
//   * the parts of a layer's dilated conv that do not depend on this step's chain (taps t-d, t-2d from the HBM dilation
//     rings) are issued into a second TMEM accumulator in the two hand-off gaps of the previous layer;
//   * the local-condition projections (wavenet_ops.py:198-209; 30 layers + postprocess1) change once per `ratio` steps:
//     they are recomputed in float32 on the CUDA cores at each frame change into a per-CTA L2 table and added in the
//     epilogue (with the conv bias folded in);
//   * the skip path is off the chain: its rows ride in the residual MMA, are accumulated in registers across the layers,
//     and the skip start (wavenet.py:127-128) is folded into a 32-tap FIR with host-premultiplied weights;
//   * the draw is distributed: CTA s owns stream s (softmax, greedy / sequential-cumsum sample, mu-law LUT) and
//     broadcasts the new network input to the cluster.
// The dilation queues are HBM rings [2d + 1][cluster][32 planes][32 rows = 16 hi + 16 lo][8] bf16 in operand layout (slot
// t mod (2d + 1) = layer input of step t; one slot more than the reference's two chained queues of capacity d because the
// slot of step t is written while the tap t - 2d is still being read), so a tap is one 16 KB bulk copy straight into the B
// operand.  One split cluster barrier per step orders the ring stores of a step before the bulk reads of the next.  Geometry fixed to the reference's default (R = G = 256, S = 512, Q = 256,
// C = 128, 32-tap preprocess, kernel_size 3).
#pragma once
#include <cuda_bf16.h>
#include "wavenet_fp32_cluster.cuh"
#include "wavenet_bf16_cluster.cuh"

namespace vqwn {

constexpr int TC_CS = 16;
constexpr int TC_THREADS = 256;
constexpr int TC_NS = 16;                 // streams per cluster (at most)
constexpr int TC_NN = 32;                 // MMA N: 16 hi rows + 16 lo rows
constexpr int TC_R = 256, TC_G = 256, TC_S = 512, TC_Q = 256, TC_C = 128, TC_PK = 32;
constexpr int TC_MAXL = 64;
constexpr int TC_PLANE = TC_NN * 16;      // bytes of one activation plane
constexpr int TC_XB = 32 * TC_PLANE;      // K = 256 activation operand: 16 KB
constexpr int TC_ROWS1 = 64, TC_ROWS2 = 96, TC_ROWSP1 = 64, TC_ROWSP2 = 32;
constexpr int TC_W1 = 32 * TC_ROWS1 * 16;     // one tap of the gated conv: 32 KB
constexpr int TC_W2 = 32 * TC_ROWS2 * 16;     // residual + skip: 48 KB
constexpr int TC_WP1 = 64 * TC_ROWSP1 * 16;   // postprocess1 (K = 512): 64 KB
constexpr int TC_WP2 = 64 * TC_ROWSP2 * 16;   // postprocess2 (K = 512): 32 KB
constexpr int TC_LAYER_BYTES = 3 * TC_W1 + TC_W2;      // per (layer, CTA): 144 KB
// shared memory map (bytes)
constexpr int TC_OFF_WA = 0;                            // current-tap weights | first half of postprocess1
constexpr int TC_OFF_WB = TC_OFF_WA + TC_W1;            // tap t-d weights     | second half of postprocess1
constexpr int TC_OFF_WC = TC_OFF_WB + TC_W1;            // tap t-2d weights
constexpr int TC_OFF_WD = TC_OFF_WC + TC_W1;            // residual + skip     | postprocess2
constexpr int TC_OFF_XC = TC_OFF_WD + TC_W2;            // layer input         } postprocess2 input (K = 512)
constexpr int TC_OFF_XG = TC_OFF_XC + TC_XB;            // gate output         }
constexpr int TC_OFF_XT1 = TC_OFF_XG + TC_XB;           // tap t-d             } postprocess1 input (K = 512)
constexpr int TC_OFF_XT2 = TC_OFF_XT1 + TC_XB;          // tap t-2d            }
constexpr int TC_OFF_STG = TC_OFF_XT2 + TC_XB;          // push staging: 4 planes
constexpr int TC_OFF_HIST = TC_OFF_STG + 4 * TC_PLANE;  // [16][32] fp32 network-input history ring (remote-written)
constexpr int TC_OFF_LOG = TC_OFF_HIST + TC_NS * TC_PK * 4;   // [256] fp32 logits of this CTA's stream (remote-written)
constexpr int TC_OFF_US = TC_OFF_LOG + TC_Q * 4;        // [16][32] history in tap order     } 8 KB of CTA-local scratch,
constexpr int TC_OFF_CUR0 = TC_OFF_US + TC_NS * TC_PK * 4;    // [16 ch][16] fp32 FIR output  } aliased by the condition rows
constexpr int TC_OFF_SKF = TC_OFF_CUR0 + 16 * TC_NS * 4;      // [32 ch][16] skip FIR part    } [16][128] fp32 at a frame
constexpr int TC_OFF_SKX = TC_OFF_SKF + 32 * TC_NS * 4;       // [32 ch][16] skip lo sums     } change
constexpr int TC_OFF_PROB = TC_OFF_SKX + 32 * TC_NS * 4;      // [256] fp32 draw scratch      }
constexpr int TC_OFF_LAYERS = TC_OFF_PROB + TC_Q * 4;
constexpr int TC_OFF_BARS = TC_OFF_LAYERS + TC_MAXL * 48;
constexpr int TC_NBARS = 16;
constexpr int TC_OFF_MISC = TC_OFF_BARS + TC_NBARS * 8;
constexpr int TC_SMEM = TC_OFF_MISC + 16;
static_assert(TC_OFF_PROB + TC_Q * 4 - TC_OFF_US == TC_NS * TC_C * 4, "condition rows alias exactly the local scratch");
static_assert(TC_SMEM <= 232448, "shared memory budget");
// per-cluster global staging of the hand-offs that do not live in a ring: gate output (double-buffered by layer parity),
// postprocess1 input, postprocess2 input
constexpr int TC_GST_XG = 0, TC_GST_XS = 2 * TC_XB, TC_GST_XN = TC_GST_XS + 2 * TC_XB, TC_GSTAGE = TC_GST_XN + 2 * TC_XB;

struct TcLayerDev {
  const __nv_bfloat16* w;     // [16 CTAs][W_A | W_B | W_C | W_D] hi/lo plane tiles
  const float* wlc;           // gated/local_condition/kernel [C][2G] float32 (row stride 2G)
  const float* b1;            // gated/bias [2G]
  const float* bres;          // residual/bias [R]
  __nv_bfloat16* ring;        // [2d + 1][nclusters][32 planes][32][8]
  int d;
  int pad_;
};
static_assert(sizeof(TcLayerDev) == 48, "layer record size");

struct TcParams {
  int L, B, nclusters, cluster0, spc;      // spc: streams per cluster; cluster c owns streams [c*spc, c*spc + spc)
  const float *pre_k, *pre_b;              // preprocess/kernel [32][256], bias [256]
  const float *skf_k, *skf_b;              // skip start folded into the FIR: [32][512] = pre_k . skip/kernel; [512] all skip biases
  const __nv_bfloat16 *post1, *post2;      // [16][64 KB], [16][32 KB]
  const float *post1_lc, *post1_b, *post2_b;   // postprocess1/local_condition/kernel [C][S], biases
  const TcLayerDev* layers;
  float* ctab;                             // [launch cluster][16][L+1][512] condition table
  uint8_t* gstage;                         // [launch cluster][TC_GSTAGE] hand-off staging in L2
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
  int b_offset;                            // global index of stream 0 (sharded runs): keys the seeded generator
  int flags;                               // reserved
  float* audio_out;
  int* idx_out;
  float* logits_out;
  float* probs_out;
  long long* prof;
  int* err;
};

__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// element (row n, channel k of the slice) of an operand-layout block: hi copy in row n, lo copy in row 16 + n
__device__ __forceinline__ void tc_st_split(uint8_t* planes, int n, int k, float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  uint8_t* q = planes + (k >> 3) * TC_PLANE + n * 16 + (k & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(q) = hi;
  *reinterpret_cast<__nv_bfloat16*>(q + 16 * 16) = lo;
}
// gate nonlinearities on the special-function unit: ex2.approx / rcp.approx are accurate to ~2 ulp, the results to
// ~4e-7 absolute - below the 2^-17 operand split of the contractions that consume them
__device__ __forceinline__ float tc_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tc_tanh(float x) {
  const float x2 = x * x;
  // |x| < 0.25: odd Taylor polynomial (next term 62/2835 x^9 < 2e-8 relative); otherwise 2 sigmoid(2x) - 1
  const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  const float big = fmaf(2.0f, tc_sigmoid(2.0f * x), -1.0f);
  return (x2 < 0.0625f) ? poly : big;
}
__device__ __forceinline__ void tc_st_async_f32(unsigned addr, float v, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(addr), "r"(__float_as_uint(v)), "r"(mbar) : "memory");
}

// one bulk copy global -> the same shared-memory offset of every CTA in `mask`; each receiver's mbarrier (same offset)
// counts the bytes
__device__ __forceinline__ void tc_bulk_multicast(unsigned dst_smem, const void* gsrc, unsigned bytes, unsigned mbar, unsigned short mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst_smem), "l"(gsrc), "r"(bytes), "r"(mbar), "h"(mask) : "memory");
}

__device__ __forceinline__ bool tc_test_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(f32_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(TC_THREADS, 1) wavenet_tc_cluster(const TcParams p_in) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TcLayerDev* const layers_s = reinterpret_cast<TcLayerDev*>(sm + TC_OFF_LAYERS);
  for (int i = tid; i < p_in.L * (int)(sizeof(TcLayerDev) / 4); i += TC_THREADS)
    reinterpret_cast<uint32_t*>(layers_s)[i] = reinterpret_cast<const uint32_t*>(p_in.layers)[i];
  TcParams p = p_in;
  p.layers = layers_s;
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int lcluster = (int)blockIdx.x / TC_CS;
  const int cluster = p.cluster0 + lcluster;
  const int b0 = cluster * p.spc;
  const int nvalid = max(0, min(p.spc, p.B - b0));
  const int L = p.L;

  uint8_t* const xc = sm + TC_OFF_XC;
  uint8_t* const stg = sm + TC_OFF_STG;
  float* const hist = reinterpret_cast<float*>(sm + TC_OFF_HIST);
  float* const logits_s = reinterpret_cast<float*>(sm + TC_OFF_LOG);
  float* const u_s = reinterpret_cast<float*>(sm + TC_OFF_US);
  float* const cur0 = reinterpret_cast<float*>(sm + TC_OFF_CUR0);
  float* const skf = reinterpret_cast<float*>(sm + TC_OFF_SKF);
  float* const skx = reinterpret_cast<float*>(sm + TC_OFF_SKX);
  float* const prob_s = reinterpret_cast<float*>(sm + TC_OFF_PROB);
  float* const ct_rows = reinterpret_cast<float*>(sm + TC_OFF_US);     // [16][128] condition rows (alias)
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sm + TC_OFF_BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + TC_OFF_MISC);
  unsigned long long* const wbarA = bars + 0;
  unsigned long long* const wbarB = bars + 1;
  unsigned long long* const wbarC = bars + 2;
  unsigned long long* const wbarD = bars + 3;
  unsigned long long* const tapbar1 = bars + 4;
  unsigned long long* const tapbar2 = bars + 5;
  unsigned long long* const accbar = bars + 6;
  unsigned long long* const lgbar = bars + 7;
  unsigned long long* const smpbar = bars + 8;
  unsigned long long* const auxbar = bars + 9;     // the step-boundary tap MMAs completed
  unsigned long long* const xcbar = bars + 10;     // next layer input: the slices of all 16 CTAs landed
  unsigned long long* const xgbar = bars + 11;     // gate output slices landed
  unsigned long long* const xsbar = bars + 12;     // relu(skip) slices landed (postprocess1 input)
  unsigned long long* const xnbar = bars + 13;     // postprocess1 output slices landed (postprocess2 input)

  for (int i = tid; i < (TC_OFF_LAYERS - TC_OFF_XC) / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(sm + TC_OFF_XC)[i] = 0u;
  __syncthreads();
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  const unsigned RX1 = TC_CS * 2u * TC_PLANE;     // 16 senders x 16-channel slice (2 planes, all 32 rows)
  const unsigned RX2 = TC_CS * 4u * TC_PLANE;     // 16 senders x 32-channel slice
  if (tid == 0) {
    for (int i = 0; i < TC_NBARS; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(f32_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect(xcbar, RX1);
    mbar_expect(xgbar, RX1);
    mbar_expect(xsbar, RX2);
    mbar_expect(xnbar, RX2);
    if (rank < nvalid) mbar_expect(lgbar, TC_Q * 4);
    if (!ext) mbar_expect(smpbar, 4u * (unsigned)nvalid);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(f32_smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < nvalid * TC_PK; i += TC_THREADS) hist[i] = ld_cg(p.u_hist + (long long)b0 * TC_PK + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  cl_barrier();      // every CTA's barriers are initialised and armed before any remote st.async can arrive

  const float mu = (float)(TC_Q - 1);
  unsigned phA = 0u, phB = 0u, phC = 0u, phD = 0u, pht1 = 0u, pht2 = 0u, phacc = 0u, phlg = 0u, phsmp = 0u;
  unsigned phxc = 0u, phxg = 0u, phxs = 0u, phxn = 0u, phaux = 0u;
  // in-kernel cycle counters (VQWN_PROFILE=1): CTA 0 of the first cluster, thread 0 (epilogue view) and thread 128 (MMA view)
  const bool prof = (p.prof != nullptr) && blockIdx.x == 0 && p.cluster0 == 0 && (tid == 0 || tid == 128);
  long long pf[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) pf[i] = 0;
  long long pf_t = 0;
#define TC_PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define TC_PF_ADD(i) do { if (prof) { const long long n_ = clock64(); pf[(i)] += n_ - pf_t; pf_t = n_; } } while (0)
  uint32_t elected = 0;
  if (warp == 4) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(elected));
  // instruction descriptors: fp32 accumulate, bf16 x bf16, N = 32; M = 128, and M = 64 for the gated-conv tiles (64 real
  // rows: half the A-operand traffic; its accumulator occupies lanes 0-15 of each of the four TMEM lane quarters)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NN >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NN >> 3) << 17) | ((64u >> 4) << 24);
  constexpr uint32_t D1A = 0, D1B = 32, D2 = 64;     // TMEM columns: gated conv (double-buffered) | everything else
  const uint32_t sm_u32 = f32_smem_u32(sm);

  // ------------------------------------------------------------------ helpers
  auto wait_bar = [&](unsigned long long* bar, unsigned& ph) {
    (void)mbar_wait_bounded(bar, ph, p.err);
    ph ^= 1u;
  };
  auto acc_wait = [&]() {
    wait_bar(accbar, phacc);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // weight tile copies: one elected thread per loader warp
  auto issue_w = [&](int dst_off, const __nv_bfloat16* src, unsigned bytes, unsigned long long* bar) {
    mbar_expect(bar, bytes);
    cl_bulk_g2s_keep(reinterpret_cast<float*>(sm + dst_off), reinterpret_cast<const float*>(src), bytes, bar);
  };
  const long long ring_slot_elems = (long long)p.nclusters * (TC_XB / 2);
  // which: 1 = tap t-d, 2 = tap t-2d; slot of step s = s mod (2d + 1)
  auto issue_tap = [&](int l, long long t, int which) {
    const TcLayerDev& ly = p.layers[l];
    const long long nslot = 2ll * ly.d + 1;
    const long long slot = (which == 1) ? ((t + ly.d + 1) % nslot) : ((t + 1) % nslot);
    unsigned long long* bar = (which == 1) ? tapbar1 : tapbar2;
    mbar_expect(bar, TC_XB);
    bulk_g2s(reinterpret_cast<float*>(sm + (which == 1 ? TC_OFF_XT1 : TC_OFF_XT2)),
             reinterpret_cast<const float*>(ly.ring + slot * ring_slot_elems + (long long)cluster * (TC_XB / 2)), TC_XB, bar);
  };
  // this CTA's 1 KB slice of the ring slot that holds layer l's input of step t
  auto ring_slice = [&](int l, long long t) {
    const TcLayerDev& ly = p.layers[l];
    const long long slot = t % (2ll * ly.d + 1);
    return reinterpret_cast<uint8_t*>(ly.ring + slot * ring_slot_elems + (long long)cluster * (TC_XB / 2)) + rank * 2 * TC_PLANE;
  };
  auto layer_w = [&](int l, int off_bytes) {
    return reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const uint8_t*>(p.layers[l].w) +
                                                  (size_t)rank * TC_LAYER_BYTES + off_bytes);
  };
  // epilogue warps (0-3) + MMA warp (4): TMEM reads of the finished epilogues are ordered before the MMAs that overwrite
  // those accumulators, and the staging buffers are free again
  auto stage_sync5 = [&]() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 3, 160;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  auto ep_sync = [&]() { asm volatile("bar.sync 2, 128;" ::: "memory"); };
  // a chain of nks MMAs, one per K chunk of 16: D[M x 32] (TMEM column d_col) (+)= A[M x 16] . B[32 x 16]^T.  Only the
  // elected thread runs the loop (a real branch: predicating the instruction for the whole warp makes the compiler wrap
  // every MMA in a per-lane loop), the descriptors advance by a constant - an MMA of this size occupies the pipe for ~64
  // cycles whatever M and N are (tools/r2_probe.cu), the loop must not cost more than that
  auto mma_chain = [&](uint32_t d_col, int a_off, uint32_t a_rows, int b_off, int nks, bool fresh, bool m64) {
    if (elected) {
      const uint32_t a_lbo = a_rows * 16u;
      const uint32_t id = m64 ? idesc64 : idesc;
      uint64_t da = bc_desc(sm_u32 + (uint32_t)a_off, a_lbo);
      uint64_t db = bc_desc(sm_u32 + (uint32_t)b_off, TC_PLANE);
      const uint64_t sa = (uint64_t)((2u * a_lbo) >> 4), sb = (uint64_t)((2u * TC_PLANE) >> 4);
      uint32_t accf = fresh ? 0u : 1u;
#pragma unroll 4
      for (int ks = 0; ks < nks; ++ks) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem + d_col), "l"(da), "l"(db), "r"(id), "r"(accf) : "memory");
        da += sa; db += sb; accf = 1u;
      }
    }
    __syncwarp();
  };
  auto mma_commit_to = [&](unsigned long long* bar) {     // bar fires when every MMA issued so far has completed
    if (elected)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(f32_smem_u32(bar)) : "memory");
    __syncwarp();
  };
  auto operand_fence = [&]() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // warp 4: nks K chunks of a stage whose B operand is pushed by the 16 CTAs of the cluster (remote) or written locally
  auto w4_chain = [&](uint32_t d_col, int a_off, uint32_t a_rows, int b_off, int nks, unsigned long long* sbar, unsigned& sph,
                      bool remote, bool fresh, unsigned rearm_bytes, bool m64) {
    if (remote) wait_bar(sbar, sph);
    operand_fence();
    mma_chain(d_col, a_off, a_rows, b_off, nks, fresh, m64);
    if (remote && lane == 0) mbar_expect(sbar, rearm_bytes);
  };
  // warp 4: the chain-independent part of layer l's gated conv (one older tap), into that layer's accumulator
  auto w4_tap = [&](int l, int which) {
    const uint32_t dcol = (l & 1) ? D1B : D1A;
    if (which == 1) { wait_bar(wbarB, phB); wait_bar(tapbar1, pht1); }
    else { wait_bar(wbarC, phC); wait_bar(tapbar2, pht2); }
    operand_fence();
    const int a_off = which == 1 ? TC_OFF_WB : TC_OFF_WC, b_off = which == 1 ? TC_OFF_XT1 : TC_OFF_XT2;
    mma_chain(dcol, a_off, TC_ROWS1, b_off, 16, which == 1, true);
  };
  // warp 0: publish the staged slice (nplanes planes of this CTA's channels, stg) through L2: copy it to `gdst`, then ONE
  // multicast bulk copy delivers it to offset dst_off of all 16 CTAs and counts its bytes on every receiver's `sbar`
  auto publish = [&](uint8_t* gdst, int nplanes, int dst_off, unsigned long long* sbar) {
    for (int c = lane; c < nplanes * (TC_PLANE / 16); c += 32)
      *reinterpret_cast<float4*>(gdst + c * 16) = *reinterpret_cast<const float4*>(stg + c * 16);
    asm volatile("fence.proxy.async.global;" ::: "memory");     // generic-proxy stores before the bulk copy's read
    __syncwarp();
    if (lane == 0)
      tc_bulk_multicast(sm_u32 + (unsigned)dst_off, gdst, (unsigned)(nplanes * TC_PLANE), f32_smem_u32(sbar), (unsigned short)0xFFFF);
  };
  uint8_t* const gst = p.gstage + (size_t)lcluster * TC_GSTAGE;

  const uint32_t my_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float* const ctab = p.ctab + ((size_t)lcluster * TC_CS + rank) * (size_t)(L + 1) * 512;

  // ------------------------------------------------------------------ prologue: first weights and taps of the run
  if (tid == 160) {
    issue_w(TC_OFF_WA, layer_w(0, 0), TC_W1, wbarA);
    issue_w(TC_OFF_WD, layer_w(0, 3 * TC_W1), TC_W2, wbarD);
  }
  if (tid == 224) {
    issue_w(TC_OFF_WB, layer_w(0, TC_W1), TC_W1, wbarB);
    issue_w(TC_OFF_WC, layer_w(0, 2 * TC_W1), TC_W1, wbarC);
  }
  if (tid == 192) {
    issue_tap(0, p.t0, 1);
    issue_tap(0, p.t0, 2);
  }
  if (warp == 4) {
    w4_tap(0, 1);
    w4_tap(0, 2);
    mma_commit_to(auxbar);
  }
  unsigned nA = 1, nT2 = 1;      // copies issued so far on wbarA / tapbar2 (development probe)
  long long cond_frame = -1;
  float cnd[8];       // warps 0-1: condition (+ bias) terms of the next gated conv / postprocess1 epilogue
#pragma unroll
  for (int j = 0; j < 8; ++j) cnd[j] = 0.f;

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {
    const long long frame_t = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    const bool more = (t + 1 < p.t0 + p.T);
    TC_PF_START();
    // ================================================================ all threads: frame change -> condition table
    if (frame_t != cond_frame) {
      for (int idx = tid; idx < TC_NS * (TC_C / 4); idx += TC_THREADS) {
        const int n = idx >> 5, c4 = idx & 31;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < nvalid)
          x = __ldg(reinterpret_cast<const float4*>(p.cond + (long long)(b0 + n) * p.cond_bstride + frame_t * TC_C) + c4);
        reinterpret_cast<float4*>(ct_rows)[idx] = x;
      }
      __syncthreads();
      const int cl_ = tid & 31, sg = tid >> 5;
      const float* r0 = ct_rows + (2 * sg) * TC_C;
      const float* r1 = r0 + TC_C;
#pragma unroll 1
      for (int st = 0; st <= L; ++st) {
        const float* wsrc;
        int ld, col;
        float bias;
        if (st < L) {
          wsrc = p.layers[st].wlc; ld = 2 * TC_G;
          col = ((cl_ >> 4) ? TC_G : 0) + 16 * rank + (cl_ & 15);
          bias = __ldg(p.layers[st].b1 + col);
        } else {
          wsrc = p.post1_lc; ld = TC_S;
          col = 32 * rank + cl_;
          bias = __ldg(p.post1_b + col);
        }
        float a0 = bias, a1 = bias;
#pragma unroll 8
        for (int k = 0; k < TC_C; ++k) {
          const float w = __ldg(wsrc + (size_t)k * ld + col);
          a0 = fmaf(r0[k], w, a0);
          a1 = fmaf(r1[k], w, a1);
        }
        float* dst = ctab + (size_t)st * 512;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = 2 * sg + h;
          const float a = h ? a1 : a0;
          int pos;
          if (st < L) { const int ch = cl_ & 15; pos = ((ch >> 2) * 16 + 4 * (s >> 2) + (ch & 3)) * 8 + (cl_ >> 4) * 4 + (s & 3); }
          else pos = ((cl_ >> 4) * 32 + 16 * (s >> 3) + (cl_ & 15)) * 8 + (s & 7);
          __stcg(dst + pos, a);
        }
      }
      cond_frame = frame_t;
      __syncthreads();
    }
    // ================================================================ all threads: history -> FIR -> layer-0 input, skip FIR part
    {
      const int slot_t = (int)(t % TC_PK);
      if (ext) {
        if (tid < TC_NS) {
          const int b = b0 + tid;
          float x = 0.f;
          if (tid < nvalid) {
            if (p.mode == GEN_STEP) x = p.ext_audio[b];
            else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
          }
          const float u = mu_law_encode_dev(x, mu, 0.f);
          hist[tid * TC_PK + slot_t] = u;
          if (rank == 0 && tid < nvalid) st_cg(p.u_hist + (long long)b * TC_PK + slot_t, u);
        }
        __syncthreads();
      }
      for (int idx = tid; idx < TC_NS * TC_PK; idx += TC_THREADS) {
        const int i = idx / TC_PK, j = idx - i * TC_PK;
        u_s[idx] = hist[i * TC_PK + ((slot_t - j) & (TC_PK - 1))];
      }
      __syncthreads();
      // warps 0-1 fetch the first layer's condition terms while the FIR runs
      if (warp < 4 && lane < 16) {
        const float4* src = reinterpret_cast<const float4*>(ctab + (warp * 16 + lane) * 8);
        const float4 a = __ldcg(src), b = __ldcg(src + 1);
        cnd[0] = a.x; cnd[1] = a.y; cnd[2] = a.z; cnd[3] = a.w; cnd[4] = b.x; cnd[5] = b.y; cnd[6] = b.z; cnd[7] = b.w;
      }
      // h0 = (u0*K[PK-1] + b) + u1*K[PK-2] + ...   thread = channel (wavenet_ops.py:178,193: kernel[k-1] is the current sample)
      {
        float acc[TC_NS];
        const float fb = __ldg(p.pre_b + tid);
#pragma unroll
        for (int i = 0; i < TC_NS; ++i) acc[i] = fb;
#pragma unroll 4
        for (int j = 0; j < TC_PK; ++j) {
          const float w = __ldg(p.pre_k + (TC_PK - 1 - j) * TC_R + tid);
#pragma unroll
          for (int i = 0; i < TC_NS; ++i) acc[i] = fmaf(u_s[i * TC_PK + j], w, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < TC_NS; ++i) {
          tc_st_split(xc, i, tid, acc[i]);
          if ((tid >> 4) == rank) cur0[(tid & 15) * TC_NS + i] = acc[i];
        }
      }
      // skip start folded into the FIR (wavenet.py:127-128): skf[ch][s], ch = 32 rank + (tid & 31), streams 2 (tid >> 5) + {0,1}
      {
        const int ch = tid & 31, sg = tid >> 5;
        const float* kp = p.skf_k + 32 * rank + ch;
        float a0 = __ldg(p.skf_b + 32 * rank + ch), a1 = a0;
        const float* u0 = u_s + (2 * sg) * TC_PK;
#pragma unroll 8
        for (int j = 0; j < TC_PK; ++j) {
          const float w = __ldg(kp + (TC_PK - 1 - j) * TC_S);
          a0 = fmaf(u0[j], w, a0);
          a1 = fmaf(u0[TC_PK + j], w, a1);
        }
        skf[ch * TC_NS + 2 * sg] = a0;
        skf[ch * TC_NS + 2 * sg + 1] = a1;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // push_ops of layer 0: this CTA's slice of the first layer's input -> its dilation ring (loader warps 5-6)
      if (tid >= 160 && tid < 224) {
        const int c = tid - 160;
        *reinterpret_cast<float4*>(ring_slice(0, t) + c * 16) = *reinterpret_cast<const float4*>(xc + rank * 2 * TC_PLANE + c * 16);
        asm volatile("fence.proxy.async.global;" ::: "memory");     // generic-proxy ring stores vs later bulk-copy reads
      }
    }
    TC_PF_ADD(0);

    if (warp == 4) {
      // ============================================================== MMA warp
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        if (l > 0) stage_sync5();
        wait_bar(wbarA, phA);
        TC_PF_ADD(1);
        w4_chain((l & 1) ? D1B : D1A, TC_OFF_WA, TC_ROWS1, TC_OFF_XC, 16, xcbar, phxc, l > 0, false, RX1, true);
        mma_commit_to(accbar);
        TC_PF_ADD(2);
        if (!last) w4_tap(l + 1, 1);
        TC_PF_ADD(3);
        stage_sync5();
        wait_bar(wbarD, phD);
        TC_PF_ADD(4);
        w4_chain(D2, TC_OFF_WD, TC_ROWS2, TC_OFF_XG, 16, xgbar, phxg, true, true, RX1, false);
        mma_commit_to(accbar);
        TC_PF_ADD(5);
        if (!last) w4_tap(l + 1, 2);
        TC_PF_ADD(6);
      }
      cl_arrive();
      stage_sync5();
      wait_bar(wbarA, phA);
      w4_chain(D2, TC_OFF_WA, TC_ROWSP1, TC_OFF_XT1, 32, xsbar, phxs, true, true, RX2, false);
      mma_commit_to(accbar);
      cl_wait();
      stage_sync5();
      wait_bar(wbarD, phD);
      w4_chain(D1B, TC_OFF_WD, TC_ROWSP2, TC_OFF_XC, 32, xnbar, phxn, true, true, RX2, false);
      mma_commit_to(accbar);
      TC_PF_ADD(7);
      if (more) {
        // next step's layer-0 older taps run on the tensor pipe during the draw and the FIR
        w4_tap(0, 1);
        w4_tap(0, 2);
        mma_commit_to(auxbar);
      }
      TC_PF_ADD(8);
    } else if (warp < 4) {
      // ============================================================== epilogue warps
      float v[32];
      float cur[8];       // warp 0: float32 residual chain, channel 16 rank + (lane & 15), streams 8 (lane >> 4) .. +8
      float sk[16];       // warp 1 / 2: skip sums (hi / lo rows) of channel 32 rank + lane, 16 streams
      if (warp == 0) {
        const int i = lane & 15, q = lane >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = cur0[i * TC_NS + 8 * q + j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = 0.f;
      }
#pragma unroll
      for (int s = 0; s < 16; ++s) sk[s] = (warp == 1) ? skf[lane * TC_NS + s] : 0.f;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const TcLayerDev& ly = p.layers[l];
        const bool last = (l == L - 1);
        const uint32_t d1 = (l & 1) ? D1B : D1A;
        // ---------------------------------------------------------------- S1: gate
        const float bres = (warp == 0) ? __ldg(ly.bres + 16 * rank + (lane & 15)) : 0.f;
        if (l > 0) stage_sync5();
        acc_wait();
        TC_PF_ADD(1);
        {
          // M = 64 accumulator: warp w reads rows 16 w .. 16 w + 15 in its lanes 0-15; lane = 4 qq + c: row type qq (tanh-hi |
          // sigmoid-hi | tanh-lo | sigmoid-lo) of gate channel 16 rank + 4 w + c.  Lanes 16-31 carry nothing.
          tc_ld32(my_taddr + d1, v);
          const int qq = (lane >> 2) & 3;
          float k8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo_ = v[j] + v[16 + j], hi_ = v[8 + j] + v[24 + j];       // streams j and 8 + j (hi + lo copies)
            const float keep = (qq & 2) ? hi_ : lo_, send = (qq & 2) ? lo_ : hi_;
            k8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);                // hi rows + lo rows
          }
          float mine[4], other[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mine[j] = (qq & 1) ? k8[4 + j] : k8[j];
            const float send = (qq & 1) ? k8[j] : k8[4 + j];
            other[j] = __shfl_xor_sync(0xffffffffu, send, 4);                    // tanh <-> sigmoid partner
          }
          if (lane < 16) {
            const int ch = 4 * warp + (lane & 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float at = ((qq & 1) ? other[j] : mine[j]) + cnd[j];
              const float as = ((qq & 1) ? mine[j] : other[j]) + cnd[4 + j];
              const float g = tc_tanh(at) * tc_sigmoid(as);                      // wavenet_ops.py:235-236
              tc_st_split(stg + (ch >> 3) * TC_PLANE, 4 * qq + j, ch & 7, g);
            }
          }
          // condition terms of the next epilogue of this kind: next layer (all four warps, lanes 0-15), or postprocess1
          // (warps 0-1, lane = 16 q + i)
          if (!last ? (lane < 16) : (warp < 2)) {
            const float4* src = reinterpret_cast<const float4*>(ctab + (size_t)(l + 1) * 512 + (last ? warp * 32 + lane : warp * 16 + lane) * 8);
            const float4 a = __ldcg(src), b = __ldcg(src + 1);
            cnd[0] = a.x; cnd[1] = a.y; cnd[2] = a.z; cnd[3] = a.w; cnd[4] = b.x; cnd[5] = b.y; cnd[6] = b.z; cnd[7] = b.w;
          }
        }
        TC_PF_ADD(2);
        ep_sync();
        if (warp == 0) publish(gst + TC_GST_XG + (l & 1) * TC_XB + rank * 2 * TC_PLANE, 2, TC_OFF_XG + rank * 2 * TC_PLANE, xgbar);
        TC_PF_ADD(3);
        // ---------------------------------------------------------------- S2: residual + skip 1x1
        stage_sync5();
        acc_wait();
        TC_PF_ADD(4);
        if (warp == 0) {
          // lane = 16 q + i: residual rows hi | lo of channel 16 rank + i
          tc_ld32(my_taddr + D2, v);
          const int q = lane >> 4, i = lane & 15;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo_ = v[j] + v[16 + j], hi_ = v[8 + j] + v[24 + j];
            const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
            const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            const float nv = cur[j] + (r + bres);                                // wavenet.py:145
            cur[j] = nv;
            if (!last) tc_st_split(stg, 8 * q + j, i, nv);                       // the last residual is dead (wavenet.py:145)
          }
          // next layer's input of this step: one 1 KB store into its dilation ring (push_ops) that is also the source
          // of the hand-off to the cluster
          if (!last) {
            __syncwarp();
            publish(ring_slice(l + 1, t), 2, TC_OFF_XC + rank * 2 * TC_PLANE, xcbar);
          }
        } else if (warp < 3) {
          // lanes 32-63: skip rows hi, lanes 64-95: skip rows lo of channel 32 rank + lane
          tc_ld32(my_taddr + D2, v);
#pragma unroll
          for (int s = 0; s < 16; ++s) sk[s] += v[s] + v[16 + s];
          if (last && warp == 2) {
#pragma unroll
            for (int s = 0; s < 16; ++s) skx[lane * TC_NS + s] = sk[s];
          }
        }
        TC_PF_ADD(5);
        if (last) {
          // relu(skip total) -> postprocess1 input slice (4 planes)
          ep_sync();
          if (warp == 1) {
#pragma unroll
            for (int s = 0; s < 16; ++s) tc_st_split(stg, s, lane, fmaxf(sk[s] + skx[lane * TC_NS + s], 0.f));     // wavenet.py:153
          }
          ep_sync();
          if (warp == 0) publish(gst + TC_GST_XS + rank * 4 * TC_PLANE, 4, TC_OFF_XT1 + rank * 4 * TC_PLANE, xsbar);
        }
        TC_PF_ADD(6);
      }
      // every ring store of this step is issued: split cluster barrier (waited for before the next step's tap loads)
      cl_arrive();
      // ================================================================ postprocess1 (+ condition), relu
      stage_sync5();
      acc_wait();
      cl_wait();
      const float bias_p2 = (warp == 0) ? __ldg(p.post2_b + 16 * rank + (lane & 15)) : 0.f;
      if (warp < 2) {
        // lane = 16 q + i: rows hi | lo of postprocess1 channel 32 rank + 16 warp + i
        tc_ld32(my_taddr + D2, v);
        const int q = lane >> 4, i = lane & 15;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo_ = v[j] + v[16 + j], hi_ = v[8 + j] + v[24 + j];
          const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
          const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16) + cnd[j];
          tc_st_split(stg + warp * 2 * TC_PLANE, 8 * q + j, i, fmaxf(r, 0.f));     // wavenet.py:163
        }
      }
      ep_sync();
      if (warp == 0) publish(gst + TC_GST_XN + rank * 4 * TC_PLANE, 4, TC_OFF_XC + rank * 4 * TC_PLANE, xnbar);
      TC_PF_ADD(7);
      // ================================================================ postprocess2 -> logits, scattered to the drawing CTAs
      stage_sync5();
      acc_wait();
      if (warp == 0) {
        // lane = 16 q + i: rows hi | lo of logit 16 rank + i
        tc_ld32(my_taddr + D1B, v);
        const int q = lane >> 4, i = lane & 15;
        const unsigned dst = f32_smem_u32(logits_s + 16 * rank + i);
        const unsigned mb = f32_smem_u32(lgbar);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo_ = v[j] + v[16 + j], hi_ = v[8 + j] + v[24 + j];
          const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
          const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16) + bias_p2;
          const unsigned s = (unsigned)(8 * q + j);
          if ((int)s < nvalid) tc_st_async_f32(cl_mapa(dst, s), r, cl_mapa(mb, s));
        }
        TC_PF_ADD(8);
        // ============================================================== softmax + draw + mu-law decode: CTA s owns stream s
        if (rank < nvalid) {
          wait_bar(lgbar, phlg);
          if (lane == 0) mbar_expect(lgbar, TC_Q * 4);
          const int b = b0 + rank;
          float lg[8];
#pragma unroll
          for (int qi = 0; qi < 8; ++qi) lg[qi] = logits_s[lane + 32 * qi];
          const int k = warp_softmax_draw(p, TC_Q, lg, b, t, lane, prob_s);
          if (k >= 0 && lane < TC_CS) {
            const float un = __ldg(p.enc_lut + k);
            const int slot_n = (int)((t + 1) % TC_PK);
            if (lane == 0) st_cg(p.u_hist + (long long)b * TC_PK + slot_n, un);
            tc_st_async_f32(cl_mapa(f32_smem_u32(hist + rank * TC_PK + slot_n), (unsigned)lane), un,
                            cl_mapa(f32_smem_u32(smpbar), (unsigned)lane));
          }
        }
        TC_PF_ADD(9);
      }
    } else {
      // ============================================================== loader warps: 5 (W_A, W_D), 6 (taps), 7 (W_B, W_C)
      // the step-boundary tap MMAs (layer 0, issued behind the previous step's postprocess2) have released W_B / X_T1
      if (lane == 0 && L > 1) {
        if (warp == 7) { wait_bar(auxbar, phaux); issue_w(TC_OFF_WB, layer_w(1, TC_W1), TC_W1, wbarB); }
        if (warp == 6) { wait_bar(auxbar, phaux); issue_tap(1, t, 1); }
      }
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        acc_wait();      // S1 of layer l complete: W_A, W_C (its t-2d part was issued earlier) and X_T2 are free
        if (lane == 0) {
          // development probe (VQWN_PROFILE): latency of one weight tile copy and one tap copy, issue -> landed
          const bool probe = (p.prof != nullptr) && blockIdx.x == 0 && p.cluster0 == 0 && t == p.t0 + 50 && (l == 7 || l == 19);
          const long long c0 = probe ? clock64() : 0;
          if (warp == 5) {
            if (!last) issue_w(TC_OFF_WA, layer_w(l + 1, 0), TC_W1, wbarA);
            else issue_w(TC_OFF_WA, p.post1 + (size_t)rank * (TC_WP1 / 2), TC_WP1, wbarA);
            nA += 1;
            if (probe) {
              const long long c1 = clock64();
              while (!tc_test_wait(wbarA, (nA - 1) & 1u)) {}
              p.prof[40 + (l == 19 ? 4 : 0)] = c1 - c0;
              p.prof[41 + (l == 19 ? 4 : 0)] = clock64() - c0;
            }
          } else if (warp == 7) {
            if (!last) issue_w(TC_OFF_WC, layer_w(l + 1, 2 * TC_W1), TC_W1, wbarC);
            else if (more) issue_w(TC_OFF_WC, layer_w(0, 2 * TC_W1), TC_W1, wbarC);
          } else if (!last) {
            issue_tap(l + 1, t, 2);
            nT2 += 1;
            if (probe) {
              const long long c1 = clock64();
              while (!tc_test_wait(tapbar2, (nT2 - 1) & 1u)) {}
              p.prof[42 + (l == 19 ? 4 : 0)] = c1 - c0;
              p.prof[43 + (l == 19 ? 4 : 0)] = clock64() - c0;
            }
          }
        }
        acc_wait();      // S2 of layer l complete: W_D, W_B (the next layer's t-d part was issued before it) and X_T1 are free
        if (lane == 0) {
          if (warp == 5) {
            if (!last) issue_w(TC_OFF_WD, layer_w(l + 1, 3 * TC_W1), TC_W2, wbarD);
            else issue_w(TC_OFF_WD, p.post2 + (size_t)rank * (TC_WP2 / 2), TC_WP2, wbarD);
          } else if (l + 2 < L) {
            if (warp == 7) issue_w(TC_OFF_WB, layer_w(l + 2, TC_W1), TC_W1, wbarB);
            else issue_tap(l + 2, t, 1);
          }
        }
      }
      cl_arrive();
      acc_wait();        // postprocess1 complete
      cl_wait();         // ... and every CTA's ring stores of this step are visible
      if (lane == 0 && more) {
        if (warp == 5) { issue_w(TC_OFF_WA, layer_w(0, 0), TC_W1, wbarA); nA += 1; }
        else if (warp == 7) issue_w(TC_OFF_WB, layer_w(0, TC_W1), TC_W1, wbarB);
        else { issue_tap(0, t + 1, 1); issue_tap(0, t + 1, 2); nT2 += 1; }
      }
      acc_wait();        // postprocess2 complete
      if (lane == 0 && more && warp == 5) issue_w(TC_OFF_WD, layer_w(0, 3 * TC_W1), TC_W2, wbarD);
    }
    // ================================================================ all threads: the new samples of all streams are in hist
    if (!ext) {
      wait_bar(smpbar, phsmp);
      __syncthreads();
      if (tid == 0) mbar_expect(smpbar, 4u * (unsigned)nvalid);
    }
    TC_PF_ADD(10);
  }
  cl_barrier();
  if (prof) for (int i = 0; i < 12; ++i) p.prof[(tid == 0 ? 0 : 16) + i] = pf[i];
#undef TC_PF_START
#undef TC_PF_ADD
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

// float32 [K rows][ldw] row-major weights -> hi/lo bf16 K-major plane tiles per cluster CTA.
//   dst[cta * cta_stride + (plane * rows + row) * 8 + e] = hi or lo part of src[(k0 + 8 plane + e) * ldw + col(cta, row)]
// row -> (column, part) by tile kind (the epilogue lane mappings of wavenet_tc_cluster):
//   kind 0 (gated conv tap, 64 rows, M = 64 accumulator): row = 16 w + 4 qq + i: column (qq & 1 ? G : 0) + 16 cta + 4 w + i, part qq >> 1
//   kind 1 (residual | skip, 96 rows): row < 32: column 16 cta + (row & 15), part row >> 4; 32..63: column R + 32 cta + row - 32
//                                       (hi); 64..95: column R + 32 cta + row - 64 (lo)
//   kind 2 (postprocess1, 64 rows): row = 32 w + 16 q + i: column 32 cta + 16 w + i, part q
//   kind 3 (postprocess2, 32 rows): row = 16 q + i: column 16 cta + i, part q
__global__ void pack_tc_tiles_kernel(const float* __restrict__ src, int ldw, int k0, int K, int rows, int kind,
                                     size_t cta_stride, __nv_bfloat16* __restrict__ dst) {
  const long long per_cta = (long long)K * rows;
  const long long total = (long long)TC_CS * per_cta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cta = (int)(i / per_cta);
    long long r = i % per_cta;
    const int e = (int)(r % 8); r /= 8;
    const int row = (int)(r % rows);
    const int plane = (int)(r / rows);
    int col, part;
    if (kind == 0) {
      const int w = row >> 4, qq = (row >> 2) & 3, ii = row & 3;
      col = ((qq & 1) ? TC_G : 0) + 16 * cta + 4 * w + ii; part = qq >> 1;
    } else if (kind == 1) {
      if (row < 32) { col = 16 * cta + (row & 15); part = row >> 4; }
      else if (row < 64) { col = TC_R + 32 * cta + (row - 32); part = 0; }
      else { col = TC_R + 32 * cta + (row - 64); part = 1; }
    } else if (kind == 2) {
      const int w = row >> 5, q = (row >> 4) & 1, ii = row & 15;
      col = 32 * cta + 16 * w + ii; part = q;
    } else {
      col = 16 * cta + (row & 15); part = row >> 4;
    }
    const float x = src[(size_t)(k0 + plane * 8 + e) * ldw + col];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    dst[(size_t)cta * cta_stride + (size_t)(plane * rows + row) * 8 + e] = part ? __float2bfloat16_rn(x - __bfloat162float(hi)) : hi;
  }
}

// skip start folded into the preprocess FIR: skf_k[j][c] = sum_m pre_k[j][m] skip_k[m][c] (float64 accumulation),
// skf_b[c] = sum_m pre_b[m] skip_k[m][c] + skip_b[c] + sum over layers of their skip biases (b2[l][R + c])
__global__ void fold_skip_start_kernel(const float* __restrict__ pre_k, const float* __restrict__ pre_b,
                                       const float* __restrict__ skip_k, const float* __restrict__ skip_b,
                                       const float* const* __restrict__ layer_b2, int L,
                                       float* __restrict__ skf_k, float* __restrict__ skf_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= TC_S) return;
  for (int j = 0; j <= TC_PK; ++j) {
    double a = 0.0;
    const float* row = (j < TC_PK) ? pre_k + (size_t)j * TC_R : pre_b;
    for (int m = 0; m < TC_R; ++m) a += (double)row[m] * (double)skip_k[(size_t)m * TC_S + c];
    if (j < TC_PK) skf_k[(size_t)j * TC_S + c] = (float)a;
    else {
      a += (double)skip_b[c];
      for (int l = 0; l < L; ++l) a += (double)layer_b2[l][TC_R + c];
      skf_b[c] = (float)a;
    }
  }
}

}  // namespace vqwn
