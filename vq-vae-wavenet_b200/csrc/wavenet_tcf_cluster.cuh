// WaveNet fast generation on the 5th-generation tensor cores at float32-grade accuracy, ONE hand-off per layer
// (VQWN_PREC_TC).  Same reference semantics as the float32 kernels (wavenet.py:103-172, wavenet_ops.py:163-267,
// utils.py:13-46, mu_law_ops.py:5-31).
//
// Arithmetic: split bfloat16 as in wavenet_tc_cluster.cuh (x = hi + lo, weight tiles stack hi and lo rows along M, the
// activation operand stacks the hi and lo copies of the streams along N, fp32 accumulation in TMEM; biases, residual
// chain, skip sum, gate, softmax and the draw are float32).
//
// What is new against wavenet_tc_cluster.cuh - the dependency chain of a time step is 30 hand-offs, not 60:
//   * the reference's layer is gate_l = f(W2_l x_l + taps), x_{l+1} = x_l + Wres_l gate_l (wavenet_ops.py:240-267): two
//     contractions that each need an all-gather of their input across the cluster.  Here the current-tap term is
//     expanded once, W2_l x_l = W2_l x_{l-1} + (W2_l Wres_{l-1}) gate_{l-1} + W2_l bres_{l-1}, with P_l = W2_l Wres_{l-1}
//     premultiplied on the host side (float64) and the bias term folded into the gated bias.  gate_{l-1} and x_{l-1}
//     arrive together (both come out of stage l-1), so stage l needs ONE all-gather: [gate_{l-1} | x_{l-1}];
//   * an MMA of this size costs its issuing thread ~64 cycles whatever M <= 128 and N <= 128 are (tools/r2_probe.cu; most
//     of it is the descriptor transfer loop ptxas emits per instruction, the pipe itself is busy ~26 cycles: DESIGN.md 8),
//     so work is packed per instruction, not per FLOP: the activation operand stacks TWO inputs along N (64 columns), the
//     weight tile puts the rows that multiply the first input in lanes 0-15 and those for the second in lanes 16-31 of
//     every TMEM lane quarter: [P_l | W2_l] x [gate_{l-1} | x_{l-1}] is 16 instructions, both older taps
//     [W1_l | W0_l] x [x_l(t-d) | x_l(t-2d)] another 16 (issued a stage ahead into the same accumulator), residual + skip
//     rows 16 more; the off-diagonal blocks are never read;
//   * weights stream through ONE FIFO of 16 KB chunks (4 instructions each) in issue order: a chunk's slot is released by
//     tcgen05.commit and refilled by the loader lanes, so a layer's 176 KB of tiles never have to be resident at once;
//   * hand-offs go through L2 as one multicast bulk copy per slice (see wavenet_tc_cluster.cuh); the dilation queues are
//     HBM rings of PAIR blocks [2d + 1][cluster][16 senders][x(t-d) | x(t-2d)][1 KB]: a step's layer input is stored
//     twice (1 KB each) so that both taps of a later step are ONE contiguous 32 KB bulk copy into the N-stacked operand.
// Operand layout everywhere: per 16 channels (one MMA K step) a block [row groups of 8][2 x 8 channels][8 rows x 16 B];
// descriptors: no swizzle, K-direction stride 128 B, row-group stride 256 B.
// Geometry fixed to the reference's default (R = G = 256, S = 512, Q = 256, C = 128, 32-tap preprocess, kernel_size 3).
#pragma once
#include <cuda_bf16.h>
#include "wavenet_tc_cluster.cuh"

namespace vqwn {

constexpr int TF_CS = 16;
constexpr int TF_THREADS = 384;           // warps 0-3 epilogue | 4-7 MMA issue | 8-9 weight loaders | 10 tap loader | 11 layer-input publisher
constexpr int TF_NISSUE = 4;
constexpr int TF_NS = 16;                  // streams per cluster (at most)
constexpr int TF_R = 256, TF_G = 256, TF_S = 512, TF_Q = 256, TF_C = 128, TF_PK = 32;
constexpr int TF_MAXL = 64;
// order of a stage's two chains on the tensor pipe: residual + skip of layer l-1 first, then the gate of layer l (default,
// -4 % per step: x_l reaches the cluster ~1.4k cycles into the stage and the gate path alone is critical); -DTF_ORDER_GA
// restores gate first (the round-2 default until the end of the round)
#if !defined(TF_ORDER_GA) && !defined(TF_ORDER_RA)
#define TF_ORDER_RA 1
#endif
#ifdef TF_ISSUE_CALL
#define TF_ISSUE_INLINE __noinline__
#else
#define TF_ISSUE_INLINE __forceinline__
#endif
constexpr int TF_L2_KEEP_LAYERS = 255;     // default: every stage's weight tiles ask to stay in the L2 (see the loader lanes)
constexpr int TF_TRACE_N = 320;            // trace events per warp (profile build)
constexpr int TF_BLK = 1024;               // one K step (16 channels) of a 32-row activation operand
constexpr int TF_PAIR = TF_CS * 2 * TF_BLK;    // N-stacked operand, K = 256: [16 senders][first | second][1 KB] = 32 KB
// weight FIFO: a chunk is one bulk copy = TF_CPW x 4 instructions (TF_CPW issuing warps share it); the copy unit of an SM
// takes a few hundred cycles per bulk copy whatever its size up to 48 KB (tools/r2_probe.cu), so fewer, larger chunks
#ifdef TF_CHUNK32
constexpr int TF_CPW = 2;
constexpr int TF_NSLOT = 3;
#else
constexpr int TF_CPW = 1;
constexpr int TF_NSLOT = 7;
#endif
constexpr int TF_SLOT = 16384 * TF_CPW;    // slot: TF_CPW x 4 instructions x 128 rows x 32 B
constexpr int TF_CHUNK_A = 4 * 128 * 32, TF_CHUNK_R = 4 * 96 * 32, TF_CHUNK_P1 = 4 * 64 * 32, TF_CHUNK_P2 = 4 * 32 * 32;
// per-CTA weight stream of one time step (bytes): T_0 | stage 0: A_0, T_1 | stage l: A_l, R_{l-1}, T_{l+1} | ... | tail
constexpr int TF_TILE_A = 4 * TF_CHUNK_A, TF_TILE_R = 4 * TF_CHUNK_R, TF_TILE_P1 = 8 * TF_CHUNK_P1, TF_TILE_P2 = 8 * TF_CHUNK_P2;
__host__ __device__ constexpr size_t tf_stream_bytes(int L) {
  return (size_t)TF_TILE_A + (size_t)L * TF_TILE_A + (size_t)L * TF_TILE_R + (size_t)(L - 1) * TF_TILE_A + TF_TILE_P1 + TF_TILE_P2;
}
// TF_FIFO2: the two per-CTA streams of a time step: gate / tap tiles, and everything else
__host__ __device__ constexpr size_t tf_big_bytes(int L) { return (size_t)2 * L * TF_TILE_A; }
__host__ __device__ constexpr size_t tf_small_bytes(int L) { return (size_t)L * TF_TILE_R + TF_TILE_P1 + TF_TILE_P2; }
// shared memory map (bytes)
constexpr int TF_OFF_W = 0;                                  // weight FIFO
constexpr int TF_OFF_B1 = TF_OFF_W + TF_NSLOT * TF_SLOT;     // [2][gate | layer input] operands, by stage parity
constexpr int TF_OFF_B2 = TF_OFF_B1 + 2 * TF_PAIR;           // [tap t-d | tap t-2d] operand; postprocess1 input
constexpr int TF_OFF_STG = TF_OFF_B2 + TF_PAIR;              // publish staging: 2 blocks + 2 layer-input blocks (by layer parity)
constexpr int TF_OFF_HIST = TF_OFF_STG + 4 * TF_BLK;         // [16][32] fp32 network-input history ring (remote-written)
constexpr int TF_OFF_LOG = TF_OFF_HIST + TF_NS * TF_PK * 4;  // [256] fp32 logits of this CTA's stream (remote-written)
constexpr int TF_OFF_US = TF_OFF_LOG + TF_Q * 4;             // [16][32] history in tap order     } 8 KB of CTA-local scratch,
constexpr int TF_OFF_CUR0 = TF_OFF_US + TF_NS * TF_PK * 4;   // [16 ch][16] fp32 FIR output       } aliased by the condition
constexpr int TF_OFF_SKF = TF_OFF_CUR0 + 16 * TF_NS * 4;     // [32 ch][16] skip FIR part         } rows [16][128] fp32 at a
constexpr int TF_OFF_SKX = TF_OFF_SKF + 32 * TF_NS * 4;      // [32 ch][16] skip lo sums          } frame change
constexpr int TF_OFF_PROB = TF_OFF_SKX + 32 * TF_NS * 4;     // [256] fp32 draw scratch           }
constexpr int TF_OFF_LAYERS = TF_OFF_PROB + TF_Q * 4;
constexpr int TF_OFF_BARS = TF_OFF_LAYERS + TF_MAXL * 48;
constexpr int TF_NBARS = 40;
constexpr int TF_WFULL1 = 32;               // second set of weight-arrival barriers (odd tenancies of a slot)
constexpr int TF_OFF_MISC = TF_OFF_BARS + TF_NBARS * 8;
constexpr int TF_OFF_PROF = TF_OFF_MISC + 16;            // [16] cycle counters of the issuing thread, by chain kind
constexpr int TF_SMEM = TF_OFF_PROF + 16 * 8;
static_assert(TF_OFF_PROB + TF_Q * 4 - TF_OFF_US == TF_NS * TF_C * 4, "condition rows alias exactly the local scratch");
static_assert(TF_SMEM <= 232448, "shared memory budget");
// per-cluster global staging of the hand-offs that do not live in a ring: gate output (double-buffered by layer parity),
// postprocess1 input, postprocess2 input
constexpr int TF_GST_XG = 0, TF_GST_XS = 2 * TF_CS * TF_BLK, TF_GST_XN = TF_GST_XS + 2 * TF_CS * TF_BLK,
              TF_GSTAGE = TF_GST_XN + 2 * TF_CS * TF_BLK;

struct TfLayerDev {
  const float* wlc;           // gated/local_condition/kernel [C][2G] float32 (row stride 2G)
  const float* b1;            // gated/bias [2G] + W2_l . residual/bias of layer l-1
  const float* bres;          // residual/bias [R] (first R entries of the layer's [R + S] bias vector)
  uint8_t* ring;              // [2d + 1][nclusters][TF_PAIR]
  int d;
  int pad_[3];
};
static_assert(sizeof(TfLayerDev) == 48, "layer record size");

struct TfParams {
  int L, B, nclusters, cluster0, spc;      // spc: streams per cluster; cluster c owns streams [c*spc, c*spc + spc)
  const float *pre_k, *pre_b;              // preprocess/kernel [32][256], bias [256]
  const float *skf_k, *skf_b;              // skip start folded into the FIR: [32][512] = pre_k . skip/kernel; [512] all skip biases
  const uint8_t* wstream;                  // [16 CTAs][tf_stream_bytes(L)]
  const float *post1_lc, *post1_b, *post2_b;   // postprocess1/local_condition/kernel [C][S], biases
  const TfLayerDev* layers;
  float* ctab;                             // [launch cluster][16][L+1][512] condition table
  uint8_t* gstage;                         // [launch cluster][TF_GSTAGE] hand-off staging in L2
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
  int flags;                               // bit 0: reproducible accumulation order (issuing warps take turns); bits 8-15: stages whose weight tiles ask to stay in the L2
  float* audio_out;
  int* idx_out;
  float* logits_out;
  float* probs_out;
  long long* prof;
  int* err;
};

// element (row n, channel k of the 16-channel block) of an operand block: hi copy in row n, lo copy in row 16 + n
__device__ __forceinline__ void tf_st_split(uint8_t* blk, int n, int k, float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  uint8_t* q = blk + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(q) = hi;
  *reinterpret_cast<__nv_bfloat16*>(q + 2 * 256) = lo;
}
__device__ __forceinline__ uint64_t tf_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;           // K direction: the two 8-channel halves of a K step
  d |= (uint64_t)(256 >> 4) << 32;           // M / N direction: 8-row groups
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void tf_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// zero 32 accumulator columns of the warp's TMEM lane quarter
__device__ __forceinline__ void tf_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tf_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f32_smem_u32(bar)) : "memory");
}

// one copy of the bounded wait loop for the whole kernel: the three warp roles run disjoint code and share the
// instruction cache, every inlined copy costs all of them
__device__ __noinline__ void tf_wait_slow(unsigned bar_addr, unsigned parity, int* err) {
  unsigned long long t_start = 0;
  bool noted = false;
#pragma unroll 1
  for (unsigned spin = 0;; ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {
      // bounded by wall time (the spin rate depends on how many lanes wait): a waiter stuck for 0.25 s records what it
      // waits for in host-mapped memory (VQWN_DEBUG=1 prints the table after a failed launch), after 2 s the launch is
      // declared broken
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_start == 0) t_start = now;
      if (!noted && now - t_start > 250000000ull && (threadIdx.x & 31) == 0) {
        volatile int* d = reinterpret_cast<volatile int*>(err) + 64 + (blockIdx.x * 12 + (threadIdx.x >> 5)) * 2;
        d[0] = (int)bar_addr; d[1] = (int)parity + 100;
        __threadfence_system();
        noted = true;
      }
      if (now - t_start > 2000000000ull) break;
    }
  }
  // A wait that never completes is a broken launch: record which one (barrier offset in shared memory, parity, thread,
  // block - read back by the host for the error message) and stop the grid instead of running on with missing operands
  if (atomicCAS(err, 0, 2) == 0) { err[1] = (int)bar_addr; err[2] = (int)parity; err[3] = (int)threadIdx.x; err[4] = (int)blockIdx.x; }
  __threadfence_system();
  __trap();
}
// The wait every role uses: a short polling loop INLINE, the bounded loop above only when that runs out.  A call of the
// non-inlined function costs ~290 cycles even when the phase has long completed (measured in the loader lanes with
// -DTF_WAIT_COST: call 352 - 64 cycles of trace overhead, the bare instruction 115 - 64) - the caller saves and restores
// its live registers around the call - and a stage's critical path goes through four or five waits.
__device__ __forceinline__ void tf_wait(unsigned bar_addr, unsigned parity, int* err) {
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    if (ok) return;
  }
  tf_wait_slow(bar_addr, parity, err);
}

// Weight FIFO barriers and parity waits.  An mbarrier parity wait is valid only while the waiter is at most one phase
// away from the barrier (a waiter that asks for phase n while phase n - 1 is still open sees "the other parity" and
// passes at once), and the FIFO's producers and consumers are only loosely coupled: two loader lanes, four issuing warps.
//   * release barriers wfree[slot]: every slot belongs to ONE loader lane (even slots: warp 8, odd slots: warp 9), which
//     therefore sees every phase of its barriers, in order;
//   * arrival barriers: TWO per slot, used alternately (wfull[tenancy & 1][slot]).  Consecutive phases of one barrier are
//     chunks ci and ci + 14; chunk ci - 14 lies three or four chains back, and every chain that far back has completed
//     in all four issuing warps before any of them can reach chunk ci (gate chain l+1 needs the gather of gate l, i.e.
//     all four accA commits of chain l; the residual chain of l needs x_l, i.e. accB of l-1; the tap chain of l+2 needs
//     its pair block, loaded after all four tapfree commits of l+1; the tail is ordered the same way by xsbar / b1bar).
//     With one barrier per slot the distance would be 7 chunks - less than two chains - and nothing orders those.
#ifdef TF_FIFO2
// TWO weight FIFOs in the same 112 KB: class 0 = 4 slots x 16 KB for the gate and tap tiles, class 1 = 4 slots x 12 KB for
// the residual + skip tiles (and the postprocess tiles, 8 / 4 KB chunks).  With one FIFO of seven 16 KB slots the eighth
// chunk of a stage - the last quarter of the residual + skip tile - is requested only when the gate chain has released a
// slot and lands ~2k cycles into the stage; with the 12 KB class beside it a stage's gate tile AND residual + skip tile
// (64 + 48 KB) are resident when its gather completes.  A chunk id is (class << 31) | index within the class; each class
// has one loader lane (warp 8: class 0, warp 9: class 1), so every release barrier is still watched by one lane, and
// consecutive phases of an arrival barrier are 8 chunks = two chains of the class apart (see above).
constexpr int TF_SLOT_R = 12288;
constexpr int TF_OFF_WR = TF_OFF_W + 4 * 16384;
static_assert(TF_CPW == 1 && TF_OFF_WR + 4 * TF_SLOT_R == TF_OFF_W + TF_NSLOT * TF_SLOT, "two FIFOs fill the weight area exactly");
__device__ __forceinline__ unsigned tf_slot_off(unsigned cid) {
  const unsigned slot = cid & 3u;
  return (cid >> 31) ? (unsigned)TF_OFF_WR + slot * TF_SLOT_R : (unsigned)TF_OFF_W + slot * 16384u;
}
__device__ __forceinline__ unsigned tf_wfull_bar(unsigned sm_u32, unsigned cid) {
  const unsigned idx = cid & 0x7fffffffu;
  return sm_u32 + TF_OFF_BARS + ((((idx >> 2) & 1u) ? TF_WFULL1 : 0u) + (cid >> 31) * 4u + (idx & 3u)) * 8u;
}
__device__ __forceinline__ unsigned tf_wfull_parity(unsigned cid) { return ((cid & 0x7fffffffu) >> 3) & 1u; }
__device__ __forceinline__ unsigned tf_wfree_idx(unsigned cid) { return 8u + (cid >> 31) * 4u + (cid & 3u); }
#else
__device__ __forceinline__ unsigned tf_slot_off(unsigned cid) { return (unsigned)TF_OFF_W + (cid % TF_NSLOT) * TF_SLOT; }
__device__ __forceinline__ unsigned tf_wfull_bar(unsigned sm_u32, unsigned ci) {
  const unsigned u = ci / TF_NSLOT, slot = ci - u * TF_NSLOT;
  return sm_u32 + TF_OFF_BARS + ((u & 1u) ? (TF_WFULL1 + slot) : slot) * 8u;
}
__device__ __forceinline__ unsigned tf_wfull_parity(unsigned ci) { return ((ci / TF_NSLOT) >> 1) & 1u; }
__device__ __forceinline__ unsigned tf_wfree_idx(unsigned cid) { return 8u + cid % TF_NSLOT; }
#endif

// One weight chunk (4 K steps): wait for it, issue its 4 MMAs D[128 x N] += A . B^T, release its FIFO slot.  Inline
// (TF_ISSUE_INLINE): as a call it kept the three roles' code small, but every call cost more than the instruction
// fetches it saved (90.6 -> 86.1 us per step).  Every MMA costs ~17 instructions of descriptor arithmetic, register ->
// uniform-register moves and the per-thread issue loop; -DTF_ELECT_AT_SITE removes most of them.
// d_tmem: accumulator address; b_addr: B operand of the chunk's first K step.
__device__ TF_ISSUE_INLINE void tf_issue_chunk(unsigned ci, unsigned cid, uint32_t a_step, uint32_t d_tmem, uint32_t idesc, uint32_t b_addr,
                                            uint32_t b_step, uint32_t sm_u32, uint32_t elected, int* err, bool have_weights, unsigned* turn_ptr) {
  // ci: position in the issue sequence of the step chain (orders the reproducible mode); cid: the quarter-tile's place in
  // the weight FIFO(s): the chunk that holds it is cid / TF_CPW, at sub-position cid % TF_CPW
#ifdef TF_FIFO2
  const unsigned chunk = cid, sub = 0u;
#else
  const unsigned chunk = cid / TF_CPW, sub = cid - chunk * TF_CPW;
#endif
  if (!have_weights) tf_wait(tf_wfull_bar(sm_u32, chunk), tf_wfull_parity(chunk), err);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef TF_ELECT_AT_SITE
  // a fresh elect.sync at the site: the compiler may know that exactly one lane runs the block
  {
    uint32_t e_;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(e_));
    elected = e_;
  }
#endif
  if (elected) {
    uint64_t da = tf_desc(sm_u32 + tf_slot_off(chunk) + sub * 4u * a_step);
    uint64_t db = tf_desc(b_addr);
    const uint64_t sa = (uint64_t)(a_step >> 4), sb = (uint64_t)(b_step >> 4);
    // The chunks of a chain accumulate into one TMEM tile from four different threads, and float32 accumulation depends
    // on the order in which the tensor pipe receives them (last-bit differences from run to run, ~1e-6 relative).
    // Reproducible mode (vqwn_set_reproducible): the issuing warps take turns in FIFO order, a counter in shared memory,
    // so the pipe sees the MMAs in exactly the order one thread would have issued them; the waits and the descriptor
    // arithmetic of the four warps still overlap, the issue itself (~80 cycles per MMA) no longer does.
    volatile unsigned* const turn = reinterpret_cast<volatile unsigned*>(turn_ptr);
    if (turn) while (*turn != ci) { }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc) : "memory");
      da += sa; db += sb;
    }
    if (turn) *turn = ci + 1u;
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm_u32 + TF_OFF_BARS + tf_wfree_idx(chunk) * 8u) : "memory");
  }
  __syncwarp();
}
// the barrier fires when every MMA this thread has issued so far has completed
__device__ TF_ISSUE_INLINE void tf_commit(uint32_t bar_addr, uint32_t elected) {
#ifdef TF_ELECT_AT_SITE
  {
    uint32_t e_;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(e_));
    elected = e_;
  }
#endif
  if (elected)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  __syncwarp();
}

// PROF: in-kernel cycle counters (VQWN_PROFILE=1).  A separate instantiation: the issuing warps are bound by the length of
// their own instruction stream, every time stamp costs them
template <bool PROF>
__global__ void __launch_bounds__(TF_THREADS, 1) wavenet_tcf_cluster(const TfParams p_in) {
  extern __shared__ __align__(1024) uint8_t sm[];
  // the warp index goes through a shuffle so that the compiler knows it (and the FIFO positions, slot addresses and MMA
  // descriptors derived from it) is uniform across the warp: uniform registers instead of per-MMA register -> uniform-register
  // transfer loops in the issuing warps
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  TfLayerDev* const layers_s = reinterpret_cast<TfLayerDev*>(sm + TF_OFF_LAYERS);
  for (int i = tid; i < p_in.L * (int)(sizeof(TfLayerDev) / 4); i += TF_THREADS)
    reinterpret_cast<uint32_t*>(layers_s)[i] = reinterpret_cast<const uint32_t*>(p_in.layers)[i];
  TfParams p = p_in;
  p.layers = layers_s;
  unsigned rank_u;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int lcluster = (int)blockIdx.x / TF_CS;
  const int cluster = p.cluster0 + lcluster;
  const int b0 = cluster * p.spc;
  const int nvalid = max(0, min(p.spc, p.B - b0));
  const int L = p.L;

  uint8_t* const b1buf = sm + TF_OFF_B1;
  uint8_t* const b2buf = sm + TF_OFF_B2;
  uint8_t* const stg = sm + TF_OFF_STG;
  float* const hist = reinterpret_cast<float*>(sm + TF_OFF_HIST);
  float* const logits_s = reinterpret_cast<float*>(sm + TF_OFF_LOG);
  float* const u_s = reinterpret_cast<float*>(sm + TF_OFF_US);
  float* const cur0 = reinterpret_cast<float*>(sm + TF_OFF_CUR0);
  float* const skf = reinterpret_cast<float*>(sm + TF_OFF_SKF);
  float* const skx = reinterpret_cast<float*>(sm + TF_OFF_SKX);
  float* const prob_s = reinterpret_cast<float*>(sm + TF_OFF_PROB);
  float* const ct_rows = reinterpret_cast<float*>(sm + TF_OFF_US);     // [16][128] condition rows (alias)
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sm + TF_OFF_BARS);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(sm + TF_OFF_MISC);
  // reproducible mode (flags bit 0): next FIFO chunk whose MMAs may be issued
  unsigned* const turn_s = (p.flags & 1) ? reinterpret_cast<unsigned*>(sm + TF_OFF_MISC + 8) : nullptr;
  // bars 0..6 and 32..38: weight chunk landed (even / odd tenancy of the slot, see tf_wfull_bar)
  unsigned long long* const wfree = bars + 8;          // [7] the MMAs that read the slot completed
  unsigned long long* const b1bar = bars + 16;         // [2] gathered operand of a stage landed (by stage parity)
  unsigned long long* const tapbar = bars + 18;        // pair block of the next layer's taps landed in B2
  unsigned long long* const tapfree = bars + 19;       // the MMAs that read B2 completed
  unsigned long long* const xsbar = bars + 20;         // postprocess1 input slices landed in B2
  unsigned long long* const accA = bars + 21;          // gate pre-activations / postprocess accumulators complete
  unsigned long long* const accB = bars + 22;          // residual + skip accumulator complete
  unsigned long long* const e1done = bars + 23;        // the gate epilogue has read its accumulator (4 warps)
  unsigned long long* const e2done = bars + 24;        // the residual / skip epilogue has read its accumulator (3 warps)
  unsigned long long* const lgbar = bars + 25;
  unsigned long long* const smpbar = bars + 26;

  for (int i = tid; i < (TF_OFF_LAYERS - TF_OFF_W) / 16; i += TF_THREADS) reinterpret_cast<uint4*>(sm + TF_OFF_W)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const bool ext = (p.mode == GEN_STEP || p.mode == GEN_TEACHER);
  if (tid == 0) {
    *reinterpret_cast<unsigned*>(sm + TF_OFF_MISC + 8) = 0u;
    for (int i = 0; i < TF_NBARS; ++i) {
      unsigned cnt = (bars + i == e1done) ? 4u : ((bars + i == e2done) ? 3u : 1u);
      if (bars + i == accA || bars + i == accB || bars + i == tapfree) cnt = TF_NISSUE;      // one commit per issuing warp
      if (i >= 8 && i < 8 + TF_NSLOT) cnt = TF_CPW;                                          // wfree: one commit per issuing warp of the chunk
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f32_smem_u32(&bars[i])), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // first use of the receive barriers in a step: stage 0 (layer input only, 16 KB) on parity 0, stage 1 on parity 1
    mbar_expect(&b1bar[0], TF_CS * TF_BLK);
    mbar_expect(&b1bar[1], 2 * TF_CS * TF_BLK);
    mbar_expect(xsbar, 2 * TF_CS * TF_BLK);
    if (rank < nvalid) mbar_expect(lgbar, TF_Q * 4);
    if (!ext) mbar_expect(smpbar, 4u * (unsigned)nvalid);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(f32_smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < nvalid * TF_PK; i += TF_THREADS) hist[i] = ld_cg(p.u_hist + (long long)b0 * TF_PK + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // every MMA chain accumulates (see consume): the accumulators start at zero and every epilogue re-zeroes what it read
  if (warp < 4) {
#pragma unroll
    for (int c = 0; c < 224; c += 32) tf_zero32(tmem + ((uint32_t)(warp * 32) << 16) + c);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cl_barrier();      // every CTA's barriers are initialised and armed before any remote copy can arrive

  const float mu = (float)(TF_Q - 1);
  // in-kernel cycle counters (VQWN_PROFILE=1): CTA 0 of the first cluster, thread 0 (epilogue view) and thread 128 (MMA view)
  const bool prof = PROF && (p.prof != nullptr) && blockIdx.x == 0 && p.cluster0 == 0 && (tid == 0 || tid == 128);
  long long pf[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) pf[i] = 0;
  long long pf_t = 0;
  // event trace of ONE time step (VQWN_PROFILE=1, step p.t0 + 300 of CTA 0 of cluster 0): lane 0 of every warp appends
  // (clock << 16 | event << 8 | layer) to its row of p.prof[256 + warp * TF_TRACE_N ...]; tools/tcf_trace.py prints it
  int tr_n = 0;
  bool tracing = false;
#define TF_TR(ev, l_) do { if (PROF && tracing && lane == 0 && tr_n < TF_TRACE_N) { \
    p.prof[256 + warp * TF_TRACE_N + tr_n] = (clock64() << 16) | ((long long)(ev) << 8) | (long long)(l_); ++tr_n; } } while (0)
#define TF_PF_START() do { if (prof) pf_t = clock64(); } while (0)
#define TF_PF_ADD(i) do { if (prof) { const long long n_ = clock64(); pf[(i)] += n_ - pf_t; pf_t = n_; } } while (0)
#ifdef TF_DEBUG_MARKS
  // development: every warp of every CTA leaves its last program point in host-mapped memory (posted stores, no fence)
#define TF_MARK(code) do { if (p.prof && lane == 0) { \
    reinterpret_cast<volatile long long*>(p.prof)[1024 + blockIdx.x * 12 + warp] = (long long)(t_mark * 1000 + (code)); } } while (0)
#else
#define TF_MARK(code) do { } while (0)
#endif
  long long t_mark = -1;
  uint32_t elected = 0;
  const bool issuer = (warp >= 4 && warp < 4 + TF_NISSUE);
  const unsigned iss = (unsigned)(warp - 4);
  if (issuer) asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(elected));
  // instruction descriptors: fp32 accumulate, bf16 x bf16, M = 128, N = 64 (stacked operands) or 32
  const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t ACC0 = 0, ACC1 = 64, ACCR = 128, ACCP1 = 160, ACCP2 = 192;     // TMEM columns
  const uint32_t sm_u32 = f32_smem_u32(sm);

  // ------------------------------------------------------------------ helpers
  auto wait_bar = [&](unsigned long long* bar, unsigned& ph) {
    tf_wait(f32_smem_u32(bar), ph, p.err);
    ph ^= 1u;
  };
  const long long ring_slot_bytes = (long long)p.nclusters * TF_PAIR;
  // pair block u of layer l: [16 senders][x(u - d) | x(u - 2d)][1 KB]
  // (time indices stay below 2^31 - checked by the host - so the slot is a 32-bit modulo, not a 64-bit division routine
  // inlined at every use)
  auto pair_block = [&](int l, long long u) {
    const TfLayerDev& ly = p.layers[l];
    return ly.ring + (long long)((unsigned)u % (2u * (unsigned)ly.d + 1u)) * ring_slot_bytes + (long long)cluster * TF_PAIR;
  };
  // warp-level: publish `nblk` staged blocks through L2: copy them to `gdst` (and `gdst2` when given), then ONE multicast
  // bulk copy delivers them to offset dst_off of all 16 CTAs and counts the bytes on every receiver's `sbar`
  auto publish = [&](const uint8_t* src, uint8_t* gdst, uint8_t* gdst2, int nblk, int dst_off, unsigned long long* sbar) {
    for (int c = lane; c < nblk * (TF_BLK / 16); c += 32) {
      const float4 x = *reinterpret_cast<const float4*>(src + c * 16);
      *reinterpret_cast<float4*>(gdst + c * 16) = x;
      if (gdst2) *reinterpret_cast<float4*>(gdst2 + c * 16) = x;
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");     // generic-proxy stores before the bulk copies' reads
    __syncwarp();
    if (lane == 0 && sbar)
      tc_bulk_multicast(sm_u32 + (unsigned)dst_off, gdst, (unsigned)(nblk * TF_BLK), f32_smem_u32(sbar), (unsigned short)0xFFFF);
  };
  uint8_t* const gst = p.gstage + (size_t)lcluster * TF_GSTAGE;
  const uint32_t my_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float* const ctab = p.ctab + ((size_t)lcluster * TF_CS + rank) * (size_t)(L + 1) * 512;
  const uint8_t* const wsrc = p.wstream + (size_t)rank * tf_stream_bytes(L);

  // ------------------------------------------------------------------ role state
  // MMA thread: FIFO position and barrier phases
  // (the FIFO position is NOT carried in a variable across the role branches: anything assigned under `if (warp == 4)`
  // counts as thread-dependent, and slot addresses derived from it would need a register -> uniform-register transfer per
  // MMA operand; it is recomputed per step from the step index: 4 prologue chunks + 12 L + 16 chunks per step)
  unsigned ph_b10 = 0u, ph_b11 = 0u, ph_tap = 0u, ph_xs = 0u, ph_e1 = 0u, ph_e2 = 0u;
  unsigned n_r = 0;                      // residual/skip chains issued so far
  // this warp's chunk(s) of a chain of `nch` chunks (4 or 8) of `rows`-row tiles that starts at FIFO position ci0: chunk
  // iss (and 4 + iss).  The chunks of a chain go to the four issuing warps: MMAs of different threads are not ordered, so
  // every chain ACCUMULATES - the epilogue that reads an accumulator leaves it zeroed.
  // FIFO positions: one sequence (cid = ci), or per class with TF_FIFO2 (cb: 16 KB chunks of the gate / tap tiles, cr: the
  // others); a chain takes its `nch` places with take(class, nch)
  unsigned cb = 0u, cr = 0u;
  auto take = [&](unsigned ci_now, int cls, unsigned nch) -> unsigned {
#ifdef TF_FIFO2
    unsigned r;
    if (cls) { r = 0x80000000u | cr; cr += nch; } else { r = cb; cb += nch; }
    return r;
#else
    return ci_now;
#endif
  };
  auto consume = [&](unsigned ci0, unsigned cid0, int nch, int rows, uint32_t d_col, bool n64, uint32_t b_addr, uint32_t b_step, bool have_weights = false) {
    const uint32_t id = n64 ? idesc64 : idesc32;
    tf_issue_chunk(ci0 + iss, cid0 + iss, (uint32_t)rows * 32u, tmem + d_col, id, b_addr + 4u * iss * b_step, b_step, sm_u32, elected, p.err, have_weights, turn_s);
    if (nch == 8)
      tf_issue_chunk(ci0 + 4u + iss, cid0 + 4u + iss, (uint32_t)rows * 32u, tmem + d_col, id, b_addr + 4u * (4u + iss) * b_step, b_step, sm_u32, elected, p.err, false, turn_s);
  };
  // the weight chunk of the chain that is waiting for a gather: checked BEFORE the gather wait, off the critical path
  auto weights_ready = [&](unsigned cid0) {
#ifdef TF_FIFO2
    const unsigned c_ = cid0 + iss;
#else
    const unsigned c_ = (cid0 + iss) / TF_CPW;
#endif
    tf_wait(tf_wfull_bar(sm_u32, c_), tf_wfull_parity(c_), p.err);
  };
  auto wait_b1 = [&](int par) {
    if (par) wait_bar(&b1bar[1], ph_b11); else wait_bar(&b1bar[0], ph_b10);
  };
  auto commit_to = [&](unsigned long long* bar) { tf_commit(f32_smem_u32(bar), elected); };
  auto operand_fence = [&]() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // loader lanes (warp 8 lane 0: even FIFO slots, warp 9 lane 0: odd slots): chunks issued so far, stream position
  unsigned li = 0;
  size_t woff = 0;
  // (three loader lanes - slot s owned by lane s mod 3 - measured no faster, twice: 99.3 us against 98.6, and 93.6 against 90.6; the
  // feed is bound by bytes in flight, not by the issue rate of the loader lanes)
  constexpr unsigned NLOADERS = 2;
  const bool wloader = (lane == 0) && (warp == 8 || warp == 9);
  const unsigned wmine = (warp == 9) ? 1u : 0u;
  const bool tloader = (lane == 0) && (warp == 10);
  // L2 policy of the weight copies: the tiles of the first `keep_layers` stages are kept (evict_last), the others stream
  // (evict_first).  All clusters read the same 88 MB once per time step - a cyclic pattern larger than the part of the L2
  // that keeps it, i.e. no tile survives until the next step if all of them ask to stay; a subset that fits does.
  unsigned long long wpol_keep = 0, wpol_stream = 0, wpol = 0;
  const int keep_layers = (p.flags >> 8) & 0xff;
  if (wloader) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(wpol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(wpol_stream));
    wpol = keep_layers > 0 ? wpol_keep : wpol_stream;
  }
#ifdef TF_FIFO2
  // a lane loads the chunks of ITS class only (cls == wmine); li and woff are the lane's position in its class and stream
  const uint8_t* const wsrc2 = wmine ? p.wstream + (size_t)TF_CS * tf_big_bytes(L) + (size_t)rank * tf_small_bytes(L)
                                     : p.wstream + (size_t)rank * tf_big_bytes(L);
  auto load_chunks = [&](int nch, int bytes, unsigned cls = 0u) {
    if (cls != wmine) return;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      const unsigned cid = (cls << 31) | li;
      if (li >= 4u) tf_wait(sm_u32 + TF_OFF_BARS + tf_wfree_idx(cid) * 8u, ((li >> 2) - 1u) & 1u, p.err);
      const unsigned fb = tf_wfull_bar(sm_u32, cid);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((unsigned)bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                   ::"r"(sm_u32 + tf_slot_off(cid)), "l"(wsrc2 + woff), "r"(bytes), "r"(fb), "l"(wpol) : "memory");
      li += 1;
      woff += (size_t)bytes;
    }
  };
#else
  auto load_chunks = [&](int nch, int bytes, unsigned cls = 0u) {
    (void)cls;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      const unsigned slot = li % TF_NSLOT;
      if ((slot % NLOADERS) == wmine) {
        TF_TR(22, li & 255u);
        if (li >= TF_NSLOT) tf_wait(f32_smem_u32(&wfree[slot]), ((li / TF_NSLOT) - 1u) & 1u, p.err);
        TF_TR(20, li & 255u);
#ifdef TF_WAIT_COST
        if (PROF && li >= TF_NSLOT) {       // calibration: the same wait again (fast path), then two back-to-back trace events
          tf_wait(f32_smem_u32(&wfree[slot]), ((li / TF_NSLOT) - 1u) & 1u, p.err);
          TF_TR(23, li & 255u);
          TF_TR(24, li & 255u);
          unsigned ok_;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok_) : "r"(f32_smem_u32(&wfree[slot])), "r"(((li / TF_NSLOT) - 1u) & 1u) : "memory");
          if (ok_) TF_TR(25, li & 255u);
        }
#endif
        const unsigned fb = tf_wfull_bar(sm_u32, li);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((unsigned)bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(sm_u32 + TF_OFF_W + slot * TF_SLOT), "l"(wsrc + woff), "r"(bytes), "r"(fb), "l"(wpol) : "memory");
        TF_TR(21, li & 255u);
      }
      li += 1;
      woff += (size_t)bytes;
    }
  };
#endif
  // tap loader (warp 6 lane 0)
  unsigned ph_tapfree = 0u;
  auto load_taps = [&](int l, long long t) {
    mbar_expect(tapbar, TF_PAIR);
    bulk_g2s(reinterpret_cast<float*>(b2buf), reinterpret_cast<const float*>(pair_block(l, t)), TF_PAIR, tapbar);
  };
  // epilogue phases
  unsigned ph_accA = 0u, ph_accB = 0u, ph_lg = 0u, ph_smp = 0u;

  // ------------------------------------------------------------------ prologue: the first step's layer-0 taps
  if (tloader) load_taps(0, p.t0);
  if (wloader) load_chunks(8 / TF_CPW, TF_CHUNK_A * TF_CPW);          // T_0 and A_0 sit at the head of the stream
  if (issuer) {
    wait_bar(tapbar, ph_tap);
    operand_fence();
    consume(0u, take(0u, 0, 4u), 4, 128, ACC0, true, sm_u32 + TF_OFF_B2, 2 * TF_BLK);
    commit_to(tapfree);
  }
  TF_MARK(1);
  long long cond_frame = -1;
  float cnd[8];       // condition (+ bias) terms of the next gate / postprocess1 epilogue
#pragma unroll
  for (int j = 0; j < 8; ++j) cnd[j] = 0.f;

  for (long long t = p.t0; t < p.t0 + p.T; ++t) {
    const long long frame_t = (p.ratio > 0) ? (t - p.t0) / p.ratio : 0;
    const bool more = (t + 1 < p.t0 + p.T);
    t_mark = t - p.t0;
    TF_MARK(2);
    tracing = PROF && p.prof != nullptr && blockIdx.x == 0 && p.cluster0 == 0 && (t - p.t0) == 300;
    TF_TR(1, 0);
    TF_PF_START();
    // ================================================================ all threads: frame change -> condition table
    if (frame_t != cond_frame) {
      for (int idx = tid; idx < TF_NS * (TF_C / 4); idx += TF_THREADS) {
        const int n = idx >> 5, c4 = idx & 31;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < nvalid)
          x = __ldg(reinterpret_cast<const float4*>(p.cond + (long long)(b0 + n) * p.cond_bstride + frame_t * TF_C) + c4);
        reinterpret_cast<float4*>(ct_rows)[idx] = x;
      }
      __syncthreads();
      const int cl_ = tid & 31, sg = (tid >> 5) & 7;
      const float* r0 = ct_rows + (2 * sg) * TF_C;
      const float* r1 = r0 + TF_C;
#pragma unroll 1
      for (int st = 0; st <= (tid < 256 ? L : -1); ++st) {
        const float* wsrc_;
        int ld, col;
        float bias;
        if (st < L) {
          wsrc_ = p.layers[st].wlc; ld = 2 * TF_G;
          col = ((cl_ >> 4) ? TF_G : 0) + 16 * rank + (cl_ & 15);
          bias = __ldg(p.layers[st].b1 + col);
        } else {
          wsrc_ = p.post1_lc; ld = TF_S;
          col = 32 * rank + cl_;
          bias = __ldg(p.post1_b + col);
        }
        float a0 = bias, a1 = bias;
#pragma unroll 8
        for (int k = 0; k < TF_C; ++k) {
          const float w = __ldg(wsrc_ + (size_t)k * ld + col);
          a0 = fmaf(r0[k], w, a0);
          a1 = fmaf(r1[k], w, a1);
        }
        float* dst = ctab + (size_t)st * 512;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = 2 * sg + h;
          const float a = h ? a1 : a0;
          int pos;
          if (st < L) {
            // gate epilogue: warp w = ch >> 2, lane = 16 blk + 4 qq + c handles channel 4 w + c, streams
            // 8 blk + 4 (qq >> 1) + 2 (qq & 1) + {0, 1}; entry 2 type + e
            const int ch = cl_ & 15, ty = cl_ >> 4;
            const int ln = 16 * (s >> 3) + 4 * ((s >> 1) & 3) + (ch & 3);
            pos = ((ch >> 2) * 32 + ln) * 4 + 2 * ty + (s & 1);
          } else pos = ((cl_ >> 4) * 32 + 16 * (s >> 3) + (cl_ & 15)) * 8 + (s & 7);
          __stcg(dst + pos, a);
        }
      }
      cond_frame = frame_t;
      __syncthreads();
    }
    // ================================================================ all threads: history -> FIR -> layer-0 input, skip FIR part
    {
      const int slot_t = (int)(t % TF_PK);
      if (ext) {
        if (tid < TF_NS) {
          const int b = b0 + tid;
          float x = 0.f;
          if (tid < nvalid) {
            if (p.mode == GEN_STEP) x = p.ext_audio[b];
            else x = (t > p.t0) ? p.ext_audio[(long long)b * p.T + (t - p.t0 - 1)] : 0.f;
          }
          const float u = mu_law_encode_dev(x, mu, 0.f);
          hist[tid * TF_PK + slot_t] = u;
          if (rank == 0 && tid < nvalid) st_cg(p.u_hist + (long long)b * TF_PK + slot_t, u);
        }
        __syncthreads();
      }
      for (int idx = tid; idx < TF_NS * TF_PK; idx += TF_THREADS) {
        const int i = idx / TF_PK, j = idx - i * TF_PK;
        u_s[idx] = hist[i * TF_PK + ((slot_t - j) & (TF_PK - 1))];
      }
      __syncthreads();
      // the first layer's condition terms are fetched while the FIR runs
      if (warp < 4) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(ctab + (warp * 32 + lane) * 4));
        cnd[0] = a.x; cnd[1] = a.y; cnd[2] = a.z; cnd[3] = a.w;
      }
      // Warps 4-11 (256 threads; the epilogue warps keep their registers): this CTA's OWN 16 channels of the preprocess
      // FIR - the cluster gets them like every other layer input, as one 1 KB slice - and its 32 channels of the skip
      // start.  h0 = (u0*K[PK-1] + b) + u1*K[PK-2] + ... (wavenet_ops.py:178,193: kernel[k-1] is the current sample); the
      // tap weights of a thread are fetched together (one L2 round trip), not one dependent load per tap.
      if (tid >= 128) {
        const int q = tid - 128;
        {
          const int c = q & 15, s_ = q >> 4;
          float w[TF_PK];
#pragma unroll
          for (int j = 0; j < TF_PK; ++j) w[j] = __ldg(p.pre_k + (TF_PK - 1 - j) * TF_R + 16 * rank + c);
          float a = __ldg(p.pre_b + 16 * rank + c);
          const float* u0 = u_s + s_ * TF_PK;
#pragma unroll
          for (int j = 0; j < TF_PK; ++j) a = fmaf(u0[j], w[j], a);
          cur0[c * TF_NS + s_] = a;
          tf_st_split(stg, s_, c, a);
        }
        {
          // skip start folded into the FIR (wavenet.py:127-128): skf[ch][s], ch = 32 rank + (q & 31), streams 2 (q >> 5) + {0,1}
          const int ch = q & 31, sg = q >> 5;
          const float* kp = p.skf_k + 32 * rank + ch;
          float w[TF_PK];
#pragma unroll
          for (int j = 0; j < TF_PK; ++j) w[j] = __ldg(kp + (TF_PK - 1 - j) * TF_S);
          float a0 = __ldg(p.skf_b + 32 * rank + ch), a1 = a0;
          const float* u0 = u_s + (2 * sg) * TF_PK;
#pragma unroll
          for (int j = 0; j < TF_PK; ++j) {
            a0 = fmaf(u0[j], w[j], a0);
            a1 = fmaf(u0[TF_PK + j], w[j], a1);
          }
          skf[ch * TF_NS + 2 * sg] = a0;
          skf[ch * TF_NS + 2 * sg + 1] = a1;
        }
      }
      __syncthreads();
      // push_ops of layer 0 + hand-off (warp 11): the slice goes to the dilation ring once per tap position, and from there
      // to the layer-input half of BOTH stage operands (stage 0 pairs it with zero weights, stage 1 with gate_0)
      if (warp == 11) {
        const int d0 = p.layers[0].d;
        uint8_t* g1 = pair_block(0, t + d0) + rank * 2 * TF_BLK;
        uint8_t* g2 = pair_block(0, t + 2 * d0) + rank * 2 * TF_BLK + TF_BLK;
        publish(stg, g1, g2, 1, TF_OFF_B1 + rank * 2 * TF_BLK + TF_BLK, &b1bar[0]);
        if (lane == 0)
          tc_bulk_multicast(sm_u32 + (unsigned)(TF_OFF_B1 + TF_PAIR + rank * 2 * TF_BLK + TF_BLK), g1, TF_BLK, f32_smem_u32(&b1bar[1]),
                            (unsigned short)0xFFFF);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    TF_PF_ADD(0);
    TF_MARK(7);

    if (issuer) {
      // ============================================================== MMA issue warps (all four run the same sequence)
      // Accumulator reuse needs no barrier of its own: the gather a stage waits for contains this CTA's own slices, which
      // the epilogue publishes only after it has read (and zeroed) the accumulators of the previous stage.  The one
      // exception is the tail's skip chain (e2done).
      unsigned ci = 4u + (unsigned)(t - p.t0) * (12u * (unsigned)L + 16u);       // chunks consumed before this step
      cb = 4u + (unsigned)(t - p.t0) * (8u * (unsigned)L);                       // TF_FIFO2: the same, per class
      cr = (unsigned)(t - p.t0) * (4u * (unsigned)L + 16u);
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const int par = l & 1;
        const uint32_t b1a = sm_u32 + TF_OFF_B1 + par * TF_PAIR;
        TF_MARK(100 + l);
#ifdef TF_ORDER_RA
        const unsigned cid0 = take(ci, l > 0 ? 1 : 0, 4u);      // the stage's first chain: residual + skip (gate for layer 0)
#else
        const unsigned cid0 = take(ci, 0, 4u);                  // the stage's first chain: the gate tile
#endif
        weights_ready(cid0);
        TF_TR(2, l);
        wait_b1(par);
        TF_TR(3, l);
        // next use of this parity's barrier: stage l + 2 (gate + layer input, 32 KB); else the tail's gate (parity L & 1,
        // 16 KB), postprocess2's input (the other parity if it is 1, 32 KB), or the next step's stage 0 (16 KB)
        if (lane == 0 && iss == 0)
          mbar_expect(&b1bar[par], (l + 2 < L || (par != (L & 1) && par == 1)) ? 2 * TF_CS * TF_BLK : TF_CS * TF_BLK);
        operand_fence();
#ifdef TF_ORDER_RA
        // residual + skip chain first: its result x_l is the LATER of the two hand-offs of the stage (it is published
        // behind the gate epilogue) and its operand gate_{l-1} is in the same gather
        if (l > 0) {
          consume(ci, cid0, 4, 96, ACCR, false, b1a, 2 * TF_BLK, true);         // residual + skip rows of layer l-1 x gate_{l-1}
          commit_to(accB);
          TF_TR(5, l);
          ci += 4u;
        }
        consume(ci, l == 0 ? cid0 : take(ci, 0, 4u), 4, 128, par ? ACC1 : ACC0, true, b1a, 2 * TF_BLK, l == 0);   // [P_l | W2_l] x [gate_{l-1} | x_{l-1}] on top of the taps
        commit_to(accA);
        TF_TR(4, l);
        ci += 4u;
#else
        consume(ci, cid0, 4, 128, par ? ACC1 : ACC0, true, b1a, 2 * TF_BLK, true);   // [P_l | W2_l] x [gate_{l-1} | x_{l-1}] on top of the taps
        commit_to(accA);
        TF_TR(4, l);
        ci += 4u;
        if (l > 0) {
          consume(ci, take(ci, 1, 4u), 4, 96, ACCR, false, b1a, 2 * TF_BLK);    // residual + skip rows of layer l-1 x gate_{l-1}
          commit_to(accB);
          TF_TR(5, l);
          ci += 4u;
        }
#endif
        if (l + 1 < L) {
          wait_bar(tapbar, ph_tap);
          TF_TR(6, l);
          operand_fence();
          consume(ci, take(ci, 0, 4u), 4, 128, par ? ACC0 : ACC1, true, sm_u32 + TF_OFF_B2, 2 * TF_BLK);      // taps of layer l+1
          commit_to(tapfree);
          TF_TR(7, l);
          ci += 4u;
        }
      }
      // ---- tail: skip rows of the last layer, postprocess1, postprocess2, next step's layer-0 taps
      {
        const int par = L & 1;
        TF_MARK(150);
        wait_b1(par);
        // parity L & 1 next: postprocess2 input (parity 1, 32 KB), or stage 0 of the next step (16 KB)
        if (lane == 0 && iss == 0) mbar_expect(&b1bar[par], par == 1 ? 2 * TF_CS * TF_BLK : TF_CS * TF_BLK);
        wait_bar(e2done, ph_e2);       // the residual / skip epilogue of layer L-2 has read ACCR (nothing it publishes is waited for)
        operand_fence();
        consume(ci, take(ci, 1, 4u), 4, 96, ACCR, false, sm_u32 + TF_OFF_B1 + par * TF_PAIR, 2 * TF_BLK);
        commit_to(accB);
        ci += 4u;
        cl_arrive();
        TF_MARK(151);
        wait_bar(xsbar, ph_xs);
        if (lane == 0 && iss == 0) mbar_expect(xsbar, 2 * TF_CS * TF_BLK);
        operand_fence();
        consume(ci, take(ci, 1, 8u), 8, 64, ACCP1, false, sm_u32 + TF_OFF_B2, TF_BLK);
        commit_to(accA);
        commit_to(tapfree);
        ci += 8u;
        TF_MARK(152);
        cl_wait();
        TF_MARK(153);
        wait_b1(1);
        if (lane == 0 && iss == 0) mbar_expect(&b1bar[1], 2 * TF_CS * TF_BLK);      // stage 1 of the next step
        operand_fence();
        consume(ci, take(ci, 1, 8u), 8, 32, ACCP2, false, sm_u32 + TF_OFF_B1 + TF_PAIR, TF_BLK);
        commit_to(accA);
        ci += 8u;
        TF_MARK(154);
        if (more) {
          wait_bar(tapbar, ph_tap);
          operand_fence();
          consume(ci, take(ci, 0, 4u), 4, 128, ACC0, true, sm_u32 + TF_OFF_B2, 2 * TF_BLK);
          commit_to(tapfree);
        }
      }
    } else if (warp < 4) {
      // ============================================================== epilogue warps
      uint32_t v0[32], v1[32];
      float cur[8];       // warp 0: float32 residual chain, channel 16 rank + (lane & 15), streams 8 (lane >> 4) .. +8
      float sk[16];       // warp 1 / 2: skip sums (hi / lo rows) of channel 32 rank + lane, 16 streams
      if (warp == 0) {
        const int i = lane & 15, q = lane >> 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = cur0[i * TF_NS + 8 * q + j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = 0.f;
      }
#pragma unroll
      for (int s = 0; s < 16; ++s) sk[s] = (warp == 1) ? skf[lane * TF_NS + s] : 0.f;
      const int blk = lane >> 4, qq = (lane >> 2) & 3, cc = lane & 3;
      // residual + skip epilogue of layer `lr` (its rows were multiplied with gate_lr): x_{lr+1} = x_lr + res + b
      auto res_skip_epilogue = [&](int lr) {
        const bool dead = (lr == L - 1);                      // the last residual is dead (wavenet.py:145)
        const float bres = (warp == 0 && !dead) ? __ldg(p.layers[lr].bres + 16 * rank + (lane & 15)) : 0.f;
        const long long r0_ = prof ? clock64() : 0;
        wait_bar(accB, ph_accB);
        TF_TR(13, lr);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long r1_ = prof ? clock64() : 0;
        if (warp < 3 && !(dead && warp == 0)) {
          tf_ld32_issue(my_taddr + ACCR, v0);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (warp < 3) {
          tf_zero32(my_taddr + ACCR);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0 && lr == L - 2) tf_mbar_arrive(e2done);       // waited for by the tail's skip chain only
          // warps 0-2 have all read and zeroed their rows before warp 0 publishes the layer input that lets the next
          // residual + skip chain start
          asm volatile("bar.sync 3, 96;" ::: "memory");
        }
        const long long r2_ = prof ? clock64() : 0;
        if (prof) { pf[9] += r1_ - r0_; pf[11] += r2_ - r1_; }       // accB wait | TMEM read + zero + 3-warp barrier
        if (warp == 0 && !dead) {
          // lane = 16 q + i: residual rows hi | lo of channel 16 rank + i
          const int q = lane >> 4, i = lane & 15;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float lo_ = __uint_as_float(v0[j]) + __uint_as_float(v0[16 + j]);
            const float hi_ = __uint_as_float(v0[8 + j]) + __uint_as_float(v0[24 + j]);
            const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
            const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            const float nv = cur[j] + (r + bres);                                // wavenet.py:145
            cur[j] = nv;
            tf_st_split(stg + (2 + (lr & 1)) * TF_BLK, 8 * q + j, i, nv);
          }
          // warp 11 publishes the staged slice (ring stores + hand-off, below): warp 0 goes straight on to the gate
          // epilogue.  The staging block alternates by layer parity: block lr & 1 is rewritten at layer lr + 2, whose
          // residual chain needed the cluster-wide gather of x_{lr+1}, i.e. warp 11 has published x_lr before.
          __threadfence_block();
          asm volatile("bar.arrive 4, 64;" ::: "memory");
          TF_TR(14, lr);
        } else if (warp == 1 || warp == 2) {
          // lanes 32-63: skip rows hi, lanes 64-95: skip rows lo of channel 32 rank + lane
#pragma unroll
          for (int s = 0; s < 16; ++s) sk[s] += __uint_as_float(v0[s]) + __uint_as_float(v0[16 + s]);
        }
      };
#pragma unroll 1
      for (int l = 0; l <= L; ++l) {
        const uint32_t acc = (l & 1) ? ACC1 : ACC0;
        TF_MARK(200 + l);
#ifdef TF_ORDER_RA
        if (l > 0) res_skip_epilogue(l - 1);       // its chain is issued first in the stage
#endif
        if (l < L) {
        // ---------------------------------------------------------------- gate of layer l
        wait_bar(accA, ph_accA);
        TF_TR(10, l);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef TF_ORDER_RA
        asm volatile("bar.sync 2, 128;" ::: "memory");       // warp 3 has published the previous gate: the staging buffer is free
#endif
        TF_PF_ADD(1);
        tf_ld32_issue(my_taddr + acc, v0);
        tf_ld32_issue(my_taddr + acc + 32, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tf_zero32(my_taddr + acc);
        tf_zero32(my_taddr + acc + 32);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        {
          // lanes 0-15 of the warp hold the rows that multiplied the first stacked input (columns 0-31), lanes 16-31 those
          // for the second (columns 32-63); columns = 16 hi copies | 16 lo copies of the streams
          float a8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float m0 = blk ? (__uint_as_float(v1[j]) + __uint_as_float(v1[16 + j])) : (__uint_as_float(v0[j]) + __uint_as_float(v0[16 + j]));
            const float m1 = blk ? (__uint_as_float(v1[8 + j]) + __uint_as_float(v1[24 + j])) : (__uint_as_float(v0[8 + j]) + __uint_as_float(v0[24 + j]));
            const float keep = blk ? m1 : m0, send = blk ? m0 : m1;             // keep streams 8 blk + j
            a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);              // first + second stacked product
          }
          float b4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float keep = (qq & 2) ? a8[4 + j] : a8[j], send = (qq & 2) ? a8[j] : a8[4 + j];
            b4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);               // hi rows + lo rows; streams 8 blk + 4 (qq >> 1) + j
          }
          // tanh lanes (qq & 1 = 0) finish streams j = 0, 1, sigmoid lanes streams j = 2, 3
          const float s0 = (qq & 1) ? b4[0] : b4[2], s1 = (qq & 1) ? b4[1] : b4[3];
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
          const float at0 = ((qq & 1) ? r0 : b4[0]) + cnd[0], at1 = ((qq & 1) ? r1 : b4[1]) + cnd[1];
          const float as0 = ((qq & 1) ? b4[2] : r0) + cnd[2], as1 = ((qq & 1) ? b4[3] : r1) + cnd[3];
          const int sbase = 8 * blk + 4 * (qq >> 1) + 2 * (qq & 1);
          tf_st_split(stg, sbase, 4 * warp + cc, tc_tanh(at0) * tc_sigmoid(as0));       // wavenet_ops.py:235-236
          tf_st_split(stg, sbase + 1, 4 * warp + cc, tc_tanh(at1) * tc_sigmoid(as1));
          // condition terms of the next epilogue of this kind: next layer, or postprocess1 (warps 0-1, lane = 16 q + i)
          if (l + 1 < L) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(ctab + (size_t)(l + 1) * 512 + (warp * 32 + lane) * 4));
            cnd[0] = a.x; cnd[1] = a.y; cnd[2] = a.z; cnd[3] = a.w;
          } else if (warp < 2) {
            const float4* src = reinterpret_cast<const float4*>(ctab + (size_t)L * 512 + (warp * 32 + lane) * 8);
            const float4 a = __ldcg(src), b = __ldcg(src + 1);
            cnd[0] = a.x; cnd[1] = a.y; cnd[2] = a.z; cnd[3] = a.w; cnd[4] = b.x; cnd[5] = b.y; cnd[6] = b.z; cnd[7] = b.w;
          }
        }
        TF_PF_ADD(2);
        TF_TR(11, l);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        // gate_l: first input of stage l+1's stacked pair (and of the last layer's skip rows in the tail)
        // (warp 3 publishes; warp 0 goes straight on to the residual rows - the layer input it produces is the later of the
        // two slices the next stage waits for)
        if (warp == 3)
          publish(stg, gst + TF_GST_XG + (l & 1) * TF_CS * TF_BLK + rank * TF_BLK, nullptr, 1,
                  TF_OFF_B1 + ((l + 1) & 1) * TF_PAIR + rank * 2 * TF_BLK, &b1bar[(l + 1) & 1]);
        TF_TR(12, l);
        TF_PF_ADD(3);
        }
        // ---------------------------------------------------------------- residual + skip of layer l-1 (l = L: the tail's
        // skip rows of the last layer)
#ifndef TF_ORDER_RA
        if (l > 0) res_skip_epilogue(l - 1);
        TF_PF_ADD(4);
        if (l < L) asm volatile("bar.sync 2, 128;" ::: "memory");       // the staging buffer is free for the next gate epilogue
#endif
      }
      // ================================================================ tail: relu(skip total) -> postprocess1
      // every ring store of this step is issued: split cluster barrier (waited for before the next step's tap loads)
      TF_MARK(250);
      cl_arrive();
      if (warp == 2) {
#pragma unroll
        for (int s = 0; s < 16; ++s) skx[lane * TF_NS + s] = sk[s];
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp == 1) {
        // skip channel 32 rank + lane = K step 2 rank + (lane >> 4) of postprocess1's input
#pragma unroll
        for (int s = 0; s < 16; ++s)
          tf_st_split(stg + (lane >> 4) * TF_BLK, s, lane & 15, fmaxf(sk[s] + skx[lane * TF_NS + s], 0.f));     // wavenet.py:153
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp == 0) publish(stg, gst + TF_GST_XS + rank * 2 * TF_BLK, nullptr, 2, TF_OFF_B2 + rank * 2 * TF_BLK, xsbar);
      TF_PF_ADD(5);
      // ================================================================ postprocess1 (+ condition), relu
      TF_MARK(251);
      wait_bar(accA, ph_accA);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      TF_MARK(252);
      cl_wait();
      TF_MARK(253);
      const float bias_p2 = (warp == 0) ? __ldg(p.post2_b + 16 * rank + (lane & 15)) : 0.f;
      if (warp < 2) {
        // lane = 16 q + i: rows hi | lo of postprocess1 channel 32 rank + 16 warp + i
        tf_ld32_issue(my_taddr + ACCP1, v0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tf_zero32(my_taddr + ACCP1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const int q = lane >> 4, i = lane & 15;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo_ = __uint_as_float(v0[j]) + __uint_as_float(v0[16 + j]);
          const float hi_ = __uint_as_float(v0[8 + j]) + __uint_as_float(v0[24 + j]);
          const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
          const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16) + cnd[j];
          tf_st_split(stg + warp * TF_BLK, 8 * q + j, i, fmaxf(r, 0.f));     // wavenet.py:163
        }
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp == 0) publish(stg, gst + TF_GST_XN + rank * 2 * TF_BLK, nullptr, 2, TF_OFF_B1 + TF_PAIR + rank * 2 * TF_BLK, &b1bar[1]);
      TF_PF_ADD(6);
      // ================================================================ postprocess2 -> logits, scattered to the drawing CTAs
      TF_MARK(254);
      wait_bar(accA, ph_accA);
      TF_MARK(255);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (warp == 0) {
        // lane = 16 q + i: rows hi | lo of logit 16 rank + i
        tf_ld32_issue(my_taddr + ACCP2, v0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tf_zero32(my_taddr + ACCP2);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const int q = lane >> 4, i = lane & 15;
        const unsigned dst = f32_smem_u32(logits_s + 16 * rank + i);
        const unsigned mb = f32_smem_u32(lgbar);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float lo_ = __uint_as_float(v0[j]) + __uint_as_float(v0[16 + j]);
          const float hi_ = __uint_as_float(v0[8 + j]) + __uint_as_float(v0[24 + j]);
          const float keep = q ? hi_ : lo_, send = q ? lo_ : hi_;
          const float r = keep + __shfl_xor_sync(0xffffffffu, send, 16) + bias_p2;
          const unsigned s = (unsigned)(8 * q + j);
          if ((int)s < nvalid) tc_st_async_f32(cl_mapa(dst, s), r, cl_mapa(mb, s));
        }
        TF_PF_ADD(7);
        // ============================================================== softmax + draw + mu-law decode: CTA s owns stream s
        if (rank < nvalid) {
          TF_MARK(256);
          wait_bar(lgbar, ph_lg);
          TF_MARK(257);
          if (lane == 0) mbar_expect(lgbar, TF_Q * 4);
          const int b = b0 + rank;
          float lg[8];
#pragma unroll
          for (int qi = 0; qi < 8; ++qi) lg[qi] = logits_s[lane + 32 * qi];
          const int k = warp_softmax_draw(p, TF_Q, lg, b, t, lane, prob_s);
          if (k >= 0 && lane < TF_CS) {
            const float un = __ldg(p.enc_lut + k);
            const int slot_n = (int)((t + 1) % TF_PK);
            if (lane == 0) st_cg(p.u_hist + (long long)b * TF_PK + slot_n, un);
            tc_st_async_f32(cl_mapa(f32_smem_u32(hist + rank * TF_PK + slot_n), (unsigned)lane), un,
                            cl_mapa(f32_smem_u32(smpbar), (unsigned)lane));
          }
        }
        TF_PF_ADD(8);
      }
    } else {
      // ============================================================== loader warps: 8 / 9 (lane 0) weight stream, 10 (lane 0) taps
      // their part of the step's ring stores is done: arrive at the step's cluster barrier first - the MMA warps wait on it
      // in the tail while the weight FIFO is still being fed
      if (warp != 11) cl_arrive();
      else {
        // warp 11: publish x_{lr+1} = layer lr+1's input of this step, staged by warp 0's residual epilogue: stored once per
        // tap position of its dilation ring (push_ops); the first copy is also the source of the hand-off to the cluster
        // (second operand of stage lr+2's stacked pair).  Its ring stores precede its arrival at the step's cluster barrier.
#pragma unroll 1
        for (int lr = 0; lr + 1 < L; ++lr) {
          asm volatile("bar.sync 4, 64;" ::: "memory");
          const int ln = lr + 1;
          const int dn = p.layers[ln].d;
          uint8_t* g1 = pair_block(ln, t + dn) + rank * 2 * TF_BLK;
          uint8_t* g2 = pair_block(ln, t + 2 * dn) + rank * 2 * TF_BLK + TF_BLK;
          const bool needed = (ln + 1 < L);
          publish(stg + (2 + (lr & 1)) * TF_BLK, g1, g2, 1, TF_OFF_B1 + ((ln + 1) & 1) * TF_PAIR + rank * 2 * TF_BLK + TF_BLK, needed ? &b1bar[(ln + 1) & 1] : nullptr);
          TF_TR(15, lr);
        }
        cl_arrive();
      }
      if (wloader) {
        // the same chain order as the MMA warps' (op = 3 l + k, tail 3 L ..); the first gate tile of a step was requested
        // at the end of the previous one, and this step ends with the next step's T_0 and A_0 (stream offset wraps to 0)
#pragma unroll 1
#ifdef TF_ORDER_RA
        constexpr int K_RES = 0, OP0 = 2;       // stage l: R_{l-1}, A_l, T_{l+1}; A_0 came with the previous step's tail
#else
        constexpr int K_RES = 1, OP0 = 1;       // stage l: A_l, R_{l-1}, T_{l+1}
#endif
        for (int op = OP0; op < 3 * L + 3; ++op) {
          const int l = (op < 3 * L) ? op / 3 : L;
          const int k = (op < 3 * L) ? op - 3 * l : 3 + (op - 3 * L);
          if ((k == K_RES && l == 0) || (k == 2 && l + 1 >= L)) continue;
          const int nch = (k == 4 || k == 5) ? 8 : 4;
          TF_MARK(300 + op);
          const int bytes = (k == K_RES || k == 3) ? TF_CHUNK_R : (k == 4 ? TF_CHUNK_P1 : (k == 5 ? TF_CHUNK_P2 : TF_CHUNK_A));
          wpol = (l < keep_layers) ? wpol_keep : wpol_stream;
          load_chunks(nch / TF_CPW, bytes * TF_CPW, (bytes == TF_CHUNK_A) ? 0u : 1u);
        }
        wpol = keep_layers > 0 ? wpol_keep : wpol_stream;
        if (more) { woff = 0; load_chunks(8 / TF_CPW, TF_CHUNK_A * TF_CPW); }           // T_0 and A_0 of the next step
      }
      if (tloader) {
#pragma unroll 1
        for (int l = 0; l + 1 < L; ++l) {
          TF_MARK(400 + l);
          wait_bar(tapfree, ph_tapfree);       // the previous pair block (layer l's taps) has been consumed
          TF_TR(30, l);
          load_taps(l + 1, t);
          TF_TR(31, l);
        }
        wait_bar(tapfree, ph_tapfree);         // layer L-1's taps consumed: B2 now receives postprocess1's input
        wait_bar(tapfree, ph_tapfree);         // postprocess1 consumed it
      }
      TF_MARK(450);
      __syncwarp();
      cl_wait();                               // every CTA's ring stores of this step are visible
      if (tloader && more) load_taps(0, t + 1);
    }
    // ================================================================ all threads: the new samples of all streams are in hist
    TF_MARK(500);
    if (!ext) wait_bar(smpbar, ph_smp);
    TF_MARK(501);
    // Free-running modes: a CTA leaves the step only with the new samples of ALL streams, i.e. after every CTA's
    // postprocess2 - nobody can publish into the next step's operands (or count bytes on a receive barrier) while a
    // neighbour still works on this step.  With external inputs nothing couples the CTAs: a cluster barrier does.
    if (ext) cl_barrier();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();       // also: the step's scratch (skip sums, draw) is free before the next step rewrites it
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ext && tid == 0) mbar_expect(smpbar, 4u * (unsigned)nvalid);
    TF_PF_ADD(10);
  }
  TF_MARK(600);
  cl_barrier();
  TF_MARK(601);
  if (prof) for (int i = 0; i < 12; ++i) p.prof[(tid == 0 ? 0 : 16) + i] = pf[i];
  if (prof && tid == 128) for (int i = 0; i < 14; ++i) p.prof[32 + i] = reinterpret_cast<long long*>(sm + TF_OFF_PROF)[i];
#undef TF_PF_START
#undef TF_PF_ADD
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------------------------------------
// weight stream builder
// ---------------------------------------------------------------------------------------------------------------------
// P[k][n] = sum_m Wres[k][m] W2[m][n] (float64 accumulation): Wres = w2prev [G][ldr] columns 0..R-1 (gate channel k ->
// residual channel m), W2 = current-tap rows of w1cur [R][2G]
__global__ void tf_premultiply_kernel(const float* __restrict__ w2prev, int ldr, const float* __restrict__ w1cur,
                                      float* __restrict__ P) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;     // 0 .. 2G-1
  const int k = blockIdx.y;                                // 0 .. G-1
  if (n >= 2 * TF_G) return;
  double a = 0.0;
  for (int m = 0; m < TF_R; ++m) a += (double)w2prev[(size_t)k * ldr + m] * (double)w1cur[(size_t)m * 2 * TF_G + n];
  P[(size_t)k * 2 * TF_G + n] = (float)a;
}
// b1adj[n] = b1[n] + sum_m bres_prev[m] W2[m][n]   (bres_prev null: copy)
__global__ void tf_fold_bias_kernel(const float* __restrict__ b1, const float* __restrict__ bres_prev,
                                    const float* __restrict__ w1cur, float* __restrict__ b1adj) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= 2 * TF_G) return;
  double a = (double)b1[n];
  if (bres_prev)
    for (int m = 0; m < TF_R; ++m) a += (double)bres_prev[m] * (double)w1cur[(size_t)m * 2 * TF_G + n];
  b1adj[n] = (float)a;
}
// one tile of `ninstr` K steps x `rows` rows for every cluster CTA:
//   dst[cta * cta_stride + kk * rows * 32 + (m >> 3) * 256 + (e >> 3) * 128 + (m & 7) * 16 + (e & 7) * 2] = hi or lo part of
//   the weight that multiplies input channel kk * 16 + e in row m.  Row -> (source, column, part) by tile kind:
//   kind 0 (stacked gate tile, 128 rows): m = 32 w + 16 blk + 4 qq + c: column (qq & 1 ? G : 0) + 16 cta + 4 w + c, part
//           qq >> 1, source src0 for blk 0 and src1 for blk 1 (either may be null = zero rows), both [K][2G]
//   kind 1 (residual | skip, 96 rows): m < 32: column 16 cta + (m & 15), part m >> 4; 32..63: column R + 32 cta + m - 32 (hi);
//           64..95: column R + 32 cta + m - 64 (lo); src0 [K][R + S]
//   kind 2 (postprocess1, 64 rows): m = 32 w + 16 q + i: column 32 cta + 16 w + i, part q; src0 [K][S]
//   kind 3 (postprocess2, 32 rows): m = 16 q + i: column 16 cta + i, part q; src0 [K][Q]
__global__ void tf_pack_kernel(const float* __restrict__ src0, const float* __restrict__ src1, int ldw, int ninstr, int rows,
                               int kind, size_t cta_stride, uint8_t* __restrict__ dst) {
  const long long per_cta = (long long)ninstr * rows * 16;
  const long long total = (long long)TF_CS * per_cta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cta = (int)(i / per_cta);
    long long r = i % per_cta;
    const int e = (int)(r % 16); r /= 16;
    const int m = (int)(r % rows);
    const int kk = (int)(r / rows);
    int col, part;
    const float* src = src0;
    if (kind == 0) {
      const int w = m >> 5, blk = (m >> 4) & 1, qq = (m >> 2) & 3, c = m & 3;
      col = ((qq & 1) ? TF_G : 0) + 16 * cta + 4 * w + c; part = qq >> 1;
      src = blk ? src1 : src0;
    } else if (kind == 1) {
      if (m < 32) { col = 16 * cta + (m & 15); part = m >> 4; }
      else if (m < 64) { col = TF_R + 32 * cta + (m - 32); part = 0; }
      else { col = TF_R + 32 * cta + (m - 64); part = 1; }
    } else if (kind == 2) {
      const int w = m >> 5, q = (m >> 4) & 1, ii = m & 15;
      col = 32 * cta + 16 * w + ii; part = q;
    } else {
      col = 16 * cta + (m & 15); part = m >> 4;
    }
    const float x = src ? src[(size_t)(kk * 16 + e) * ldw + col] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 v = part ? __float2bfloat16_rn(x - __bfloat162float(hi)) : hi;
    *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)cta * cta_stride + (size_t)kk * rows * 32 + (m >> 3) * 256 + (e >> 3) * 128 +
                                      (m & 7) * 16 + (e & 7) * 2) = v;
  }
}

}  // namespace vqwn
