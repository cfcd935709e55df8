// VQ nearest-codebook lookup on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
// Reference: model.py:57-74 (direct-form distance, lowest-index argmin, gather, straight-through).
//
// Results are IDENTICAL to vq_direct_kernel (csrc/vq.cuh): the tensor cores only rank the codes.
//   D'[v,k] = z_v . e_k - 0.5*||e_k||^2 + c_v        (kind::tf32, fp32 accumulate in TMEM)
// is maximal where ||z_v - e_k||^2 is minimal.  The -0.5*||e||^2 term (split hi/lo) and the
// per-vector offset c_v (makes D' > 0 so float bits order as integers) ride in 8 extra K columns.
// Every code whose D' lies within `thr` of the running maximum is recorded; thr is twice a
// rigorous bound on |D' - exact| (tf32 operand truncation 2^-10 each, fp32 accumulation, the
// direct form's own rounding), so the exact winner is always recorded.  A vector with a single
// candidate is decided; otherwise the candidates are re-evaluated with the SAME float32
// instruction sequence as vq_direct_kernel (sequential __fsub_rn/__fmaf_rn over d, lowest index
// on ties), reading the untouched fp32 rows that already sit in shared memory as MMA operands.
//
// One persistent CTA per SM, 416 threads:
//   warps 0-7   epilogue: thread = vector = TMEM lane, two warps per lane quadrant (one per accumulator half = 256
//               codes); two passes over the half: maximum, then the codes within the threshold of the maximum;
//               rows with several candidates are queued and re-evaluated exactly, one (row, candidate) pair per thread
//   warp  8     MMA issue (one elected lane), TMEM allocation
//   warps 9-12  loaders: cp.async z tiles (fp32, K-chunk plane layout), ||z||, K-augmentation; 32 rows each
// Shared memory: codebook [512 x 72] fp32 as the B operand (144 KB, loaded once), 2 stages of
// z tile [128 x 72] fp32 as the A operand (36 KB each), candidate lists (6 KB).
// TMEM: 512 columns = two 256-column accumulator halves (codes 0-255 / 256-511), double-buffered
// against the MMA of the next tile.
// Algorithmic bytes per vector: 256 in + 256 (z_q) or 512 (condition row) out + 8 index.
#pragma once
#include "common.cuh"

namespace vqwn {

constexpr int VT_EPI_WARPS = 8;          // 2 warps per TMEM lane quadrant: one per accumulator half
constexpr int VT_MMA_WARP = 8;
constexpr int VT_LOAD_WARP0 = 9;         // 4 loader warps, 32 rows of the tile each
constexpr int VT_LOAD_WARPS = 4;
constexpr int VT_THREADS = 32 * (VT_EPI_WARPS + 1 + VT_LOAD_WARPS);   // 416
constexpr int VT_TILE = 128;             // vectors per tile (MMA M)
constexpr int VT_K = 512;                // codes
constexpr int VT_D = 64;
constexpr int VT_KA = 72;                // augmented K (64 + 8)
// K-major no-swizzle operand tiles stored as K-chunk PLANES: plane c (4 consecutive k) holds the
// 8-row x 16-byte core matrices of all row groups back to back (SBO = 128 B); planes are
// LBO = rows*16 + 16 bytes apart.  The 16-byte pad makes the 18 chunks of one row fall into
// different banks (a row read is 2-way instead of 16-way conflicted) and a thread-per-row read
// of one plane is a contiguous 512 B per warp.
constexpr int VT_SBO = 128;
constexpr int VT_LBO_E = VT_K * 16 + 16;                // 8208
constexpr int VT_LBO_Z = VT_TILE * 16 + 16;             // 2064
constexpr int VT_SE_BYTES = (VT_KA / 4) * VT_LBO_E;     // 147744
constexpr int VT_SZ_BYTES = (VT_KA / 4) * VT_LBO_Z;     // 37152
constexpr int VT_LIST = 4;               // candidates kept per vector per accumulator half
constexpr size_t VT_SMEM = VT_SE_BYTES + 2 * VT_SZ_BYTES + 2 * (VT_LIST + 1) * VT_TILE * 4 + 6 * VT_TILE * 4 + 256;

__device__ __forceinline__ uint32_t vt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) in an operand tile whose planes are `lbo` bytes apart
template <int LBO>
__device__ __forceinline__ uint32_t vt_off(int row, int k) {
  return (uint32_t)((k >> 2) * LBO + row * 16 + (k & 3) * 4);
}
#define VT_OFF_E(row, k) vt_off<VT_LBO_E>(row, k)
#define VT_OFF_Z(row, k) vt_off<VT_LBO_Z>(row, k)

template <int LBO>
__device__ __forceinline__ uint64_t vt_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LBO >> 4) << 16;
  d |= (uint64_t)(VT_SBO >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void vt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void vt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vt_smem_u32(bar)) : "memory");
}
// bounded spin: a broken pipeline must surface as an error, not as a hung GPU
__device__ __forceinline__ bool vt_mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  const uint32_t addr = vt_smem_u32(bar);
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return true;
  }
  atomicExch(err, 1);
  return false;
}

__device__ __forceinline__ void vt_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// running maximum of 32 accumulator values (D' > 0: float bits order as unsigned ints)
#define VT_MAX32(v, m)                                                                        \
  do {                                                                                        \
    _Pragma("unroll") for (int i_ = 0; i_ < 32; ++i_) m = max(m, v[i_]);                      \
  } while (0)

// record every value of a 32-column block that reaches thr_key (group maxima gate the rare path; the body is
// branch-free: predicated store into slot min(cnt, VT_LIST), the last slot being a dummy)
#define VT_COLLECT32(v, kbase)                                                                \
  do {                                                                                        \
    _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) {                                        \
      uint32_t m_ = v[8 * i_];                                                                \
      _Pragma("unroll") for (int j_ = 1; j_ < 8; ++j_) m_ = max(m_, v[8 * i_ + j_]);          \
      if (m_ >= thr_key) {                                                                    \
        uint32_t mask_ = 0;   /* independent compares; the serial part is the rare loop below */ \
        _Pragma("unroll") for (int j_ = 0; j_ < 8; ++j_) mask_ |= (v[8 * i_ + j_] >= thr_key ? 1u : 0u) << j_; \
        while (mask_) {                                                                       \
          const int j_ = __ffs(mask_) - 1;                                                    \
          mask_ &= mask_ - 1;                                                                 \
          mylist[min(cnt, VT_LIST) * VT_TILE] = (uint32_t)((kbase) + 8 * i_ + j_);            \
          ++cnt;                                                                              \
        }                                                                                     \
      }                                                                                       \
    }                                                                                         \
  } while (0)

__device__ __forceinline__ void vt_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(VT_THREADS, 1)
vq_tc_kernel(const float* __restrict__ z, const float* __restrict__ E, long long N,
             long long* __restrict__ idx_out, float* __restrict__ zq_out, int out_stride,
             const float* __restrict__ spk_table, const int* __restrict__ spk_idx, int spk_dim, int F,
             const float* __restrict__ emax_p, int* __restrict__ err, long long* __restrict__ prof, int out_code) {
  extern __shared__ __align__(1024) uint8_t vt_smem_raw[];
  uint8_t* smem = vt_smem_raw;
  uint8_t* sE = smem;
  uint8_t* sZ0 = smem + VT_SE_BYTES;
  uint32_t* list_p = reinterpret_cast<uint32_t*>(sZ0 + 2 * VT_SZ_BYTES);          // [half][slot][row] code indices
  uint32_t* hmax_s = list_p + 2 * (VT_LIST + 1) * VT_TILE;                        // [half][row] maximum key of the half
  uint32_t* cnt_s = hmax_s + 2 * VT_TILE;                                         // [half][row] candidates found
  float* zn = reinterpret_cast<float*>(cnt_s + 2 * VT_TILE);                      // [stage][row] ||z||
  uint64_t* bars = reinterpret_cast<uint64_t*>(zn + 2 * VT_TILE);
  uint64_t* z_full = bars;        // [2] loaders -> MMA, epilogue
  uint64_t* z_empty = bars + 2;   // [2] epilogue + MMA -> loaders
  uint64_t* acc_full = bars + 4;  // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 6; // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  __shared__ int bestk_s[VT_TILE];
  __shared__ int fail_s;
  __shared__ int rq_n, ov_n;                          // rows queued for exact re-evaluation / for a full exact scan
  __shared__ unsigned char rq_rows[VT_TILE];          // queued rows from the front, full-scan rows from the back
  if (threadIdx.x == 0) { fail_s = 0; rq_n = 0; ov_n = 0; }

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float emax = __ldg(emax_p) * 1.000001f;
  const long long ntiles = (N + VT_TILE - 1) / VT_TILE;
  const bool pf = prof != nullptr && blockIdx.x == 0 && lane == 0;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long pt = clock64();
#define VT_PF(i) do { if (pf) { long long n_ = clock64(); pc[i] += n_ - pt; pt = n_; } } while (0)

  // ---- one-time setup: codebook -> shared memory (fp32 bits untouched) + K augmentation
#pragma unroll 4
  for (int i = tid; i < VT_K * (VT_D / 4); i += VT_THREADS) {
    const int k = i >> 4, c = i & 15;
    const float4 v = __ldg(reinterpret_cast<const float4*>(E + (size_t)k * VT_D) + c);
    *reinterpret_cast<float4*>(sE + VT_OFF_E(k, 4 * c)) = v;
  }
  __syncthreads();
  for (int k = tid; k < VT_K; k += VT_THREADS) {
    float ne = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 e = *reinterpret_cast<const float4*>(sE + VT_OFF_E(k, 4 * c));
      ne = fmaf(e.x, e.x, ne); ne = fmaf(e.y, e.y, ne); ne = fmaf(e.z, e.z, ne); ne = fmaf(e.w, e.w, ne);
    }
    const float hi = __uint_as_float(__float_as_uint(ne) & 0xFFFFE000u);   // exactly tf32-representable
    const float lo = ne - hi;
    *reinterpret_cast<float4*>(sE + VT_OFF_E(k, 64)) = make_float4(hi, lo, 1.0f, 0.f);
    *reinterpret_cast<float4*>(sE + VT_OFF_E(k, 68)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      vt_mbar_init(&z_full[i], VT_LOAD_WARPS);
      vt_mbar_init(&z_empty[i], VT_EPI_WARPS + 1);   // epilogue warps + MMA commit
      vt_mbar_init(&acc_full[i], 1);
      vt_mbar_init(&acc_empty[i], 4);                // the 4 epilogue warps of that half
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == VT_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(vt_smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  VT_PF(0);

  if (warp >= VT_LOAD_WARP0) {
    // ================================================================= loaders (32 rows each)
    const int r0 = (warp - VT_LOAD_WARP0) * 32;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      uint8_t* sZ = sZ0 + s * VT_SZ_BYTES;
      if (it >= 2 && !vt_mbar_wait(&z_empty[s], ((it >> 1) - 1) & 1, err)) break;
      const long long v0 = tile * VT_TILE;
#pragma unroll 4
      for (int m = 0; m < 16; ++m) {
        const int i = lane + 32 * m;
        const int r = r0 + (i >> 4), c = i & 15;
        const bool valid = (v0 + r) < N;
        const float* src = valid ? (z + (size_t)(v0 + r) * VT_D + 4 * c) : z;
        const uint32_t dst = vt_smem_u32(sZ + VT_OFF_Z(r, 4 * c));
        const int nbytes = valid ? 16 : 0;    // zero-fill rows past the end
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      {
        // ||z|| and the augmented K columns [-0.5, -0.5, c_v, 0 | 0, 0, 0, 0]; lane = row: conflict-free plane reads
        const int r = r0 + lane;
        float nz = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(sZ + VT_OFF_Z(r, 4 * c));
          nz = fmaf(v.x, v.x, nz); nz = fmaf(v.y, v.y, nz); nz = fmaf(v.z, v.z, nz); nz = fmaf(v.w, v.w, nz);
        }
        const float nrm = sqrtf(nz);
        const float cv = 1.02f * (nrm * emax + 0.5f * emax * emax) + 1e-30f;
        *reinterpret_cast<float4*>(sZ + VT_OFF_Z(r, 64)) = make_float4(-0.5f, -0.5f, cv, 0.f);
        *reinterpret_cast<float4*>(sZ + VT_OFF_Z(r, 68)) = make_float4(0.f, 0.f, 0.f, 0.f);
        zn[s * VT_TILE + r] = nrm;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) vt_mbar_arrive(&z_full[s]);
    }
  } else if (warp == VT_MMA_WARP) {
    // ================================================================= MMA issue
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(elected));
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (!vt_mbar_wait(&z_full[s], (it >> 1) & 1, err)) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_base = vt_smem_u32(sZ0 + s * VT_SZ_BYTES);
      bool ok = true;
      for (int h = 0; h < 2 && ok; ++h) {
        if (it >= 1) ok = vt_mbar_wait(&acc_empty[h], (it - 1) & 1, err);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_base = vt_smem_u32(sE) + (uint32_t)h * (256 / 8) * VT_SBO;   // codes 256.. start 32 row groups in
        const uint32_t d_addr = tmem + (uint32_t)h * 256;
#pragma unroll
        for (int ks = 0; ks < VT_KA / 8; ++ks) {
          const uint64_t da = vt_desc<VT_LBO_Z>(a_base + ks * 2 * VT_LBO_Z);
          const uint64_t db = vt_desc<VT_LBO_E>(b_base + ks * 2 * VT_LBO_E);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                       "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(d_addr), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(elected) : "memory");
        }
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                     ::"r"(vt_smem_u32(&acc_full[h])), "r"(elected) : "memory");
      }
      if (!ok) break;
      // the z stage is free for the loaders once these MMAs have read it (and the epilogue is done)
      asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                   "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                   ::"r"(vt_smem_u32(&z_empty[s])), "r"(elected) : "memory");
    }
  } else {
    // ================================================================= epilogue
    // warp = hsel*4 + q: TMEM lane quadrant q (vectors 32q..32q+31), accumulator half hsel (codes 256*hsel..)
    const int q = warp & 3, hsel = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * 256);
    uint32_t* mylist = list_p + hsel * (VT_LIST + 1) * VT_TILE + row;
    int it = 0;
    bool ok = true;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint8_t* sZ = sZ0 + s * VT_SZ_BYTES;
      const long long v0 = tile * VT_TILE;
      ok = ok && vt_mbar_wait(&z_full[s], (it >> 1) & 1, err);   // zn[] and the fp32 rows are in place
      VT_PF(1);
      const float nrm = zn[s * VT_TILE + row];
      // thr = 2 x bound on |D' - exact| (see header): tf32 operand truncation, accumulation slop,
      // and the float32 direct form's own rounding
      const float ze = nrm * emax;
      const float bound = ze * (1.0f / 512.0f) + (2.f * ze + emax * emax) * (1.0f / 8192.0f) +
                          (nrm + emax) * (nrm + emax) * (1.0f / 131072.0f);
      const float thr = 2.0f * bound;
      ok = ok && vt_mbar_wait(&acc_full[hsel], it & 1, err);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      VT_PF(2);
      // ---- pass 1: maximum key of this half (8 blocks of 32 columns, next TMEM load in flight)
      uint32_t va[32], vb[32];
      uint32_t hm = 0;
      vt_ld32(lane_base, va);
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        vt_ld32(lane_base + (uint32_t)(c0 + 32), vb);
        VT_MAX32(va, hm);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        vt_ld32(lane_base + (uint32_t)((c0 + 64) & 255), va);     // wraps to block 0: first block of pass 2
        VT_MAX32(vb, hm);
      }
      hmax_s[hsel * VT_TILE + row] = hm;
      vt_epi_sync();
      const uint32_t gmax = max(hmax_s[row], hmax_s[VT_TILE + row]);
      const float tthr = __uint_as_float(gmax) - thr;
      const uint32_t thr_key = tthr > 0.f ? __float_as_uint(tthr) : 0u;
      VT_PF(3);
      // ---- pass 2: codes of this half within thr of the global maximum
      int cnt = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        vt_ld32(lane_base + (uint32_t)(c0 + 32), vb);
        VT_COLLECT32(va, hsel * 256 + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 64 < 256) vt_ld32(lane_base + (uint32_t)(c0 + 64), va);
        VT_COLLECT32(vb, hsel * 256 + c0 + 32);
      }
      cnt_s[hsel * VT_TILE + row] = (uint32_t)cnt;
      if (!ok) fail_s = 1;        // a timed-out wait: every epilogue warp leaves after this tile (uniform decision)
      // this accumulator half may be overwritten by the next tile's MMA
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) vt_mbar_arrive(&acc_empty[hsel]);
      vt_epi_sync();
      VT_PF(4);

      // ---- decide: single candidate -> done.  Rows with several candidates (about one in five) are queued and their
      //      (row, candidate) pairs are re-evaluated exactly, ONE PAIR PER THREAD over all 256 epilogue threads: the
      //      divergent per-row loops this replaces made every warp pay for its worst row.  Winner per row through a
      //      64-bit shared-memory atomicMin on (distance bits << 32 | code): smallest distance, lowest code on ties.
      {
        const int n0 = (int)cnt_s[row], n1 = (int)cnt_s[VT_TILE + row];
        const int total = n0 + n1;
        const bool overflow = n0 > VT_LIST || n1 > VT_LIST || total == 0;   // overflow: exact scan of the whole codebook
        const bool need = overflow || total != 1;
        unsigned long long* rowbest = reinterpret_cast<unsigned long long*>(hmax_s);     // [row], hmax_s is free now
        if (hsel == 0) {
          if (!need) {
            const int best_k = (n0 > 0) ? (int)list_p[row] : (int)list_p[(VT_LIST + 1) * VT_TILE + row];
            bestk_s[row] = best_k;
            if (idx_out != nullptr && v0 + row < N) idx_out[v0 + row] = (long long)best_k;
          } else {
            rowbest[row] = ~0ULL;
            if (overflow) rq_rows[VT_TILE - 1 - atomicAdd(&ov_n, 1)] = (unsigned char)row;
            else rq_rows[atomicAdd(&rq_n, 1)] = (unsigned char)row;
          }
        }
        vt_epi_sync();
        auto exact = [&](int r, int k) {
          // the float32 instruction sequence of vq_direct_kernel (sequential over d)
          float dist = 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float4 z4 = *reinterpret_cast<const float4*>(sZ + VT_OFF_Z(r, 4 * c));
            const float4 e4 = *reinterpret_cast<const float4*>(sE + VT_OFF_E(k, 4 * c));
            float t;
            t = __fsub_rn(z4.x, e4.x); dist = __fmaf_rn(t, t, dist);
            t = __fsub_rn(z4.y, e4.y); dist = __fmaf_rn(t, t, dist);
            t = __fsub_rn(z4.z, e4.z); dist = __fmaf_rn(t, t, dist);
            t = __fsub_rn(z4.w, e4.w); dist = __fmaf_rn(t, t, dist);
          }
          atomicMin(&rowbest[r], ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(unsigned)k);
        };
        const int nq = rq_n, no = ov_n;
        for (int pidx = tid; pidx < nq * 2 * VT_LIST; pidx += 32 * VT_EPI_WARPS) {
          const int r = (int)rq_rows[pidx / (2 * VT_LIST)], j = pidx % (2 * VT_LIST);
          const int m0 = (int)cnt_s[r], m1 = (int)cnt_s[VT_TILE + r];
          if (j < m0 + m1)
            exact(r, (j < m0) ? (int)list_p[j * VT_TILE + r] : (int)list_p[((VT_LIST + 1) + (j - m0)) * VT_TILE + r]);
        }
        for (int q2 = 0; q2 < no; ++q2) {
          const int r = (int)rq_rows[VT_TILE - 1 - q2];
          for (int k = tid; k < VT_K; k += 32 * VT_EPI_WARPS) exact(r, k);
        }
        vt_epi_sync();
        if (hsel == 0 && need) {
          const int best_k = (int)(unsigned)(rowbest[row] & 0xffffffffULL);
          bestk_s[row] = best_k;
          if (idx_out != nullptr && v0 + row < N) idx_out[v0 + row] = (long long)best_k;
        }
        if (tid == 0) { rq_n = 0; ov_n = 0; }      // for the next tile (two barriers away from the next use)
      }
      vt_epi_sync();
      VT_PF(5);

      // ---- fused gather + straight-through + speaker concat: warp w writes rows 16w..16w+15, 2 per iteration
      if (zq_out != nullptr) {
        const int sub = lane >> 4, ch = lane & 15;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int orow = warp * 16 + 2 * j + sub;
          const int kb = bestk_s[orow];
          const long long gv = v0 + orow;
          const float4 zz = *reinterpret_cast<const float4*>(sZ + VT_OFF_Z(orow, 4 * ch));
          const float4 ee = *reinterpret_cast<const float4*>(sE + VT_OFF_E(kb, 4 * ch));
          float4 o;
          o.x = __fadd_rn(zz.x, __fsub_rn(ee.x, zz.x));               // model.py:73
          o.y = __fadd_rn(zz.y, __fsub_rn(ee.y, zz.y));
          o.z = __fadd_rn(zz.z, __fsub_rn(ee.z, zz.z));
          o.w = __fadd_rn(zz.w, __fsub_rn(ee.w, zz.w));
          if (gv < N) *reinterpret_cast<float4*>(zq_out + (size_t)gv * out_stride + 4 * ch) = out_code ? ee : o;     // Magenta/config.py:242
        }
        if (spk_dim > 0) {
          for (int j = 0; j < 8; ++j) {
            const long long gv = v0 + warp * 16 + 2 * j + sub;
            if (gv < N) {
              const float* srow = spk_table + (size_t)spk_idx[(int)(gv / F)] * spk_dim;
              float* op = zq_out + (size_t)gv * out_stride + VT_D;
              for (int c = 4 * ch; c < spk_dim; c += 64)
                *reinterpret_cast<float4*>(op + c) = __ldg(reinterpret_cast<const float4*>(srow + c));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) vt_mbar_arrive(&z_empty[s]);
      VT_PF(6);
      if (fail_s) break;
    }
    if (pf && warp == 0) for (int i = 0; i < 8; ++i) prof[16 + i] = pc[i];
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == VT_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// max_k ||e_k|| (one block, K <= 1024 threads); feeds the candidate threshold of vq_tc_kernel
__global__ void vq_emax_kernel(const float* __restrict__ E, int K, int D, float* __restrict__ out) {
  __shared__ float red[32];
  float nrm = 0.f;
  if ((int)threadIdx.x < K) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) { const float e = E[(size_t)threadIdx.x * D + d]; s = fmaf(e, e, s); }
    nrm = sqrtf(s);
  }
  for (int off = 16; off > 0; off >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(0xffffffffu, nrm, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (threadIdx.x == 0) out[0] = v;
  }
}

}  // namespace vqwn
