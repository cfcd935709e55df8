// C-ABI implementation (include/vqwn.h) over the sm_100a kernels in this directory.
// Host side only does bookkeeping: tensor registry by reference variable name, weight
// packing (device-to-device copies), state allocation, launches and timing.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vqwn.h"
#include "common.cuh"
#include "vq.cuh"
#include "vq_tc.cuh"
#include "vq_tc2.cuh"
#include "wavenet_fp32.cuh"
#include "wavenet_fp32_cluster.cuh"
#include "wavenet_bf16_cluster.cuh"
#include "wavenet_tc_cluster.cuh"
#include "wavenet_tcf_cluster.cuh"
#include "sample.cuh"
#include "encoder.cuh"

using namespace vqwn;

namespace {

thread_local std::string g_create_error;

struct TensorSlot {
  std::string name;
  std::vector<int64_t> shape;
  size_t numel = 0;
  float* dev = nullptr;
  bool set = false;
  bool required = true;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct vqwn_handle {
  vqwn_config cfg;
  int device = 0;
  int max_batch = 0, Bp_max = 0;
  int num_sms = 0;
  int precision = VQWN_PREC_FP32;
  int vq_kernel = VQWN_VQ_AUTO;
  float* emax_dev = nullptr;
  int* vq_err = nullptr;
  bool emax_valid = false;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<TensorSlot> tensors;
  std::unordered_map<std::string, int> index;
  // derived dims
  int L = 0, R = 0, G = 0, S = 0, Q = 0, C = 0, PK = 0, K = 0, D = 0, SPK = 0;
  int actA_floats = 0, actB_floats = 0, wfloatsA = 0, wfloatsB = 0;
  size_t smem_fp32 = 0;
  float* wtiles = nullptr;          // all tile-major weights, one allocation (L2 persistence window)
  size_t wtiles_floats = 0;
  std::vector<size_t> off_w1t, off_w2t;
  size_t off_skip0t = 0, off_post1t = 0, off_post2t = 0;
  int* gen_err = nullptr;
  int vq_out_code = 0;             // 1: the VQ emits e_k (Magenta/config.py:242) instead of z_e + (e_k - z_e) (model.py:73)
  int* tf_err_host = nullptr;      // host-mapped error record of the one-hand-off tensor-core kernel (readable after a trap)
  int* tf_err_dev = nullptr;
  int gen_kernel = 0;                      // 0 auto (cluster kernel when it applies), 1 barrier, 3 cluster
  // cluster kernel (wavenet_fp32_cluster.cuh)
  bool cl_ok = false;                      // geometry constraints hold and 16-CTA clusters can be scheduled
  int cl_max_clusters = 0;                 // co-resident clusters of 16 CTAs (measured 7 on a B200)
  float* wcl = nullptr;                    // cluster-tiled weights, same block order as wtiles
  ClLayerDev* cl_layers_dev = nullptr;
  std::vector<ClLayerDev> cl_layers_host;
  int cl_w1_floats = 0, cl_w2_floats = 0;
  // bf16 tensor-core cluster kernel (wavenet_bf16_cluster.cuh)
  bool bc_ok = false;                      // default geometry and 8-CTA clusters schedulable
  int bc_max_clusters = 0;
  __nv_bfloat16* wbc = nullptr;            // bf16 K-major plane tiles, same block order as wtiles
  BcLayerDev* bc_layers_dev = nullptr;
  std::vector<BcLayerDev> bc_layers_host;
  // split-bf16 (float32-grade) tensor-core cluster kernel (wavenet_tc_cluster.cuh)
  bool tc_ok = false;                      // default geometry and 16-CTA clusters schedulable
  int tc_max_clusters = 0;
  __nv_bfloat16* wtc = nullptr;            // [L][16][144 KB] hi/lo tiles, then postprocess1 [16][64 KB], postprocess2 [16][32 KB]
  size_t wtc_bytes = 0;
  size_t tc_persist_bytes = 0;             // persisting L2 set aside for the weight tiles (0: not available)
  size_t tc_max_window = 0;                // cudaDevAttrMaxAccessPolicyWindowSize
  float *tc_skf_k = nullptr, *tc_skf_b = nullptr, *tc_ctab = nullptr;
  uint8_t* tc_gstage = nullptr;            // [co-resident cluster][TC_GSTAGE] hand-off staging
  // one-hand-off-per-layer variant (wavenet_tcf_cluster.cuh): the default VQWN_PREC_TC kernel; VQWN_TC_KERNEL=v1 selects the older one
  bool tf_use = true;
  bool tc_reproducible = false;            // vqwn_set_reproducible: fixed accumulation order in the tensor-core kernel
  uint8_t* wtf = nullptr;                  // [16][tf_stream_bytes(L)] per-CTA weight streams
  size_t wtf_bytes = 0;
  float* tf_ptmp = nullptr;                // [G][2G] premultiplied W2_l . Wres_{l-1} of the layer being packed
  float* tf_b1adj = nullptr;               // [L][2G] gated biases with W2_l . bres_{l-1} folded in
  uint8_t* tf_gstage = nullptr;
  long long* dbg_host = nullptr;
  TfLayerDev* tf_layers_dev = nullptr;
  std::vector<TfLayerDev> tf_layers_host;
  const float** tc_b2_ptrs = nullptr;
  TcLayerDev* tc_layers_dev = nullptr;
  std::vector<TcLayerDev> tc_layers_host;
  int stream_offset = 0;                   // global index of stream 0 (vqwn_set_stream_offset)
  // packed fp32 weights
  bool packed = false;
  std::vector<float*> w1, b1, w2, b2;
  float *post1_w = nullptr;
  LayerDev* layers_dev = nullptr;
  std::vector<LayerDev> layers_host;
  float *enc_lut = nullptr, *dec_lut = nullptr;
  // state
  float *u_hist = nullptr, *cur = nullptr, *g = nullptr, *skip = nullptr, *n1 = nullptr, *logits = nullptr;
  float* ring_base = nullptr;
  size_t ring_floats = 0;
  unsigned long long* barrier = nullptr;
  long long* prof = nullptr;
  bool profile = false;
  int B = 0;            // streams of the current run (vqwn_reset)
  long long t = 0;      // steps since reset
  // resident / staging buffers
  DevBuf cond_res, uni_res, audio_res, idx_res, logits_res, x_res, small_a, small_b, small_c, small_d;
  DevBuf vq_z, vq_idx, vq_out, spk_idx;
  DevBuf enc_x, enc_a, enc_b, enc_fold, enc_z, enc_c, enc_d, enc_e, enc_f;
  int cond_B = 0, cond_F = 0;
  long long uni_T = 0; int uni_B = 0;
  long long out_T = 0; int out_B = 0;
  long long vq_n = 0;
  // instrumentation
  std::string err;
  double last_ms = 0.0;
  int64_t launches = 0;
  const char* last_kernel = "";
};

namespace {

int fail(vqwn_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

#define CK(h, call)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      char buf_[512];                                                                    \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                      \
      return fail(h, e_ == cudaErrorMemoryAllocation ? VQWN_ERR_NOMEM : VQWN_ERR_CUDA, buf_); \
    }                                                                                    \
  } while (0)

int ensure(vqwn_handle* h, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return VQWN_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr; b.bytes = 0;
  if (bytes == 0) bytes = 16;
  CK(h, cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return VQWN_OK;
}

void add_tensor(vqwn_handle* h, const std::string& name, std::vector<int64_t> shape, bool required = true) {
  TensorSlot s;
  s.name = name;
  s.shape = shape;
  s.numel = 1;
  for (auto d : shape) s.numel *= (size_t)d;
  s.required = required;
  h->index[name] = (int)h->tensors.size();
  h->tensors.push_back(s);
}

std::string layer_scope(const vqwn_config& c, int i) {
  char buf[64];
  snprintf(buf, sizeof buf, "decoder/cycle_%d/layer_%d", 1 + i / c.num_cycle_layers, 1 + i % c.num_cycle_layers);
  return buf;
}

float* TP(vqwn_handle* h, const std::string& name) { return h->tensors[h->index.at(name)].dev; }

// default LUTs (mu_law_ops.py:26-31 and :5-8 in float32 libm); the Python shim overrides them
// with NumPy-computed tables ("lut/mu_law_decode", "lut/mu_law_encode") for bit parity with
// the reference's NumPy decode.
void default_luts(int Q, std::vector<float>& dec, std::vector<float>& enc) {
  const float mu = (float)(Q - 1);
  dec.resize(Q + 1); enc.resize(Q + 1);
  for (int k = 0; k <= Q; ++k) {
    float y = (2.0f * (float)k / mu) - 1.0f;
    float s = (y > 0.f) ? 1.f : ((y < 0.f) ? -1.f : 0.f);
    float x = s * (powf(1.0f + mu, fabsf(y)) - 1.0f) / mu;
    dec[k] = x;
    float xc = fminf(fmaxf(x, -1.f), 1.f);
    float sx = (xc > 0.f) ? 1.f : ((xc < 0.f) ? -1.f : 0.f);
    enc[k] = sx * log1pf(mu * fabsf(xc)) / log1pf(mu);
  }
}

int pack_weights(vqwn_handle* h) {
  if (h->packed) return VQWN_OK;
  for (auto& s : h->tensors)
    if (s.required && !s.set) return fail(h, VQWN_ERR_STATE, "tensor not set: " + s.name);
  const int R = h->R, G = h->G, S = h->S, C = h->C;
  const size_t f = sizeof(float);
  for (int l = 0; l < h->L; ++l) {
    const std::string sc = layer_scope(h->cfg, l);
    const float* gk = TP(h, sc + "/gated/kernel");          // [3,R,2G]
    // rows: kernel[2] (current) ; kernel[1] (t-d) ; kernel[0] (t-2d) ; local_condition
    for (int tap = 0; tap < 3; ++tap)
      CK(h, cudaMemcpyAsync(h->w1[l] + (size_t)tap * R * 2 * G, gk + (size_t)(2 - tap) * R * 2 * G,
                            (size_t)R * 2 * G * f, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->w1[l] + (size_t)3 * R * 2 * G, TP(h, sc + "/gated/local_condition/kernel"),
                          (size_t)C * 2 * G * f, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->b1[l], TP(h, sc + "/gated/bias"), (size_t)2 * G * f, cudaMemcpyDeviceToDevice, h->stream));
    // cols: residual | skip
    CK(h, cudaMemcpy2DAsync(h->w2[l], (size_t)(R + S) * f, TP(h, sc + "/residual/kernel"), (size_t)R * f,
                            (size_t)R * f, G, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpy2DAsync(h->w2[l] + R, (size_t)(R + S) * f, TP(h, sc + "/skip/kernel"), (size_t)S * f,
                            (size_t)S * f, G, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->b2[l], TP(h, sc + "/residual/bias"), (size_t)R * f, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->b2[l] + R, TP(h, sc + "/skip/bias"), (size_t)S * f, cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(h, cudaMemcpyAsync(h->post1_w, TP(h, "decoder/postprocess1/kernel"), (size_t)S * S * f,
                        cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->post1_w + (size_t)S * S, TP(h, "decoder/postprocess1/local_condition/kernel"),
                        (size_t)C * S * f, cudaMemcpyDeviceToDevice, h->stream));
  // tile-major copies for the bulk-copy engine (one contiguous block per stage tile)
  {
    auto pack = [&](const float* src, int ldw, int K, int NC, int ntiles, int pairedG, float* dst) {
      const long long total = (long long)ntiles * K * NC;
      int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
      pack_tiles_kernel<<<grid, 256, 0, h->stream>>>(src, ldw, K, NC, ntiles, pairedG, dst);
      h->launches += 1;
    };
    pack(TP(h, "decoder/skip/kernel"), S, R, 16, S / 16, 0, h->wtiles + h->off_skip0t);
    for (int l = 0; l < h->L; ++l) {
      pack(h->w1[l], 2 * G, 3 * R + C, 16, G / 8, G, h->wtiles + h->off_w1t[l]);
      pack(h->w2[l], R + S, G, 32, (R + S) / 32, 0, h->wtiles + h->off_w2t[l]);
    }
    pack(h->post1_w, S, S + C, 16, S / 16, 0, h->wtiles + h->off_post1t);
    pack(TP(h, "decoder/postprocess2/kernel"), h->Q, S, 16, h->Q / 16, 0, h->wtiles + h->off_post2t);
    CK(h, cudaGetLastError());
  }
  if (h->cl_ok) {
    // per-CTA column slices of the cluster kernel: [16][K][columns of CTA r]
    auto packc = [&](const float* src, int ldw, int K, int n0, int base1, int n1, float* dst) {
      const long long total = (long long)CL_CS * K * (n0 + n1);
      int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
      pack_cluster_kernel<<<grid, 256, 0, h->stream>>>(src, ldw, K, n0, base1, n1, CL_CS, dst);
      h->launches += 1;
    };
    const int Qn = h->Q;
    packc(TP(h, "decoder/skip/kernel"), S, R, S / CL_CS, 0, 0, h->wcl + h->off_skip0t);
    for (int l = 0; l < h->L; ++l) {
      packc(h->w1[l], 2 * G, 3 * R + C, G / CL_CS, G, G / CL_CS, h->wcl + h->off_w1t[l]);
      packc(h->w2[l], R + S, G, R / CL_CS, R, S / CL_CS, h->wcl + h->off_w2t[l]);
    }
    packc(h->post1_w, S, S + C, S / CL_CS, 0, 0, h->wcl + h->off_post1t);
    packc(TP(h, "decoder/postprocess2/kernel"), Qn, S, Qn / CL_CS, 0, 0, h->wcl + h->off_post2t);
    CK(h, cudaGetLastError());
  }
  if (h->bc_ok) {
    auto packb = [&](const float* src, int ldw, int K, int rows, int mode, int n0, int base1, __nv_bfloat16* dst) {
      const long long total = (long long)BC_CS * K * rows;
      int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
      pack_bf16_planes_kernel<<<grid, 256, 0, h->stream>>>(src, ldw, K, rows, mode, n0, base1, BC_CS, dst);
      h->launches += 1;
    };
    packb(TP(h, "decoder/skip/kernel"), S, R, BC_NSK, 0, 0, 0, h->wbc + h->off_skip0t);
    for (int l = 0; l < h->L; ++l) {
      packb(h->w1[l], 2 * G, 3 * R + C, BC_ROWS_S1, 1, 0, G, h->wbc + h->off_w1t[l]);
      packb(h->w2[l], R + S, G, BC_ROWS_S2, 2, BC_NR, R, h->wbc + h->off_w2t[l]);
    }
    packb(h->post1_w, S, S + C, BC_NSK, 0, 0, 0, h->wbc + h->off_post1t);
    packb(TP(h, "decoder/postprocess2/kernel"), h->Q, S, BC_NQ, 0, 0, 0, h->wbc + h->off_post2t);
    CK(h, cudaGetLastError());
  }
  if (h->tc_ok) {
    auto packt = [&](const float* src, int ldw, int k0, int K, int rows, int kind, size_t cta_stride_bytes, uint8_t* dst) {
      const long long total = (long long)TC_CS * K * rows;
      int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
      pack_tc_tiles_kernel<<<grid, 256, 0, h->stream>>>(src, ldw, k0, K, rows, kind, cta_stride_bytes / 2,
                                                       reinterpret_cast<__nv_bfloat16*>(dst));
      h->launches += 1;
    };
    uint8_t* base = reinterpret_cast<uint8_t*>(h->wtc);
    for (int l = 0; l < h->L; ++l) {
      uint8_t* lb = base + (size_t)l * TC_CS * TC_LAYER_BYTES;
      // w1 rows: [0,R) current tap (kernel[2]) | [R,2R) tap t-d (kernel[1]) | [2R,3R) tap t-2d (kernel[0]) | condition
      packt(h->w1[l], 2 * G, 0, R, TC_ROWS1, 0, TC_LAYER_BYTES, lb);
      packt(h->w1[l], 2 * G, R, R, TC_ROWS1, 0, TC_LAYER_BYTES, lb + TC_W1);
      packt(h->w1[l], 2 * G, 2 * R, R, TC_ROWS1, 0, TC_LAYER_BYTES, lb + 2 * TC_W1);
      packt(h->w2[l], R + S, 0, G, TC_ROWS2, 1, TC_LAYER_BYTES, lb + 3 * TC_W1);
    }
    uint8_t* p1 = base + (size_t)h->L * TC_CS * TC_LAYER_BYTES;
    uint8_t* p2 = p1 + (size_t)TC_CS * TC_WP1;
    packt(h->post1_w, S, 0, S, TC_ROWSP1, 2, TC_WP1, p1);
    packt(TP(h, "decoder/postprocess2/kernel"), h->Q, 0, S, TC_ROWSP2, 3, TC_WP2, p2);
    std::vector<const float*> b2p(h->L);
    for (int l = 0; l < h->L; ++l) b2p[l] = h->b2[l];
    CK(h, cudaMemcpyAsync(h->tc_b2_ptrs, b2p.data(), sizeof(const float*) * h->L, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    fold_skip_start_kernel<<<(TC_S + 127) / 128, 128, 0, h->stream>>>(
        TP(h, "decoder/preprocess/kernel"), TP(h, "decoder/preprocess/bias"), TP(h, "decoder/skip/kernel"),
        TP(h, "decoder/skip/bias"), h->tc_b2_ptrs, h->L, h->tc_skf_k, h->tc_skf_b);
    h->launches += 1;
    CK(h, cudaGetLastError());
    // weight streams of the one-hand-off-per-layer kernel: per cluster CTA, tiles in issue order
    //   T_0 | A_0 T_1 | A_1 R_0 T_2 | ... | A_{L-1} R_{L-2} | R_{L-1} postprocess1 postprocess2
    //   A_l = [P_l = W2_l . Wres_{l-1} | W2_l] (stacked inputs [gate_{l-1} | x_{l-1}]), T_l = [W1_l | W0_l] (taps t-d | t-2d)
    {
      const size_t SB = tf_stream_bytes(h->L);
#ifdef TF_FIFO2
      // two streams per CTA (the kernel's two weight FIFOs): [16][gate / tap tiles in issue order], then [16][the rest]
      const size_t SBig = tf_big_bytes(h->L), SSmall = tf_small_bytes(h->L);
      uint8_t* const base_small = h->wtf + (size_t)TF_CS * SBig;
      size_t offb = 0, offs = 0;
      auto pk = [&](const float* s0, const float* s1, int ldw, int ninstr, int rows, int kind, size_t bytes) {
        const bool big = (kind == 0);
        const long long total = (long long)TF_CS * ninstr * rows * 16;
        int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
        tf_pack_kernel<<<grid, 256, 0, h->stream>>>(s0, s1, ldw, ninstr, rows, kind, big ? SBig : SSmall,
                                                    big ? h->wtf + offb : base_small + offs);
        if (big) offb += bytes; else offs += bytes;
        h->launches += 1;
      };
#else
      size_t off = 0;
      auto pk = [&](const float* s0, const float* s1, int ldw, int ninstr, int rows, int kind, size_t bytes) {
        const long long total = (long long)TF_CS * ninstr * rows * 16;
        int grid = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
        tf_pack_kernel<<<grid, 256, 0, h->stream>>>(s0, s1, ldw, ninstr, rows, kind, SB, h->wtf + off);
        off += bytes;
        h->launches += 1;
      };
#endif
      const size_t tap = (size_t)R * 2 * G;      // w1 rows: current tap | tap t-d | tap t-2d | condition
      pk(h->w1[0] + tap, h->w1[0] + 2 * tap, 2 * G, 16, 128, 0, TF_TILE_A);
      for (int l = 0; l < h->L; ++l) {
#ifdef TF_ORDER_RA
        // stage l: R_{l-1} before A_l (the kernel issues the residual + skip chain first)
        if (l >= 1) pk(h->w2[l - 1], nullptr, R + S, 16, 96, 1, TF_TILE_R);
#endif
        if (l == 0) {
          pk(nullptr, h->w1[0], 2 * G, 16, 128, 0, TF_TILE_A);
          tf_fold_bias_kernel<<<(2 * G + 127) / 128, 128, 0, h->stream>>>(h->b1[0], nullptr, h->w1[0], h->tf_b1adj);
        } else {
          tf_premultiply_kernel<<<dim3((2 * G + 127) / 128, G), 128, 0, h->stream>>>(h->w2[l - 1], R + S, h->w1[l], h->tf_ptmp);
          pk(h->tf_ptmp, h->w1[l], 2 * G, 16, 128, 0, TF_TILE_A);
          tf_fold_bias_kernel<<<(2 * G + 127) / 128, 128, 0, h->stream>>>(h->b1[l], h->b2[l - 1], h->w1[l],
                                                                           h->tf_b1adj + (size_t)l * 2 * G);
        }
        h->launches += 2;
#ifndef TF_ORDER_RA
        if (l >= 1) pk(h->w2[l - 1], nullptr, R + S, 16, 96, 1, TF_TILE_R);
#endif
        if (l + 1 < h->L) pk(h->w1[l + 1] + tap, h->w1[l + 1] + 2 * tap, 2 * G, 16, 128, 0, TF_TILE_A);
      }
      pk(h->w2[h->L - 1], nullptr, R + S, 16, 96, 1, TF_TILE_R);
      pk(h->post1_w, nullptr, S, 32, 64, 2, TF_TILE_P1);
      pk(TP(h, "decoder/postprocess2/kernel"), nullptr, h->Q, 32, 32, 3, TF_TILE_P2);
#ifdef TF_FIFO2
      const size_t off = (offb == SBig && offs == SSmall) ? SB : 0;
#endif
      if (off != SB) return fail(h, VQWN_ERR_INVALID, "tensor-core kernel: weight stream size mismatch");
      CK(h, cudaGetLastError());
    }
  }
  CK(h, cudaMemcpyAsync(h->enc_lut, TP(h, "lut/mu_law_encode"), (size_t)(h->Q + 1) * f, cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->dec_lut, TP(h, "lut/mu_law_decode"), (size_t)(h->Q + 1) * f, cudaMemcpyDeviceToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->packed = true;
  return VQWN_OK;
}

size_t tc_ring_bytes(const vqwn_handle* h, int B);

int do_reset(vqwn_handle* h, int B) {
  if (B < 1 || B > h->max_batch) return fail(h, VQWN_ERR_INVALID, "batch out of range (1..max_batch)");
  const size_t Bp = h->Bp_max;
  // only the part of the ring storage this padded batch uses (layout: launch_fp32)
  const size_t Bp_run = (size_t)(B + FP32_TB - 1) / FP32_TB * FP32_TB;
  size_t ring_clear = h->ring_floats / Bp * Bp_run * sizeof(float);
  if (h->tc_ok && tc_ring_bytes(h, B) > ring_clear) ring_clear = tc_ring_bytes(h, B);
  if (ring_clear > h->ring_floats * sizeof(float)) ring_clear = h->ring_floats * sizeof(float);
  CK(h, cudaMemsetAsync(h->ring_base, 0, ring_clear, h->stream));
  CK(h, cudaMemsetAsync(h->u_hist, 0, Bp * h->PK * sizeof(float), h->stream));
  CK(h, cudaMemsetAsync(h->cur, 0, Bp * h->R * sizeof(float), h->stream));
  CK(h, cudaMemsetAsync(h->g, 0, Bp * h->G * sizeof(float), h->stream));
  CK(h, cudaMemsetAsync(h->skip, 0, Bp * h->S * sizeof(float), h->stream));
  CK(h, cudaMemsetAsync(h->n1, 0, Bp * h->S * sizeof(float), h->stream));
  CK(h, cudaMemsetAsync(h->logits, 0, Bp * h->Q * sizeof(float), h->stream));
  h->B = B;
  h->t = 0;
  return VQWN_OK;
}

// ---------------------------------------------------------------------------------------
// cluster kernel plumbing
// ---------------------------------------------------------------------------------------
typedef void (*cl_kernel_t)(const ClParams);
cl_kernel_t cl_kernel_for(int MS) {
  switch (MS) {
    case 2: return wavenet_fp32_cluster<2>;
    case 4: return wavenet_fp32_cluster<4>;
    case 6: return wavenet_fp32_cluster<6>;
    case 8: return wavenet_fp32_cluster<8>;
    default: return wavenet_fp32_cluster<10>;
  }
}

// K-groups of a stage: as many as fit the 256 threads, dividing K into multiples of 4, with the partial sums
// fitting the (aliased) weight buffer
int cl_kgroups(int K, int NC, int MS, int buf_floats) {
  int kg = 256 / (NC / 4);
  if (kg > K / 4) kg = K / 4;
  for (; kg > 1; --kg)
    if (K % (4 * kg) == 0 && kg * MS * NC <= buf_floats) break;
  return kg < 1 ? 1 : kg;
}

size_t cl_smem_bytes(const vqwn_handle* h, int MS) {
  const int R = h->R, G = h->G, S = h->S, Q = h->Q, C = h->C, PK = h->PK;
  const int NSK = S / CL_CS, NC2 = R / CL_CS + NSK;
  int stage_cols = NC2 > 2 * G / CL_CS ? NC2 : 2 * G / CL_CS;
  if (NSK > stage_cols) stage_cols = NSK;
  size_t fl = (size_t)h->cl_w1_floats + h->cl_w2_floats + (size_t)MS * (G + 3 * R + C + Q + 2 * PK + NSK + stage_cols) + 1;
  return fl * sizeof(float) + 4 * sizeof(unsigned long long) + 16;
}

// streams per cluster for a batch: the smallest supported size that covers B with the co-resident clusters
int cl_streams_per_cluster(const vqwn_handle* h, int B) {
  if (!h->cl_ok || h->cl_max_clusters < 1) return 0;
  const int need = (B + h->cl_max_clusters - 1) / h->cl_max_clusters;
  for (int ms = 2; ms <= CL_MAX_MS; ms += 2)
    if (ms >= need) return ms;
  return CL_MAX_MS;      // more streams than the co-resident clusters hold: several launches of full clusters
}

int launch_cluster(vqwn_handle* h, int MS, int mode, long long T, const float* cond, long long cond_bstride, int ratio,
                   const float* ext_audio, const double* uniforms, uint64_t seed, float* audio_out, int* idx_out,
                   float* logits_out, float* probs_out) {
  ClParams p;
  memset(&p, 0, sizeof p);
  const int R = h->R, G = h->G, S = h->S, Q = h->Q, C = h->C;
  p.L = h->L; p.R = R; p.G = G; p.S = S; p.Q = Q; p.C = C; p.PK = h->PK;
  p.B = h->B;
  p.Bp = (h->B + FP32_TB - 1) / FP32_TB * FP32_TB;      // ring layout shared with the barrier kernel
  p.w1_floats = h->cl_w1_floats; p.w2_floats = h->cl_w2_floats;
  const int NSK = S / CL_CS, NC1 = 2 * G / CL_CS, NC2 = R / CL_CS + NSK, NQ = Q / CL_CS;
  p.kg_s0 = cl_kgroups(R, NSK, MS, h->cl_w2_floats);
  p.kg_s1 = cl_kgroups(3 * R + C, NC1, MS, h->cl_w1_floats);
  p.kg_s2 = cl_kgroups(G, NC2, MS, h->cl_w2_floats);
  p.kg_p1 = cl_kgroups(S + C, NSK, MS, h->cl_w1_floats);
  p.kg_p2 = cl_kgroups(S, NQ, MS, h->cl_w2_floats);
  p.pre_k = TP(h, "decoder/preprocess/kernel"); p.pre_b = TP(h, "decoder/preprocess/bias");
  p.skip0c = h->wcl + h->off_skip0t; p.skip0_b = TP(h, "decoder/skip/bias");
  p.post1c = h->wcl + h->off_post1t; p.post1_b = TP(h, "decoder/postprocess1/bias");
  p.post2c = h->wcl + h->off_post2t; p.post2_b = TP(h, "decoder/postprocess2/bias");
  size_t off = 0;
  for (int l = 0; l < h->L; ++l) {
    h->cl_layers_host[l].ring = h->ring_base + off;
    off += (size_t)2 * h->cfg.dilations[l] * p.Bp * R;
  }
  CK(h, cudaMemcpyAsync(h->cl_layers_dev, h->cl_layers_host.data(), sizeof(ClLayerDev) * h->L, cudaMemcpyHostToDevice, h->stream));
  p.layers = h->cl_layers_dev;
  p.enc_lut = h->enc_lut; p.dec_lut = h->dec_lut;
  p.u_hist = h->u_hist;
  p.t0 = h->t; p.T = T; p.mode = mode;
  p.cond = cond; p.cond_bstride = cond_bstride; p.ratio = ratio;
  p.ext_audio = ext_audio; p.uniforms = uniforms; p.seed = seed; p.b_offset = h->stream_offset;
  p.audio_out = audio_out; p.idx_out = idx_out; p.logits_out = logits_out; p.probs_out = probs_out;
  p.prof = h->profile ? h->prof : nullptr;
  p.err = h->gen_err;
  CK(h, cudaMemsetAsync(h->gen_err, 0, sizeof(int), h->stream));
  const int nclusters = (h->B + MS - 1) / MS;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = cl_smem_bytes(h, MS);
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(h, cudaEventRecord(h->ev0, h->stream));
  // clusters are independent (streams never interact): a batch larger than the co-resident clusters runs as
  // consecutive launches over disjoint stream groups
  for (int c0 = 0; c0 < nclusters; c0 += h->cl_max_clusters) {
    const int nc = (nclusters - c0 < h->cl_max_clusters) ? (nclusters - c0) : h->cl_max_clusters;
    p.cluster0 = c0;
    cfg.gridDim = dim3(nc * CL_CS);
    CK(h, cudaLaunchKernelEx(&cfg, cl_kernel_for(MS), p));
    h->launches += 1;
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "wavenet_fp32_cluster";
  h->t += T;
  return VQWN_OK;
}

int launch_bf16(vqwn_handle* h, int mode, long long T, const float* cond, long long cond_bstride, int ratio,
                const float* ext_audio, const double* uniforms, uint64_t seed, float* audio_out, int* idx_out,
                float* logits_out, float* probs_out) {
  const int nclusters = (h->B + BC_NS - 1) / BC_NS;
  BcParams p;
  memset(&p, 0, sizeof p);
  p.L = h->L; p.B = h->B; p.nclusters = nclusters;
  p.pre_k = TP(h, "decoder/preprocess/kernel"); p.pre_b = TP(h, "decoder/preprocess/bias");
  p.skip0 = h->wbc + h->off_skip0t; p.skip0_b = TP(h, "decoder/skip/bias");
  p.post1 = h->wbc + h->off_post1t; p.post1_b = TP(h, "decoder/postprocess1/bias");
  p.post2 = h->wbc + h->off_post2t; p.post2_b = TP(h, "decoder/postprocess2/bias");
  // bf16 rings in plane layout live in the same storage as the float32 rings (half the bytes)
  size_t off = 0;
  __nv_bfloat16* rb = reinterpret_cast<__nv_bfloat16*>(h->ring_base);
  for (int l = 0; l < h->L; ++l) {
    h->bc_layers_host[l].ring = rb + off;
    off += (size_t)2 * h->cfg.dilations[l] * nclusters * BC_R * BC_NS;
  }
  if (off * sizeof(__nv_bfloat16) > h->ring_floats * sizeof(float)) return fail(h, VQWN_ERR_INVALID, "bf16 kernel: ring storage too small");
  CK(h, cudaMemcpyAsync(h->bc_layers_dev, h->bc_layers_host.data(), sizeof(BcLayerDev) * h->L, cudaMemcpyHostToDevice, h->stream));
  p.layers = h->bc_layers_dev;
  p.enc_lut = h->enc_lut; p.dec_lut = h->dec_lut;
  p.u_hist = h->u_hist;
  p.t0 = h->t; p.T = T; p.mode = mode;
  p.cond = cond; p.cond_bstride = cond_bstride; p.ratio = ratio;
  p.ext_audio = ext_audio; p.uniforms = uniforms; p.seed = seed; p.b_offset = h->stream_offset;
  p.audio_out = audio_out; p.idx_out = idx_out; p.logits_out = logits_out; p.probs_out = probs_out;
  p.prof = h->profile ? h->prof : nullptr;
  p.err = h->gen_err;
  CK(h, cudaMemsetAsync(h->gen_err, 0, sizeof(int), h->stream));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(BC_THREADS);
  cfg.dynamicSmemBytes = BC_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = BC_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int c0 = 0; c0 < nclusters; c0 += h->bc_max_clusters) {      // disjoint stream groups, one launch per co-resident set
    const int nc = (nclusters - c0 < h->bc_max_clusters) ? (nclusters - c0) : h->bc_max_clusters;
    p.cluster0 = c0;
    cfg.gridDim = dim3(nc * BC_CS);
    CK(h, cudaLaunchKernelEx(&cfg, wavenet_bf16_cluster, p));
    h->launches += 1;
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "wavenet_bf16_cluster";
  h->t += T;
  return VQWN_OK;
}

// streams per cluster / clusters of a batch for the split-bf16 kernel: spread the streams over the co-resident
// clusters (fewer streams per cluster = less hand-off traffic), at most 16 per cluster
int tc_spc(const vqwn_handle* h, int B) {
  int spc = (B + h->tc_max_clusters - 1) / h->tc_max_clusters;
  if (spc > TC_NS) spc = TC_NS;
  if (spc < 1) spc = 1;
  return spc;
}
size_t tc_ring_bytes(const vqwn_handle* h, int B) {
  const int spc = tc_spc(h, B);
  const size_t ncl = (size_t)(B + spc - 1) / spc;
  size_t dsum = 0;
  for (int l = 0; l < h->L; ++l) dsum += (size_t)h->cfg.dilations[l];
  // 2d + 1 slots per layer; the one-hand-off kernel stores pair blocks (both tap positions) of 32 KB
  return (2 * dsum + (size_t)h->L) * ncl * (size_t)TF_PAIR;
}

int launch_tcf(vqwn_handle* h, int mode, long long T, const float* cond, long long cond_bstride, int ratio,
               const float* ext_audio, const double* uniforms, uint64_t seed, float* audio_out, int* idx_out,
               float* logits_out, float* probs_out) {
  const int spc = tc_spc(h, h->B);
  const int nclusters = (h->B + spc - 1) / spc;
  if (h->t + T > 0x7fffffffLL) return fail(h, VQWN_ERR_INVALID, "tensor-core kernel: time index beyond 2^31 samples; call vqwn_reset");
  if (T > 10000000LL) return fail(h, VQWN_ERR_INVALID, "tensor-core kernel: at most 10^7 time steps per call (32-bit weight-FIFO position)");
  TfParams p;
  memset(&p, 0, sizeof p);
  p.L = h->L; p.B = h->B; p.nclusters = nclusters; p.spc = spc;
  p.pre_k = TP(h, "decoder/preprocess/kernel"); p.pre_b = TP(h, "decoder/preprocess/bias");
  p.skf_k = h->tc_skf_k; p.skf_b = h->tc_skf_b;
  p.wstream = h->wtf;
  p.flags = h->tc_reproducible ? 1 : 0;
  {
    // weight tiles of the first N stages ask to stay in the L2 (evict_last), the rest stream (evict_first); VQWN_TC_L2_LAYERS
    int keep = TF_L2_KEEP_LAYERS;
    if (const char* e = getenv("VQWN_TC_L2_LAYERS")) keep = atoi(e);
    if (keep < 0) keep = 0;
    if (keep > 255) keep = 255;
    p.flags |= keep << 8;
  }
  p.post1_lc = h->post1_w + (size_t)h->S * h->S;
  p.post1_b = TP(h, "decoder/postprocess1/bias"); p.post2_b = TP(h, "decoder/postprocess2/bias");
  size_t off = 0;
  uint8_t* rb = reinterpret_cast<uint8_t*>(h->ring_base);
  for (int l = 0; l < h->L; ++l) {
    h->tf_layers_host[l].ring = rb + off;
    off += ((size_t)2 * h->cfg.dilations[l] + 1) * nclusters * TF_PAIR;
  }
  if (off > h->ring_floats * sizeof(float)) return fail(h, VQWN_ERR_INVALID, "tensor-core kernel: ring storage too small");
  CK(h, cudaMemcpyAsync(h->tf_layers_dev, h->tf_layers_host.data(), sizeof(TfLayerDev) * h->L, cudaMemcpyHostToDevice, h->stream));
  p.layers = h->tf_layers_dev;
  p.ctab = h->tc_ctab;
  p.gstage = h->tf_gstage;
  p.enc_lut = h->enc_lut; p.dec_lut = h->dec_lut;
  p.u_hist = h->u_hist;
  p.t0 = h->t; p.T = T; p.mode = mode;
  p.cond = cond; p.cond_bstride = cond_bstride; p.ratio = ratio;
  p.ext_audio = ext_audio; p.uniforms = uniforms; p.seed = seed; p.b_offset = h->stream_offset;
  p.audio_out = audio_out; p.idx_out = idx_out; p.logits_out = logits_out; p.probs_out = probs_out;
  p.prof = h->profile ? h->prof : nullptr;
#ifdef TF_DEBUG_MARKS
  static long long* dbg_host = nullptr;
  if (!dbg_host) { cudaHostAlloc(&dbg_host, 131072, cudaHostAllocMapped); memset(dbg_host, 0, 131072); }
  { long long* dp = nullptr; cudaHostGetDevicePointer(&dp, dbg_host, 0); p.prof = dp; }
  h->dbg_host = dbg_host;
#endif
  // the error record lives in host-mapped memory: a wait that times out traps, and device memory is unreadable afterwards
  CK(h, cudaMemsetAsync(h->gen_err, 0, sizeof(int), h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  memset(h->tf_err_host, 0, 4096 * sizeof(int));
  p.err = h->tf_err_dev;
#ifdef TF_DEBUG_MARKS
  { memset(dbg_host + 1024, 0, 3072 * 8); memset(dbg_host + 4096, 0, 32768); p.err = reinterpret_cast<int*>(p.prof + 4096); }
#endif
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(TF_THREADS);
  cfg.dynamicSmemBytes = TF_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TF_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (h->tc_persist_bytes > 0 && !getenv("VQWN_TC_NO_PERSIST")) {
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    // (a 50-layer stream - the Magenta topology - is larger than the largest window the device accepts)
    const size_t win = h->wtf_bytes < h->tc_max_window ? h->wtf_bytes : h->tc_max_window;
    attr[1].val.accessPolicyWindow.base_ptr = h->wtf;
    attr[1].val.accessPolicyWindow.num_bytes = win;
    attr[1].val.accessPolicyWindow.hitRatio = (float)((double)h->tc_persist_bytes >= (double)win ? 1.0 : (double)h->tc_persist_bytes / (double)win);
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 2;
  }
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int c0 = 0; c0 < nclusters; c0 += h->tc_max_clusters) {      // disjoint stream groups, one launch per co-resident set
    const int nc = (nclusters - c0 < h->tc_max_clusters) ? (nclusters - c0) : h->tc_max_clusters;
    p.cluster0 = c0;
    cfg.gridDim = dim3(nc * TF_CS);
    if (h->profile) CK(h, cudaLaunchKernelEx(&cfg, wavenet_tcf_cluster<true>, p));
    else CK(h, cudaLaunchKernelEx(&cfg, wavenet_tcf_cluster<false>, p));
    h->launches += 1;
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "wavenet_tcf_cluster";
  h->t += T;
  return VQWN_OK;
}

int launch_tc(vqwn_handle* h, int mode, long long T, const float* cond, long long cond_bstride, int ratio,
              const float* ext_audio, const double* uniforms, uint64_t seed, float* audio_out, int* idx_out,
              float* logits_out, float* probs_out) {
  const int spc = tc_spc(h, h->B);
  const int nclusters = (h->B + spc - 1) / spc;
  TcParams p;
  memset(&p, 0, sizeof p);
  p.L = h->L; p.B = h->B; p.nclusters = nclusters; p.spc = spc;
  p.pre_k = TP(h, "decoder/preprocess/kernel"); p.pre_b = TP(h, "decoder/preprocess/bias");
  p.skf_k = h->tc_skf_k; p.skf_b = h->tc_skf_b;
  uint8_t* base = reinterpret_cast<uint8_t*>(h->wtc);
  p.post1 = reinterpret_cast<const __nv_bfloat16*>(base + (size_t)h->L * TC_CS * TC_LAYER_BYTES);
  p.post2 = reinterpret_cast<const __nv_bfloat16*>(base + (size_t)h->L * TC_CS * TC_LAYER_BYTES + (size_t)TC_CS * TC_WP1);
  p.post1_lc = h->post1_w + (size_t)h->S * h->S;
  p.post1_b = TP(h, "decoder/postprocess1/bias"); p.post2_b = TP(h, "decoder/postprocess2/bias");
  size_t off = 0;
  uint8_t* rb = reinterpret_cast<uint8_t*>(h->ring_base);
  for (int l = 0; l < h->L; ++l) {
    h->tc_layers_host[l].ring = reinterpret_cast<__nv_bfloat16*>(rb + off);
    off += ((size_t)2 * h->cfg.dilations[l] + 1) * nclusters * TC_XB;
  }
  if (off > h->ring_floats * sizeof(float)) return fail(h, VQWN_ERR_INVALID, "tensor-core kernel: ring storage too small");
  CK(h, cudaMemcpyAsync(h->tc_layers_dev, h->tc_layers_host.data(), sizeof(TcLayerDev) * h->L, cudaMemcpyHostToDevice, h->stream));
  p.layers = h->tc_layers_dev;
  p.ctab = h->tc_ctab;
  p.gstage = h->tc_gstage;
  p.enc_lut = h->enc_lut; p.dec_lut = h->dec_lut;
  p.u_hist = h->u_hist;
  p.t0 = h->t; p.T = T; p.mode = mode;
  p.cond = cond; p.cond_bstride = cond_bstride; p.ratio = ratio;
  p.ext_audio = ext_audio; p.uniforms = uniforms; p.seed = seed; p.b_offset = h->stream_offset;
  p.audio_out = audio_out; p.idx_out = idx_out; p.logits_out = logits_out; p.probs_out = probs_out;
  p.prof = h->profile ? h->prof : nullptr;
  p.err = h->gen_err;
  CK(h, cudaMemsetAsync(h->gen_err, 0, sizeof(int), h->stream));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TC_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (h->tc_persist_bytes > 0 && !getenv("VQWN_TC_NO_PERSIST")) {
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow.base_ptr = h->wtc;
    const size_t win = h->wtc_bytes < h->tc_max_window ? h->wtc_bytes : h->tc_max_window;
    attr[1].val.accessPolicyWindow.num_bytes = win;
    attr[1].val.accessPolicyWindow.hitRatio = (float)((double)h->tc_persist_bytes >= (double)win ? 1.0 : (double)h->tc_persist_bytes / (double)win);
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 2;
  }
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int c0 = 0; c0 < nclusters; c0 += h->tc_max_clusters) {      // disjoint stream groups, one launch per co-resident set
    const int nc = (nclusters - c0 < h->tc_max_clusters) ? (nclusters - c0) : h->tc_max_clusters;
    p.cluster0 = c0;
    cfg.gridDim = dim3(nc * TC_CS);
    CK(h, cudaLaunchKernelEx(&cfg, wavenet_tc_cluster, p));
    h->launches += 1;
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "wavenet_tc_cluster";
  h->t += T;
  return VQWN_OK;
}

// ring layout depends on the padded batch of the run; rebuilt at every launch
int launch_fp32(vqwn_handle* h, int mode, long long T, const float* cond, long long cond_bstride, int ratio,
                const float* ext_audio, const double* uniforms, uint64_t seed, float* audio_out, int* idx_out,
                float* logits_out, float* probs_out) {
  if (h->precision == VQWN_PREC_TC && h->tf_use)
    return launch_tcf(h, mode, T, cond, cond_bstride, ratio, ext_audio, uniforms, seed, audio_out, idx_out, logits_out,
                      probs_out);
  if (h->precision == VQWN_PREC_TC)
    return launch_tc(h, mode, T, cond, cond_bstride, ratio, ext_audio, uniforms, seed, audio_out, idx_out, logits_out,
                     probs_out);
  if (h->precision == VQWN_PREC_BF16)
    return launch_bf16(h, mode, T, cond, cond_bstride, ratio, ext_audio, uniforms, seed, audio_out, idx_out, logits_out,
                       probs_out);
  if (h->gen_kernel == 0 || h->gen_kernel == 3) {
    const int ms = cl_streams_per_cluster(h, h->B);
    if (ms > 0)
      return launch_cluster(h, ms, mode, T, cond, cond_bstride, ratio, ext_audio, uniforms, seed, audio_out, idx_out,
                            logits_out, probs_out);
    if (h->gen_kernel == 3) return fail(h, VQWN_ERR_INVALID, "cluster kernel does not apply to this geometry / batch");
  }
  GenParams p;
  memset(&p, 0, sizeof p);
  p.L = h->L; p.R = h->R; p.G = h->G; p.S = h->S; p.Q = h->Q; p.C = h->C; p.PK = h->PK;
  p.B = h->B;
  p.Bp = (h->B + FP32_TB - 1) / FP32_TB * FP32_TB;
  p.actA_floats = h->actA_floats; p.actB_floats = h->actB_floats; p.wfloatsA = h->wfloatsA; p.wfloatsB = h->wfloatsB;
  p.pre_k = TP(h, "decoder/preprocess/kernel"); p.pre_b = TP(h, "decoder/preprocess/bias");
  p.skip0t = h->wtiles + h->off_skip0t; p.skip0_b = TP(h, "decoder/skip/bias");
  p.post1t = h->wtiles + h->off_post1t; p.post1_b = TP(h, "decoder/postprocess1/bias");
  p.post2t = h->wtiles + h->off_post2t; p.post2_b = TP(h, "decoder/postprocess2/bias");
  // per-layer ring bases for this padded batch
  size_t off = 0;
  for (int l = 0; l < h->L; ++l) {
    h->layers_host[l].ring = h->ring_base + off;
    off += (size_t)2 * h->cfg.dilations[l] * p.Bp * h->R;
  }
  CK(h, cudaMemcpyAsync(h->layers_dev, h->layers_host.data(), sizeof(LayerDev) * h->L, cudaMemcpyHostToDevice, h->stream));
  p.layers = h->layers_dev;
  p.enc_lut = h->enc_lut; p.dec_lut = h->dec_lut;
  p.u_hist = h->u_hist; p.cur = h->cur; p.g = h->g; p.skip = h->skip; p.n1 = h->n1; p.logits = h->logits;
  p.t0 = h->t; p.T = T; p.mode = mode;
  p.cond = cond; p.cond_bstride = cond_bstride; p.ratio = ratio;
  p.ext_audio = ext_audio; p.uniforms = uniforms; p.seed = seed; p.b_offset = h->stream_offset;
  p.audio_out = audio_out; p.idx_out = idx_out; p.logits_out = logits_out; p.probs_out = probs_out;
  p.barrier = h->barrier;
  p.prof = h->profile ? h->prof : nullptr;
  p.err = h->gen_err;
  CK(h, cudaMemsetAsync(h->gen_err, 0, sizeof(int), h->stream));
  CK(h, cudaMemsetAsync(h->barrier, 0, 32 * sizeof(unsigned long long), h->stream));
  CK(h, cudaEventRecord(h->ev0, h->stream));
  void* args[] = {&p};
  CK(h, cudaLaunchCooperativeKernel((const void*)wavenet_fp32_persistent, dim3(h->num_sms), dim3(FP32_THREADS),
                                    args, h->smem_fp32, h->stream));
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->launches += 1;
  h->last_kernel = "wavenet_fp32_persistent";
  h->t += T;
  return VQWN_OK;
}

#ifdef TF_DEBUG_MARKS
// development build: what the host-mapped debug area says after a failed launch
static void tf_debug_dump(vqwn_handle* h) {
  if (!h->dbg_host) return;
  const int* e = reinterpret_cast<const int*>(h->dbg_host + 4096);
  fprintf(stderr, "[tcf first timeout] code %d barrier index %d parity %d thread %d block %d\n", e[0],
          ((e[1] & 0x3ffff) - 1024 - (int)TF_OFF_BARS) / 8, e[2], e[3], e[4]);
  for (int c = 0; c < 112; ++c) {
    int any = 0;
    for (int w = 0; w < 12; ++w) any |= e[64 + (c * 12 + w) * 2 + 1];
    if (!any) continue;
    fprintf(stderr, "[tcf stuck] cta %3d:", c);
    for (int w = 0; w < 12; ++w) {
      const int a = e[64 + (c * 12 + w) * 2], pr = e[64 + (c * 12 + w) * 2 + 1];
      if (pr) fprintf(stderr, " w%d:bar%d/p%d", w, ((a & 0x3ffff) - 1024 - (int)TF_OFF_BARS) / 8, pr - 100);
    }
    fprintf(stderr, "\n");
  }
  for (int c = 0; c < 112; ++c) {
    const long long* m = h->dbg_host + 1024 + c * 12;
    long long any = 0;
    for (int w = 0; w < 12; ++w) any |= m[w];
    if (!any) continue;
    fprintf(stderr, "[tcf marks] cta %3d:", c);
    for (int w = 0; w < 12; ++w) fprintf(stderr, " %lld", m[w]);
    fprintf(stderr, "\n");
  }
}
#endif

int finish_timing(vqwn_handle* h) {
#ifdef TF_DEBUG_MARKS
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) tf_debug_dump(h);
#endif
  {
    const cudaError_t se = cudaStreamSynchronize(h->stream);
    if (se != cudaSuccess && strcmp(h->last_kernel, "wavenet_tcf_cluster") == 0 && h->tf_err_host && h->tf_err_host[0]) {
      const int* ev = h->tf_err_host;
      if (getenv("VQWN_DEBUG")) {
        for (int c = 0; c < 112; ++c) {
          int any = 0;
          for (int w = 0; w < 12; ++w) any |= ev[64 + (c * 12 + w) * 2 + 1];
          if (!any) continue;
          fprintf(stderr, "[tcf stuck] cta %3d:", c);
          for (int w = 0; w < 12; ++w) {
            const int a = ev[64 + (c * 12 + w) * 2], pr = ev[64 + (c * 12 + w) * 2 + 1];
            if (pr) fprintf(stderr, " w%d:bar%d/p%d", w, ((a & 0x3ffff) - 1024 - (int)TF_OFF_BARS) / 8, pr - 100);
          }
          fprintf(stderr, "\n");
        }
      }
      char msg[320];
      snprintf(msg, sizeof msg, "generation kernel stopped: wait timed out (barrier index %d, parity %d, thread %d, block %d): %s",
               ((ev[1] & 0x3ffff) - 1024 - (int)TF_OFF_BARS) / 8, ev[2], ev[3], ev[4], cudaGetErrorString(se));
      return fail(h, VQWN_ERR_CUDA, msg);
    }
    CK(h, se);
  }
  float ms = 0.f;
  CK(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_ms = ms;
  if (strncmp(h->last_kernel, "wavenet_", 8) == 0 && strcmp(h->last_kernel, "wavenet_tcf_cluster") != 0) {
    int ev[8] = {0};
    CK(h, cudaMemcpy(ev, h->gen_err, sizeof ev, cudaMemcpyDeviceToHost));
    const int e = ev[0];
    if (e) return fail(h, VQWN_ERR_CUDA, e == 2 ? "generation kernel: operand wait timed out" : (e == 4 ? "generation kernel: packet wait timed out" : "generation kernel: grid barrier timed out"));
  }
  if (h->profile && strncmp(h->last_kernel, "vq_tc", 5) == 0) {
    long long pf[32];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "[vqwn profile] vq_tc CTA0 epilogue-warp0 cycles: setup=%lld wait_z=%lld wait_acc=%lld pass1=%lld pass2=%lld decide=%lld output=%lld (kernel %.3f ms)\n",
              pf[16], pf[17], pf[18], pf[19], pf[20], pf[21], pf[22], ms);
  }
  if (h->profile && strcmp(h->last_kernel, "wavenet_fp32_cluster") == 0) {
    long long pf[24];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess) {
      const char* cls[3] = {"S1", "S2", "other"};
      for (int c = 0; c < 3; ++c)
        fprintf(stderr, "[vqwn profile] cluster CTA0 %s cycles: fir=%lld operand_wait=%lld contraction=%lld reduce=%lld math/draw=%lld push=%lld prefetch_issue=%lld recv_wait=%lld (kernel %.3f ms)\n",
                cls[c], pf[8 * c + 0], pf[8 * c + 1], pf[8 * c + 2], pf[8 * c + 7], pf[8 * c + 5], pf[8 * c + 3], pf[8 * c + 6], pf[8 * c + 4], ms);
    }
  }
  if (h->profile && strcmp(h->last_kernel, "wavenet_bf16_cluster") == 0) {
    long long pf[24];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess) {
      const char* cls[3] = {"S1", "S2", "other"};
      for (int c = 0; c < 3; ++c)
        fprintf(stderr, "[vqwn profile] bf16 CTA0 %s cycles: fir=%lld recv_wait=%lld operand_wait=%lld mma_chain=%lld tmem_ld=%lld ep_math=%lld ep_sync+queue=%lld push=%lld (kernel %.3f ms)\n",
                cls[c], pf[8 * c + 0], pf[8 * c + 4], pf[8 * c + 1], pf[8 * c + 2], pf[8 * c + 7], pf[8 * c + 6], pf[8 * c + 3], pf[8 * c + 5], ms);
    }
  }
  if (h->profile && strcmp(h->last_kernel, "wavenet_tcf_cluster") == 0) {
    long long pf[48];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "[vqwn profile] tcf CTA0 epilogue thread 0 cycles: step_start=%lld gate_acc_wait=%lld gate_epilogue=%lld gate_publish=%lld res_skip=%lld tail_skip=%lld post1=%lld post2=%lld draw=%lld sample_wait=%lld | inside res_skip: accB_wait=%lld tmem_read_zero_sync=%lld (kernel %.3f ms)\n",
              pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7], pf[8], pf[10], pf[9], pf[11], ms);
      const char* kn[7] = {"gate", "res+skip", "taps", "tail_skip", "post1", "post2", "next_taps"};
      fprintf(stderr, "[vqwn profile] tcf CTA0 issuing thread (warp 4) cycles, wait / issue per chain kind:");
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %s=%lld/%lld", kn[k], pf[32 + 2 * k], pf[32 + 2 * k + 1]);
      fprintf(stderr, " | step_start=%lld weight_chunk_wait=%lld operand_fence=%lld mma_issue=%lld chunk_commit=%lld chain_commits=%lld consume_total=%lld\n", pf[16], pf[25], pf[19], pf[20], pf[21], pf[22], pf[23]);
    }
    if (const char* tp = getenv("VQWN_TRACE_FILE")) {
      // one time step's event trace of CTA 0 (kernel: TF_TR): "warp event layer cycles" lines for tools/tcf_trace.py
      std::vector<long long> tr(12 * TF_TRACE_N);
      if (cudaMemcpy(tr.data(), h->prof + 256, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
        if (FILE* f = fopen(tp, "w")) {
          for (int w = 0; w < 12; ++w)
            for (int i = 0; i < TF_TRACE_N; ++i) {
              const long long v = tr[(size_t)w * TF_TRACE_N + i];
              if (v) fprintf(f, "%d %lld %lld %lld\n", w, (v >> 8) & 255, v & 255, v >> 16);
            }
          fclose(f);
        }
      }
    }
  }
  if (h->profile && strcmp(h->last_kernel, "wavenet_tc_cluster") == 0) {
    long long pf[48];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "[vqwn profile] tc bulk copies, cycles (issue / issue->landed): weight tile 32 KB layer 8: %lld / %lld, layer 20: %lld / %lld; tap 16 KB layer 8 (d=256): %lld / %lld, layer 20 (d=1): %lld / %lld\n",
              pf[40], pf[41], pf[44], pf[45], pf[42], pf[43], pf[46], pf[47]);
      fprintf(stderr, "[vqwn profile] tc CTA0 epilogue thread 0 cycles: step_start=%lld S1_acc_wait=%lld S1_epilogue=%lld S1_push=%lld S2_acc_wait=%lld S2_epilogue=%lld S2_queue+push=%lld post1=%lld post2=%lld draw=%lld sample_wait=%lld (kernel %.3f ms)\n",
              pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7], pf[8], pf[9], pf[10], ms);
      fprintf(stderr, "[vqwn profile] tc CTA0 MMA thread cycles: step_start=%lld S1_weight_wait=%lld S1_slices+chain=%lld tap1=%lld S2_sync+weight_wait=%lld S2_slices+chain=%lld tap2=%lld post=%lld tail_taps=%lld sample_wait=%lld\n",
              pf[16], pf[17], pf[18], pf[19], pf[20], pf[21], pf[22], pf[23], pf[24], pf[26]);
    }
  }
  if (h->profile && strcmp(h->last_kernel, "wavenet_fp32_persistent") == 0) {
    long long pf[8];
    if (cudaMemcpy(pf, h->prof, sizeof pf, cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "[vqwn profile] CTA0 cycles: barrier=%lld act_wait=%lld compute=%lld epilogue=%lld draw=%lld issue=%lld arrive=%lld prefetch=%lld (kernel %.3f ms)\n",
              pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7], ms);
  }
  return VQWN_OK;
}

int launch_vq(vqwn_handle* h, const float* z, long long n, long long* idx, float* out, int out_stride,
              const int* spk_idx, int spk_dim, int F) {
  const int K = h->K;
  const float* E = TP(h, "embedding/embedding");
  const float* spk = spk_dim > 0 ? TP(h, "speaker_embedding") : nullptr;
  const bool tensor_ok = (K == VT_K && h->D == VT_D);
  if ((h->vq_kernel == VQWN_VQ_TENSOR || h->vq_kernel == VQWN_VQ_TENSOR_BF16) && !tensor_ok)
    return fail(h, VQWN_ERR_NOTIMPL, "tensor-core VQ kernel needs k = 512 and latent_dim = 64");
  const bool use_tensor = tensor_ok && h->vq_kernel != VQWN_VQ_DIRECT && h->vq_kernel != VQWN_VQ_EXPANDED;
  if (use_tensor && !h->emax_valid) {
    vq_emax_kernel<<<1, (K + 31) / 32 * 32, 0, h->stream>>>(E, K, h->D, h->emax_dev);
    CK(h, cudaGetLastError());
    h->launches += 1;
    h->emax_valid = true;
  }
  CK(h, cudaEventRecord(h->ev0, h->stream));
  if (use_tensor) {
    const long long ntiles = (n + VT_TILE - 1) / VT_TILE;
    int grid = (int)(ntiles < (long long)h->num_sms ? ntiles : (long long)h->num_sms);
    if (grid < 1) grid = 1;
    CK(h, cudaMemsetAsync(h->vq_err, 0, sizeof(int), h->stream));
    if (h->vq_kernel != VQWN_VQ_TENSOR_BF16) {
      vq_tc_kernel<<<grid, VT_THREADS, VT_SMEM, h->stream>>>(z, E, n, idx, out, out_stride, spk, spk_idx, spk_dim, F,
                                                             h->emax_dev, h->vq_err, h->profile ? h->prof : nullptr, h->vq_out_code);
      h->last_kernel = "vq_tc_kernel";
    } else {
      vq_tc2_kernel<<<grid, VT_THREADS, V2_SMEM, h->stream>>>(z, E, n, idx, out, out_stride, spk, spk_idx, spk_dim, F,
                                                              h->emax_dev, h->vq_err, h->profile ? h->prof : nullptr, h->vq_out_code);
      h->last_kernel = "vq_tc2_kernel";
    }
  } else {
    // at least 128 threads: the kernel stages 4 vectors with VB*D/4 <= 64 threads and reduces with one warp per vector
    // (VB = 4 warps); threads beyond k hold no code (has_code guards)
    int threads = (K + 31) / 32 * 32;
    if (threads < 128) threads = 128;
    const long long nblocks = (n + 3) / 4;
    int grid = (int)((nblocks < (long long)h->num_sms * 1) ? nblocks : (long long)h->num_sms);
    if (grid < 1) grid = 1;
    if (h->D == 64)
      vq_direct_kernel<64, 4><<<grid, threads, 0, h->stream>>>(z, E, K, n, idx, out, out_stride, spk, spk_idx, spk_dim, F, h->vq_out_code, h->vq_kernel == VQWN_VQ_EXPANDED);
    else if (h->D == 32)
      vq_direct_kernel<32, 4><<<grid, threads, 0, h->stream>>>(z, E, K, n, idx, out, out_stride, spk, spk_idx, spk_dim, F, h->vq_out_code,
                                                               h->vq_kernel == VQWN_VQ_EXPANDED);
    else
      return fail(h, VQWN_ERR_NOTIMPL, "latent_dim must be 32 or 64");
    h->last_kernel = "vq_direct_kernel";
  }
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->launches += 1;
  return VQWN_OK;
}

int check_vq_error(vqwn_handle* h) {
  if (strncmp(h->last_kernel, "vq_tc", 5) != 0) return VQWN_OK;
  int e = 0;
  CK(h, cudaMemcpy(&e, h->vq_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return fail(h, VQWN_ERR_CUDA, "tensor-core VQ kernel: pipeline wait timed out");
  return VQWN_OK;
}

int check_tensor_ready(vqwn_handle* h, const char* name) {
  auto it = h->index.find(name);
  if (it == h->index.end()) return fail(h, VQWN_ERR_INVALID, std::string("unknown tensor: ") + name);
  if (!h->tensors[it->second].set) return fail(h, VQWN_ERR_STATE, std::string("tensor not set: ") + name);
  return VQWN_OK;
}

}  // namespace

#define ENTER(h)                                                              \
  if (!(h)) return fail(nullptr, VQWN_ERR_INVALID, "null handle");            \
  { cudaError_t e0_ = cudaSetDevice((h)->device);                             \
    if (e0_ != cudaSuccess) return fail((h), VQWN_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e0_)); }

extern "C" {

const char* vqwn_version(void) { return "vqwn 0.1 (sm_100a)"; }

const char* vqwn_last_error(const vqwn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vqwn_create(const vqwn_config* cfg, int device, int max_batch, vqwn_handle** out) {
  if (!cfg || !out) return fail(nullptr, VQWN_ERR_INVALID, "null argument");
  *out = nullptr;
  const vqwn_config& c = *cfg;
  if (c.num_layers < 1 || c.num_layers > VQWN_MAX_LAYERS) return fail(nullptr, VQWN_ERR_INVALID, "num_layers out of range");
  if (c.kernel_size != 3) return fail(nullptr, VQWN_ERR_NOTIMPL, "only kernel_size 3 is implemented");
  if (c.quantization_channels != 256)
    return fail(nullptr, VQWN_ERR_NOTIMPL, "only quantization_channels 256 works end to end (reference utils.py:41, wavenet.py:113)");
  if (c.pre_filters != c.residual_filters) return fail(nullptr, VQWN_ERR_INVALID, "preprocess.filters must equal residual_filters");
  if (c.dilation_filters != c.residual_filters) return fail(nullptr, VQWN_ERR_INVALID, "dilation_filters must equal residual_filters");
  const int Ccond = c.latent_dim + c.speaker_dim;
  if (c.residual_filters % 128 || c.skip_filters % 128 || Ccond % 128 || c.dilation_filters % 128)
    return fail(nullptr, VQWN_ERR_INVALID, "channel counts (residual, skip, gate, latent+speaker) must be multiples of 128");
  if ((c.residual_filters & (c.residual_filters - 1)) || (c.skip_filters & (c.skip_filters - 1)))
    return fail(nullptr, VQWN_ERR_NOTIMPL, "residual_filters and skip_filters must be powers of two");
  if (c.pre_kernel_size < 1 || c.pre_kernel_size > 64) return fail(nullptr, VQWN_ERR_INVALID, "preprocess.kernel_size out of range");
  if (c.use_vq && (c.k < 1 || c.k > 512)) return fail(nullptr, VQWN_ERR_NOTIMPL, "k must be <= 512");
  if (c.latent_dim != 32 && c.latent_dim != 64) return fail(nullptr, VQWN_ERR_NOTIMPL, "latent_dim must be 32 or 64");
  if (c.encoder != VQWN_ENCODER_NONE && c.encoder != VQWN_ENCODER_64 && c.encoder != VQWN_ENCODER_MAGENTA &&
      c.encoder != VQWN_ENCODER_2019)
    return fail(nullptr, VQWN_ERR_NOTIMPL, "encoders on the device: Encoder_64 (64), Encoder_Magenta (1), or none (0)");
  if (c.encoder != VQWN_ENCODER_NONE && c.latent_dim != 64)
    return fail(nullptr, VQWN_ERR_NOTIMPL, "the device encoders need latent_dim = 64");
  if (c.num_cycle_layers < 1) return fail(nullptr, VQWN_ERR_INVALID, "num_cycle_layers must be >= 1");
  for (int i = 0; i < c.num_layers; ++i)
    if (c.dilations[i] < 1) return fail(nullptr, VQWN_ERR_INVALID, "dilation must be >= 1");
  if (max_batch < 1) return fail(nullptr, VQWN_ERR_INVALID, "max_batch must be >= 1");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, VQWN_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, VQWN_ERR_INVALID, "device index out of range");

  vqwn_handle* h = new vqwn_handle();
  h->cfg = c;
  h->device = device;
  h->max_batch = max_batch;
  h->Bp_max = (max_batch + FP32_TB - 1) / FP32_TB * FP32_TB;
  h->L = c.num_layers; h->R = c.residual_filters; h->G = c.dilation_filters; h->S = c.skip_filters;
  h->Q = c.quantization_channels; h->C = Ccond; h->PK = c.pre_kernel_size; h->K = c.k; h->D = c.latent_dim;
  h->SPK = c.speaker_dim;

#define CKC(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      std::string m_ = std::string(#call) + " failed: " + cudaGetErrorString(e_);         \
      vqwn_destroy(h);                                                                    \
      return fail(nullptr, e_ == cudaErrorMemoryAllocation ? VQWN_ERR_NOMEM : VQWN_ERR_CUDA, m_); \
    }                                                                                     \
  } while (0)

  CKC(cudaSetDevice(device));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    vqwn_destroy(h);
    return fail(nullptr, VQWN_ERR_CUDA, "device is not sm_100 (Blackwell); this library is built for sm_100a only");
  }
  h->num_sms = prop.multiProcessorCount;
  CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));

  // tensor registry (reference variable names)
  const int R = h->R, G = h->G, S = h->S, Q = h->Q, C = h->C;
  add_tensor(h, "embedding/embedding", {c.k, c.latent_dim}, c.use_vq != 0);
  add_tensor(h, "speaker_embedding", {c.num_speakers > 0 ? c.num_speakers : 1, c.speaker_dim > 0 ? c.speaker_dim : 1}, c.speaker_dim > 0);
  add_tensor(h, "decoder/preprocess/kernel", {h->PK, 1, R});
  add_tensor(h, "decoder/preprocess/bias", {R});
  add_tensor(h, "decoder/skip/kernel", {1, R, S});
  add_tensor(h, "decoder/skip/bias", {S});
  for (int l = 0; l < h->L; ++l) {
    const std::string sc = layer_scope(c, l);
    add_tensor(h, sc + "/gated/kernel", {3, R, 2 * G});
    add_tensor(h, sc + "/gated/bias", {2 * G});
    add_tensor(h, sc + "/gated/local_condition/kernel", {1, C, 2 * G});
    add_tensor(h, sc + "/skip/kernel", {1, G, S});
    add_tensor(h, sc + "/skip/bias", {S});
    add_tensor(h, sc + "/residual/kernel", {1, G, R});
    add_tensor(h, sc + "/residual/bias", {R});
  }
  add_tensor(h, "decoder/postprocess1/kernel", {1, S, S});
  add_tensor(h, "decoder/postprocess1/bias", {S});
  add_tensor(h, "decoder/postprocess1/local_condition/kernel", {1, C, S});
  add_tensor(h, "decoder/postprocess2/kernel", {1, S, Q});
  add_tensor(h, "decoder/postprocess2/bias", {Q});
  if (c.encoder == 64) {
    int cin = 1;
    for (int i = 0; i < 7; ++i) {
      const std::string sfx = i == 0 ? "" : "_" + std::to_string(i);
      const int cout = i < 6 ? 768 : c.latent_dim;
      const int k = i < 6 ? 5 : 1;
      add_tensor(h, "encoder/conv1d" + sfx + "/kernel", {k, cin, cout}, false);
      add_tensor(h, "encoder/conv1d" + sfx + "/bias", {cout}, false);
      add_tensor(h, "encoder/batch_normalization" + sfx + "/gamma", {cout}, false);
      add_tensor(h, "encoder/batch_normalization" + sfx + "/beta", {cout}, false);
      add_tensor(h, "encoder/batch_normalization" + sfx + "/moving_mean", {cout}, false);
      add_tensor(h, "encoder/batch_normalization" + sfx + "/moving_variance", {cout}, false);
      cin = cout;
    }
  }
  if (c.encoder == VQWN_ENCODER_MAGENTA) {
    // Encoder/encoder.py:37-64: conv1d_v2 variables ("kernel" [k,in,out], "bias") in their scopes under "encoder/"
    const int MC = 128, MK = 5;
    add_tensor(h, "encoder/preprocess/kernel", {MK, 1, MC}, false);
    add_tensor(h, "encoder/preprocess/bias", {MC}, false);
    for (int i = 0; i < 6; ++i) {
      const std::string sc = "encoder/cycle_1/layer_" + std::to_string(i + 1);
      add_tensor(h, sc + "/dilated/kernel", {1, MC, MC}, false);
      add_tensor(h, sc + "/dilated/bias", {MC}, false);
      add_tensor(h, sc + "/gate/kernel", {MK, MC, MC}, false);
      add_tensor(h, sc + "/gate/bias", {MC}, false);
      add_tensor(h, sc + "/filter/kernel", {MK, MC, MC}, false);
      add_tensor(h, sc + "/filter/bias", {MC}, false);
      add_tensor(h, sc + "/residual/kernel", {1, MC, MC}, false);
      add_tensor(h, sc + "/residual/bias", {MC}, false);
    }
    add_tensor(h, "encoder/postprocess/kernel", {1, MC, c.latent_dim}, false);
    add_tensor(h, "encoder/postprocess/bias", {c.latent_dim}, false);
  }
  if (c.encoder == VQWN_ENCODER_2019) {
    // Encoder/encoder.py:75-96: ten unnamed keras Conv1D layers in creation order
    const int shp[10][3] = {{3, MFCC_COEFS, 768}, {3, 768, 768}, {4, 768, 768}, {3, 768, 768}, {3, 768, 768}, {3, 768, 768},
                            {3, 768, 768}, {3, 768, 768}, {3, 768, 768}, {1, 768, c.latent_dim}};
    for (int i = 0; i < 10; ++i) {
      const std::string sfx = i == 0 ? "" : "_" + std::to_string(i);
      add_tensor(h, "encoder/conv1d" + sfx + "/kernel", {shp[i][0], shp[i][1], shp[i][2]}, false);
      add_tensor(h, "encoder/conv1d" + sfx + "/bias", {shp[i][2]}, false);
    }
  }
  add_tensor(h, "lut/mu_law_decode", {Q + 1}, false);
  add_tensor(h, "lut/mu_law_encode", {Q + 1}, false);
  for (auto& s : h->tensors) CKC(cudaMalloc(&s.dev, s.numel * sizeof(float)));
  {
    std::vector<float> dec, enc;
    default_luts(Q, dec, enc);
    CKC(cudaMemcpy(TP(h, "lut/mu_law_decode"), dec.data(), dec.size() * sizeof(float), cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(TP(h, "lut/mu_law_encode"), enc.data(), enc.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->tensors[h->index["lut/mu_law_decode"]].set = true;
    h->tensors[h->index["lut/mu_law_encode"]].set = true;
  }

  // packed weights
  h->w1.resize(h->L); h->b1.resize(h->L); h->w2.resize(h->L); h->b2.resize(h->L);
  h->layers_host.resize(h->L);
  for (int l = 0; l < h->L; ++l) {
    CKC(cudaMalloc(&h->w1[l], (size_t)(3 * R + C) * 2 * G * sizeof(float)));
    CKC(cudaMalloc(&h->b1[l], (size_t)2 * G * sizeof(float)));
    CKC(cudaMalloc(&h->w2[l], (size_t)G * (R + S) * sizeof(float)));
    CKC(cudaMalloc(&h->b2[l], (size_t)(R + S) * sizeof(float)));
    h->layers_host[l] = LayerDev{nullptr, h->b1[l], nullptr, h->b2[l], nullptr, c.dilations[l], 0};
  }
  CKC(cudaMalloc(&h->post1_w, (size_t)(S + C) * S * sizeof(float)));
  CKC(cudaMalloc(&h->layers_dev, sizeof(LayerDev) * h->L));
  CKC(cudaMalloc(&h->enc_lut, (Q + 1) * sizeof(float)));
  CKC(cudaMalloc(&h->emax_dev, sizeof(float)));
  CKC(cudaMalloc(&h->vq_err, sizeof(int)));
  CKC(cudaFuncSetAttribute((const void*)vq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VT_SMEM));
  CKC(cudaFuncSetAttribute((const void*)vq_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V2_SMEM));
  CKC(cudaMalloc(&h->dec_lut, (Q + 1) * sizeof(float)));

  // state
  const size_t Bp = h->Bp_max;
  size_t dsum = 0;
  for (int l = 0; l < h->L; ++l) dsum += (size_t)c.dilations[l];
  h->ring_floats = 2 * dsum * Bp * R;
  // the split-bf16 kernel spreads a batch over up to 7 clusters of 16 ring rows each
  h->tc_ok = R == TC_R && G == TC_G && S == TC_S && Q == TC_Q && C == TC_C && h->PK == TC_PK && c.kernel_size == 3 &&
             h->L >= 2 && h->L <= TC_MAXL;
  if (h->tc_ok) {
    CKC(cudaFuncSetAttribute((const void*)wavenet_tc_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    CKC(cudaFuncSetAttribute((const void*)wavenet_tc_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(TC_CS); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TC_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, (const void*)wavenet_tc_cluster, &cfg) != cudaSuccess) { nc = 0; (void)cudaGetLastError(); }
    h->tc_max_clusters = nc;
    if (nc < 1) h->tc_ok = false;
  }
  if (h->tc_ok) {
    size_t worst = 0;
    for (int b = 1; b <= max_batch; ++b) { const size_t x = tc_ring_bytes(h, b); if (x > worst) worst = x; }
    if (worst > h->ring_floats * sizeof(float)) h->ring_floats = (worst + 3) / 4;
    h->wtc_bytes = (size_t)h->L * TC_CS * TC_LAYER_BYTES + (size_t)TC_CS * (TC_WP1 + TC_WP2);
    CKC(cudaMalloc(&h->wtc, h->wtc_bytes));
    // every cluster streams all weight tiles once per time step (cyclic reuse, 72 MB): without help they fall out of the
    // L2 between two steps and every tile pays DRAM latency (measured 83 MB of DRAM reads per step).  Reserve the
    // persisting part of the L2 for them; launch_tc attaches the access-policy window.
    {
      int max_persist = 0;
      if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess && max_persist > 0) {
        size_t want = (size_t)TF_CS * tf_stream_bytes(h->L) + (4u << 20);      // the larger of the two kernels' weight sets
        if (want > (size_t)max_persist) want = (size_t)max_persist;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) h->tc_persist_bytes = want;
        else (void)cudaGetLastError();
        int max_win = 0;
        if (cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device) == cudaSuccess && max_win > 0)
          h->tc_max_window = (size_t)max_win;
        else { (void)cudaGetLastError(); h->tc_persist_bytes = 0; }
      } else (void)cudaGetLastError();
    }
    CKC(cudaMalloc(&h->tc_skf_k, (size_t)TC_PK * TC_S * sizeof(float)));
    CKC(cudaMalloc(&h->tc_skf_b, (size_t)TC_S * sizeof(float)));
    CKC(cudaMalloc(&h->tc_ctab, (size_t)h->tc_max_clusters * TC_CS * (h->L + 1) * 512 * sizeof(float)));
    {
      CKC(cudaFuncSetAttribute((const void*)wavenet_tcf_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM));
      CKC(cudaFuncSetAttribute((const void*)wavenet_tcf_cluster<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      CKC(cudaFuncSetAttribute((const void*)wavenet_tcf_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM));
      CKC(cudaFuncSetAttribute((const void*)wavenet_tcf_cluster<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      const char* tk = getenv("VQWN_TC_KERNEL");
      h->tf_use = !(tk && strcmp(tk, "v1") == 0);
      if (const char* rp = getenv("VQWN_TC_REPRODUCIBLE")) h->tc_reproducible = (atoi(rp) != 0);
      h->wtf_bytes = (size_t)TF_CS * tf_stream_bytes(h->L);
      CKC(cudaMalloc(&h->wtf, h->wtf_bytes));
      CKC(cudaMalloc(&h->tf_ptmp, (size_t)G * 2 * G * sizeof(float)));
      CKC(cudaMalloc(&h->tf_b1adj, (size_t)h->L * 2 * G * sizeof(float)));
      CKC(cudaMalloc(&h->tf_gstage, (size_t)h->tc_max_clusters * TF_GSTAGE));
      CKC(cudaMemset(h->tf_gstage, 0, (size_t)h->tc_max_clusters * TF_GSTAGE));
      CKC(cudaMalloc(&h->tf_layers_dev, sizeof(TfLayerDev) * h->L));
      h->tf_layers_host.resize(h->L);
      for (int l = 0; l < h->L; ++l) {
        TfLayerDev ld;
        memset(&ld, 0, sizeof ld);
        ld.wlc = h->w1[l] + (size_t)3 * R * 2 * G;
        ld.b1 = h->tf_b1adj + (size_t)l * 2 * G;
        ld.bres = h->b2[l];
        ld.ring = nullptr;
        ld.d = c.dilations[l];
        h->tf_layers_host[l] = ld;
      }
    }
    CKC(cudaMalloc(&h->tc_gstage, (size_t)h->tc_max_clusters * TC_GSTAGE));
    CKC(cudaMemset(h->tc_gstage, 0, (size_t)h->tc_max_clusters * TC_GSTAGE));
    CKC(cudaMalloc(&h->tc_b2_ptrs, sizeof(const float*) * h->L));
    CKC(cudaMalloc(&h->tc_layers_dev, sizeof(TcLayerDev) * h->L));
    h->tc_layers_host.resize(h->L);
    for (int l = 0; l < h->L; ++l)
      h->tc_layers_host[l] = TcLayerDev{reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<uint8_t*>(h->wtc) + (size_t)l * TC_CS * TC_LAYER_BYTES),
                                        h->w1[l] + (size_t)3 * R * 2 * G, h->b1[l], h->b2[l], nullptr, c.dilations[l], 0};
  }
  CKC(cudaMalloc(&h->ring_base, h->ring_floats * sizeof(float)));
  CKC(cudaMalloc(&h->u_hist, Bp * h->PK * sizeof(float)));
  CKC(cudaMalloc(&h->cur, Bp * R * sizeof(float)));
  CKC(cudaMalloc(&h->g, Bp * G * sizeof(float)));
  CKC(cudaMalloc(&h->skip, Bp * S * sizeof(float)));
  CKC(cudaMalloc(&h->n1, Bp * S * sizeof(float)));
  CKC(cudaMalloc(&h->logits, Bp * Q * sizeof(float)));
  CKC(cudaMalloc(&h->barrier, 32 * sizeof(unsigned long long)));
  CKC(cudaMalloc(&h->prof, (256 + 12 * TF_TRACE_N) * sizeof(long long)));       // counters + one step's event trace (VQWN_PROFILE=1)
  CKC(cudaMemset(h->prof, 0, (256 + 12 * TF_TRACE_N) * sizeof(long long)));
  CKC(cudaMemset(h->prof, 0, 256 * sizeof(long long)));
  h->profile = getenv("VQWN_PROFILE") != nullptr;

  // tile-major weight block
  {
    size_t off = 0;
    h->off_skip0t = off; off += (size_t)R * S;
    h->off_w1t.resize(h->L); h->off_w2t.resize(h->L);
    for (int l = 0; l < h->L; ++l) {
      h->off_w1t[l] = off; off += (size_t)(3 * R + C) * 2 * G;
      h->off_w2t[l] = off; off += (size_t)G * (R + S);
    }
    h->off_post1t = off; off += (size_t)(S + C) * S;
    h->off_post2t = off; off += (size_t)S * Q;
    h->wtiles_floats = off;
    CKC(cudaMalloc(&h->wtiles, off * sizeof(float)));
    for (int l = 0; l < h->L; ++l) {
      h->layers_host[l].w1t = h->wtiles + h->off_w1t[l];
      h->layers_host[l].w2t = h->wtiles + h->off_w2t[l];
    }
  }
  // cluster kernel: geometry constraints, weight block, shared-memory sizes, schedulable clusters
  {
    const bool pow2 = (R & (R - 1)) == 0 && (S & (S - 1)) == 0 && (G & (G - 1)) == 0;
    const bool div = R % CL_CS == 0 && S % CL_CS == 0 && G % CL_CS == 0 && Q % CL_CS == 0 &&
                     (G / CL_CS) % 4 == 0 && (R / CL_CS) % 4 == 0 && (S / CL_CS) % 4 == 0 && (Q / CL_CS) % 4 == 0 &&
                     C % 4 == 0 && Q <= 256 && Q % 32 == 0 &&
                     // power-of-two slices (index masks), tiles of at most 512 outputs, one gate output per thread
                     ((2 * G / CL_CS) & (2 * G / CL_CS - 1)) == 0 && ((S / CL_CS) & (S / CL_CS - 1)) == 0 &&
                     ((Q / CL_CS) & (Q / CL_CS - 1)) == 0 && CL_MAX_MS * (R / CL_CS + S / CL_CS) <= 512 &&
                     CL_MAX_MS * (G / CL_CS) <= 256 && CL_MAX_MS * (S / CL_CS) / 4 <= 256;
    h->cl_ok = pow2 && div && S <= 2 * R && S <= G + R && h->L <= 64;
    if (h->cl_ok) {
      const int NSK = S / CL_CS, NC1 = 2 * G / CL_CS, NC2 = R / CL_CS + NSK, NQ = Q / CL_CS;
      h->cl_w1_floats = (3 * R + C) * NC1;
      if ((S + C) * NSK > h->cl_w1_floats) h->cl_w1_floats = (S + C) * NSK;
      h->cl_w2_floats = G * NC2;
      if (S * NQ > h->cl_w2_floats) h->cl_w2_floats = S * NQ;
      if (R * NSK > h->cl_w2_floats) h->cl_w2_floats = R * NSK;
      if (cl_smem_bytes(h, CL_MAX_MS) + 4096 > (size_t)prop.sharedMemPerBlockOptin) h->cl_ok = false;   // + static table
    }
    if (h->cl_ok) {
      CKC(cudaMalloc(&h->wcl, h->wtiles_floats * sizeof(float)));
      CKC(cudaMalloc(&h->cl_layers_dev, sizeof(ClLayerDev) * h->L));
      h->cl_layers_host.resize(h->L);
      for (int l = 0; l < h->L; ++l)
        h->cl_layers_host[l] = ClLayerDev{h->wcl + h->off_w1t[l], h->b1[l], h->wcl + h->off_w2t[l], h->b2[l], nullptr,
                                          c.dilations[l], 0};
      h->cl_max_clusters = 1 << 30;
      for (int ms = 2; ms <= CL_MAX_MS; ms += 2) {
        cl_kernel_t k = cl_kernel_for(ms);
        CKC(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cl_smem_bytes(h, ms)));
        CKC(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(CL_CS); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = cl_smem_bytes(h, ms);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, (const void*)k, &cfg) != cudaSuccess) { nc = 0; (void)cudaGetLastError(); }
        if (nc < h->cl_max_clusters) h->cl_max_clusters = nc;
      }
      if (h->cl_max_clusters < 1) h->cl_ok = false;
    }
  }
  // bf16 tensor-core cluster kernel: reference default geometry only
  h->bc_ok = R == BC_R && G == BC_G && S == BC_S && Q == BC_Q && C == BC_C && h->PK == BC_PK && c.kernel_size == 3 && h->L <= 64;
  if (h->bc_ok) {
    CKC(cudaMalloc(&h->wbc, h->wtiles_floats * sizeof(__nv_bfloat16)));
    CKC(cudaMalloc(&h->bc_layers_dev, sizeof(BcLayerDev) * h->L));
    h->bc_layers_host.resize(h->L);
    for (int l = 0; l < h->L; ++l)
      h->bc_layers_host[l] = BcLayerDev{h->wbc + h->off_w1t[l], h->b1[l], h->wbc + h->off_w2t[l], h->b2[l], nullptr,
                                        c.dilations[l], 0};
    CKC(cudaFuncSetAttribute((const void*)wavenet_bf16_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BC_SMEM));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(BC_CS); cfg.blockDim = dim3(BC_THREADS); cfg.dynamicSmemBytes = BC_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = BC_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, (const void*)wavenet_bf16_cluster, &cfg) != cudaSuccess) { nc = 0; (void)cudaGetLastError(); }
    h->bc_max_clusters = nc;
    if (nc < 1) h->bc_ok = false;
  }
  CKC(cudaMalloc(&h->gen_err, 8 * sizeof(int)));
  CKC(cudaMemset(h->gen_err, 0, 8 * sizeof(int)));
  CKC(cudaHostAlloc(&h->tf_err_host, 4096 * sizeof(int), cudaHostAllocMapped));
  memset(h->tf_err_host, 0, 4096 * sizeof(int));
  CKC(cudaHostGetDevicePointer(&h->tf_err_dev, h->tf_err_host, 0));
  if (const char* gk = getenv("VQWN_GEN_KERNEL")) h->gen_kernel = (strcmp(gk, "barrier") == 0) ? 1 : (strcmp(gk, "cluster") == 0 ? 3 : 0);
  h->actA_floats = FP32_TB * 3 * R;                       // gated conv: current | t-d | t-2d segments
  if (FP32_TB * S > h->actA_floats) h->actA_floats = FP32_TB * S;   // post1: relu(skip)
  h->actB_floats = FP32_TB * S;                           // post2: relu(n1)
  if (FP32_TB * R > h->actB_floats) h->actB_floats = FP32_TB * R;
  if (FP32_TB * G > h->actB_floats) h->actB_floats = FP32_TB * G;
  h->wfloatsA = (3 * R + C) * 16;
  if ((S + C) * 16 > h->wfloatsA) h->wfloatsA = (S + C) * 16;
  h->wfloatsB = G * 32;
  if (S * 16 > h->wfloatsB) h->wfloatsB = S * 16;
  if (R * 16 > h->wfloatsB) h->wfloatsB = R * 16;
  h->smem_fp32 = ((size_t)h->wfloatsA + h->wfloatsB + (size_t)h->actA_floats + h->actB_floats + (size_t)FP32_TB * C +
                  FP32_RED_FLOATS + (size_t)FP32_TB * h->PK + (size_t)FP32_WARPS * Q) * sizeof(float) + 64;
  if (h->smem_fp32 > (size_t)prop.sharedMemPerBlockOptin) {
    vqwn_destroy(h);
    return fail(nullptr, VQWN_ERR_INVALID, "configuration needs more shared memory than the device offers");
  }
  CKC(cudaFuncSetAttribute((const void*)wavenet_fp32_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_fp32));
  int occ = 0;
  CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)wavenet_fp32_persistent, FP32_THREADS, h->smem_fp32));
  if (occ < 1) {
    vqwn_destroy(h);
    return fail(nullptr, VQWN_ERR_CUDA, "persistent kernel cannot be made co-resident");
  }
#undef CKC
  *out = h;
  return VQWN_OK;
}

int vqwn_destroy(vqwn_handle* h) {
  if (!h) return VQWN_OK;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  for (auto& s : h->tensors) if (s.dev) cudaFree(s.dev);
  for (auto p : h->w1) if (p) cudaFree(p);
  for (auto p : h->b1) if (p) cudaFree(p);
  for (auto p : h->w2) if (p) cudaFree(p);
  for (auto p : h->b2) if (p) cudaFree(p);
  void* singles[] = {h->post1_w, h->layers_dev, h->enc_lut, h->dec_lut, h->ring_base, h->u_hist, h->cur, h->g,
                     h->skip, h->n1, h->logits, h->barrier, h->prof, h->emax_dev, h->vq_err, h->wtiles, h->gen_err, h->wcl, h->cl_layers_dev, h->wbc, h->bc_layers_dev,
                     h->wtc, h->tc_skf_k, h->tc_skf_b, h->tc_ctab, h->tc_gstage, (void*)h->tc_b2_ptrs, h->tc_layers_dev,
                     h->wtf, h->tf_ptmp, h->tf_b1adj, h->tf_gstage, h->tf_layers_dev};
  for (void* p : singles) if (p) cudaFree(p);
  DevBuf* bufs[] = {&h->cond_res, &h->uni_res, &h->audio_res, &h->idx_res, &h->logits_res, &h->x_res, &h->small_a,
                    &h->small_b, &h->small_c, &h->small_d, &h->vq_z, &h->vq_idx, &h->vq_out, &h->spk_idx,
                    &h->enc_x, &h->enc_a, &h->enc_b, &h->enc_fold, &h->enc_z, &h->enc_c, &h->enc_d, &h->enc_e, &h->enc_f};
  for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
  if (h->tf_err_host) cudaFreeHost(h->tf_err_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return VQWN_OK;
}

int vqwn_set_stream(vqwn_handle* h, void* cuda_stream) {
  ENTER(h);
  CK(h, cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return VQWN_OK;
}

int vqwn_set_reproducible(vqwn_handle* h, int on) {
  if (!h) return VQWN_ERR_INVALID;
  h->tc_reproducible = (on != 0);
  return VQWN_OK;
}

int vqwn_set_precision(vqwn_handle* h, int precision) {
  ENTER(h);
  // the dilation-queue layout differs between the two paths: a change of precision invalidates the queue state, the
  // step API then asks for vqwn_reset (vqwn_generate / vqwn_teacher_forced reset by themselves)
  if (precision == VQWN_PREC_FP32) {
    if (h->precision != precision) h->B = 0;
    h->precision = precision;
    return VQWN_OK;
  }
  if (precision == VQWN_PREC_BF16) {
    if (!h->bc_ok)
      return fail(h, VQWN_ERR_NOTIMPL, "bf16 tensor-core path is built for the reference's default WaveNet geometry only");
    if (h->precision != precision) h->B = 0;
    h->precision = precision;
    return VQWN_OK;
  }
  if (precision == VQWN_PREC_TC) {
    if (!h->tc_ok)
      return fail(h, VQWN_ERR_NOTIMPL, "split-bf16 tensor-core path is built for the reference's default WaveNet geometry only");
    if (h->precision != precision) h->B = 0;
    h->precision = precision;
    return VQWN_OK;
  }
  return fail(h, VQWN_ERR_INVALID, "unknown precision id");
}

int vqwn_set_stream_offset(vqwn_handle* h, int64_t offset) {
  ENTER(h);
  if (offset < 0 || offset > 0x7fffffff) return fail(h, VQWN_ERR_INVALID, "stream offset out of range");
  h->stream_offset = (int)offset;
  return VQWN_OK;
}

int vqwn_set_vq_kernel(vqwn_handle* h, int kernel) {
  ENTER(h);
  if (kernel < VQWN_VQ_AUTO || kernel > VQWN_VQ_EXPANDED) return fail(h, VQWN_ERR_INVALID, "unknown VQ kernel id");
  h->vq_kernel = kernel;
  return VQWN_OK;
}

int vqwn_set_vq_output(vqwn_handle* h, int output) {
  ENTER(h);
  if (output != VQWN_VQ_OUT_STRAIGHT_THROUGH && output != VQWN_VQ_OUT_CODE) return fail(h, VQWN_ERR_INVALID, "unknown VQ output id");
  h->vq_out_code = output == VQWN_VQ_OUT_CODE ? 1 : 0;
  return VQWN_OK;
}

int vqwn_set_tensor(vqwn_handle* h, const char* tf_name, const float* host, const int64_t* shape, int ndim) {
  ENTER(h);
  if (!tf_name || !host || !shape) return fail(h, VQWN_ERR_INVALID, "null argument");
  auto it = h->index.find(tf_name);
  if (it == h->index.end()) return fail(h, VQWN_ERR_INVALID, std::string("unknown tensor: ") + tf_name);
  TensorSlot& s = h->tensors[it->second];
  bool ok = (ndim == (int)s.shape.size());
  for (int i = 0; ok && i < ndim; ++i) ok = (shape[i] == s.shape[i]);
  if (!ok) {
    std::string m = std::string("shape mismatch for ") + tf_name + ": expected [";
    for (size_t i = 0; i < s.shape.size(); ++i) m += (i ? "," : "") + std::to_string(s.shape[i]);
    m += "]";
    return fail(h, VQWN_ERR_INVALID, m);
  }
  CK(h, cudaMemcpyAsync(s.dev, host, s.numel * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  s.set = true;
  h->packed = false;
  if (it->first == "embedding/embedding") h->emax_valid = false;
  return VQWN_OK;
}

int vqwn_get_tensor(vqwn_handle* h, const char* tf_name, float* host_out, int64_t capacity) {
  ENTER(h);
  if (!tf_name || !host_out) return fail(h, VQWN_ERR_INVALID, "null argument");
  int rc = check_tensor_ready(h, tf_name);
  if (rc) return rc;
  TensorSlot& s = h->tensors[h->index[tf_name]];
  if ((int64_t)s.numel > capacity) return fail(h, VQWN_ERR_INVALID, "output buffer too small");
  CK(h, cudaMemcpyAsync(host_out, s.dev, s.numel * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return VQWN_OK;
}

int vqwn_num_tensors(const vqwn_handle* h) { return h ? (int)h->tensors.size() : 0; }

int vqwn_tensor_info(const vqwn_handle* h, int i, char* name_out, int name_cap, int64_t* shape_out, int* ndim_out,
                     int* is_set) {
  if (!h || i < 0 || i >= (int)h->tensors.size()) return VQWN_ERR_INVALID;
  const TensorSlot& s = h->tensors[i];
  if (name_out && name_cap > 0) { strncpy(name_out, s.name.c_str(), name_cap - 1); name_out[name_cap - 1] = 0; }
  if (shape_out) for (size_t k = 0; k < s.shape.size() && k < 4; ++k) shape_out[k] = s.shape[k];
  if (ndim_out) *ndim_out = (int)s.shape.size();
  if (is_set) *is_set = s.set ? 1 : 0;
  return VQWN_OK;
}

// ---------------------------------------------------------------------------------- encoder
// Encoder_Magenta (Encoder/encoder.py:37-64) on the device: shift_right + mu_law_encode, causal k=5 preprocess conv,
// 6 x [1x1 stride-2 conv, causal dilated k=5 gate / filter convs, tanh*sigmoid, 1x1 residual], 1x1 postprocess.
static int encode_magenta(vqwn_handle* h, const float* x, int B, int64_t T, float* z_e_out) {
  int rc;
  const int MC = 128, MK = 5, D = h->D;
  const int dil[6] = {1, 2, 4, 8, 16, 16};
  const char* parts[] = {"/kernel", "/bias"};
  for (const char* pt : parts) {
    if ((rc = check_tensor_ready(h, (std::string("encoder/preprocess") + pt).c_str()))) return rc;
    if ((rc = check_tensor_ready(h, (std::string("encoder/postprocess") + pt).c_str()))) return rc;
    for (int i = 0; i < 6; ++i)
      for (const char* sub : {"/dilated", "/gate", "/filter", "/residual"})
        if ((rc = check_tensor_ready(h, ("encoder/cycle_1/layer_" + std::to_string(i + 1) + sub + pt).c_str()))) return rc;
  }
  // identity scale / shift for the shared conv epilogue
  if ((rc = ensure(h, h->enc_fold, (size_t)7 * 2 * 768 * sizeof(float)))) return rc;
  {
    std::vector<float> ident(2 * 768, 0.f);
    for (int i = 0; i < 768; ++i) ident[i] = 1.f;
    CK(h, cudaMemcpyAsync(h->enc_fold.p, ident.data(), ident.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  const float* one = (const float*)h->enc_fold.p;
  const float* zero = one + 768;
  // per stream: u [T], en [T,128] (ping), and at half rate d, g, f, gated/res [T/2,128] each
  const size_t per_stream = (size_t)T * sizeof(float) + (size_t)T * MC * sizeof(float) + 4 * (size_t)(T / 2) * MC * sizeof(float);
  size_t gsz = ((size_t)1 << 30) / per_stream;
  if (gsz < 1) gsz = 1;
  const int group = gsz > (size_t)B ? B : (int)gsz;
  if ((rc = ensure(h, h->enc_x, (size_t)group * T * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_a, (size_t)group * T * sizeof(float)))) return rc;               // u
  if ((rc = ensure(h, h->enc_b, (size_t)group * T * MC * sizeof(float)))) return rc;          // en
  if ((rc = ensure(h, h->enc_c, (size_t)group * (T / 2) * MC * sizeof(float)))) return rc;    // d
  if ((rc = ensure(h, h->enc_d, (size_t)group * (T / 2) * MC * sizeof(float)))) return rc;    // gate conv
  if ((rc = ensure(h, h->enc_e, (size_t)group * (T / 2) * MC * sizeof(float)))) return rc;    // filter conv
  if ((rc = ensure(h, h->enc_f, (size_t)group * (T / 2) * MC * sizeof(float)))) return rc;    // gated, then residual conv
  if ((rc = ensure(h, h->enc_z, (size_t)group * (T / 64) * D * sizeof(float)))) return rc;
  auto ew_grid = [](long long n) { return (int)((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535); };
  auto conv = [&](const float* in, const std::string& scope, float* out, int g, int Tin, int cin, int Tout, int cout, int k,
                  int stride, int dl) {
    const long long M = (long long)g * Tout;
    dim3 grid((unsigned)((M + ENC_BM - 1) / ENC_BM), (unsigned)(cout / ENC_BN));
    conv1d_gemm_kernel<<<grid, 256, 0, h->stream>>>(in, TP(h, scope + "/kernel"), TP(h, scope + "/bias"), one, zero, out, g, Tin,
                                                    cin, Tout, cout, k, stride, dl * (k - 1), 0, dl);
    h->launches += 1;
  };
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int b0 = 0; b0 < B; b0 += group) {
    const int g = (B - b0 < group) ? (B - b0) : group;
    CK(h, cudaMemcpyAsync(h->enc_x.p, x + (size_t)b0 * T, (size_t)g * T * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    float* u = (float*)h->enc_a.p;
    float* en = (float*)h->enc_b.p;
    float* d = (float*)h->enc_c.p;
    float* gc = (float*)h->enc_d.p;
    float* fc = (float*)h->enc_e.p;
    float* tmp = (float*)h->enc_f.p;
    magenta_pre_kernel<<<ew_grid((long long)g * T), 256, 0, h->stream>>>((const float*)h->enc_x.p, u, g, (int)T, 255.0f);
    {
      const long long total = (long long)g * T * MC;
      conv1d_in1_kernel<<<ew_grid(total), 256, 0, h->stream>>>(u, TP(h, "encoder/preprocess/kernel"), TP(h, "encoder/preprocess/bias"),
                                                               one, zero, en, g, (int)T, (int)T, MC, MK, 1, MK - 1, 0);
    }
    h->launches += 2;
    int Tin = (int)T;
    for (int i = 0; i < 6; ++i) {
      const std::string sc = "encoder/cycle_1/layer_" + std::to_string(i + 1);
      const int Tout = Tin / 2;
      const long long n = (long long)g * Tout * MC;
      conv(en, sc + "/dilated", d, g, Tin, MC, Tout, MC, 1, 2, 1);             // 1x1, stride 2: samples 0, 2, 4, ...
      conv(d, sc + "/gate", gc, g, Tout, MC, Tout, MC, MK, 1, dil[i]);
      conv(d, sc + "/filter", fc, g, Tout, MC, Tout, MC, MK, 1, dil[i]);
      magenta_gate_kernel<<<ew_grid(n), 256, 0, h->stream>>>(gc, fc, tmp, n);
      conv(tmp, sc + "/residual", gc, g, Tout, MC, Tout, MC, 1, 1, 1);
      magenta_add_kernel<<<ew_grid(n), 256, 0, h->stream>>>(d, gc, en, n);
      h->launches += 2;
      Tin = Tout;
    }
    conv(en, "encoder/postprocess", (float*)h->enc_z.p, g, Tin, MC, Tin, D, 1, 1, 1);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(z_e_out + (size_t)b0 * (T / 64) * D, h->enc_z.p, (size_t)g * (T / 64) * D * sizeof(float),
                          cudaMemcpyDeviceToHost, h->stream));
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "conv1d_gemm_kernel";
  return finish_timing(h);
}

// Encoder_2019 (Encoder/encoder.py:72-98) on the device: MFCC front end (one fused kernel), then the conv stack through
// the implicit-GEMM kernel.  The front end's constants (hann window, mel weights, DCT matrix) are built here in
// float64 from the published definitions of tf.contrib.signal (Encoder/encoder_ops.py:14-43) and rounded to float32.
static int encode_2019(vqwn_handle* h, const float* x, int B, int64_t T, float* z_e_out) {
  int rc;
  const int D = h->D;
  for (int i = 0; i < 10; ++i) {
    const std::string base = "encoder/conv1d" + (i == 0 ? std::string() : "_" + std::to_string(i));
    if ((rc = check_tensor_ready(h, (base + "/kernel").c_str()))) return rc;
    if ((rc = check_tensor_ready(h, (base + "/bias").c_str()))) return rc;
  }
  const int F1 = (int)((T + MFCC_STEP - 1) / MFCC_STEP), F2 = (F1 + 1) / 2;
  // constants + identity scale / shift + the first kernel padded to 16 input channels: one small buffer
  const size_t n_win = MFCC_FRAME, n_mel = (size_t)MFCC_BINS * MFCC_MELS, n_dct = (size_t)MFCC_MELS * MFCC_COEFS;
  const size_t n_w0 = (size_t)3 * MFCC_CPAD * 768;
  const size_t off_win = 2 * 768, off_mel = off_win + n_win, off_dct = off_mel + n_mel, off_w0 = off_dct + n_dct + 4;
  if ((rc = ensure(h, h->enc_fold, (off_w0 + n_w0) * sizeof(float)))) return rc;
  {
    std::vector<float> c(off_w0, 0.f);
    for (int i = 0; i < 768; ++i) c[i] = 1.f;
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < MFCC_FRAME; ++n) c[off_win + n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / MFCC_FRAME));
    auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
    const double lo = mel(20.0), hi = mel(8000.0);
    for (int k = 1; k < MFCC_BINS; ++k) {                       // the DC bin stays zero
      const double sm = mel(8000.0 * k / (MFCC_BINS - 1));
      for (int m = 0; m < MFCC_MELS; ++m) {
        const double e0 = lo + (hi - lo) * m / (MFCC_MELS + 1), e1 = lo + (hi - lo) * (m + 1) / (MFCC_MELS + 1),
                     e2 = lo + (hi - lo) * (m + 2) / (MFCC_MELS + 1);
        const double w = fmin((sm - e0) / (e1 - e0), (e2 - sm) / (e2 - e1));
        c[off_mel + (size_t)k * MFCC_MELS + m] = (float)(w > 0.0 ? w : 0.0);
      }
    }
    for (int m = 0; m < MFCC_MELS; ++m)
      for (int q = 0; q < MFCC_COEFS; ++q)
        c[off_dct + (size_t)m * MFCC_COEFS + q] = (float)(2.0 * cos(PI * q * (2 * m + 1) / (2.0 * MFCC_MELS)) / sqrt(2.0 * MFCC_MELS));
    CK(h, cudaMemcpyAsync(h->enc_fold.p, c.data(), c.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
  }
  const float* one = (const float*)h->enc_fold.p;
  const float* zero = one + 768;
  float* w0p = (float*)h->enc_fold.p + off_w0;
  pad_cin_kernel<<<64, 256, 0, h->stream>>>(TP(h, "encoder/conv1d/kernel"), w0p, 3, MFCC_COEFS, MFCC_CPAD, 768);
  h->launches += 1;
  const size_t per_stream = (size_t)T * sizeof(float) + 3 * (size_t)F1 * 768 * sizeof(float);
  size_t gsz = ((size_t)1 << 30) / per_stream;
  if (gsz < 1) gsz = 1;
  const int group = gsz > (size_t)B ? B : (int)gsz;
  if ((rc = ensure(h, h->enc_x, (size_t)group * T * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_a, (size_t)group * F1 * 768 * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_b, (size_t)group * F1 * 768 * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_c, (size_t)group * F1 * 768 * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_z, (size_t)group * F2 * D * sizeof(float)))) return rc;
  auto ew_grid = [](long long n) { return (int)((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535); };
  // keras Conv1D(padding='same'): Tout = ceil(Tin / stride), total pad = max((Tout-1) stride + k - Tin, 0), left = total / 2
  auto conv = [&](const float* in, int layer, const float* Wk, float* out, int g, int Tin, int cin, int cout, int k, int stride, int relu) {
    const int Tout = (Tin + stride - 1) / stride;
    int total_pad = (Tout - 1) * stride + k - Tin;
    if (total_pad < 0) total_pad = 0;
    const std::string base = "encoder/conv1d" + (layer == 0 ? std::string() : "_" + std::to_string(layer));
    const long long M = (long long)g * Tout;
    dim3 grid((unsigned)((M + ENC_BM - 1) / ENC_BM), (unsigned)(cout / ENC_BN));
    conv1d_gemm_kernel<<<grid, 256, 0, h->stream>>>(in, Wk ? Wk : TP(h, base + "/kernel"), TP(h, base + "/bias"), one, zero, out, g, Tin,
                                                    cin, Tout, cout, k, stride, total_pad / 2, relu);
    h->launches += 1;
    return Tout;
  };
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int b0 = 0; b0 < B; b0 += group) {
    const int g = (B - b0 < group) ? (B - b0) : group;
    CK(h, cudaMemcpyAsync(h->enc_x.p, x + (size_t)b0 * T, (size_t)g * T * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    float* a = (float*)h->enc_a.p;
    float* b = (float*)h->enc_b.p;
    float* c = (float*)h->enc_c.p;
    mfcc_kernel<<<g * F1, 256, 0, h->stream>>>((const float*)h->enc_x.p, one + off_win, one + off_mel, one + off_dct, c, (int)T, F1);
    h->launches += 1;
    conv(c, 0, w0p, a, g, F1, MFCC_CPAD, 768, 3, 1, 1);                         // net = conv_3_768(mfcc)
    conv(a, 1, nullptr, b, g, F1, 768, 768, 3, 1, 1);                            // conv = conv_3_768(net)
    enc_add_kernel<<<ew_grid((long long)g * F1 * 768), 256, 0, h->stream>>>(b, a, a, (long long)g * F1 * 768);    // net = conv + net
    conv(a, 2, nullptr, b, g, F1, 768, 768, 4, 2, 1);                            // strided_conv_4_768 -> b [F2]
    float* net = b;
    float* tmp = a;
    const long long n2 = (long long)g * F2 * 768;
    for (int i = 3; i < 5; ++i) {                                                // 2 x (conv + net)
      conv(net, i, nullptr, tmp, g, F2, 768, 768, 3, 1, 1);
      enc_add_kernel<<<ew_grid(n2), 256, 0, h->stream>>>(tmp, net, net, n2);
    }
    for (int i = 5; i < 9; ++i) {                                                // 4 x (relu + relu), quirk Q17
      conv(net, i, nullptr, tmp, g, F2, 768, 768, 3, 1, 1);
      enc_add_kernel<<<ew_grid(n2), 256, 0, h->stream>>>(tmp, tmp, tmp, n2);
      float* t_ = net; net = tmp; tmp = t_;
    }
    h->launches += 7;
    conv(net, 9, nullptr, (float*)h->enc_z.p, g, F2, 768, D, 1, 1, 0);           // linear_64
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(z_e_out + (size_t)b0 * F2 * D, h->enc_z.p, (size_t)g * F2 * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "conv1d_gemm_kernel";
  return finish_timing(h);
}

int vqwn_encode_audio(vqwn_handle* h, const float* x, int B, int64_t T, float* z_e_out) {
  ENTER(h);
  if (!x || !z_e_out || B < 1 || T < 64) return fail(h, VQWN_ERR_INVALID, "bad argument");
  if (h->cfg.encoder == VQWN_ENCODER_NONE)
    return fail(h, VQWN_ERR_NOTIMPL, "no encoder configured on the device (vqwn_config.encoder = 64, 1 or 2019)");
  if (h->cfg.encoder == VQWN_ENCODER_2019) {
    if (T % (2 * MFCC_STEP) != 0) return fail(h, VQWN_ERR_INVALID, "T must be a multiple of 320 (Encoder_2019 hop)");
    return encode_2019(h, x, B, T, z_e_out);
  }
  if (T % 64 != 0) return fail(h, VQWN_ERR_INVALID, "T must be a multiple of 64 (encoder hop)");
  if (h->cfg.encoder == VQWN_ENCODER_MAGENTA) return encode_magenta(h, x, B, T, z_e_out);
  int rc;
  std::vector<std::string> sfx(7);
  for (int i = 0; i < 7; ++i) {
    sfx[i] = i == 0 ? "" : "_" + std::to_string(i);
    const char* parts[] = {"/kernel", "/bias"};
    for (const char* pt : parts)
      if ((rc = check_tensor_ready(h, ("encoder/conv1d" + sfx[i] + pt).c_str()))) return rc;
    const char* bparts[] = {"/gamma", "/beta", "/moving_mean", "/moving_variance"};
    for (const char* pt : bparts)
      if ((rc = check_tensor_ready(h, ("encoder/batch_normalization" + sfx[i] + pt).c_str()))) return rc;
  }
  // folded BatchNorm (scale, shift) per layer
  const int D = h->D;
  if ((rc = ensure(h, h->enc_fold, (size_t)7 * 2 * 768 * sizeof(float)))) return rc;
  float* fold = (float*)h->enc_fold.p;
  for (int i = 0; i < 7; ++i) {
    const int cout = i < 6 ? 768 : D;
    const std::string bn = "encoder/batch_normalization" + sfx[i];
    bn_fold_kernel<<<(cout + 255) / 256, 256, 0, h->stream>>>(TP(h, bn + "/gamma"), TP(h, bn + "/beta"), TP(h, bn + "/moving_mean"),
                                                              TP(h, bn + "/moving_variance"), 1e-3f, fold + (size_t)i * 1536,
                                                              fold + (size_t)i * 1536 + 768, cout);
    h->launches += 1;
  }
  // utterances are processed in groups that keep the two ping-pong activation buffers under ~1 GiB
  const size_t per_stream = (size_t)(T / 2) * 768 * sizeof(float) + (size_t)(T / 4) * 768 * sizeof(float);
  size_t gsz = per_stream ? (((size_t)1 << 30) / per_stream) : 1;
  if (gsz < 1) gsz = 1;
  int group = gsz > (size_t)B ? B : (int)gsz;
  if (group > B) group = B;
  if ((rc = ensure(h, h->enc_x, (size_t)group * T * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_a, (size_t)group * (T / 2) * 768 * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_b, (size_t)group * (T / 4) * 768 * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->enc_z, (size_t)group * (T / 64) * D * sizeof(float)))) return rc;
  CK(h, cudaEventRecord(h->ev0, h->stream));
  for (int b0 = 0; b0 < B; b0 += group) {
    const int g = (B - b0 < group) ? (B - b0) : group;
    CK(h, cudaMemcpyAsync(h->enc_x.p, x + (size_t)b0 * T, (size_t)g * T * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    const float* in = (const float*)h->enc_x.p;
    int Tin = (int)T, cin = 1;
    for (int i = 0; i < 7; ++i) {
      const int cout = i < 6 ? 768 : D;
      const int k = i < 6 ? 5 : 1, stride = i < 6 ? 2 : 1;
      const int Tout = (Tin + stride - 1) / stride;
      int total_pad = (Tout - 1) * stride + k - Tin;
      if (total_pad < 0) total_pad = 0;
      const int left = total_pad / 2;                 // TF 'same': 1 left / 2 right for even T, k = 5, stride 2
      float* out = (i == 6) ? (float*)h->enc_z.p : ((i & 1) ? (float*)h->enc_b.p : (float*)h->enc_a.p);
      const float* Wk = TP(h, "encoder/conv1d" + sfx[i] + "/kernel");
      const float* bk = TP(h, "encoder/conv1d" + sfx[i] + "/bias");
      const float* sc = fold + (size_t)i * 1536;
      const float* sh = sc + 768;
      if (cin == 1) {
        const long long total = (long long)g * Tout * cout;
        int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
        conv1d_in1_kernel<<<grid, 256, 0, h->stream>>>(in, Wk, bk, sc, sh, out, g, Tin, Tout, cout, k, stride, left);
      } else {
        const long long M = (long long)g * Tout;
        dim3 grid((unsigned)((M + ENC_BM - 1) / ENC_BM), (unsigned)(cout / ENC_BN));
        conv1d_gemm_kernel<<<grid, 256, 0, h->stream>>>(in, Wk, bk, sc, sh, out, g, Tin, cin, Tout, cout, k, stride, left,
                                                        i < 6 ? 1 : 0);
      }
      CK(h, cudaGetLastError());
      h->launches += 1;
      in = out; Tin = Tout; cin = cout;
    }
    CK(h, cudaMemcpyAsync(z_e_out + (size_t)b0 * (T / 64) * D, h->enc_z.p, (size_t)g * (T / 64) * D * sizeof(float),
                          cudaMemcpyDeviceToHost, h->stream));
  }
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->last_kernel = "conv1d_gemm_kernel";
  return finish_timing(h);
}

// ---------------------------------------------------------------------------------- VQ
int vqwn_vq_upload(vqwn_handle* h, const float* z_e, int64_t n) {
  ENTER(h);
  if (!z_e || n < 0) return fail(h, VQWN_ERR_INVALID, "bad argument");
  int rc;
  if ((rc = ensure(h, h->vq_z, (size_t)n * h->D * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->vq_idx, (size_t)n * sizeof(long long)))) return rc;
  if ((rc = ensure(h, h->vq_out, (size_t)n * h->D * sizeof(float)))) return rc;
  CK(h, cudaMemcpyAsync(h->vq_z.p, z_e, (size_t)n * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->vq_n = n;
  return VQWN_OK;
}

int vqwn_vq_resident(vqwn_handle* h, int64_t n) {
  ENTER(h);
  if (!h->cfg.use_vq) return fail(h, VQWN_ERR_STATE, "use_vq is false");
  if (n < 0 || n > h->vq_n) return fail(h, VQWN_ERR_STATE, "no resident z_e of that size (vqwn_vq_upload first)");
  int rc = check_tensor_ready(h, "embedding/embedding");
  if (rc) return rc;
  if (n == 0) { h->last_ms = 0; return VQWN_OK; }
  rc = launch_vq(h, (const float*)h->vq_z.p, n, (long long*)h->vq_idx.p, (float*)h->vq_out.p, h->D, nullptr, 0, 1);
  if (rc) return rc;
  if ((rc = finish_timing(h))) return rc;
  return check_vq_error(h);
}

int vqwn_vq_download(vqwn_handle* h, int64_t n, int64_t* idx_out, float* zq_out) {
  ENTER(h);
  if (n < 0 || n > h->vq_n) return fail(h, VQWN_ERR_STATE, "no resident result of that size");
  if (idx_out) CK(h, cudaMemcpyAsync(idx_out, h->vq_idx.p, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  if (zq_out) CK(h, cudaMemcpyAsync(zq_out, h->vq_out.p, (size_t)n * h->D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return VQWN_OK;
}

int vqwn_vq_lookup(vqwn_handle* h, const float* z_e, int64_t n, int64_t* idx_out, float* zq_out) {
  int rc;
  if ((rc = vqwn_vq_upload(h, z_e, n))) return rc;
  if ((rc = vqwn_vq_resident(h, n))) return rc;
  return vqwn_vq_download(h, n, idx_out, zq_out);
}

static int upload_speakers(vqwn_handle* h, const int32_t* speaker_idx, int B) {
  int rc;
  if ((rc = ensure(h, h->spk_idx, (size_t)B * sizeof(int)))) return rc;
  if (h->SPK > 0) {
    if (!speaker_idx) return fail(h, VQWN_ERR_INVALID, "speaker_idx is required when speaker_embedding > 0");
    for (int b = 0; b < B; ++b)
      if (speaker_idx[b] < 0 || speaker_idx[b] >= h->cfg.num_speakers)
        return fail(h, VQWN_ERR_INVALID, "speaker index out of range");
    CK(h, cudaMemcpyAsync(h->spk_idx.p, speaker_idx, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  return VQWN_OK;
}

int vqwn_build_condition(vqwn_handle* h, const float* z_q, const int32_t* speaker_idx, int B, int F, float* cond_out) {
  ENTER(h);
  if (!z_q || !cond_out || B < 1 || F < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  int rc;
  if (h->SPK > 0 && (rc = check_tensor_ready(h, "speaker_embedding"))) return rc;
  if ((rc = upload_speakers(h, speaker_idx, B))) return rc;
  const long long n = (long long)B * F;
  if ((rc = ensure(h, h->vq_z, (size_t)n * h->D * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->cond_res, (size_t)n * h->C * sizeof(float)))) return rc;
  CK(h, cudaMemcpyAsync(h->vq_z.p, z_q, (size_t)n * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaEventRecord(h->ev0, h->stream));
  const long long total = n * h->C;
  int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  build_condition_kernel<<<grid, 256, 0, h->stream>>>((const float*)h->vq_z.p, h->SPK > 0 ? TP(h, "speaker_embedding") : nullptr,
                                                      (const int*)h->spk_idx.p, h->D, h->SPK, F, n, (float*)h->cond_res.p);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->launches += 1;
  h->last_kernel = "build_condition_kernel";
  h->cond_B = B; h->cond_F = F;
  CK(h, cudaMemcpyAsync(cond_out, h->cond_res.p, (size_t)n * h->C * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  return finish_timing(h);
}

int vqwn_encode_condition(vqwn_handle* h, const float* z_e, const int32_t* speaker_idx, int B, int F,
                          int64_t* idx_out, float* cond_out) {
  ENTER(h);
  if (!z_e || !cond_out || B < 1 || F < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  if (!h->cfg.use_vq) {
    // model.py:140-142: z_q = z_e
    if (idx_out) return fail(h, VQWN_ERR_STATE, "use_vq is false: there are no indices");
    return vqwn_build_condition(h, z_e, speaker_idx, B, F, cond_out);
  }
  int rc;
  if ((rc = check_tensor_ready(h, "embedding/embedding"))) return rc;
  if (h->SPK > 0 && (rc = check_tensor_ready(h, "speaker_embedding"))) return rc;
  if ((rc = upload_speakers(h, speaker_idx, B))) return rc;
  const long long n = (long long)B * F;
  if ((rc = ensure(h, h->vq_z, (size_t)n * h->D * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->vq_idx, (size_t)n * sizeof(long long)))) return rc;
  if ((rc = ensure(h, h->cond_res, (size_t)n * h->C * sizeof(float)))) return rc;
  h->vq_n = 0;
  CK(h, cudaMemcpyAsync(h->vq_z.p, z_e, (size_t)n * h->D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  rc = launch_vq(h, (const float*)h->vq_z.p, n, (long long*)h->vq_idx.p, (float*)h->cond_res.p, h->C,
                 (const int*)h->spk_idx.p, h->SPK, F);
  if (rc) return rc;
  h->cond_B = B; h->cond_F = F;
  if (idx_out) CK(h, cudaMemcpyAsync(idx_out, h->vq_idx.p, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaMemcpyAsync(cond_out, h->cond_res.p, (size_t)n * h->C * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = finish_timing(h))) return rc;
  return check_vq_error(h);
}

// ---------------------------------------------------------------------------------- WaveNet
int vqwn_reset(vqwn_handle* h, int B) {
  ENTER(h);
  int rc = do_reset(h, B);
  if (rc) return rc;
  CK(h, cudaStreamSynchronize(h->stream));
  return VQWN_OK;
}

int vqwn_step(vqwn_handle* h, const float* audio_t, const float* cond_t, float* logits_out, float* probs_out) {
  ENTER(h);
  if (!audio_t || !cond_t) return fail(h, VQWN_ERR_INVALID, "null argument");
  if (h->B < 1) return fail(h, VQWN_ERR_STATE, "vqwn_reset must be called first (init_ops)");
  int rc;
  if ((rc = pack_weights(h))) return rc;
  const int B = h->B;
  if ((rc = ensure(h, h->small_a, (size_t)B * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->small_b, (size_t)B * h->C * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->small_c, (size_t)B * h->Q * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->small_d, (size_t)B * h->Q * sizeof(float)))) return rc;
  CK(h, cudaMemcpyAsync(h->small_a.p, audio_t, (size_t)B * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->small_b.p, cond_t, (size_t)B * h->C * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  rc = launch_fp32(h, GEN_STEP, 1, (const float*)h->small_b.p, h->C, 0, (const float*)h->small_a.p, nullptr, 0,
                   nullptr, nullptr, (float*)h->small_c.p, (float*)h->small_d.p);
  if (rc) return rc;
  if (logits_out) CK(h, cudaMemcpyAsync(logits_out, h->small_c.p, (size_t)B * h->Q * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (probs_out) CK(h, cudaMemcpyAsync(probs_out, h->small_d.p, (size_t)B * h->Q * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  return finish_timing(h);
}

int vqwn_decode(vqwn_handle* h, const float* probs, int B, int mode, const double* uniforms, int32_t* idx_out,
                float* audio_out) {
  ENTER(h);
  if (!probs || B < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  if (mode != VQWN_MODE_GREEDY && mode != VQWN_MODE_SAMPLE)
    return fail(h, VQWN_ERR_NOTIMPL, "decode mode not implemented");     // utils.py:46
  if (mode == VQWN_MODE_SAMPLE && !uniforms) return fail(h, VQWN_ERR_INVALID, "sample mode needs uniforms[B]");
  int rc;
  if ((rc = pack_weights(h))) return rc;
  const int Q = h->Q;
  if ((rc = ensure(h, h->small_c, (size_t)B * Q * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->small_a, (size_t)B * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->small_b, (size_t)B * (sizeof(double) + sizeof(int))))) return rc;
  double* d_u = (double*)h->small_b.p;
  int* d_idx = (int*)((char*)h->small_b.p + (size_t)B * sizeof(double));
  CK(h, cudaMemcpyAsync(h->small_c.p, probs, (size_t)B * Q * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (mode == VQWN_MODE_SAMPLE)
    CK(h, cudaMemcpyAsync(d_u, uniforms, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaEventRecord(h->ev0, h->stream));
  decode_kernel<<<(B + 7) / 8, 256, 8 * Q * sizeof(float), h->stream>>>((const float*)h->small_c.p, B, Q, mode, d_u, h->dec_lut,
                                                                       d_idx, (float*)h->small_a.p);
  CK(h, cudaGetLastError());
  CK(h, cudaEventRecord(h->ev1, h->stream));
  h->launches += 1;
  h->last_kernel = "decode_kernel";
  if (idx_out) CK(h, cudaMemcpyAsync(idx_out, d_idx, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (audio_out) CK(h, cudaMemcpyAsync(audio_out, h->small_a.p, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  return finish_timing(h);
}

int vqwn_upload_condition(vqwn_handle* h, const float* cond, int B, int F) {
  ENTER(h);
  if (!cond || B < 1 || F < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  int rc;
  const size_t bytes = (size_t)B * F * h->C * sizeof(float);
  if ((rc = ensure(h, h->cond_res, bytes))) return rc;
  CK(h, cudaMemcpyAsync(h->cond_res.p, cond, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->cond_B = B; h->cond_F = F;
  return VQWN_OK;
}

int vqwn_upload_uniforms(vqwn_handle* h, const double* uniforms, int64_t T, int B) {
  ENTER(h);
  if (!uniforms || T < 1 || B < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  int rc;
  const size_t bytes = (size_t)T * B * sizeof(double);
  if ((rc = ensure(h, h->uni_res, bytes))) return rc;
  CK(h, cudaMemcpyAsync(h->uni_res.p, uniforms, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  h->uni_T = T; h->uni_B = B;
  return VQWN_OK;
}

static int generate_common(vqwn_handle* h, int B, int F, int64_t T, int mode, bool have_uniforms, uint64_t seed) {
  if (mode != VQWN_MODE_GREEDY && mode != VQWN_MODE_SAMPLE)
    return fail(h, VQWN_ERR_NOTIMPL, "decode mode not implemented");     // utils.py:46
  if (B < 1 || B > h->max_batch) return fail(h, VQWN_ERR_INVALID, "batch out of range (1..max_batch)");
  if (F < 1 || T < 1 || T % F != 0) return fail(h, VQWN_ERR_INVALID, "T must be a positive multiple of F (generate.py:107)");
  if (h->cond_B != B || h->cond_F != F) return fail(h, VQWN_ERR_STATE, "resident condition does not match B,F");
  if (have_uniforms && (h->uni_B != B || h->uni_T < T)) return fail(h, VQWN_ERR_STATE, "resident uniforms do not match T,B");
  int rc;
  if ((rc = pack_weights(h))) return rc;
  if ((rc = ensure(h, h->audio_res, (size_t)B * T * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->idx_res, (size_t)B * T * sizeof(int)))) return rc;
  if ((rc = do_reset(h, B))) return rc;
  rc = launch_fp32(h, mode == VQWN_MODE_GREEDY ? GEN_GREEDY : GEN_SAMPLE, T, (const float*)h->cond_res.p,
                   (long long)F * h->C, (int)(T / F), nullptr,
                   (mode == VQWN_MODE_SAMPLE && have_uniforms) ? (const double*)h->uni_res.p : nullptr, seed,
                   (float*)h->audio_res.p, (int*)h->idx_res.p, nullptr, nullptr);
  if (rc) return rc;
  h->out_B = B; h->out_T = T;
  return VQWN_OK;
}

int vqwn_generate_resident(vqwn_handle* h, int B, int F, int64_t T, int mode, uint64_t seed) {
  ENTER(h);
  int rc = generate_common(h, B, F, T, mode, h->uni_T > 0 && mode == VQWN_MODE_SAMPLE, seed);
  if (rc) return rc;
  return finish_timing(h);
}

int vqwn_download_output(vqwn_handle* h, int B, int64_t T, float* audio_out, int32_t* idx_out) {
  ENTER(h);
  if (h->out_B != B || h->out_T != T) return fail(h, VQWN_ERR_STATE, "no resident output of that shape");
  if (audio_out) CK(h, cudaMemcpyAsync(audio_out, h->audio_res.p, (size_t)B * T * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (idx_out) CK(h, cudaMemcpyAsync(idx_out, h->idx_res.p, (size_t)B * T * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(h, cudaStreamSynchronize(h->stream));
  return VQWN_OK;
}

int vqwn_generate(vqwn_handle* h, const float* cond, int B, int F, int64_t T, int mode, const double* uniforms,
                  uint64_t seed, float* audio_out, int32_t* idx_out) {
  ENTER(h);
  if (!cond || !audio_out) return fail(h, VQWN_ERR_INVALID, "null argument");
  if (mode != VQWN_MODE_GREEDY && mode != VQWN_MODE_SAMPLE)
    return fail(h, VQWN_ERR_NOTIMPL, "decode mode not implemented");
  if (B < 1 || F < 1 || T < 1) return fail(h, VQWN_ERR_INVALID, "bad argument");
  int rc;
  const size_t cbytes = (size_t)B * F * h->C * sizeof(float);
  if ((rc = ensure(h, h->cond_res, cbytes))) return rc;
  CK(h, cudaMemcpyAsync(h->cond_res.p, cond, cbytes, cudaMemcpyHostToDevice, h->stream));
  h->cond_B = B; h->cond_F = F;
  bool have_u = false;
  if (mode == VQWN_MODE_SAMPLE && uniforms) {
    const size_t ubytes = (size_t)T * B * sizeof(double);
    if ((rc = ensure(h, h->uni_res, ubytes))) return rc;
    CK(h, cudaMemcpyAsync(h->uni_res.p, uniforms, ubytes, cudaMemcpyHostToDevice, h->stream));
    h->uni_T = T; h->uni_B = B;
    have_u = true;
  }
  if ((rc = generate_common(h, B, F, T, mode, have_u, seed))) return rc;
  CK(h, cudaMemcpyAsync(audio_out, h->audio_res.p, (size_t)B * T * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (idx_out) CK(h, cudaMemcpyAsync(idx_out, h->idx_res.p, (size_t)B * T * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return finish_timing(h);
}

int vqwn_teacher_forced(vqwn_handle* h, const float* x, const float* cond, int B, int F, int64_t T, float* logits_out) {
  ENTER(h);
  if (!x || !cond || !logits_out) return fail(h, VQWN_ERR_INVALID, "null argument");
  if (B < 1 || B > h->max_batch) return fail(h, VQWN_ERR_INVALID, "batch out of range (1..max_batch)");
  if (F < 1 || T < 1 || T % F != 0) return fail(h, VQWN_ERR_INVALID, "T must be a positive multiple of F");
  int rc;
  if ((rc = pack_weights(h))) return rc;
  const size_t cbytes = (size_t)B * F * h->C * sizeof(float);
  const size_t lbytes = (size_t)B * T * h->Q * sizeof(float);
  if ((rc = ensure(h, h->cond_res, cbytes))) return rc;
  if ((rc = ensure(h, h->x_res, (size_t)B * T * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->logits_res, lbytes))) return rc;
  CK(h, cudaMemcpyAsync(h->cond_res.p, cond, cbytes, cudaMemcpyHostToDevice, h->stream));
  CK(h, cudaMemcpyAsync(h->x_res.p, x, (size_t)B * T * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  h->cond_B = B; h->cond_F = F;
  if ((rc = do_reset(h, B))) return rc;
  rc = launch_fp32(h, GEN_TEACHER, T, (const float*)h->cond_res.p, (long long)F * h->C, (int)(T / F),
                   (const float*)h->x_res.p, nullptr, 0, nullptr, nullptr, (float*)h->logits_res.p, nullptr);
  if (rc) return rc;
#ifdef TF_DEBUG_MARKS
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) tf_debug_dump(h);
#endif
  CK(h, cudaMemcpyAsync(logits_out, h->logits_res.p, lbytes, cudaMemcpyDeviceToHost, h->stream));
  return finish_timing(h);
}

double vqwn_last_kernel_ms(const vqwn_handle* h) { return h ? h->last_ms : 0.0; }
int64_t vqwn_launch_count(const vqwn_handle* h) { return h ? h->launches : 0; }
const char* vqwn_last_kernel_name(const vqwn_handle* h) { return h ? h->last_kernel : ""; }

}  // extern "C"
