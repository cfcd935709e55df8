// Encoder_64 forward (reference: Encoder/encoder.py:8-26) -- SURVEY 8f #1, the step right before the hot path.
// 6 x [Conv1D(768, k=5, strides=2, 'same') + ReLU + BatchNormalization] + Conv1D(latent_dim, k=1) + BatchNormalization,
// BatchNormalization in inference form (moving statistics, epsilon 1e-3).  Runs once per utterance.
//
// Each layer is an implicit GEMM in float32: rows m = (utterance, output frame), K index = (tap, input channel),
// N = output channel; channels-last activations [B][T][C] make every K chunk of a row a contiguous 64-byte read.
// 64x64 output tile per 256-thread CTA, 4x4 register tile per thread, K chunks of 16 through shared memory;
// bias + ReLU + folded BatchNorm (scale, shift) in the epilogue.
#pragma once
#include "common.cuh"

namespace vqwn {

// BatchNorm inference form folded to y = x*scale + shift
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s;
  }
}

// first layer: one input channel (K = 5): out[b][t][co] = relu(sum_j x[b][2t+j-left] * W[j][0][co] + bias) * scale + shift
__global__ void conv1d_in1_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                  const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ y,
                                  int B, int Tin, int Tout, int Cout, int ksize, int stride, int left, int relu = 1) {
  const long long total = (long long)B * Tout * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long m = i / Cout;
    const int t = (int)(m % Tout);
    const int b = (int)(m / Tout);
    float acc = 0.f;
    for (int j = 0; j < ksize; ++j) {
      const int ti = stride * t + j - left;
      const float xv = (ti >= 0 && ti < Tin) ? x[(long long)b * Tin + ti] : 0.f;
      acc = fmaf(xv, __ldg(W + j * Cout + co), acc);
    }
    acc += __ldg(bias + co);
    if (relu) acc = fmaxf(acc, 0.f);
    y[i] = fmaf(acc, __ldg(scale + co), __ldg(shift + co));
  }
}

constexpr int ENC_BM = 64, ENC_BN = 64, ENC_BK = 16;

// general layer: Cin % 16 == 0, Cout % 64 == 0
__global__ void __launch_bounds__(256) conv1d_gemm_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ y, int B,
                                                          int Tin, int Cin, int Tout, int Cout, int ksize, int stride,
                                                          int left, int relu, int dil = 1) {
  __shared__ __align__(16) float As[ENC_BK][ENC_BM + 4];
  __shared__ __align__(16) float Bs[ENC_BK][ENC_BN];
  const int tid = threadIdx.x;
  const long long M = (long long)B * Tout;
  const long long m0 = (long long)blockIdx.x * ENC_BM;
  const int n0 = blockIdx.y * ENC_BN;
  const int ty = tid >> 4, tx = tid & 15;
  // A loader: row ar of the tile, 4 consecutive input channels aq*4..
  const int ar = tid >> 2, aq = tid & 3;
  const long long am = m0 + ar;
  const bool arow_ok = am < M;
  const int ab = arow_ok ? (int)(am / Tout) : 0;
  const int at = arow_ok ? (int)(am % Tout) : 0;
  // B loader
  const int bk = tid >> 4, bc = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int K = ksize * Cin;
  for (int kk = 0; kk < K; kk += ENC_BK) {
    const int j = kk / Cin, ci0 = kk - j * Cin;
    const int ti = stride * at + j * dil - left;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow_ok && ti >= 0 && ti < Tin)
      av = __ldg(reinterpret_cast<const float4*>(x + ((long long)ab * Tin + ti) * Cin + ci0 + 4 * aq));
    const float4 bv = __ldg(reinterpret_cast<const float4*>(W + (long long)(kk + bk) * Cout + n0 + 4 * bc));
    __syncthreads();
    As[4 * aq + 0][ar] = av.x; As[4 * aq + 1][ar] = av.y; As[4 * aq + 2][ar] = av.z; As[4 * aq + 3][ar] = av.w;
    *reinterpret_cast<float4*>(&Bs[bk][4 * bc]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ENC_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][4 * ty]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * tx]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
      acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
      acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
  }
  const float4 bi = __ldg(reinterpret_cast<const float4*>(bias + n0 + 4 * tx));
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + n0 + 4 * tx));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + n0 + 4 * tx));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + 4 * ty + i;
    if (m < M) {
      float4 v = make_float4(acc[i][0] + bi.x, acc[i][1] + bi.y, acc[i][2] + bi.z, acc[i][3] + bi.w);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
      *reinterpret_cast<float4*>(y + m * Cout + n0 + 4 * tx) = v;
    }
  }
}

// ---- Encoder_Magenta pieces (Encoder/encoder.py:37-64)
// u[b][t] = mu_law_encode(x[b][t-1]), x[b][-1] = 0   (shift_right, wavenet_ops.py:9-14; float path of mu_law_ops.py:5-8)
__global__ void magenta_pre_kernel(const float* __restrict__ x, float* __restrict__ u, int B, int T, float mu) {
  const long long total = (long long)B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    u[i] = mu_law_encode_dev(t > 0 ? x[i - 1] : 0.f, mu, 0.f);
  }
}
// gated = tanh(gate) * sigmoid(filter)   (encoder.py:57)
__global__ void magenta_gate_kernel(const float* __restrict__ g, const float* __restrict__ f, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = tanhf(g[i]) * __fdiv_rn(1.0f, 1.0f + expf(-f[i]));
}
// en = d + r   (encoder.py:59)
__global__ void magenta_add_kernel(const float* __restrict__ d, const float* __restrict__ r, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = d[i] + r[i];
}

// ---- Encoder_2019 pieces (Encoder/encoder.py:66-98, Encoder/encoder_ops.py:14-43)
// MFCC front end, one CTA per (utterance, 10 ms frame): 400-sample frame (zero past the end: pad_end) x periodic hann
// window -> |DFT| at 201 bins (direct form, twiddles cos/sin(2 pi j / 400) from a shared-memory table indexed by k n
// mod 400) -> 80 mel bands (weights [201][80]) -> log(. + 1e-6) -> 13 DCT-II coefficients (matrix [80][13], factor
// 2 / sqrt(160) folded in).  Output channels-last [B][frames][16]: 13 coefficients + 3 zeros, so the first
// convolution runs through the same implicit-GEMM kernel as every other layer (its K chunk is 16 channels).
constexpr int MFCC_FRAME = 400, MFCC_STEP = 160, MFCC_BINS = 201, MFCC_MELS = 80, MFCC_COEFS = 13, MFCC_CPAD = 16;
__global__ void __launch_bounds__(256) mfcc_kernel(const float* __restrict__ x, const float* __restrict__ window,
                                                   const float* __restrict__ melw, const float* __restrict__ dct,
                                                   float* __restrict__ out, int T, int frames) {
  __shared__ float fr[MFCC_FRAME], ct[MFCC_FRAME], st[MFCC_FRAME], mag[MFCC_BINS + 3], lm[MFCC_MELS];
  const int f = blockIdx.x % frames, b = blockIdx.x / frames, tid = threadIdx.x;
  for (int n = tid; n < MFCC_FRAME; n += 256) {
    const long long ti = (long long)f * MFCC_STEP + n;
    fr[n] = (ti < T ? x[(long long)b * T + ti] : 0.f) * window[n];
    ct[n] = cospif((float)n * (1.0f / 200.0f));
    st[n] = sinpif((float)n * (1.0f / 200.0f));
  }
  __syncthreads();
  if (tid < MFCC_BINS) {
    float re = 0.f, im = 0.f;
    int j = 0;
    for (int n = 0; n < MFCC_FRAME; ++n) {
      re = fmaf(fr[n], ct[j], re);
      im = fmaf(fr[n], st[j], im);
      j += tid;
      if (j >= MFCC_FRAME) j -= MFCC_FRAME;
    }
    mag[tid] = sqrtf(re * re + im * im);
  }
  __syncthreads();
  if (tid < MFCC_MELS) {
    float a = 0.f;
    for (int k = 0; k < MFCC_BINS; ++k) a = fmaf(mag[k], __ldg(melw + k * MFCC_MELS + tid), a);
    lm[tid] = logf(a + 1e-6f);
  }
  __syncthreads();
  if (tid < MFCC_CPAD) {
    float a = 0.f;
    if (tid < MFCC_COEFS)
      for (int m = 0; m < MFCC_MELS; ++m) a = fmaf(lm[m], __ldg(dct + m * MFCC_COEFS + tid), a);
    out[((long long)b * frames + f) * MFCC_CPAD + tid] = a;
  }
}
// W [k][cin][cout] -> Wp [k][cin_pad][cout] with zero rows for the padded input channels
__global__ void pad_cin_kernel(const float* __restrict__ W, float* __restrict__ Wp, int k, int cin, int cin_pad, int cout) {
  const int total = k * cin_pad * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin_pad, j = i / (cout * cin_pad);
    Wp[i] = ci < cin ? W[((long long)j * cin + ci) * cout + co] : 0.f;
  }
}
// out = a + b (Encoder_2019's residual adds; a == b gives its `relu + relu`)
__global__ void enc_add_kernel(const float* a, const float* b, float* out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

}  // namespace vqwn
