/* vqwn.h -- C ABI of the B200-native VQ-VAE-WaveNet inference hot path.
 *
 * One shared library (libvqwn.so), plain pointers and sizes, no torch / TensorFlow types.
 * The reference (StanislavParovoy/VQ-VAE-WaveNet) has no FFI layer: its "operator API" for
 * this path is the set of TF graph handles generate.py touches.  Each entry point below
 * names the reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - every function returns an int status: 0 = VQWN_OK, negative = error class; the text
 *     is available from vqwn_last_error().  Nothing aborts or throws across the ABI.
 *   - the caller owns every host buffer (C-contiguous, row-major, float32 unless stated);
 *     the library owns all device memory.  Calls block until outputs are on the host.
 *   - a handle is bound to one CUDA device and is not re-entrant; distinct handles may be
 *     driven from distinct threads or processes (one per GPU).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     VQWN_ERR_CUDA.
 */
#ifndef VQWN_H_
#define VQWN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQWN_OK            0
#define VQWN_ERR_INVALID  (-1)  /* bad argument / shape / unsupported configuration       */
#define VQWN_ERR_CUDA     (-2)  /* CUDA runtime error (no device, launch failure, ...)     */
#define VQWN_ERR_STATE    (-3)  /* call order: weights missing, no resident inputs, ...    */
#define VQWN_ERR_NOTIMPL  (-4)  /* mirrors the reference's NotImplementedError paths       */
#define VQWN_ERR_NOMEM    (-5)

#define VQWN_MAX_LAYERS 64

/* decode modes: utils.py:30-46 ("sample" | "greedy"; anything else -> NotImplementedError) */
#define VQWN_MODE_GREEDY 0
#define VQWN_MODE_SAMPLE 1

/* vqwn_config.encoder */
#define VQWN_ENCODER_NONE 0
#define VQWN_ENCODER_MAGENTA 1   /* Encoder_Magenta (Encoder/encoder.py:29-64) */
#define VQWN_ENCODER_64 64       /* Encoder_64 (Encoder/encoder.py:8-26) */
#define VQWN_ENCODER_2019 2019   /* Encoder_2019 (Encoder/encoder.py:66-98; MFCC front end Encoder/encoder_ops.py:14-43), hop 320 */

/* arithmetic of the decoder step (vqwn_set_precision) */
#define VQWN_PREC_FP32 0   /* fp32 CUDA-core contraction; the parity anchor                 */
#define VQWN_PREC_BF16 1   /* bf16 tcgen05 contraction, fp32 accumulate (tolerance 2e-2)    */
#define VQWN_PREC_TC   2   /* split-bf16 (hi + lo) tcgen05 contraction, fp32 accumulate: float32-grade (logits ~1e-5) */

/* VQ kernel selection (vqwn_set_vq_kernel); AUTO / DIRECT / TENSOR / TENSOR_BF16 give identical indices and z_q */
#define VQWN_VQ_AUTO   0   /* tensor-core kernel when k = 512 and latent_dim = 64, else direct  */
#define VQWN_VQ_DIRECT 1   /* float32 CUDA-core direct form                                      */
#define VQWN_VQ_TENSOR 2   /* tcgen05 tf32 ranking + exact float32 re-evaluation of near-minima  */
#define VQWN_VQ_TENSOR_BF16 3   /* experiment (vq_tc2.cuh): split-bf16 ranking, 15x fewer near-minima to re-evaluate, same results;
                                 * measured slower (its float32 rows come from L2, not shared memory) - DESIGN.md 4.1 */
#define VQWN_VQ_EXPANDED 4      /* Magenta/sonnet.py:91-98: ||z||^2 - 2 z.w + ||w||^2 in float32, first minimum; picks the direct
                                 * form's code except on near-ties of the two formulations */
/* what the VQ hands to the decoder (vqwn_set_vq_output) */
#define VQWN_VQ_OUT_STRAIGHT_THROUGH 0   /* z_e + (e_k - z_e): model.py:73,85-87 (differs from e_k in the last bit) */
#define VQWN_VQ_OUT_CODE 1               /* e_k itself: Magenta/config.py:240-242 (`self.encoding = e_k`)                     */

/* model_parameters.json + wavenet_parameters.json (generate.py:63-64, wavenet.py:10-21) */
typedef struct vqwn_config {
  int32_t quantization_channels; /* wavenet_parameters.json "quantization_channels" (256)  */
  int32_t num_layers;            /* len(dilation_rates) = num_cycles * num_cycle_layers     */
  int32_t num_cycle_layers;      /* only used to form variable names cycle_c/layer_l        */
  int32_t dilations[VQWN_MAX_LAYERS];
  int32_t kernel_size;           /* 3 (only 3 is supported)                                 */
  int32_t dilation_filters;      /* G: gated conv produces 2*G channels                     */
  int32_t skip_filters;          /* S                                                       */
  int32_t residual_filters;      /* R (must equal preprocess.filters)                       */
  int32_t pre_kernel_size;       /* preprocess.kernel_size (32)                             */
  int32_t pre_filters;           /* preprocess.filters (256)                                */
  int32_t k;                     /* model_parameters.json "k" (512)                         */
  int32_t latent_dim;            /* "latent_dim" (64)                                       */
  int32_t speaker_dim;           /* "speaker_embedding" (64; 0 = no speaker condition)      */
  int32_t num_speakers;          /* 109 / 340 / 251 (generate.py:46-57)                     */
  int32_t use_vq;                /* "use_vq"                                                */
  int32_t encoder;               /* "encoder": VQWN_ENCODER_64 / VQWN_ENCODER_MAGENTA run on the device, 0 = encoder output is supplied */
} vqwn_config;

typedef struct vqwn_handle vqwn_handle;

/* ---- lifetime ------------------------------------------------------------------------ */
const char* vqwn_version(void);
/* replaces: building the TF graph + tf.Session (generate.py:42,63-86).  max_batch bounds B
 * of every later call (queues are sized for it, like the FIFOQueue shapes of
 * wavenet_ops.py:181-183). */
int vqwn_create(const vqwn_config* cfg, int device, int max_batch, vqwn_handle** out);
int vqwn_destroy(vqwn_handle* h);
/* message of the last failing call on this handle (h == NULL: last failing vqwn_create) */
const char* vqwn_last_error(const vqwn_handle* h);
/* run the library's kernels on a caller stream (a cudaStream_t) instead of its own, so the
 * caller can time them with its own CUDA events.  NULL restores the private stream. */
int vqwn_set_stream(vqwn_handle* h, void* cuda_stream);
/* arithmetic of the generation loop (wavenet.py:103-172 evaluated by vqwn_generate / vqwn_step / vqwn_teacher_forced).
 * VQWN_PREC_FP32 (default): every contraction in float32, the parity anchor (logits 1e-6 from the restatement).
 * VQWN_PREC_TC: tcgen05 tensor-core kernel at float32-grade accuracy: every contraction operand is split into two
 * bfloat16 numbers (hi + lo) and all four partial products are accumulated in float32 (logits ~1e-5 from the
 * restatement; greedy / same-uniform sequences as the float32 path); default WaveNet geometry only.
 * VQWN_PREC_BF16: tcgen05 tensor-core kernel, weights and contraction inputs rounded to bfloat16, float32 accumulation
 * and float32 residual / skip / softmax (logits within 2e-2); the reference's default WaveNet geometry only, otherwise
 * VQWN_ERR_NOTIMPL.  Set it before vqwn_reset / the first generate call of a run: the dilation-queue layout differs. */
int vqwn_set_precision(vqwn_handle* h, int precision);
/* VQWN_PREC_TC only.  The tensor-core kernel splits every contraction over four MMA-issuing warps that accumulate into
 * one TMEM tile; by default the tensor pipe receives their instructions in arrival order, so float32 rounding differs in
 * the last bits from run to run (~1e-6 relative on the logits; a draw can flip only on such a near-tie).  on != 0: the
 * warps take turns in a fixed order - results are bit-reproducible (a run equals its prefix, a shard equals the unsharded
 * run, the step API equals the loop) at ~25 % lower throughput.  The float32 path (VQWN_PREC_FP32) is always
 * bit-reproducible.  The reference makes no such promise either way (TensorFlow CPU MatMul threading). */
int vqwn_set_reproducible(vqwn_handle* h, int on);
int vqwn_set_vq_kernel(vqwn_handle* h, int kernel);
/* replaces the choice between model.py:73 (`z_q = z_e + stop_gradient(e_k - z_e)`, what conditions the default decoder)
 * and Magenta/config.py:242 (`self.encoding = e_k`, what Magenta/generate.py:67 feeds): applies to vqwn_vq_lookup's
 * zq_out and vqwn_encode_condition's condition rows */
int vqwn_set_vq_output(vqwn_handle* h, int output);
/* sharded runs (generate.py:34: the batch is the list of -speakers, cut into contiguous slices per GPU): global index of
 * this handle's stream 0.  It keys the counter-based generator that stands in for np.random.rand (utils.py:22) when
 * vqwn_generate gets no uniforms, so that a slice draws exactly what the unsharded run draws for the same streams. */
int vqwn_set_stream_offset(vqwn_handle* h, int64_t offset);

/* ---- weights: replaces tf.train.Saver(ema.variables_to_restore()).restore (generate.py:88-90)
 * Tensors are addressed by the reference's variable names, e.g.
 *   "embedding/embedding" [k,D]                         model.py:47-49
 *   "speaker_embedding" [num_speakers,spk]              model.py:23-26
 *   "decoder/preprocess/kernel" [32,1,256] ".../bias"   wavenet_ops.py:173-176
 *   "decoder/cycle_1/layer_1/gated/kernel" [3,R,2G]     wavenet_ops.py:173-176 via :228
 *   "decoder/cycle_1/layer_1/gated/local_condition/kernel" [1,C,2G]   wavenet_ops.py:208
 *   ".../skip/kernel" [1,G,S]  ".../residual/kernel" [1,G,R]          wavenet_ops.py:261-265
 *   "decoder/postprocess1/kernel" ... "decoder/postprocess2/bias"     wavenet.py:152-167
 * shape must match exactly. */
int vqwn_set_tensor(vqwn_handle* h, const char* tf_name, const float* host,
                    const int64_t* shape, int ndim);
/* replaces sess.run(model.embedding) / sess.run(model.speaker_embedding) (generate.py:96-101) */
int vqwn_get_tensor(vqwn_handle* h, const char* tf_name, float* host_out, int64_t capacity);
int vqwn_num_tensors(const vqwn_handle* h);
/* name / shape of the i-th expected tensor; is_set tells whether vqwn_set_tensor was called */
int vqwn_tensor_info(const vqwn_handle* h, int i, char* name_out, int name_cap,
                     int64_t* shape_out /*[4]*/, int* ndim_out, int* is_set);

/* ---- encoder ----------------------------------------------------------- */
/* replaces Encoder_64.build / Encoder_Magenta.build (Encoder/encoder.py:8-26, 29-64; model.py:36-42): x [B,T] float
 * audio -> z_e_out [B, T/64, latent_dim].  T must be a multiple of 64 (Encoder_2019: 320, see below).
 * cfg.encoder == VQWN_ENCODER_64: keras variables "encoder/conv1d[_i]/{kernel,bias}", "encoder/batch_normalization[_i]/
 *   {gamma,beta,moving_mean,moving_variance}" (i = 1..6, the first without suffix; BatchNorm in inference form, eps 1e-3).
 * cfg.encoder == VQWN_ENCODER_MAGENTA: "encoder/preprocess/{kernel,bias}", "encoder/cycle_1/layer_l/{dilated,gate,filter,
 *   residual}/{kernel,bias}" (l = 1..6), "encoder/postprocess/{kernel,bias}"; shift_right + mu_law_encode on the input.
 * cfg.encoder == VQWN_ENCODER_2019 (Encoder/encoder.py:66-98): keras variables "encoder/conv1d[_i]/{kernel,bias}", i = 1..9:
 *   MFCC front end (25 ms periodic-hann frames every 10 ms, pad_end, |DFT| 201 bins, 80 HTK mel bands 20-8000 Hz,
 *   log(. + 1e-6), 13 DCT-II coefficients x rsqrt(160): Encoder/encoder_ops.py:14-43), conv k3 + (conv k3 + skip),
 *   conv k4 stride 2, 2 x (conv k3 + skip), 4 x (conv k3, doubled: the reference's `relu + relu`), 1x1 to latent_dim.
 *   Hop 320: T must be a multiple of 320 and z_e_out is [B, T/320, latent_dim]. */
int vqwn_encode_audio(vqwn_handle* h, const float* x, int B, int64_t T, float* z_e_out);

/* ---- VQ bottleneck + conditioning ---------------------------------------------------- */
/* replaces VQVAE._discretise (model.py:57-74): direct-form squared distance, lowest-index
 * argmin, gather, z_q = z_e + (e_k - z_e).  z_e [n,D]; idx_out [n] int64 (tf.argmin dtype);
 * zq_out [n,D] (may be NULL). */
int vqwn_vq_lookup(vqwn_handle* h, const float* z_e, int64_t n, int64_t* idx_out, float* zq_out);
/* replaces the speaker lookup + decoder_ops.concat (model.py:19-27, decoder_ops.py:39-43,
 * decoder.py:49-50): cond[b,f,:] = [ z_q[b,f,:] | speaker_embedding[speaker_idx[b],:] ].
 * speaker_idx [B] int32 (the argmax of the one-hot; the all-zero "None" row is 0). */
int vqwn_build_condition(vqwn_handle* h, const float* z_q, const int32_t* speaker_idx,
                         int B, int F, float* cond_out);
/* fused: what sess.run(model.encoding) returns (generate.py:92) from the encoder output:
 * VQ lookup + gather + straight-through + speaker concat in one kernel.
 * z_e [B,F,D] -> idx_out [B,F] int64 (may be NULL), cond_out [B,F,D+spk].
 * With use_vq = 0 the VQ is the identity (model.py:140-142). */
int vqwn_encode_condition(vqwn_handle* h, const float* z_e, const int32_t* speaker_idx,
                          int B, int F, int64_t* idx_out, float* cond_out);

/* ---- WaveNet fast generation --------------------------------------------------------- */
/* replaces sess.run(wavenet.init_ops) (generate.py:105; wavenet_ops.py:181-184): zero-fill
 * every dilation queue for B streams and rewind the step counter. */
int vqwn_reset(vqwn_handle* h, int B);
/* replaces one sess.run([wavenet.predictions, wavenet.push_ops], {input_t, local_condition_t})
 * (generate.py:109-110; wavenet.py:103-172).  audio_t [B] float in [-1,1]; cond_t [B,C];
 * logits_out [B,q] (may be NULL; pre-softmax, for parity); probs_out [B,q] (may be NULL). */
int vqwn_step(vqwn_handle* h, const float* audio_t, const float* cond_t,
              float* logits_out, float* probs_out);
/* replaces utils.decode(probs, mode, quantization_channels) (utils.py:13-46) on the device:
 * greedy = first argmax of probs; sample = sequential float32 cumsum + searchsorted-left of
 * uniforms[b] (float64; the np.random.rand(B) of utils.py:22 made injectable; index may be q).
 * idx_out [B] int32 (may be NULL), audio_out [B] float32 = mu_law_decode_np(idx). */
int vqwn_decode(vqwn_handle* h, const float* probs, int B, int mode, const double* uniforms,
                int32_t* idx_out, float* audio_out);
/* replaces init_ops + the whole host sample loop (generate.py:103-113) with one persistent
 * kernel.  cond [B,F,C]; T % F == 0 (ratio = T / F, generate.py:107); uniforms [T,B] float64
 * for sample mode (NULL: a counter-based generator seeded with `seed`); audio_out [B,T]
 * float32 (= to_write, generate.py:112); idx_out [B,T] int32 (may be NULL). */
int vqwn_generate(vqwn_handle* h, const float* cond, int B, int F, int64_t T, int mode,
                  const double* uniforms, uint64_t seed, float* audio_out, int32_t* idx_out);
/* teacher-forced run of the same loop (the oracle is Wavenet.build, wavenet.py:24-100):
 * step t is fed x[b,t-1] (0 at t=0, wavenet_ops.py:9-14) instead of the model's own draw.
 * x [B,T]; logits_out [B,T,q]. */
int vqwn_teacher_forced(vqwn_handle* h, const float* x, const float* cond, int B, int F,
                        int64_t T, float* logits_out);

/* ---- resident (device-side) variants, for timing without host<->device copies -------- */
int vqwn_upload_condition(vqwn_handle* h, const float* cond, int B, int F);
int vqwn_upload_uniforms(vqwn_handle* h, const double* uniforms, int64_t T, int B);
/* runs init + T steps on the resident condition; outputs stay on the device */
int vqwn_generate_resident(vqwn_handle* h, int B, int F, int64_t T, int mode, uint64_t seed);
int vqwn_download_output(vqwn_handle* h, int B, int64_t T, float* audio_out, int32_t* idx_out);
int vqwn_vq_upload(vqwn_handle* h, const float* z_e, int64_t n);
int vqwn_vq_resident(vqwn_handle* h, int64_t n);
int vqwn_vq_download(vqwn_handle* h, int64_t n, int64_t* idx_out, float* zq_out);

/* ---- instrumentation ------------------------------------------------------------------ */
/* device time (ms, CUDA events on the launching stream) of the kernels of the last call */
double vqwn_last_kernel_ms(const vqwn_handle* h);
/* number of kernel launches issued by this handle since creation */
int64_t vqwn_launch_count(const vqwn_handle* h);
/* name of the kernel that did the work of the last generate/step/vq call (for reports) */
const char* vqwn_last_kernel_name(const vqwn_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* VQWN_H_ */
