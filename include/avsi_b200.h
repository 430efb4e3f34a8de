/*
 * avsi_b200.h -- C ABI of the B200-native audio-visual speech-inpainting hot path.
 *
 * Drop-in boundary.  The reference
 * (dr-pato/audio-visual-speech-inpainting, paths below are relative to
 * av_speech_inpainting/) has no FFI of its own: its hot path is reached through
 * Python functions / classes whose arithmetic is delegated to TensorFlow library
 * kernels.  Each entry point here replaces one such group of TF call sites and is
 * bound from Python with ctypes (audio-visual-speech-inpainting_b200/_lib.py); the
 * Python shims keep the reference's names and signatures.
 *
 * Conventions
 *   - every pointer is CALLER-OWNED DEVICE memory (row-major), no allocation inside;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns immediately; 0 = ok, negative = error (avsi_last_error() has the text);
 *   - fp16 tensors are IEEE binary16 (`uint16_t` storage); "time-major" means row
 *     index r = t * B + b;
 *   - thread-safe per stream; no global state except lazily cached TMA descriptors.
 */
#ifndef AVSI_B200_H_
#define AVSI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVSI_OK 0
#define AVSI_ERR_INVALID (-1)
#define AVSI_ERR_CUDA (-2)
#define AVSI_ERR_UNSUPPORTED (-3)

/* Text of the last error raised on the calling thread ("" if none). */
const char* avsi_last_error(void);
/* Library version / build info string (arch, build flags). */
const char* avsi_version(void);
/* Number of kernel launches issued by this library since load (all threads). */
int64_t avsi_launch_count(void);
/* The AVSI_* tuning / test switches (AVSI_LSTM_FWD / AVSI_LSTM_BWD = mma | l4 kernel selection, AVSI_LSTM_ACT,
 * AVSI_GEMM_2SM, AVSI_GEMM_BN, phase timers) are read from the environment once and cached; call this after
 * changing them inside a running process. */
void avsi_reload_env(void);
/* Debug: cycles per phase summed over the steps of the last tcgen05 recurrence launch with AVSI_L4_TIMING=1 /
 * AVSI_B4_TIMING=1 (profiles/bench_lstm.py).  out16 / out24: HOST arrays. */
int avsi_debug_lstm4_timing(unsigned long long* out16);
int avsi_debug_lstm4_bwd_timing(unsigned long long* out24);
/* sizeof() of the argument structs below, for FFI bindings to self-check their mirrors. */
int avsi_sizeof_frontend_args(void);
int avsi_sizeof_istft_args(void);

/* ------------------------------------------------------------------------------------
 * Spectrogram front end.  Replaces, in ONE kernel:
 *   get_stft            audio_processing.py:25-42  (tf.contrib.signal.stft: frame, Hann, rFFT-512)
 *   get_spectrogram     audio_processing.py:45-56  (abs, **power, log(x + 1e-6))
 *   get_log_mel_spectrogram audio_processing.py:59-72 (mel projection + log)
 *   normalise + mask    models.py:30-35            ((spec - mean) / std, * masks)
 *   mask on complex STFT masking.py:42, models.py:186
 *   concat + transpose  models.py:45,102           (audio ++ video, time-major)
 * Frames never touch HBM.  nfft must be 512 and frame_len <= 512.
 *
 *   wav        [B,N] f32            window [frame_len] f32 (periodic Hann, host-made)
 *   twiddle    [512] complex64      exp(-2*pi*i*m/512), host-made from float64
 *   T, F       frames / bins to emit (the reference's out_shape slice; F <= 257)
 *   mean,std   [F] f32 or NULL      (NULL: no normalisation)
 *   mask       [B,T,F] f32 or NULL  (1 = reliable, 0 = hole)
 *   video      [B,T,V] f32 or NULL
 * Outputs (each may be NULL = not produced):
 *   stft_out   [B,T,F] complex64    STFT (multiplied by mask if stft_masked != 0)
 *   spec_out   [B,T,F] f32          |X|^power, log'd if log_flag, normalised if mean/std
 *                                   (= target_spec_norm for power=1, log_flag=1)
 *   feat_out   [B,T,F+V] f32        spec_out * mask  ++ video         (= net_inputs)
 *   xh_out     [T*B, ldx] f16       same as feat_out, time-major, zero padded to ldx
 *   logmel_out [B,T,n_mel] f32      log(|X|^power . mel_w + mel_eps); mel_w [257,n_mel] f32
 */
typedef struct {
  const float* wav; int B; int N;
  int frame_len; int hop; int nfft; int T; int F;
  const float* window; const float* twiddle;
  const float* mean; const float* stdev; const float* mask;
  const float* video; int V;
  float power; int log_flag; int stft_masked;
  float* stft_out; float* spec_out; float* feat_out;
  uint16_t* xh_out; int ldx;
  float* logmel_out; const float* mel_w; int n_mel; float mel_eps;
  double* hole_count;  /* optional [1] f64 device accumulator: += sum(1 - mask) (caller zeroes); a double so that the
                        * count of a {0,1} mask is exact -- and the same bits -- in whatever order the warps arrive */
  int xh_video_only;   /* input='v' (models.py:42-43): xh_out holds only the video columns, from column 0 */
  int mel_masked;      /* != 0: the power spectrum is multiplied by the mask before the mel projection
                          (models_asr.py:33-36, apply_mask) */
  int xh_skip_pad;     /* != 0: the zero padding columns [I, ldx) of xh_out are NOT rewritten (the caller zeroed the
                          buffer once and nothing else writes them): saves 12 % of the row at ldx = 448 */
  const void* mel_bands; /* optional band form of mel_w (each mel filter is non-zero on one contiguous range of bins):
                          int32 lo[128] (first bin of filter j), int32 off[128] (offset of its weights; off[n_mel] =
                          total count <= 1024), then the weights as f32.  Enables the fused log-mel kernel
                          (power = 2, log_flag = 0, logmel_out only); NULL = dense projection in the general kernel */
} avsi_frontend_args;
int avsi_frontend_fwd(const avsi_frontend_args* args, void* stream);

/* Remaining feature transforms of audio_processing.py (next row 8f.4); f32, caller-owned device memory.
 *   avsi_preemphasis     y[b,n] = x[b,n] - alpha * x[b,n-1], x[b,-1] = 0          preemphasis, audio_processing.py:19-22
 *   avsi_mfcc            DCT-II of the log-mel rows * rsqrt(2 n_mel), first n_mfcc   get_mfcc :74-81
 *   avsi_delta_features  sum_{i<=N} i (f[t+i] - f[t-i]) / (2 sum i^2), edge frames replicated   delta :84-93
 *                        (src/dst rows of ld_src/ld_dst floats: add_delta_features writes the orders side by side) */
int avsi_preemphasis(const float* src, int B, int N, float alpha, float* dst, void* stream);
/* get_spectrogram / get_log_mel_spectrogram as stand-alone ops on tensors the caller already holds (audio_processing.py:45-50,
 * :59-66; inside the models the same arithmetic is fused into avsi_frontend_fwd):
 *   avsi_spectrogram  out = |stft| ** power, then log(. + 1e-6) if log_flag; stft_c64 = n complex64 values
 *   avsi_log_mel      out[r, m] = log(sum_k spec[r, k] mel_w[k, m] + eps); spec [rows, nbins], mel_w [nbins, n_mel] f32 */
int avsi_spectrogram(const float* stft_c64, int64_t n, float power, int log_flag, float* out, void* stream);
int avsi_log_mel(const float* spec, const float* mel_w, int64_t rows, int nbins, int n_mel, float eps, float* out,
                 void* stream);
int avsi_mfcc(const float* logmel, int64_t rows, int n_mel, int n_mfcc, float* out, void* stream);
/* downsampling(samples, sample_rate, downsample_rate), audio_processing.py:9-16 = scipy.signal.resample(x, num) for real x
 * (dataset preparation: audio_feat_preprocessing.py:82,189; tfrecord_utils.py): forward DFT of the n_in samples, first
 * min(n_in, n_out)/2 + 1 bins with scipy's Nyquist rule, inverse real DFT of length n_out, times n_out / n_in.  Any lengths
 * (both transforms are chirp-z convolutions on power-of-two FFTs), complex double arithmetic, `batch` recordings of equal
 * length.  x [batch, n_in] f64, y [batch, n_out] f64, workspace: avsi_resample_workspace_bytes() bytes of device memory. */
int64_t avsi_resample_workspace_bytes(int batch, int64_t n_in, int64_t n_out);
int avsi_resample_fft(const double* x, int batch, int64_t n_in, double* y, int64_t n_out, void* workspace,
                      int64_t workspace_bytes, void* stream);
int avsi_delta_features(const float* src, int ld_src, float* dst, int ld_dst, int B, int T, int F, int N, void* stream);
/* out = keep != 0 ? a : b over n complex64 values (b NULL = 0): the phase source of one consistency iteration of the phase
 * refinement that stands in for lws.run_lws in inference.py:143-154 (known spectrum in the reliable bins, re-analysed one
 * in the holes); the iteration itself is avsi_istft_fwd + avsi_frontend_fwd (phase_reconstruction.py). */
int avsi_select_c64(const float* a_c64, const float* b_c64, const float* keep, int64_t n, float* out_c64, void* stream);

/* Waveform reconstruction (next row 8f.1): get_sources / reconstruct_sources
 * audio_processing.py:145-164, enhanced_sources models.py:181-197.
 *   mag [B,T,F] f32 (or pred with denorm: mag = exp(pred*std+mean) when mean != NULL)
 *   phase_src [B,T,F] complex64 (angle taken inside; multiplied by mask first if mask != NULL)
 *   inv_window [frame_len] f32 (inverse_stft_window_fn), out [B,num_samples] f32.  */
typedef struct {
  const float* mag; const float* phase_src; const float* mask;
  const float* mean; const float* stdev;
  const float* inv_window; const float* twiddle;
  int B; int T; int F; int frame_len; int hop; int nfft; int num_samples;
  float* out;
} avsi_istft_args;
int avsi_istft_fwd(const avsi_istft_args* args, void* stream);

/* Normalised features -> time-major fp16 network input (models_asr.py:37-49, the ASR model's input assembly):
 *   x0[t*B + b, 0:F] = (feat[b,t,:] - mean) / std ; x0[.., F:F+V] = video[b,t,:] (video may be NULL); columns up to
 *   ldx are zeroed.  feat [B,T,F] f32, mean/std [F] f32 or both NULL (features taken as they are: the two-step model's
 *   audio_features = video_model.prediction, models.py:258-262). */
int avsi_features_to_x0(const float* feat, const float* mean, const float* stdev, const float* video, int B, int T,
                        int F, int V, uint16_t* x0, int ldx, void* stream);

/* Speaker-embedding sub-network of the SSNN models (models.py:800-842): what sits between its dense layers (which run
 * on avsi_gemm_f16), forward and backward.  Rows batch-major (r = b*T + t).
 *   avsi_cast_pad_f16          f32 [rows, cols] (pitch ld_src) -> f16 [rows, ld_dst], columns >= cols zero
 *   avsi_leaky_relu            tf.nn.leaky_relu(z, alpha) -> f16          avsi_leaky_relu_bwd   dz = da * (z > 0 ? 1 : alpha) -> f16
 *   avsi_masked_time_mean      out[b,n] = sum_t x[b,t,n] m[b,t] / (sum_t m[b,t] + 1), m[b,t] = mask[(b*T+t)*mask_ld]  (:834-836)
 *   avsi_masked_time_mean_bwd  dx[b,t,n] = d_out[b,n] m[b,t] inv_den[b] -> f16
 *   avsi_time_sum              out[b,n] = sum_t x[(t*B+b)*ld + n]: gradient of a vector replicated over the frames */
int avsi_cast_pad_f16(const float* src, int ld_src, int cols, uint16_t* dst, int ld_dst, int64_t rows, void* stream);
int avsi_leaky_relu(const float* z, int64_t n, float alpha, uint16_t* out, void* stream);
int avsi_leaky_relu_bwd(const float* z, const float* da, int64_t n, float alpha, uint16_t* dz, void* stream);
int avsi_masked_time_mean(const float* x, const float* mask, int mask_ld, int B, int T, int N, float* out,
                          float* inv_den, void* stream);
int avsi_masked_time_mean_bwd(const float* d_out, const float* mask, int mask_ld, const float* inv_den, int B, int T,
                              int N, uint16_t* dx, void* stream);
int avsi_time_sum(const float* x, int ld, int T, int B, int N, float* out, void* stream);

/* Per-utterance vector replicated over the frames into columns [col0, col0 + E) of the time-major fp16 network input:
 * tf.tile(tf.expand_dims(embeddings, 1), [1, T, 1]) + tf.concat([net_inputs, tiles], 2) of the embedding models
 * (models.py:1204-1206; the SSNN models' speaker embedding, models.py:846-849).  emb [B,E] f32. */
int avsi_tile_embedding(const float* emb, int B, int T, int E, uint16_t* x0, int ldx, int col0, void* stream);

/* ------------------------------------------------------------------------------------
 * Landmark stream -> network video features.  Replaces inc_fps /
 * sync_audio_visual_features (av_sync.py:7-40), get_motion_vector
 * (face_landmarks.py:30-39) and the z-normalisation of tfrecord_utils.py:104-107.
 *   landmarks [B,L,D] f32 (already start-padded to L frames), vmean/vstd [B,D] f32
 *   out [B,T,D] f32: lerp to T frames (clamped), first difference (row 0 = 0), z-norm. */
int avsi_video_features(const float* landmarks, const float* vmean, const float* vstd,
                        int B, int L, int D, int T, float* out, void* stream);

/* Dense mask from intervals: get_intrusions_mask tail, dataset_generator.py:43-46.
 *   intervals [B,K,2] i32 (onset, length; length 0 = unused) -> mask [B,T,F] f32 of {0,1}. */
int avsi_expand_mask(const int32_t* intervals, int B, int K, int T, int F, float* mask, void* stream);

/* Feature statistics (hot-path row a15): the float64 accumulation of compute_mean_std_features
 * (audio_feat_preprocessing.py:76-115).  x [rows, ldx] f32 features; mask [rows, ldm] f32 or NULL
 * (feat * mask, frame count += mask[:, 0], :87-92,104-107).  sum/sumsq [F] f64 and count [1] f64 are
 * ACCUMULATED (caller zeroes them once and calls this per batch). */
int avsi_feature_stats(const float* x, int ldx, const float* mask, int ldm, int64_t rows, int F, double* sum,
                       double* sumsq, double* count, void* stream);

/* ------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma kind::f16, fp16 operands, fp32 accumulate in TMEM,
 * operands staged by TMA).  Replaces the cuBLAS / cuDNN contractions behind
 * CudnnLSTM input projections (models.py:95-104), tf.matmul heads (models.py:117-123,
 * 1902-1912) and their gradients.
 *   C[M,N] (+)= A . B^T      trans == 0: A [M,K] (lda), B [N,K] (ldb), both K-contiguous
 *   C[M,N] (+)= A^T . B      trans == 1: A [K,M] (lda), B [K,N] (ldb), both MN-contiguous
 *   out_mode 0: C f16 = acc ; 1: C f32 = acc + bias[N] (bias may be NULL) ; 2: C f32 += acc (atomic)
 *   split_k > 1 only with out_mode 2.  lda/ldb multiples of 8 elements; A,B 16-byte aligned.
 *   layout bit 0: A is stored interleaved ("IL": [rows/32][lda/8][32][8] halves, rows padded to 32 with
 *   zeros) instead of row-major; layout bit 1: the f16 output C (out_mode 0) is written interleaved with
 *   row length ldc; layout bit 2: B is stored interleaved likewise (trans == 1 only).  IL is the layout of the
 *   gate tensors G / dG, of dL/dy and (optionally) of the layer outputs y shared with the recurrence kernels. */
int avsi_gemm_f16(const uint16_t* A, int lda, const uint16_t* B, int ldb, void* C, int ldc,
                  const float* bias, int M, int N, int K, int trans, int out_mode, int split_k,
                  int layout, void* stream);

/* ------------------------------------------------------------------------------------
 * Bit-reproducible reductions.  TensorFlow's gradient kernels behind models.py:161-196 (compute_gradients: cuDNN /
 * cuBLAS weight gradients, BiasAddGrad, the reduce_sum of the losses) are not run-to-run deterministic, and neither
 * is this library by default: split-K weight-gradient GEMMs, bias-gradient column sums and the loss sums meet in
 * floating-point atomics.  Once a device scratch buffer is registered here, avsi_gemm_f16 (out_mode 2 with k splits),
 * avsi_colsum_f16 and avsi_masked_l1 write per-split / per-block partial sums into it and add them in a FIXED order
 * (a second tiny kernel, or the block that draws the last ticket), and avsi_lstm_bwd does the same with its own
 * `scratch` argument: two runs of a training step on the same inputs then produce the same bits.
 *   scratch: device memory, 256-byte aligned, >= avsi_reduce_scratch_min_bytes() and >= avsi_gemm_f16_scratch_bytes(...)
 *   of every split-K problem that will run (a too small buffer is an error, never a silent fallback); NULL unregisters.
 *   One registration per device; all users must be ordered on ONE stream (they share the buffer).  The first 4 KB hold
 *   tickets that are zero between launches (zeroed here on `stream`).  Not covered: avsi_ctc_loss (posterior
 *   scatter in shared-memory atomics), the fp32 hole counter of the front end, avsi_istft_fwd's overlap-add. */
int avsi_set_reduce_scratch(void* scratch, int64_t bytes, void* stream);
int64_t avsi_reduce_scratch_min_bytes(void);
int64_t avsi_gemm_f16_scratch_bytes(int M, int N, int K, int trans, int out_mode, int split_k);

/* ------------------------------------------------------------------------------------
 * Persistent bidirectional LSTM recurrence.  Replaces cudnnRNNForwardTraining /
 * BackwardData of tf.contrib.cudnn_rnn.CudnnLSTM (models.py:95-104) and the
 * CudnnCompatibleLSTMCell while-loop (models.py:106-115).  Both directions of one
 * layer run in one launch; W_hh stays on chip across all T steps (shared memory of a 4-CTA
 * thread-block cluster feeding tcgen05.mma, or registers of an 8-CTA cluster for small batches).
 * HP = 256 (H padded), gate columns are laid out [dir][unit][i,g,f,o] (column = dir*1024 + unit*4 + gate).
 *   gates [T*B, 2048] f16   INTERLEAVED storage ([rows/32][2048/8][32][8], rows padded to 32 with zeros):
 *                           in: x.W_ih^T pre-activations with the i, f, o columns pre-halved (avsi_gemm_f16 layout
 *                           bit 1 on the prescaled weight copy, see avsi_cast_weights) ;
 *                           out: activated gates i, g, f, o (in place)
 *   whh   [2,1024,256] f16  recurrent weights, [dir][gate column][h_in], i, f, o rows pre-halved likewise
 *   bias  [2048] f32        prescaled likewise (avsi_gate_bias_prescale).  The tcgen05 kernel feeds it through the
 *                           tensor core (padding unit 255 of h is held at 1.0, column 255 of its W_hh copy = bias)
 *   y     [T*B, 512] f16    out: h_t, columns dir*256 + unit; row-major, or with y_il != 0 INTERLEAVED like the gates
 *                           (rows padded to 32; the consumers are avsi_gemm_f16 with layout bits 0 / 2)
 *   cst   [T*B, 512] f32    out: c_t (stash for BPTT), INTERLEAVED with 4-float chunks ([rows/32][512/4][32][4])
 * Backward: dy [T*B,512] f16 (scaled dL/dy), INTERLEAVED like the gates ([rows/32][512/8][32][8]: a warp's 32 rows x
 * 8 units are 512 contiguous bytes; avsi_gemm_f16 writes it with layout bit 1), whhT [256,2048] f16 (whh^T) -> gates becomes dgates
 * (in place), dbias[2048] f32 += column sums.  scratch (16-byte aligned, >= avsi_lstm_bwd_scratch_bytes(B)) receives one
 * row of column sums per batch tile, added to dbias in tile order (bit-reproducible); scratch == NULL: fp32 atomics. */
int avsi_lstm_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst,
                  int T, int B, int y_il, void* stream);
int64_t avsi_lstm_bwd_scratch_bytes(int B);
int avsi_lstm_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy,
                  float* dbias, void* scratch, int T, int B, void* stream);

/* ------------------------------------------------------------------------------------
 * Masked-L1 loss + gradient.  Replaces prediction / loss of models.py:127-159 (SI, mode 0)
 * and models.py:1920-1963 (MTL, mode 1) and their autodiff mirror.
 *   logits [T*B, ldl] f32 time-major (head GEMM output), target/mask [B,T,F] f32, seq_len [B] i32
 *   sums[8] f64 += { sum|d|(1-m), sum(1-m), sum|d|m, sum m, sum|d|, count, 0, 0 }
 *   prediction [B,T,F] f32 (optional)
 *   dlogits [T*B, ldd] f16 (optional): grad_scale * dL/dlogits of the UNnormalised sums, i.e.
 *     mode 0: sign(pred - target) * seqmask            (caller folds 1/(B*T*F) into grad_scale)
 *     mode 1: sign(pred - target) * (1-m) * seqmask    (caller folds 1/sum(1-m)) */
int avsi_masked_l1(const float* logits, int ldl, const float* target, const float* mask,
                   const int32_t* seq_len, int B, int T, int F, int mode, float grad_scale,
                   const float* grad_scale_dev, double* sums, float* prediction, uint16_t* dlogits,
                   int ldd, void* stream);
/* The training step's fusion of the inpainting head with that loss: logits = A[M, K] . W[N, K]^T + bias (the tf.matmul head of
 * models.py:117-123 / 1902-1912) are consumed in the GEMM's epilogue by the arithmetic of avsi_masked_l1 -- the six sums are
 * ADDED to sums[0..5], dlogits[:, :F] is written -- and never stored, except the columns >= logits_from_col (the phone head
 * read by avsi_ctc_loss; pass N or more for none).  M = B * T time-major rows; a_layout 1 = A stored interleaved;
 * partial_ws: avsi_head_l1_workspace_bytes() bytes of device scratch. */
int avsi_head_l1(const uint16_t* A, int lda, int a_layout, const uint16_t* W, int ldw, const float* bias, int M, int N, int K,
                 float* logits, int ldc, int logits_from_col, const float* target, const float* mask, const int32_t* seq_len,
                 int B, int T, int F, int mode, float grad_scale, const float* grad_scale_dev, double* sums,
                 double* partial_ws, uint16_t* dlogits, int ldd, void* stream);
int avsi_head_l1_workspace_bytes(void);
/* grad_scale_dev (here and below): optional device scalar multiplied into grad_scale, so that
 * data-dependent normalisers (sum(1-m) of the MTL loss) never need a host round trip. */

/* MTL gradient scales from the hole count: out[0] = S (L1 dlogits scale, a power of two),
 * out[1] = S*(ctc_weight/B)*holes (CTC dlogits scale), out[2] = 1/(S*holes) (optimiser unscale),
 * out[3] = holes.  loss = loss_hole + ctc_weight * mean_b(nll_b)  (models.py:1955). */
/* guard (optional): the overflow-guard state below; its dynamic scale multiplies out[0] and out[1]. */
int avsi_mtl_scales(const double* hole_count, int B, float ctc_weight, float* out, const int32_t* guard,
                    void* stream);

/* Inverted dropout of the BLSTM outputs ahead of the head(s): tf.nn.dropout(rnn_outputs, rate=dropout_rate) at
 * models.py:117, :1901, models_asr.py:120.  dst[r,c] = keep ? src[r,c] / (1 - rate) : 0 with keep = (u >= rate),
 * u = Philox-4x32-10(seed; offset, element) -- a pure function of its arguments, so the backward pass applies the
 * same call to dY.  src/dst f16 [rows, ld] (may alias), cols % 8 == 0; keep_out optional u8 [rows, cols];
 * layout bit 0 / bit 1: src / dst is stored INTERLEAVED like the gate tensor (then ld = cols). */
int avsi_dropout_f16(const void* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, float rate,
                     uint64_t seed, uint64_t offset, void* keep_out, int layout, void* stream);

/* Column sums: out[n] += sum_r X[r, col0 + n] (f16 in, f32 out) -- bias gradients. */
int avsi_colsum_f16(const uint16_t* X, int ldx, int rows, int col0, int ncols, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * CTC loss + gradient.  Replaces tf.nn.ctc_loss (CPU-only op in TF1) as called at
 * models.py:1950-1953: unnormalised logits, blank = C-1, merge_repeated.
 *   logits [T*B, ldl] f32 time-major, classes at columns col0 .. col0+C-1
 *   labels [B,Lmax] i32, lab_len/seq_len [B] i32
 *   nll [B] f32 ; dlogits [T*B, ldd] f16 at columns dcol0.. : grad_scale * d nll_b / d logits
 *   workspace >= avsi_ctc_workspace_bytes(B,T,Lmax) */
int64_t avsi_ctc_workspace_bytes(int B, int T, int Lmax);
int avsi_ctc_loss(const float* logits, int ldl, int col0, int C, const int32_t* labels, int Lmax,
                  const int32_t* lab_len, const int32_t* seq_len, int B, int T, float grad_scale,
                  const float* grad_scale_dev, float* nll, uint16_t* dlogits, int ldd, int dcol0,
                  void* workspace, void* stream);

/* CTC prefix beam search, HOST memory in and out (the reference's decoder is a CPU op too and runs at logging /
 * validation steps only): tf.nn.ctc_beam_search_decoder(tm_logits, sequence_lengths, beam_width=20) at
 * models.py:1627, :2027, models_asr.py:139.  logits [T*B, ldl] f32 time-major on the host, classes at columns
 * col0 .. col0+C-1, blank = C-1; out [B, max_out] i32 padded with -1 (tf.sparse.to_dense(default_value=-1)),
 * out_len [B], log_prob [B] optional.  merge_repeated != 0 collapses consecutive equal labels of the emitted
 * path (TF-1 default).  n_threads <= 0: all host cores. */
int avsi_ctc_beam_search_host(const float* logits, int T, int B, int ldl, int col0, int C, const int* seq_len,
                              int beam_width, int merge_repeated, int max_out, int* out, int* out_len,
                              float* log_prob, int n_threads);

/* One serialized tf.train.SequenceExample of the reference's TFRecords (tfrecord_utils.py:19-41, read by
 * dataset_reader.py:62-79) parsed on the HOST straight into caller buffers (capacities in floats):
 *   wav <- context target_audio_wav; mask / video <- feature lists mask / video_features, row after row;
 *   labels <- feature list labels (one float per row); path <- context sample_path (not NUL-terminated).
 *   meta[9] = {sequence_length, labels_length, n_wav, mask_rows, mask_cols, video_rows, video_cols, n_labels, path_len}.
 * Returns non-zero on malformed input, ragged rows or a buffer too small.  Thread safe. */
int avsi_parse_av_sample_host(const void* rec, uint64_t len, float* wav, int64_t wav_cap, float* mask, int64_t mask_cap,
                              float* video, int64_t video_cap, float* labels, int64_t labels_cap, char* path,
                              int path_cap, int64_t* meta);

/* CRC-32C (Castagnoli) of n bytes of HOST memory, continuing from `crc` (0 to start): the checksum of TFRecord
 * frames (tfrecord_utils.py:19-41 via tf.python_io.TFRecordWriter) and of tf.train.Saver tensor bundles
 * (training.py:114,267,335).  Known answer: "123456789" -> 0xE3069283. */
uint32_t avsi_crc32c_host(const void* data, uint64_t n, uint32_t crc);

/* ------------------------------------------------------------------------------------
 * Optimiser.  Replaces tf.train.AdamOptimizer ApplyAdam (models.py:168,178), TF epsilon-hat form.
 *   g is multiplied by grad_unscale first; l2 adds l2 * theta to the gradient (models.py:153-158).
 *   step = number of this update (>= 1), or 0 to take it from the guard's device-resident call count (guard word 3, bumped
 *   by avsi_grad_guard_update): the form a captured CUDA graph of the training step replays. */
int avsi_adam_tf(float* theta, const float* g, float* m, float* v, int64_t n, double lr, double b1,
                 double b2, double eps, int step, float grad_unscale, const float* grad_unscale_dev,
                 float l2, const int32_t* guard, void* stream);
/* tf.train.GradientDescentOptimizer (accum == NULL) / MomentumOptimizer(momentum = 0.9) (models.py:169-173):
 * accum = momentum * accum + g ; theta -= lr * accum.  lr is the (host-evaluated) staircase exponential decay. */
int avsi_sgd_momentum(float* theta, const float* g, float* accum, int64_t n, double lr, double momentum,
                      float grad_unscale, const float* grad_unscale_dev, float l2, const int32_t* guard,
                      void* stream);
/* Overflow guard of the fp16 gradient path (no counterpart in the reference, whose graph is fp32: there a run-away
 * gradient ends in the NaN exit of training.py:244-249; here dlogits / dY / dG are loss-scaled fp16 and would
 * saturate first).  guard = 8 device words: [0] i32 non-finite flag of this step, [1] i32 steps skipped so far,
 * [2] i32 finite steps since the last scale change, [4] f32 dynamic scale s (power of two <= 1, fed to the loss
 * kernels as grad_scale_dev / to avsi_mtl_scales), [5] f32 1/s (applied by the optimiser kernels).
 *   avsi_grad_guard_init    s = 1, counters 0 (synchronises the stream).
 *   avsi_grad_guard_check   [0] = any(!isfinite(g[0..n)))   -- run on the all-reduced gradient, before the optimiser;
 *                           the optimiser kernels leave theta / m / v untouched when [0] != 0 (guard may be NULL).
 *   avsi_grad_guard_update  after the optimiser: skipped step -> s /= 2, else after growth_interval finite steps
 *                           s = min(2 s, 1).  All on the device: the step needs no host synchronisation. */
int avsi_grad_guard_init(int32_t* guard, void* stream);
int avsi_grad_guard_check(const float* g, int64_t n, int32_t* guard, void* stream);
int avsi_grad_guard_update(int32_t* guard, int growth_interval, void* stream);
/* Widen a batch that was copied to the device in its storage dtype (src_type 0 = int16 samples, 1 = uint8 masks,
 * 2 = int32 samples as dataset_reader.py:78 yields them) to the fp32 tensor of the feed contract (training.py:69-71). */
int avsi_cast_to_f32(const void* src, int src_type, int64_t n, float* dst, void* stream);
/* fp32 -> fp16 copies of a weight matrix W [R,C]: w16 [R,C] and (optional) w16t [C,R]. */
int avsi_cast_weights(const float* w, int R, int C, uint16_t* w16, uint16_t* w16t, int halve_sigmoid_rows,
                      void* stream);
/* halve_sigmoid_rows != 0 (gate matrices W_ih, W_hh, rows = gate columns [unit][i,g,f,o]): rows i, f, o of w16 are
 * multiplied by 1/2 (exact), so that the forward recurrence evaluates sigma(z) = 1/2 tanh(z') + 1/2 on the
 * pre-halved z' without a multiply; w16t (read by the backward GEMMs) is never scaled.
 * avsi_gate_bias_prescale: out[n] = bias[n] * (1/2 for i, f, o columns, 1 for g) -- the bias vector avsi_lstm_fwd takes. */
int avsi_gate_bias_prescale(const float* bias, int n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVSI_B200_H_ */
