// The remaining feature transforms of audio_processing.py / audio_feat_preprocessing.py, thin HBM-bound
// kernels on top of the fused front end:
//   avsi_preemphasis      preemphasis(sources, alpha)            audio_processing.py:19-22
//   avsi_mfcc             get_mfcc(log_mel, num_mfccs)           audio_processing.py:74-81 (tf.signal.mfccs_from_log_mel_spectrograms)
//   avsi_delta_features   delta(features, N)                     audio_processing.py:84-93 (add_delta_features :96-103 loops it)
#include "common.cuh"

namespace avsi {

__global__ void __launch_bounds__(256)
preemphasis_kernel(const float* __restrict__ x, int B, int N, float alpha, float* __restrict__ y) {
  const long long total = (long long)B * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const float prev = n > 0 ? __ldg(x + i - 1) : 0.f;          // the shifted copy starts with a zero column
    y[i] = __ldg(x + i) - alpha * prev;
  }
}

// mfcc[k] = rsqrt(2 M) * 2 * sum_n logmel[n] cos(pi k (2 n + 1) / (2 M)),  k < n_mfcc  (DCT-II, norm=None, then TF's scale)
__global__ void __launch_bounds__(256)
mfcc_kernel(const float* __restrict__ logmel, long long rows, int M, int n_mfcc, float* __restrict__ out) {
  extern __shared__ float ctab[];                                // [n_mfcc][M]
  for (int i = threadIdx.x; i < n_mfcc * M; i += blockDim.x) {
    const int k = i / M, n = i - k * M;
    ctab[i] = (float)(2.0 * cos(3.14159265358979323846 * (double)k * (double)(2 * n + 1) / (double)(2 * M)) /
                      sqrt(2.0 * (double)M));
  }
  __syncthreads();
  const long long total = rows * n_mfcc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n_mfcc;
    const int k = (int)(i - r * n_mfcc);
    const float* src = logmel + r * M;
    const float* c = ctab + k * M;
    float acc = 0.f;
    for (int n = 0; n < M; ++n) acc = fmaf(__ldg(src + n), c[n], acc);
    out[i] = acc;
  }
}

// delta[t] = sum_{i=1..N} i (f[t+i] - f[t-i]) / (2 sum i^2), indices clamped to [0, T-1]: the reference pads by one frame
// with mode SYMMETRIC once per i, which replicates the edge frames
__global__ void __launch_bounds__(256)
delta_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst, int B, int T, int F, int N,
             float inv_den) {
  const long long total = (long long)B * T * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long bt = i / F;
    const int t = (int)(bt % T);
    const long long b0 = (bt - t);                                // b * T
    float acc = 0.f;
    for (int k = 1; k <= N; ++k) {
      const int tp = min(t + k, T - 1), tm = max(t - k, 0);
      acc += (float)k * (__ldg(src + (b0 + tp) * ld_src + f) - __ldg(src + (b0 + tm) * ld_src + f));
    }
    dst[bt * ld_dst + f] = acc * inv_den;
  }
}

inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// get_spectrogram (audio_processing.py:45-50) on an already computed complex STFT: |z| ** power, log(. + 1e-6)
__global__ void __launch_bounds__(256)
spectrogram_kernel(const float2* __restrict__ stft, long long n, float power, int log_flag, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 z = stft[i];
    float v = sqrtf(z.x * z.x + z.y * z.y);
    if (power == 2.f) v = v * v;
    else if (power != 1.f) v = powf(v, power);
    out[i] = log_flag ? logf(v + 1e-6f) : v;
  }
}

// get_log_mel_spectrogram (audio_processing.py:59-66) on a spectrogram tensor: log(spec . M + eps), M [nbins, n_mel] dense f32
// (tf.tensordot + tf.log).  One warp per row; fp32 accumulation in bin order.
__global__ void __launch_bounds__(256)
log_mel_kernel(const float* __restrict__ spec, const float* __restrict__ mel_w, long long rows, int nbins, int n_mel,
               float eps, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const float* x = spec + r * nbins;
    for (int m = lane; m < n_mel; m += 32) {
      float acc = 0.f;
      for (int k = 0; k < nbins; ++k) acc = fmaf(__ldg(x + k), __ldg(mel_w + (long long)k * n_mel + m), acc);
      out[r * n_mel + m] = logf(acc + eps);
    }
  }
}

}  // namespace avsi

extern "C" int avsi_preemphasis(const float* src, int B, int N, float alpha, float* dst, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(src && dst && src != dst, "pointers (not in place)");
  AVSI_REQUIRE(B > 0 && N > 0, "sizes");
  preemphasis_kernel<<<grid_for((long long)B * N), 256, 0, (cudaStream_t)stream>>>(src, B, N, alpha, dst);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_mfcc(const float* logmel, int64_t rows, int n_mel, int n_mfcc, float* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(logmel && out, "null pointer");
  AVSI_REQUIRE(rows > 0 && n_mel > 0 && n_mfcc > 0 && n_mfcc <= n_mel && n_mfcc * n_mel * 4 <= 48 * 1024, "sizes");
  mfcc_kernel<<<grid_for(rows * n_mfcc), 256, n_mfcc * n_mel * sizeof(float), (cudaStream_t)stream>>>(logmel, rows, n_mel,
                                                                                                    n_mfcc, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_delta_features(const float* src, int ld_src, float* dst, int ld_dst, int B, int T, int F, int N,
                                   void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(src && dst, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && F > 0 && N >= 1 && ld_src >= F && ld_dst >= F, "sizes");
  int den = 0;
  for (int i = 1; i <= N; ++i) den += 2 * i * i;
  delta_kernel<<<grid_for((long long)B * T * F), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, B, T, F, N,
                                                                                1.0f / (float)den);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_spectrogram(const float* stft_c64, int64_t n, float power, int log_flag, float* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(stft_c64 && out && n > 0, "args");
  int blocks = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 16);
  spectrogram_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(stft_c64), (long long)n, power, log_flag, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_log_mel(const float* spec, const float* mel_w, int64_t rows, int nbins, int n_mel, float eps, float* out,
                            void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(spec && mel_w && out && rows > 0 && nbins > 0 && n_mel > 0, "args");
  int blocks = (int)min((long long)(rows + 7) / 8, (long long)num_sms() * 16);
  log_mel_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(spec, mel_w, (long long)rows, nbins, n_mel, eps, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

// out[i] = keep[i] != 0 ? a[i] : b[i] over n complex64 values (b == NULL: zero): the phase source of one consistency
// iteration of phase_reconstruction.refine_phase -- the known spectrum in the reliable bins, the re-analysed one in the holes
namespace avsi {
__global__ void __launch_bounds__(256)
select_c64_kernel(const float2* __restrict__ a, const float2* __restrict__ b, const float* __restrict__ keep, long long n,
                  float2* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (keep[i] != 0.f) ? a[i] : (b ? b[i] : make_float2(0.f, 0.f));
}
}  // namespace avsi

extern "C" int avsi_select_c64(const float* a_c64, const float* b_c64, const float* keep, int64_t n, float* out_c64, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(a_c64 && keep && out_c64 && n > 0, "args");
  int blocks = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 16);
  select_c64_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(a_c64), reinterpret_cast<const float2*>(b_c64),
                                                             keep, (long long)n, reinterpret_cast<float2*>(out_c64));
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
