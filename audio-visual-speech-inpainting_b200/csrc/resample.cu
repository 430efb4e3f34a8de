// Fourier-domain resampling of whole recordings: downsampling(samples, sample_rate, downsample_rate) of
// audio_processing.py:9-16, i.e. scipy.signal.resample(x, num) for real x (dataset preparation:
// audio_feat_preprocessing.py:82,189, tfrecord_utils.py).  GRID recordings are 150 000 samples at 50 kHz
// (2^4 . 3 . 5^5 points) resampled to 48 000 (2^7 . 3 . 5^3): neither length is a power of two, so both transforms
// are evaluated as chirp-z (Bluestein) convolutions on power-of-two Stockham FFTs.  Arithmetic is complex double --
// the reference's result is float64 (pocketfft), the arrays are a few MB and live in L2, and the step is offline.
//
//   X[k]  = sum_n x[n] e^{-2 pi i k n / Nx}                          k < N/2 + 1,  N = min(num, Nx)
//   Y     = X[: N/2 + 1];  N even: Y[N/2] *= 2 (num < Nx) or 0.5 (num > Nx)      (scipy/signal/_signaltools.py, resample)
//   y[m]  = irfft(Y, num)[m] * num / Nx = Re( sum_k Z[k] e^{+2 pi i k m / num} ) / Nx,  Z = Hermitian extension of Y
//
// Chirp-z (j k = (j^2 + k^2 - (k - j)^2) / 2):  sum_j a[j] e^{s 2 pi i j k / n} = w_s[k] . sum_j (a[j] w_s[j]) w_{-s}[k - j],
// w_s[m] = e^{s i pi m^2 / n};
// the exponent m^2 is reduced mod 2n in 64-bit integers before the angle is formed, so the chirps are exact to double
// rounding for any length.
#include "common.cuh"

namespace avsi {

struct cd {
  double x, y;
};
__device__ __forceinline__ cd cmul(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cd{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return cd{a.x - b.x, a.y - b.y}; }
// e^{sign . i . pi . m^2 / n}
__device__ __forceinline__ cd chirp(long long m, long long n, double sign) {
  const long long q = (m * m) % (2 * n);
  double s, c;
  sincospi((double)q / (double)n, &s, &c);
  return cd{c, sign * s};
}

// One Stockham radix-4 pass (radix 2 when R == 2) of a length-M transform, batch = blockIdx.y.  dir = -1 forward, +1 inverse.
template <int R>
__global__ void __launch_bounds__(256) stockham_pass_kernel(const cd* __restrict__ in, cd* __restrict__ out, int M, int Ns, double dir) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = M / R;
  if (j >= q) return;
  in += (long long)blockIdx.y * M;
  out += (long long)blockIdx.y * M;
  const int k = j % Ns;
  cd v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double s, c;
    sincospi(2.0 * (double)k * (double)r / (double)(Ns * R), &s, &c);
    v[r] = cmul(in[j + r * q], cd{c, dir * s});
  }
  if (R == 2) {
    const cd a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else {
    const cd a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d0 = csub(v[1], v[3]);
    const cd d = cd{-dir * d0.y, dir * d0.x};            // (dir . i) . d0
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
  }
  const long long j0 = (long long)(j / Ns) * Ns * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) out[j0 + (long long)r * Ns] = v[r];
}

// In-place (ping-pong through tmp) power-of-two FFT of `batch` rows; returns the buffer that holds the result.
static cd* fft_pow2(cd* buf, cd* tmp, int M, int batch, double dir, cudaStream_t st) {
  int Ns = 1;
  cd *src = buf, *dst = tmp;
  while (Ns < M) {
    if (M / Ns >= 4 && (M / Ns) != 2 && ((M / Ns) & 3) == 0) {
      dim3 grid((M / 4 + 255) / 256, batch);
      stockham_pass_kernel<4><<<grid, 256, 0, st>>>(src, dst, M, Ns, dir);
      Ns *= 4;
    } else {
      dim3 grid((M / 2 + 255) / 256, batch);
      stockham_pass_kernel<2><<<grid, 256, 0, st>>>(src, dst, M, Ns, dir);
      Ns *= 2;
    }
    cd* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// a[m] = in[m] . e^{sign i pi m^2 / n} (m < n), 0 beyond (sign = sign of the DFT exponent);  in: real (REAL) or complex double
template <bool REAL>
__global__ void __launch_bounds__(256) chirp_in_kernel(const void* __restrict__ in, long long n, int M, double sign, cd* __restrict__ a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long b = blockIdx.y;
  cd v{0.0, 0.0};
  if (m < n) {
    const cd w = chirp(m, n, sign);
    if (REAL) {
      const double x = reinterpret_cast<const double*>(in)[b * n + m];
      v = cd{x * w.x, x * w.y};
    } else {
      v = cmul(reinterpret_cast<const cd*>(in)[b * n + m], w);
    }
  }
  a[b * M + m] = v;
}
// b[m] = e^{-sign i pi d^2 / n}, d = min(m, M - m) < n; 0 elsewhere (the chirp filter, wrapped)
__global__ void __launch_bounds__(256) chirp_filter_kernel(long long n, int M, double sign, cd* __restrict__ b) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long d = m < M - m ? m : M - m;
  b[m] = d < n ? chirp(d, n, -sign) : cd{0.0, 0.0};
}
__global__ void __launch_bounds__(256) pointwise_mul_kernel(cd* __restrict__ a, const cd* __restrict__ bf, int M) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  cd* row = a + (long long)blockIdx.y * M;
  row[m] = cmul(row[m], bf[m]);
}
// Z[k] of the target spectrum (Hermitian extension of Y = X[: N/2+1] with scipy's Nyquist rule) from the convolution
// result c (X[k] = conj-chirp . c[k] / M1): Z[k] = Y[k] (k <= half), conj(Y[num - k]) (k >= num - half), 0 between.
__global__ void __launch_bounds__(256)
spectrum_kernel(const cd* __restrict__ c, long long nx, int M1, long long num, cd* __restrict__ z) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= num) return;
  const long long b = blockIdx.y;
  const long long N = num < nx ? num : nx, half = N / 2;
  const long long src = k <= num - k ? k : num - k;       // bin of Y this entry mirrors
  cd v{0.0, 0.0};
  if (src <= half) {
    const cd w = chirp(src, nx, -1.0);
    cd x = cmul(c[b * M1 + src], w);
    x.x /= (double)M1;
    x.y /= (double)M1;
    if ((N & 1) == 0 && src == half) {
      const double f = num < nx ? 2.0 : (num > nx ? 0.5 : 1.0);
      x.x *= f;
      x.y *= f;
    }
    if (src == 0 || ((num & 1) == 0 && src == num / 2)) x.y = 0.0;   // irfft ignores the imaginary part of DC / Nyquist
    v = (k == src) ? x : cd{x.x, -x.y};
  }
  z[b * num + k] = v;
}
// y[m] = Re(chirp . c[m] / M2) / nx
__global__ void __launch_bounds__(256)
resample_out_kernel(const cd* __restrict__ c, long long nx, int M2, long long num, double* __restrict__ y) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= num) return;
  const long long b = blockIdx.y;
  const cd w = chirp(m, num, 1.0);
  const cd v = cmul(c[b * M2 + m], w);
  y[b * num + m] = v.x / ((double)M2 * (double)nx);
}

static int next_pow2(long long v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace avsi

using namespace avsi;

extern "C" int64_t avsi_resample_workspace_bytes(int batch, int64_t n_in, int64_t n_out) {
  if (batch < 1 || n_in < 1 || n_out < 1 || n_in > (1ll << 27) || n_out > (1ll << 27)) return -1;
  const long long M1 = next_pow2(2 * n_in - 1), M2 = next_pow2(2 * n_out - 1);
  const long long M = M1 > M2 ? M1 : M2;
  // a (batch), tmp (batch), filter + its tmp, spectrum z (batch)
  return (int64_t)sizeof(cd) * (2 * batch * M + 2 * M + batch * n_out) + 256;
}

extern "C" int avsi_resample_fft(const double* x, int batch, int64_t n_in, double* y, int64_t n_out, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  if (!x || !y || !workspace) return set_error(AVSI_ERR_INVALID, "%s: null pointer%s", __func__, "");
  const long long need = avsi_resample_workspace_bytes(batch, n_in, n_out);
  if (need < 0) return set_error(AVSI_ERR_INVALID, "%s: bad sizes%s", __func__, "");
  if (workspace_bytes < need) return set_error(AVSI_ERR_INVALID, "%s: workspace too small%s", __func__, "");
  cudaStream_t st = (cudaStream_t)stream;
  const int M1 = next_pow2(2 * n_in - 1), M2 = next_pow2(2 * n_out - 1);
  const long long M = M1 > M2 ? M1 : M2;
  cd* base = reinterpret_cast<cd*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  cd* a = base;
  cd* tmp = a + (long long)batch * M;
  cd* filt = tmp + (long long)batch * M;
  cd* ftmp = filt + M;
  cd* z = ftmp + M;

  // ---- forward DFT of the recording (length n_in) ---------------------------------------------------------------
  chirp_in_kernel<true><<<dim3((M1 + 255) / 256, batch), 256, 0, st>>>(x, n_in, M1, -1.0, a);
  chirp_filter_kernel<<<(M1 + 255) / 256, 256, 0, st>>>(n_in, M1, -1.0, filt);
  cd* A = fft_pow2(a, tmp, M1, batch, -1.0, st);
  cd* F = fft_pow2(filt, ftmp, M1, 1, -1.0, st);
  pointwise_mul_kernel<<<dim3((M1 + 255) / 256, batch), 256, 0, st>>>(A, F, M1);
  cd* C = fft_pow2(A, A == a ? tmp : a, M1, batch, 1.0, st);
  // ---- target spectrum and inverse DFT (length n_out) -------------------------------------------------------------
  spectrum_kernel<<<dim3((unsigned)((n_out + 255) / 256), batch), 256, 0, st>>>(C, n_in, M1, n_out, z);
  cd* a2 = (C == a) ? tmp : a;                          // the buffer the convolution result does not occupy
  cd* t2 = (C == a) ? a : tmp;                          // C is consumed: its buffer is scratch from here on
  chirp_in_kernel<false><<<dim3((M2 + 255) / 256, batch), 256, 0, st>>>(z, n_out, M2, 1.0, a2);
  chirp_filter_kernel<<<(M2 + 255) / 256, 256, 0, st>>>(n_out, M2, 1.0, filt);
  cd* A2 = fft_pow2(a2, t2, M2, batch, -1.0, st);
  cd* F2 = fft_pow2(filt, ftmp, M2, 1, -1.0, st);
  pointwise_mul_kernel<<<dim3((M2 + 255) / 256, batch), 256, 0, st>>>(A2, F2, M2);
  cd* C2 = fft_pow2(A2, A2 == a2 ? t2 : a2, M2, batch, 1.0, st);
  resample_out_kernel<<<dim3((unsigned)((n_out + 255) / 256), batch), 256, 0, st>>>(C2, n_in, M2, n_out, y);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
