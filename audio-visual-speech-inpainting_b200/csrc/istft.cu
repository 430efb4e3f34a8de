// Fused waveform reconstruction: (denormalise) magnitude + phase of a reference STFT ->
// inverse rFFT-512 -> inverse window -> overlap-add, one kernel.
//
// Replaces get_sources / reconstruct_sources (audio_processing.py:145-164:
// tf.complex(mag cos, mag sin) -> tf.contrib.signal.inverse_stft with
// inverse_stft_window_fn -> slice) and the enhanced_sources graph of models.py:181-197
// (exp(pred * std + mean), angle(stft * mask)).  Same 16-threads-per-frame four-step FFT as
// the forward front end, run backwards.  unit phase vector = X/|X| (1+0i where X == 0, as
// tf.angle(0) = 0), so no atan2 / sincos is needed.
#include "common.cuh"

namespace avsi {

struct cpx2 {
  float x, y;
};
__device__ __forceinline__ cpx2 c2add(cpx2 a, cpx2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx2 c2sub(cpx2 a, cpx2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx2 c2mul(cpx2 a, cpx2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cpx2 c2roti(cpx2 a) { return {-a.y, a.x}; }   // * (+i)

__device__ __forceinline__ void ifft4(cpx2& a, cpx2& b, cpx2& c, cpx2& d) {
  cpx2 s0 = c2add(a, c), s1 = c2sub(a, c), s2 = c2add(b, d), s3 = c2roti(c2sub(b, d));
  a = c2add(s0, s2);
  c = c2sub(s0, s2);
  b = c2add(s1, s3);
  d = c2sub(s1, s3);
}
// unnormalised inverse 16-point DFT (e^{+2 pi i nk/16}), natural order in and out
__device__ __forceinline__ void ifft16(cpx2 (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) ifft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  const cpx2 w1 = {c1, s1}, w2 = {r2, r2}, w3 = {s1, c1}, w6 = {-r2, r2}, w9 = {-c1, -s1};
  v[5] = c2mul(v[5], w1);
  v[6] = c2mul(v[6], w2);
  v[7] = c2mul(v[7], w3);
  v[9] = c2mul(v[9], w2);
  v[10] = c2roti(v[10]);
  v[11] = c2mul(v[11], w6);
  v[13] = c2mul(v[13], w3);
  v[14] = c2mul(v[14], w6);
  v[15] = c2mul(v[15], w9);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) ifft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  cpx2 t[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = t[i];
}

constexpr int IS_THREADS = 256;
constexpr int IS_FRAMES = 16;

struct IstftSmem {
  float2 tw[512];
  float2 X[IS_FRAMES][260];      // spectrum, then Z (natural order)
  float2 xch[IS_FRAMES][16 * 17];
};

__global__ void __launch_bounds__(IS_THREADS) istft_kernel(avsi_istft_args p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < 512; i += IS_THREADS) sm.tw[i] = reinterpret_cast<const float2*>(p.twiddle)[i];
  __syncthreads();
  const int fl = tid >> 4, q = tid & 15;
  const long long total = (long long)p.B * p.T;
  const long long out_len = (long long)(p.T - 1) * p.hop + p.frame_len;
  const long long n_keep = (p.num_samples > 0) ? min((long long)p.num_samples, out_len) : out_len;
  const long long out_stride = (p.num_samples > 0) ? p.num_samples : out_len;

  for (long long g0 = (long long)blockIdx.x * IS_FRAMES; g0 < total; g0 += (long long)gridDim.x * IS_FRAMES) {
    const long long g = g0 + fl;
    const bool live = g < total;
    const int b = live ? (int)(g / p.T) : 0;
    const int t = live ? (int)(g - (long long)b * p.T) : 0;
    float2* X = sm.X[fl];
    // ---- spectrum: mag * unit phase ---------------------------------------------------------
    const int nb = (q == 0) ? 17 : 16;
    for (int m = 0; m < nb; ++m) {
      const int k = (m < 16) ? (q + 16 * m) : 256;
      float2 v = make_float2(0.f, 0.f);
      if (live && k < p.F) {
        const long long idx = g * p.F + k;
        float mag = p.mag[idx];
        if (p.mean) mag = expf(fmaf(mag, __ldg(p.stdev + k), __ldg(p.mean + k)));
        float2 ph = reinterpret_cast<const float2*>(p.phase_src)[idx];
        if (p.mask) {
          float mk = p.mask[idx];
          ph.x *= mk;
          ph.y *= mk;
        }
        float nrm = sqrtf(ph.x * ph.x + ph.y * ph.y);
        float ux = 1.f, uy = 0.f;
        if (nrm > 0.f) {
          ux = ph.x / nrm;
          uy = ph.y / nrm;
        }
        v = make_float2(mag * ux, mag * uy);
        if (k == 0 || k == 256) v.y = 0.f;      // C2R ignores the imaginary part of DC / Nyquist
      }
      X[k] = v;
    }
    __syncwarp();
    // ---- Z[k] = E[k] + i O[k], E = (X[k] + conj X[N-k])/2, O = (X[k] - conj X[N-k])/2 * conj(W512^k)
    cpx2 zk[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int k = q + 16 * m;
      float2 a = X[k], bq = X[256 - k];
      cpx2 e = {0.5f * (a.x + bq.x), 0.5f * (a.y - bq.y)};
      cpx2 d = {0.5f * (a.x - bq.x), 0.5f * (a.y + bq.y)};
      float2 w = sm.tw[k];
      cpx2 o = c2mul(d, cpx2{w.x, -w.y});
      zk[m] = {e.x - o.y, e.y + o.x};
    }
    __syncwarp();
    // ---- inverse four-step: n = n1 + 16 n2?  use k = k1 + 16 k2 (k1 = q here), n = 16 n1 + n2:
    // z[16 n1 + n2] = sum_k1 W256^-(n2 k1) W16^-(n1 k1) [ sum_k2 Z[k1 + 16 k2] W16^-(n2 k2) ]
    // thread q = k1 holds Z[k1 + 16 k2] (k2 = m): IFFT16 over k2 -> index n2
    ifft16(zk);
    float2* xc = sm.xch[fl];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      float2 w = sm.tw[(2 * q * n2) & 511];
      cpx2 r = c2mul(zk[n2], cpx2{w.x, -w.y});
      xc[n2 * 17 + q] = make_float2(r.x, r.y);
    }
    __syncwarp();
    // thread q = n2 reads over k1, IFFT16 over k1 -> index n1
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      float2 a = xc[q * 17 + k1];
      zk[k1] = {a.x, a.y};
    }
    ifft16(zk);
    // zk[n1] = 256 * z[16 n1 + q] ; x[2n] = Re, x[2n+1] = Im
    if (live) {
      float* dst = p.out + (long long)b * out_stride;
      const long long base = (long long)t * p.hop;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int i0 = 32 * n1 + 2 * q;
        if (i0 < p.frame_len && base + i0 < n_keep)
          atomicAdd(dst + base + i0, zk[n1].x * (1.f / 256.f) * __ldg(p.inv_window + i0));
        if (i0 + 1 < p.frame_len && base + i0 + 1 < n_keep)
          atomicAdd(dst + base + i0 + 1, zk[n1].y * (1.f / 256.f) * __ldg(p.inv_window + i0 + 1));
      }
    }
    __syncwarp();
  }
}

}  // namespace avsi

extern "C" int avsi_istft_fwd(const avsi_istft_args* a, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(a != nullptr, "args");
  AVSI_REQUIRE(a->nfft == 512, "nfft must be 512");
  AVSI_REQUIRE(a->frame_len > 0 && a->frame_len <= 512 && a->hop > 0, "frame_len/hop");
  AVSI_REQUIRE(a->B > 0 && a->T > 0 && a->F > 0 && a->F <= 257, "B,T,F");
  AVSI_REQUIRE(a->mag && a->phase_src && a->inv_window && a->twiddle && a->out, "null pointer");
  AVSI_REQUIRE((a->mean == nullptr) == (a->stdev == nullptr), "mean and std together");
  cudaStream_t st = (cudaStream_t)stream;
  const long long out_len = (long long)(a->T - 1) * a->hop + a->frame_len;
  const long long out_stride = (a->num_samples > 0) ? a->num_samples : out_len;
  AVSI_CUDA(cudaMemsetAsync(a->out, 0, sizeof(float) * out_stride * a->B, st));
  const int smem = (int)sizeof(IstftSmem);
  static bool attr_done = false;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const long long groups = ((long long)a->B * a->T + IS_FRAMES - 1) / IS_FRAMES;
  long long grid = (long long)num_sms() * 4;
  if (grid > groups) grid = groups;
  istft_kernel<<<(unsigned)grid, IS_THREADS, smem, st>>>(*a);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
