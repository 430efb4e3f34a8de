// Fused spectrogram front end: frame -> Hann window -> rFFT-512 -> |.|^p -> log -> normalise
// -> mask -> (concat video) -> fp32 / time-major fp16 outputs, one kernel, frames never in HBM.
//
// Replaces get_stft / get_spectrogram / get_log_mel_spectrogram (audio_processing.py:25-72),
// the normalise + mask + concat of models.py:30-45 and the complex mask of masking.py:42.
//
// Mapping: 16 threads per frame, 16 frames per 256-thread CTA, persistent grid-stride loop
// over groups of 16 consecutive global frame indices g = b*T + t.  The 512-point real FFT is
// a 256-point complex FFT of z[n] = x[2n] + i x[2n+1] done as a four-step 16x16 FFT: every
// thread runs two 16-point FFTs in registers with ONE shared-memory exchange in between,
// then a second exchange for the real-FFT split (needs Z[k] and Z[256-k]).
// Global loads are 8-byte, 128 B contiguous per frame per instruction (the 50 % frame
// overlap is served by L1/L2, not HBM); stores are 64 B contiguous per frame per instruction.
#include "common.cuh"

namespace avsi {

struct cpx {
  float x, y;
};
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// multiply by -i (forward FFT quarter turn) / +i (inverse)
template <bool INV>
__device__ __forceinline__ cpx rot90(cpx a) {
  return INV ? cpx{-a.y, a.x} : cpx{a.y, -a.x};
}

template <bool INV>
__device__ __forceinline__ void fft4(cpx& a, cpx& b, cpx& c, cpx& d) {
  // in: x0..x3, out: X0..X3 (natural order)
  cpx s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = rot90<INV>(csub(b, d));
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = cadd(s1, s3);
  d = csub(s1, s3);
}

// 16-point FFT, in place, natural order in and out.  16 = 4 x 4:
// X[k1 + 4 k2] = sum_n2 W16^(n2 k1) [ sum_n1 x[4 n1 + n2] W4^(n1 k1) ] W4^(n2 k2)
template <bool INV>
__device__ __forceinline__ void fft16(cpx (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // v[4*k1 + n2] now holds A[k1][n2]; twiddle by W16^(n2*k1)  (conjugated for inverse)
  const float sg = INV ? 1.f : -1.f;
  const cpx w1 = {c1, sg * s1}, w2 = {r2, sg * r2}, w3 = {s1, sg * c1};
  const cpx w6 = {-r2, sg * r2}, w9 = {-c1, sg * (-s1)};
  v[4 * 1 + 1] = cmul(v[4 * 1 + 1], w1);
  v[4 * 1 + 2] = cmul(v[4 * 1 + 2], w2);
  v[4 * 1 + 3] = cmul(v[4 * 1 + 3], w3);
  v[4 * 2 + 1] = cmul(v[4 * 2 + 1], w2);
  v[4 * 2 + 2] = rot90<INV>(v[4 * 2 + 2]);  // W16^4 = -i
  v[4 * 2 + 3] = cmul(v[4 * 2 + 3], w6);
  v[4 * 3 + 1] = cmul(v[4 * 3 + 1], w3);
  v[4 * 3 + 2] = cmul(v[4 * 3 + 2], w6);
  v[4 * 3 + 3] = cmul(v[4 * 3 + 3], w9);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4<INV>(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // v[4*k1 + k2] = X[k1 + 4*k2] -> reorder to natural order
  cpx t[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = t[i];
}

constexpr int FE_THREADS = 256;
constexpr int FE_FRAMES = 16;          // frames per CTA iteration
constexpr int FE_XROW = 17;            // padded row (complex) of the 16x16 exchange
constexpr int FE_XCH = 16 * FE_XROW;   // 272 complex per frame

struct FrontendSmem {
  float2 tw[512];                       // exp(-2 pi i m / 512)
  float win[512];                       // window, zero beyond frame_len
  float2 xch[FE_FRAMES][FE_XCH];        // per-frame exchange / Z buffer
  float pw[FE_FRAMES][260];             // per-frame power spectrum (mel path only)
};

__global__ void __launch_bounds__(FE_THREADS) frontend_kernel(avsi_frontend_args p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FrontendSmem& sm = *reinterpret_cast<FrontendSmem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < 512; i += FE_THREADS) {
    sm.tw[i] = reinterpret_cast<const float2*>(p.twiddle)[i];
    sm.win[i] = (i < p.frame_len) ? p.window[i] : 0.f;
  }
  __syncthreads();

  const int fl = tid >> 4;   // local frame
  const int q = tid & 15;    // lane within the frame
  const long long total = (long long)p.B * p.T;
  const int I = p.F + (p.video ? p.V : 0);
  const bool want_mel = p.logmel_out != nullptr;
  float holes = 0.f;

  for (long long g0 = (long long)blockIdx.x * FE_FRAMES; g0 < total; g0 += (long long)gridDim.x * FE_FRAMES) {
    const long long g = g0 + fl;
    const bool live = g < total;
    const int b = live ? (int)(g / p.T) : 0;
    const int t = live ? (int)(g - (long long)b * p.T) : 0;

    // ---- load + window: thread q holds z[16 n1 + q], n1 = 0..15 -------------------------
    cpx v[16];
    {
      const long long base = (long long)b * p.N + (long long)t * p.hop;   // sample index of frame start
      const float* src = p.wav + base;
      const int avail = live ? (int)min((long long)p.frame_len, (long long)p.N - (long long)t * p.hop) : 0;
      const bool vec_ok = ((base & 1LL) == 0);
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int i0 = 32 * n1 + 2 * q;
        float x0 = 0.f, x1 = 0.f;
        if (i0 + 1 < avail) {
          if (vec_ok) {
            float2 xx = __ldg(reinterpret_cast<const float2*>(src + i0));
            x0 = xx.x;
            x1 = xx.y;
          } else {
            x0 = __ldg(src + i0);
            x1 = __ldg(src + i0 + 1);
          }
        } else if (i0 < avail) {
          x0 = __ldg(src + i0);
        }
        v[n1].x = x0 * sm.win[i0];
        v[n1].y = x1 * sm.win[i0 + 1];
      }
    }
    // ---- step 1: FFT16 over n1, twiddle W256^(q*k1), exchange ---------------------------
    fft16<false>(v);
    float2* xc = sm.xch[fl];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      float2 w = sm.tw[(2 * q * k1) & 511];
      cpx r = cmul(v[k1], cpx{w.x, w.y});
      xc[k1 * FE_XROW + q] = make_float2(r.x, r.y);
    }
    __syncwarp();
    // ---- step 2: thread q = k1 reads A[k1][n2], FFT16 over n2 -> Z[k1 + 16 k2] ----------
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      float2 a = xc[q * FE_XROW + n2];
      v[n2] = cpx{a.x, a.y};
    }
    fft16<false>(v);
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) xc[q + 16 * k2] = make_float2(v[k2].x, v[k2].y);   // natural order Z[k]
    __syncwarp();

    // ---- real-FFT split + epilogue: thread q handles bins k = q + 16 m (and 256 for q == 0) -
    const long long row_bt = g;                                   // (b*T + t)
    const long long row_tb = (long long)t * p.B + b;              // time-major row
    const int nb = (q == 0) ? 17 : 16;
    for (int m = 0; m < nb; ++m) {
      const int k = (m < 16) ? (q + 16 * m) : 256;
      float2 zk = xc[k & 255];
      float2 zn = xc[(256 - k) & 255];
      // E = (Z[k] + conj Z[N-k]) / 2 ; O = -i (Z[k] - conj Z[N-k]) / 2 ; X = E + W512^k O
      cpx e = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
      cpx d = {0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y)};
      cpx o = {d.y, -d.x};
      float2 w = sm.tw[k];   // k <= 256
      cpx X = cadd(e, cmul(o, cpx{w.x, w.y}));
      if (k >= p.F && !(want_mel && k < 257)) continue;
      float mval = 1.f;
      if (live && p.mask && k < p.F) {
        mval = __ldg(p.mask + row_bt * p.F + k);
        holes += 1.f - mval;
      }
      if (live && p.stft_out && k < p.F) {
        float sc = p.stft_masked ? mval : 1.f;
        reinterpret_cast<float2*>(p.stft_out)[row_bt * p.F + k] = make_float2(X.x * sc, X.y * sc);
      }
      float s = X.x * X.x + X.y * X.y;
      float mag;
      if (p.power == 2.f) mag = s;
      else {
        mag = sqrtf(s);
        if (p.power != 1.f) mag = powf(mag, p.power);
      }
      if (want_mel) sm.pw[fl][k] = p.mel_masked ? mag * mval : mag;
      if (!live || k >= p.F) continue;
      float val = p.log_flag ? logf(mag + 1e-6f) : mag;
      if (p.mean) val = (val - __ldg(p.mean + k)) / __ldg(p.stdev + k);
      if (p.spec_out) p.spec_out[row_bt * p.F + k] = val;
      const float mv = val * mval;
      if (p.feat_out) p.feat_out[row_bt * I + k] = mv;
      if (p.xh_out && !p.xh_video_only) p.xh_out[row_tb * p.ldx + k] = __half_as_ushort(__float2half_rn(mv));
    }
    // ---- video columns and zero padding ---------------------------------------------------
    if (live && p.video) {
      const float* vs = p.video + row_bt * p.V;
      for (int c = q; c < p.V; c += 16) {
        float x = __ldg(vs + c);
        if (p.feat_out) p.feat_out[row_bt * I + p.F + c] = x;
        if (p.xh_out) p.xh_out[row_tb * p.ldx + (p.xh_video_only ? 0 : p.F) + c] = __half_as_ushort(__float2half_rn(x));
      }
    }
    if (live && p.xh_out && !p.xh_skip_pad)
      for (int c = (p.xh_video_only ? p.V : I) + q; c < p.ldx; c += 16) p.xh_out[row_tb * p.ldx + c] = 0;
    // ---- log-mel ------------------------------------------------------------------------------
    if (want_mel) {
      __syncwarp();
      if (live) {
        for (int mb = q; mb < p.n_mel; mb += 16) {
          float acc = 0.f;
          for (int k = 0; k < 257; ++k) {
            float w = __ldg(p.mel_w + k * p.n_mel + mb);
            acc = fmaf(sm.pw[fl][k], w, acc);
          }
          p.logmel_out[row_bt * p.n_mel + mb] = logf(acc + p.mel_eps);
        }
      }
    }
    __syncwarp();
  }
  if (p.hole_count) {
    holes = warp_sum(holes);
    if ((tid & 31) == 0 && holes != 0.f) atomicAdd(p.hole_count, (double)holes);   // integer counts in a double: exact in any order
  }
}

// ---------------------------------------------------------------------------------------------------------
// Training hot path (models.py:30-45 exactly: |STFT| -> log -> normalise -> mask -> concat video, fp32 loss
// target + fp16 time-major network input), specialised so that the memory system stays busy:
//   * the samples of the NEXT group of 16 frames are fetched with cp.async (LDGSTS, zero-filled past the end of
//     the utterance) into a second shared-memory buffer while the current group is transformed: no registers,
//     no scoreboard stall at the head of the iteration;
//   * the mask row of a frame is requested before its FFT and consumed after it;
//   * zero-padded FFT inputs (samples 384..511) are compile-time zeros, the first radix-16 pass is pruned;
//   * complex arithmetic is PACKED fp32x2 (FADD2 / FMUL2 / FFMA2 on (re, im) register pairs; the half swap of a
//     quarter turn or a complex product is a free operand swizzle): half the issue slots of the scalar butterflies;
//   * sqrt / log are single MUFU ops (sqrt.approx.ftz, lg2.approx.ftz: <= 2 ulp, the 1e-5 budget is on relative L2);
//   * (b, t) of a thread's frame advance incrementally (no division in the loop); a thread past the last frame
//     recomputes the last frame (same values to the same addresses) instead of carrying predicates.
typedef unsigned long long c2;          // packed complex / pair: lo = re (or first), hi = im (or second)

__device__ __forceinline__ c2 pk(float lo, float hi) {
  c2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(c2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ c2 swp(c2 v) {
  float a, b;
  upk(v, a, b);
  return pk(b, a);
}
__device__ __forceinline__ c2 add2(c2 a, c2 b) {
  c2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 sub2(c2 a, c2 b) {
  c2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 mul2(c2 a, c2 b) {
  c2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c) {
  c2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// v * (wr + i wi) with the constant given as (wr, wr) and (-wi, wi)
__device__ __forceinline__ c2 cmul2(c2 v, c2 wrr, c2 wii) { return fma2(swp(v), wii, mul2(v, wrr)); }

__device__ __forceinline__ void fft4p(c2& a, c2& b, c2& c, c2& d, const c2 pm, const c2 mp) {
  // forward radix-4, natural order; pm = (1, -1), mp = (-1, 1): s1 -/+ i (b - d) as one FFMA2 on the swapped pair
  const c2 s0 = add2(a, c), s1 = sub2(a, c), s2 = add2(b, d), t = swp(sub2(b, d));
  a = add2(s0, s2);
  c = sub2(s0, s2);
  b = fma2(t, pm, s1);
  d = fma2(t, mp, s1);
}

__device__ __forceinline__ void fft16p(c2 (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  const c2 pm = pk(1.f, -1.f), mp = pk(-1.f, 1.f);
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) fft4p(v[n2], v[4 + n2], v[8 + n2], v[12 + n2], pm, mp);
  // forward twiddles W16^(n2 k1): w1 = (c1,-s1), w2 = (r2,-r2), w3 = (s1,-c1), w6 = (-r2,-r2), w9 = (-c1, s1)
  const c2 w1r = pk(c1, c1), w1i = pk(s1, -s1), w2r = pk(r2, r2), w2i = pk(r2, -r2), w3r = pk(s1, s1), w3i = pk(c1, -c1);
  const c2 w6r = pk(-r2, -r2), w9r = pk(-c1, -c1), w9i = pk(-s1, s1);
  v[5] = cmul2(v[5], w1r, w1i);
  v[6] = cmul2(v[6], w2r, w2i);
  v[7] = cmul2(v[7], w3r, w3i);
  v[9] = cmul2(v[9], w2r, w2i);
  v[10] = mul2(swp(v[10]), pm);           // W16^4 = -i
  v[11] = cmul2(v[11], w6r, w2i);
  v[13] = cmul2(v[13], w3r, w3i);
  v[14] = cmul2(v[14], w6r, w2i);
  v[15] = cmul2(v[15], w9r, w9i);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4p(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3], pm, mp);
  c2 t[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = t[i];
}

struct FrontendTrainSmem {
  float2 tw1p[16][16];                  // step-1 twiddles W256^(q k1) as [k1][q]
  float2 twp[128];                      // split twiddles W512^k
  ulonglong2 nrm2[128];                 // ((ln2/std_k, ln2/std_{256-k}), (-mean_k/std_k, -mean_{256-k}/std_{256-k}))
  float2 nrm128;                        // the self-mirrored bin
  c2 win2[192];                         // window as (w[2n], w[2n+1]) pairs
  c2 xch[FE_FRAMES][FE_XCH];
  c2 wavb[2][FE_FRAMES][192];           // sample pairs of the frame, double buffered
  int mlo[128], moff[128];              // log-mel variant: first bin and weight offset of every mel filter
  float mw[1024];                       // band weights
};

__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ float sqrt_ftz(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// MEL = true is the `fbanks` variant (models_asr.py:30-36, audio_feat_preprocessing.py:49-50): power spectrum (x mask
// when mel_masked) -> band-form mel projection out of shared memory -> log; the only HBM traffic is the samples in and
// [B,T,n_mel] out.
template <bool HAS_VIDEO, bool HOLES, bool MEL>
__global__ void __launch_bounds__(FE_THREADS, 2) frontend_train_kernel(avsi_frontend_args p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FrontendTrainSmem& sm = *reinterpret_cast<FrontendTrainSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const float2* twg = reinterpret_cast<const float2*>(p.twiddle);
  sm.tw1p[tid >> 4][tid & 15] = twg[(2 * (tid & 15) * (tid >> 4)) & 511];
  if (tid < 128) sm.twp[tid] = twg[tid];
  for (int i = tid; i < 192; i += FE_THREADS) sm.win2[i] = pk(p.window[2 * i], p.window[2 * i + 1]);
  if (!MEL) {
    const float ln2 = 0.69314718055994531f;
    if (tid < 128) {
      const float ia = 1.0f / p.stdev[tid], ib = 1.0f / p.stdev[256 - tid];
      sm.nrm2[tid] = make_ulonglong2(pk(ia * ln2, ib * ln2), pk(-p.mean[tid] * ia, -p.mean[256 - tid] * ib));
    }
    if (tid == 128) {
      const float is = 1.0f / p.stdev[128];
      sm.nrm128 = make_float2(is * ln2, -p.mean[128] * is);
    }
  } else {
    const int* hdr = reinterpret_cast<const int*>(p.mel_bands);
    if (tid < 128) {
      sm.mlo[tid] = hdr[tid];
      sm.moff[tid] = hdr[128 + tid];
    }
    const int nw = hdr[128 + p.n_mel];
    const float* wsrc = reinterpret_cast<const float*>(hdr + 256);
    for (int i = tid; i < nw; i += FE_THREADS) sm.mw[i] = wsrc[i];
  }
  const bool masked = !MEL || (p.mel_masked && p.mask);
  __syncthreads();

  const int fl = tid >> 4, q = tid & 15;
  const int total = p.B * p.T;                     // < 2^31 (checked by the launcher)
  const int stride = (int)gridDim.x * FE_FRAMES;
  const int dT = stride % p.T, dB = stride / p.T;
  float holes = 0.f;

  // frame of this thread in the group starting at g0; past the end it is the last frame again
  int g0 = (int)blockIdx.x * FE_FRAMES;
  int b, t;
  {
    const int g = min(g0 + fl, total - 1);
    b = g / p.T;
    t = g - b * p.T;
  }
  auto advance = [&](int& bb, int& tt, int g0n) {
    if (g0n + fl >= total) {
      bb = p.B - 1;
      tt = p.T - 1;
      return;
    }
    tt += dT;
    bb += dB;
    if (tt >= p.T) {
      tt -= p.T;
      bb += 1;
    }
  };
  auto issue = [&](int bb, int tt, int buf) {
    const long long base = (long long)bb * p.N + (long long)tt * p.hop;
    const int avail = p.N - tt * p.hop;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sm.wavb[buf][fl][q]);
    const float* src = p.wav + base + 2 * q;
    if (avail >= 384) {
#pragma unroll
      for (int n1 = 0; n1 < 12; ++n1) cp_async8(dst + (uint32_t)(128 * n1), src + 32 * n1);
    } else {
#pragma unroll
      for (int n1 = 0; n1 < 12; ++n1) {
        const int nb = max(0, min(8, (avail - 32 * n1 - 2 * q) * 4));
        cp_async8_zfill(dst + (uint32_t)(128 * n1), nb > 0 ? (const void*)(src + 32 * n1) : (const void*)p.wav, nb);
      }
    }
    cp_async_commit();
  };

  const c2 pm = pk(1.f, -1.f), mp = pk(-1.f, 1.f);
  int buf = 0;
  if (g0 < total) issue(b, t, 0);
  for (; g0 < total; g0 += stride, buf ^= 1) {
    const bool live = g0 + fl < total;
    int bn = b, tn = t;
    if (g0 + stride < total) {
      advance(bn, tn, g0 + stride);
      issue(bn, tn, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const long long row_bt = (long long)b * p.T + t, row_tb = (long long)t * p.B + b;
    // mask row of this frame: requested now, used after the transform
    // thread q owns the bin pairs (k, 256 - k), k = q + 16 m, m = 0..7 (+ the self-mirrored bin 128 for q == 0):
    // X[k] = E + W O and X[256 - k] = conj(E - W O) share Z[k], Z[256 - k] and the twiddle product
    float mva[8], mvb[8], mv128 = 1.f;
    if (masked) {
      const float* mrow = p.mask + row_bt * 257 + q;
      const float* mrev = p.mask + row_bt * 257 + 256 - q;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        mva[m] = __ldg(mrow + 16 * m);
        mvb[m] = __ldg(mrev - 16 * m);
      }
      if (q == 0) mv128 = __ldg(mrow + 128);
      if (HOLES && live) {
        float h = 1.f - mv128;
#pragma unroll
        for (int m = 0; m < 8; ++m) h += (1.f - mva[m]) + (1.f - mvb[m]);
        holes += h;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) mva[m] = mvb[m] = 1.f;
    }
    // video row of this frame (the second input stream): same early request
    float vv[9];
    if (HAS_VIDEO) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int c = q + 16 * i;
        vv[i] = (c < p.V) ? __ldg(p.video + row_bt * p.V + c) : 0.f;
      }
    }
    // ---- window; samples 384..511 are zero ----------------------------------------------------------------
    c2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 12; ++n1) v[n1] = mul2(sm.wavb[buf][fl][16 * n1 + q], sm.win2[16 * n1 + q]);
#pragma unroll
    for (int n1 = 12; n1 < 16; ++n1) v[n1] = 0ull;
    fft16p(v);
    c2* xc = sm.xch[fl];
    xc[q] = v[0];                                              // W256^0 = 1
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) {
      const float2 w = sm.tw1p[k1][q];
      xc[k1 * FE_XROW + q] = cmul2(v[k1], pk(w.x, w.x), pk(-w.y, w.y));
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = xc[q * FE_XROW + n2];
    fft16p(v);
    // thread q now holds Z[q + 16 k2] in v[k2].  The mirror Z[256 - k] of its bins k = q + 16 m, m = 0..7, is
    // v[15 - m] of lane (16 - q) & 15 (q != 0); lane 0 mirrors into itself: Z[256 - 16 m] = v[16 - m] (Z[0] for m = 0),
    // so lane 0 publishes those instead -- one shuffle per pair, no second trip through shared memory
    c2 zm[8];
    {
      const int src = (16 - q) & 15;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const c2 own = (m == 0) ? v[0] : v[16 - m];
        zm[m] = __shfl_sync(0xffffffffu, q == 0 ? own : v[15 - m], src, 16);
      }
    }
    if (MEL) __syncwarp();                                      // the power spectrum reuses the exchange buffer
    // ---- real-FFT split + |.| + log + normalise + mask ------------------------------------------------------------
    float* srow = MEL ? nullptr : p.spec_out + row_bt * 257 + q;
    float* srev = MEL ? nullptr : p.spec_out + row_bt * 257 + 256 - q;
    uint16_t* xrow = MEL ? nullptr : p.xh_out + row_tb * p.ldx + q;
    uint16_t* xrev = MEL ? nullptr : p.xh_out + row_tb * p.ldx + 256 - q;
    float* pwr = reinterpret_cast<float*>(xc);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = q + 16 * m;                                // 0 .. 127
      const c2 zk = v[m], zn = zm[m];
      const c2 e2 = fma2(zn, pm, zk);                          // 2 E = (zk.x + zn.x, zk.y - zn.y)
      const c2 d2 = fma2(zn, mp, zk);                          // 2 D = (zk.x - zn.x, zk.y + zn.y); 2 O = -i 2 D
      const float2 w = sm.twp[k];
      const c2 pp = fma2(swp(d2), pk(w.x, -w.x), mul2(d2, pk(w.y, w.y)));   // 2 W O
      const c2 x1 = add2(e2, pp), x2 = sub2(e2, pp);           // 2 X[k], 2 conj X[256 - k]
      float a0, a1, b0, b1;
      upk(mul2(x1, x1), a0, a1);
      upk(mul2(x2, x2), b0, b1);
      const float s1 = a0 + a1, s2 = b0 + b1;                  // 4 |X|^2
      if (MEL) {
        // power spectrum, parked in the .x slot of Z[k] (Z[k] and Z[256 - k] are read by this thread only, above)
        pwr[k] = 0.25f * s1 * mva[m];
        pwr[256 - k] = 0.25f * s2 * mvb[m];
        continue;
      }
      // |X| = sqrt(s) / 2; log(|X| + 1e-6); normalise (ln 2 folded into 1/std); mask
      float m1, m2;
      upk(fma2(pk(sqrt_ftz(s1), sqrt_ftz(s2)), pk(0.5f, 0.5f), pk(1e-6f, 1e-6f)), m1, m2);
      const ulonglong2 nr = sm.nrm2[k];
      const c2 val = fma2(pk(lg2_ftz(m1), lg2_ftz(m2)), nr.x, nr.y);
      float v1, v2, h1, h2;
      upk(val, v1, v2);
      upk(mul2(val, pk(mva[m], mvb[m])), h1, h2);
      srow[16 * m] = v1;
      srev[-16 * m] = v2;
      xrow[16 * m] = __half_as_ushort(__float2half_rn(h1));
      xrev[-16 * m] = __half_as_ushort(__float2half_rn(h2));
    }
    if (q == 0) {                                              // bin 128 mirrors itself: W512^128 = -i, |X| = |Z[128]|
      float zx, zy;
      upk(v[8], zx, zy);
      const float s = fmaf(zx, zx, zy * zy);
      if (MEL) {
        pwr[128] = s * mv128;
      } else {
        const float val = fmaf(lg2_ftz(sqrt_ftz(s) + 1e-6f), sm.nrm128.x, sm.nrm128.y);
        srow[128] = val;
        xrow[128] = __half_as_ushort(__float2half_rn(val * mv128));
      }
    }
    if (MEL) {
      __syncwarp();
      float* orow = p.logmel_out + row_bt * p.n_mel;
      for (int mb = q; mb < p.n_mel; mb += 16) {
        const int lo = sm.mlo[mb], o = sm.moff[mb], n = sm.moff[mb + 1] - o;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(pwr[lo + i], sm.mw[o + i], acc);
        orow[mb] = __logf(acc + p.mel_eps);
      }
    } else {
      if (HAS_VIDEO) {
        if (p.V <= 144) {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int c = q + 16 * i;
            if (c < p.V) xrow[257 - q + c] = __half_as_ushort(__float2half_rn(vv[i]));
          }
        } else {
          const float* vs = p.video + row_bt * p.V;
          for (int c = q; c < p.V; c += 16) xrow[257 - q + c] = __half_as_ushort(__float2half_rn(__ldg(vs + c)));
        }
      }
      if (!p.xh_skip_pad)
        for (int c = 257 + (HAS_VIDEO ? p.V : 0) + q; c < p.ldx; c += 16) xrow[c - q] = 0;
    }
    __syncwarp();
    b = bn;
    t = tn;
  }
  if (HOLES) {
    holes = warp_sum(holes);
    if ((tid & 31) == 0 && holes != 0.f) atomicAdd(p.hole_count, (double)holes);   // integer counts in a double: exact in any order
  }
}

}  // namespace avsi

extern "C" int avsi_frontend_fwd(const avsi_frontend_args* a, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(a != nullptr, "args");
  AVSI_REQUIRE(a->nfft == 512, "nfft must be 512");
  AVSI_REQUIRE(a->frame_len > 0 && a->frame_len <= 512, "frame_len in (0,512]");
  AVSI_REQUIRE(a->hop > 0, "hop > 0");
  AVSI_REQUIRE(a->F > 0 && a->F <= 257, "F in (0,257]");
  AVSI_REQUIRE(a->B > 0 && a->T > 0 && a->N > 0, "B,T,N > 0");
  AVSI_REQUIRE(a->wav && a->window && a->twiddle, "wav/window/twiddle");
  AVSI_REQUIRE((a->mean == nullptr) == (a->stdev == nullptr), "mean and std together");
  AVSI_REQUIRE(!a->logmel_out || (a->mel_w && a->n_mel > 0), "mel weights");
  AVSI_REQUIRE(!a->xh_out || a->ldx >= (a->xh_video_only ? 0 : a->F) + (a->video ? a->V : 0), "ldx too small");
  AVSI_REQUIRE(!a->xh_video_only || a->video, "xh_video_only needs video");
  const long long total = (long long)a->B * a->T;
  const long long groups = (total + FE_FRAMES - 1) / FE_FRAMES;
  const int smem = (int)sizeof(FrontendSmem);
  static bool attr_done = false;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  // training hot path: spec + fp16 time-major network input only, log-magnitude, GRID framing
  const bool train_path = a->frame_len == 384 && a->hop == 192 && a->F == 257 && a->power == 1.f && a->log_flag &&
                          a->mean && a->mask && a->spec_out && a->xh_out && !a->stft_out && !a->feat_out &&
                          !a->logmel_out && !a->xh_video_only && (a->N % 2 == 0) &&
                          ((uintptr_t)a->wav % 8 == 0);
  // log-mel variant of the same kernel: power-2 spectrum -> band-form mel -> log, nothing else requested
  const bool mel_path = a->frame_len == 384 && a->hop == 192 && a->power == 2.f && !a->log_flag && a->logmel_out &&
                        a->mel_bands && a->n_mel > 0 && a->n_mel < 128 && !a->spec_out && !a->xh_out && !a->stft_out &&
                        !a->feat_out && !a->hole_count && (a->N % 2 == 0) && ((uintptr_t)a->wav % 8 == 0) &&
                        (!a->mel_masked || !a->mask || a->F == 257);
  if (train_path || mel_path) {
    AVSI_REQUIRE(total < (1LL << 31) - (1LL << 24), "B*T must fit 32 bits");
    const int tsm = (int)sizeof(FrontendTrainSmem);
    static bool tattr = false;
    if (!tattr) {
      AVSI_CUDA(cudaFuncSetAttribute(frontend_train_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
      AVSI_CUDA(cudaFuncSetAttribute(frontend_train_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
      AVSI_CUDA(cudaFuncSetAttribute(frontend_train_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
      AVSI_CUDA(cudaFuncSetAttribute(frontend_train_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
      AVSI_CUDA(cudaFuncSetAttribute(frontend_train_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsm));
      tattr = true;
    }
    long long tg = (long long)num_sms() * 2;
    if (tg > groups) tg = groups;
    cudaStream_t fst = (cudaStream_t)stream;
    if (mel_path) frontend_train_kernel<false, false, true><<<(unsigned)tg, FE_THREADS, tsm, fst>>>(*a);
    else if (a->video && a->hole_count) frontend_train_kernel<true, true, false><<<(unsigned)tg, FE_THREADS, tsm, fst>>>(*a);
    else if (a->video) frontend_train_kernel<true, false, false><<<(unsigned)tg, FE_THREADS, tsm, fst>>>(*a);
    else if (a->hole_count) frontend_train_kernel<false, true, false><<<(unsigned)tg, FE_THREADS, tsm, fst>>>(*a);
    else frontend_train_kernel<false, false, false><<<(unsigned)tg, FE_THREADS, tsm, fst>>>(*a);
    AVSI_LAUNCH_CHECK();
    return AVSI_OK;
  }
  long long grid = (long long)num_sms() * 4;
  if (grid > groups) grid = groups;
  frontend_kernel<<<(unsigned)grid, FE_THREADS, smem, (cudaStream_t)stream>>>(*a);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
