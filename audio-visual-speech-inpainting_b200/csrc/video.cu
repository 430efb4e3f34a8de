// Landmark stream -> video features, and interval -> dense mask expansion.
//
// avsi_video_features replaces inc_fps / sync_audio_visual_features (av_sync.py:7-40),
// get_motion_vector(delta=1) (face_landmarks.py:30-39) and the z-normalisation of
// tfrecord_utils.py:104-107.
// avsi_expand_mask replaces the tail of get_intrusions_mask (dataset_generator.py:43-46).
#include "common.cuh"

namespace avsi {

// Interpolation abscissa of output frame t: np.linspace(0, L*(1-1/T), T)[t] = t*L/T, an exact rational, so the
// source frame floor(t*L/T) and the weight (t*L mod T)/T are computed in integers (no fp64, no rounding at the
// frame boundaries); clamped to the last frame like inc_fps.  lerp = a + w*(b - a): for integer pixel
// coordinates a and b - a are exact, and the motion vector is formed as (a1 - a0) + (w1*d1 - w0*d0).
struct LerpPos {
  int i0, i1;
  float w;
};
__device__ __forceinline__ LerpPos lerp_pos(int t, int L, int T) {
  const unsigned num = (unsigned)t * (unsigned)L;            // t < T, T * L < 2^31 checked by the launcher
  int i0 = (int)(num / (unsigned)T);
  float w = (float)(num - (unsigned)i0 * (unsigned)T) / (float)T;
  if (i0 >= L - 1) {
    i0 = L - 1;
    w = 0.f;
  }
  return {i0, min(i0 + 1, L - 1), w};
}

template <int VEC>
__global__ void __launch_bounds__(256)
video_features_kernel(const float* __restrict__ lm, const float* __restrict__ vmean, const float* __restrict__ vstd,
                      int B, int L, int D, int T, float* __restrict__ out) {
  const int dv = D / VEC;                            // vectors per frame
  // one frame (b, t) per group of `dv` consecutive threads of a block row: 32-bit index arithmetic only
  const int frames_per_block = blockDim.x / dv;
  const int fl = threadIdx.x / dv;
  if (fl >= frames_per_block) return;
  const int d = (threadIdx.x - fl * dv) * VEC;
  const long long total = (long long)B * T;
  for (long long bt0 = (long long)blockIdx.x * frames_per_block; bt0 < total; bt0 += (long long)gridDim.x * frames_per_block) {
    const long long btl = bt0 + fl;
    if (btl >= total) break;
    const int bt = (int)btl;
    const int b = bt / T, t = bt - b * T;
    const float* src = lm + (long long)b * L * D + d;
    float mv[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) mv[k] = 0.f;
    if (t > 0) {
      const LerpPos p1 = lerp_pos(t, L, T), p0 = lerp_pos(t - 1, L, T);
      float a1[VEC], b1[VEC], a0[VEC], b0[VEC];
      if (VEC == 4) {
        *reinterpret_cast<float4*>(a1) = __ldg(reinterpret_cast<const float4*>(src + (long long)p1.i0 * D));
        *reinterpret_cast<float4*>(b1) = __ldg(reinterpret_cast<const float4*>(src + (long long)p1.i1 * D));
        *reinterpret_cast<float4*>(a0) = __ldg(reinterpret_cast<const float4*>(src + (long long)p0.i0 * D));
        *reinterpret_cast<float4*>(b0) = __ldg(reinterpret_cast<const float4*>(src + (long long)p0.i1 * D));
      } else {
        a1[0] = __ldg(src + (long long)p1.i0 * D);
        b1[0] = __ldg(src + (long long)p1.i1 * D);
        a0[0] = __ldg(src + (long long)p0.i0 * D);
        b0[0] = __ldg(src + (long long)p0.i1 * D);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        mv[k] = (a1[k] - a0[k]) + (p1.w * (b1[k] - a1[k]) - p0.w * (b0[k] - a0[k]));
    }
    float o[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) o[k] = (mv[k] - __ldg(vmean + b * D + d + k)) / __ldg(vstd + b * D + d + k);
    if (VEC == 4) *reinterpret_cast<float4*>(out + (long long)bt * D + d) = *reinterpret_cast<float4*>(o);
    else out[(long long)bt * D + d] = o[0];
  }
}

// (feat - mean) / std  ++ video  ->  time-major fp16 rows (the ASR model's input assembly, models_asr.py:37-49)
__global__ void __launch_bounds__(256)
features_to_x0_kernel(const float* __restrict__ feat, const float* __restrict__ mean, const float* __restrict__ stdev,
                      const float* __restrict__ video, int B, int T, int F, int V, uint16_t* __restrict__ x0, int ldx) {
  const long long rows = (long long)B * T;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const int b = (int)(r / T), t = (int)(r - (long long)b * T);
    uint16_t* dst = x0 + ((long long)t * B + b) * ldx;
    for (int c = threadIdx.x; c < ldx; c += blockDim.x) {
      float v = 0.f;
      if (c < F) v = (feat[r * F + c] - __ldg(mean + c)) / __ldg(stdev + c);
      else if (video && c < F + V) v = video[r * V + (c - F)];
      dst[c] = __half_as_ushort(__float2half_rn(v));
    }
  }
}

__global__ void expand_mask_kernel(const int32_t* __restrict__ iv, int B, int K, int T, int F,
                                   float* __restrict__ mask) {
  const long long n = (long long)B * T * F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)((idx / F) % T);
    const int b = (int)(idx / ((long long)F * T));
    float m = 1.f;
    for (int k = 0; k < K; ++k) {
      int on = iv[(b * K + k) * 2], len = iv[(b * K + k) * 2 + 1];
      if (len > 0 && t >= on && t < on + len) m = 0.f;
    }
    mask[idx] = m;
  }
}

}  // namespace avsi

extern "C" int avsi_video_features(const float* landmarks, const float* vmean, const float* vstd, int B,
                                   int L, int D, int T, float* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(landmarks && vmean && vstd && out, "null pointer");
  AVSI_REQUIRE(B > 0 && L > 0 && D > 0 && T > 0, "sizes");
  AVSI_REQUIRE((long long)T * L < (1LL << 31) && (long long)B * T < (1LL << 31), "T * L and B * T must fit 31 bits");
  const bool vec = (D % 4 == 0) && ((uintptr_t)landmarks % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int dvh = vec ? D / 4 : D;
  AVSI_REQUIRE(dvh <= 256, "D too large");
  long long n = ((long long)B * T + (256 / dvh) - 1) / (256 / dvh);
  int blocks = (int)min(n, (long long)num_sms() * 16);
  if (vec)
    video_features_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(landmarks, vmean, vstd, B, L, D, T, out);
  else
    video_features_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(landmarks, vmean, vstd, B, L, D, T, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_features_to_x0(const float* feat, const float* mean, const float* stdev, const float* video, int B,
                                   int T, int F, int V, uint16_t* x0, int ldx, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(feat && mean && stdev && x0, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && F > 0 && V >= 0 && ldx >= F + (video ? V : 0), "sizes");
  long long rows = (long long)B * T;
  int blocks = (int)min(rows, (long long)num_sms() * 16);
  features_to_x0_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(feat, mean, stdev, video, B, T, F, V, x0, ldx);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_expand_mask(const int32_t* intervals, int B, int K, int T, int F, float* mask,
                                void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(intervals && mask, "null pointer");
  AVSI_REQUIRE(B > 0 && K > 0 && T > 0 && F > 0, "sizes");
  long long n = (long long)B * T * F;
  int blocks = (int)min((n + 255) / 256, (long long)num_sms() * 8);
  expand_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(intervals, B, K, T, F, mask);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
