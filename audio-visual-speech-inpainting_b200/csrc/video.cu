// Landmark stream -> video features, and interval -> dense mask expansion.
//
// avsi_video_features replaces inc_fps / sync_audio_visual_features (av_sync.py:7-40),
// get_motion_vector(delta=1) (face_landmarks.py:30-39) and the z-normalisation of
// tfrecord_utils.py:104-107.  The interpolation abscissa is computed in float64 exactly as
// np.linspace(0, L*(1-1/T), T) does (start + i*step), so floor() picks the reference's frames.
// avsi_expand_mask replaces the tail of get_intrusions_mask (dataset_generator.py:43-46).
#include "common.cuh"

namespace avsi {

__device__ __forceinline__ double lerp_frame(const float* lm, int L, int D, int d, double y) {
  y = fmin(y, (double)(L - 1));
  int i0 = (int)floor(y);
  if (i0 > L - 1) i0 = L - 1;
  int i1 = min(i0 + 1, L - 1);
  double w = y - (double)i0;
  return (double)lm[(long long)i0 * D + d] * (1.0 - w) + (double)lm[(long long)i1 * D + d] * w;
}

__global__ void video_features_kernel(const float* __restrict__ lm, const float* __restrict__ vmean,
                                      const float* __restrict__ vstd, int B, int L, int D, int T,
                                      float* __restrict__ out) {
  const long long n = (long long)B * T * D;
  const double stop = (double)L * (1.0 - 1.0 / (double)T);
  const double step = (T > 1) ? stop / (double)(T - 1) : 0.0;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int t = (int)((idx / D) % T);
    const int b = (int)(idx / ((long long)D * T));
    const float* src = lm + (long long)b * L * D;
    double mv = 0.0;
    if (t > 0) {
      // np.linspace: y_i = start + i*step, last point forced to `stop`
      double y1 = (t == T - 1) ? stop : (double)t * step;
      double y0 = (double)(t - 1) * step;
      mv = lerp_frame(src, L, D, d, y1) - lerp_frame(src, L, D, d, y0);
    }
    out[idx] = (float)((mv - (double)vmean[b * D + d]) / (double)vstd[b * D + d]);
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
  long long n = (long long)B * T * D;
  int blocks = (int)min((n + 255) / 256, (long long)num_sms() * 8);
  video_features_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(landmarks, vmean, vstd, B, L, D, T, out);
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
