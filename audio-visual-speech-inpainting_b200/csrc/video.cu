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
  float w = (float)(num - (unsigned)i0 * (unsigned)T) * (1.0f / (float)T);   // same expression as the incremental walk
  if (i0 >= L - 1) {
    i0 = L - 1;
    w = 0.f;
  }
  return {i0, min(i0 + 1, L - 1), w};
}

// One thread per (utterance, chunk of VF_CHUNK output frames, VEC feature columns): it walks its frames in order, so
// the interpolation position advances incrementally (no division by T per frame), the landmark rows of frame t-1 are
// the registers left by the previous iteration, and mean / std are loaded once.  A warp's lanes are consecutive column
// vectors of one row: 16-byte loads and stores, coalesced.  HBM-bound on the [B,T,D] output.

template <int VEC>
__global__ void __launch_bounds__(256)
video_features_kernel(const float* __restrict__ lm, const float* __restrict__ vmean, const float* __restrict__ vstd,
                      int B, int L, int D, int T, float* __restrict__ out, int VF_CHUNK) {
  const int dv = D / VEC;                            // column vectors per frame
  const int nchunk = (T + VF_CHUNK - 1) / VF_CHUNK;
  const long long total = (long long)B * nchunk * dv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // chunk-major order: the lanes of a warp share the chunk, hence t, the interpolation position and every
    // branch below; they differ in (utterance, column vector) only
    const int v = (int)(idx % dv);
    const long long bc = idx / dv;
    const int b = (int)(bc % B), t0 = (int)(bc / B) * VF_CHUNK, t1 = min(T, t0 + VF_CHUNK);
    const int d = v * VEC;
    const float* src = lm + (long long)b * L * D + d;
    float mean[VEC], sd[VEC], inv[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      mean[k] = __ldg(vmean + b * D + d + k);
      sd[k] = __ldg(vstd + b * D + d + k);
      inv[k] = 1.0f / sd[k];
    }
    const float inv_T = 1.0f / (float)T;
    auto load = [&](int i, float (&x)[VEC]) {
      if (VEC == 4) *reinterpret_cast<float4*>(x) = __ldg(reinterpret_cast<const float4*>(src + (long long)i * D));
      else x[0] = __ldg(src + (long long)i * D);
    };
    // state of frame t - 1 (frame t0 - 1 for the first iteration; unused when t0 == 0: the motion of frame 0 is 0)
    LerpPos pp = lerp_pos(t0 > 0 ? t0 - 1 : 0, L, T);
    float a0[VEC], b0[VEC];
    load(pp.i0, a0);
    load(pp.i1, b0);
    unsigned num = (unsigned)t0 * (unsigned)L;       // t * L, tracked as quotient / remainder by T
    int q = (int)(num / (unsigned)T);
    unsigned rem = num - (unsigned)q * (unsigned)T;
    float* dst = out + ((long long)b * T + t0) * D + d;
    for (int t = t0; t < t1; ++t) {
      LerpPos p1;
      p1.i0 = q;
      p1.w = (float)rem * inv_T;
      if (q >= L - 1) {
        p1.i0 = L - 1;
        p1.w = 0.f;
      }
      p1.i1 = min(p1.i0 + 1, L - 1);
      float a1[VEC], b1[VEC];
      if (p1.i0 == pp.i0) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          a1[k] = a0[k];
          b1[k] = b0[k];
        }
      } else if (p1.i0 == pp.i1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) a1[k] = b0[k];
        load(p1.i1, b1);
      } else {
        load(p1.i0, a1);
        load(p1.i1, b1);
      }
      float o[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // frame 0 has no motion: exactly -mean / std; elsewhere the reciprocal (1 ulp) keeps the loop free of divisions
        o[k] = (t > 0) ? ((a1[k] - a0[k]) + (p1.w * (b1[k] - a1[k]) - pp.w * (b0[k] - a0[k])) - mean[k]) * inv[k]
                       : (0.f - mean[k]) / sd[k];
      }
      if (VEC == 4) *reinterpret_cast<float4*>(dst) = *reinterpret_cast<float4*>(o);
      else dst[0] = o[0];
      dst += D;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        a0[k] = a1[k];
        b0[k] = b1[k];
      }
      pp = p1;
      rem += (unsigned)L;
      while (rem >= (unsigned)T) {
        rem -= (unsigned)T;
        ++q;
      }
    }
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
      if (c < F) {
        v = feat[r * F + c];
        if (mean) v = (v - __ldg(mean + c)) / __ldg(stdev + c);
      } else if (video && c < F + V) {
        v = video[r * V + (c - F)];
      }
      dst[c] = __half_as_ushort(__float2half_rn(v));
    }
  }
}

// per-utterance vector (speaker embedding) replicated over the frames: x0[t*B + b, col0 : col0 + E] = emb[b, :]
// (tf.tile(tf.expand_dims(embeddings, 1), [1, T, 1]) + tf.concat of models.py:1204-1206 / :846-849)
__global__ void __launch_bounds__(256)
tile_embedding_kernel(const float* __restrict__ emb, int B, int T, int E, uint16_t* __restrict__ x0, int ldx, int col0) {
  const long long n = (long long)T * B * E;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % E);
    const long long r = idx / E;                    // time-major row t*B + b
    const int b = (int)(r % B);
    x0[r * ldx + col0 + e] = __half_as_ushort(__float2half_rn(__ldg(emb + (long long)b * E + e)));
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
  static int VF_CHUNK = 0;
  if (!VF_CHUNK) {
    const char* e = getenv("AVSI_VF_CHUNK");
    VF_CHUNK = e ? atoi(e) : 32;
    if (VF_CHUNK < 1) VF_CHUNK = 32;
  }
  const long long work = (long long)B * ((T + VF_CHUNK - 1) / VF_CHUNK) * dvh;
  int blocks = (int)min((work + 255) / 256, (long long)num_sms() * 16);
  if (vec)
    video_features_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(landmarks, vmean, vstd, B, L, D, T, out, VF_CHUNK);
  else
    video_features_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(landmarks, vmean, vstd, B, L, D, T, out, VF_CHUNK);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_features_to_x0(const float* feat, const float* mean, const float* stdev, const float* video, int B,
                                   int T, int F, int V, uint16_t* x0, int ldx, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(feat && x0, "null pointer");
  AVSI_REQUIRE((mean == nullptr) == (stdev == nullptr), "mean and stdev: both or neither");
  AVSI_REQUIRE(B > 0 && T > 0 && F > 0 && V >= 0 && ldx >= F + (video ? V : 0), "sizes");
  long long rows = (long long)B * T;
  int blocks = (int)min(rows, (long long)num_sms() * 16);
  features_to_x0_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(feat, mean, stdev, video, B, T, F, V, x0, ldx);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_tile_embedding(const float* emb, int B, int T, int E, uint16_t* x0, int ldx, int col0, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(emb && x0, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && E > 0 && col0 >= 0 && ldx >= col0 + E, "sizes");
  const long long n = (long long)T * B * E;
  int blocks = (int)min((n + 255) / 256, (long long)num_sms() * 16);
  tile_embedding_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(emb, B, T, E, x0, ldx, col0);
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
