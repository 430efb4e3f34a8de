// Feature statistics: the accumulation of compute_mean_std_features (audio_feat_preprocessing.py:76-115):
// per feature column  sum x,  sum x^2  (float64, as the reference's numpy accumulators) and the frame count,
// with the optional mask product  feat * mask  and  frame_count += sum(mask[:, 0])  (:87-92, :104-107).
// HBM-bound: one pass over the features, one CTA per chunk of rows, fp64 atomics on [F] accumulators.
#include "common.cuh"

namespace avsi {

__global__ void __launch_bounds__(256)
feature_stats_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ mask, int ldm, long long R, int F,
                     double* __restrict__ sum, double* __restrict__ sumsq, double* __restrict__ count) {
  const long long rows_per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * rows_per, r1 = min(R, r0 + rows_per);
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    double s = 0.0, s2 = 0.0, n = 0.0;
    for (long long r = r0; r < r1; ++r) {
      float v = __ldg(x + r * ldx + c);
      if (mask) {
        const float m = __ldg(mask + r * ldm + c);
        v *= m;
        if (c == 0) n += (double)m;
      }
      s += (double)v;
      s2 += (double)v * (double)v;
    }
    if (r1 > r0) {
      atomicAdd(sum + c, s);
      atomicAdd(sumsq + c, s2);
      if (c == 0) atomicAdd(count, mask ? n : (double)(r1 - r0));
    }
  }
}

}  // namespace avsi

extern "C" int avsi_feature_stats(const float* x, int ldx, const float* mask, int ldm, int64_t rows, int F, double* sum,
                                  double* sumsq, double* count, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(x && sum && sumsq && count, "null pointer");
  AVSI_REQUIRE(rows > 0 && F > 0 && ldx >= F && (!mask || ldm >= F), "sizes");
  long long blocks = (rows + 63) / 64;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  feature_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, mask, ldm, rows, F, sum, sumsq, count);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
