// Element-wise / reduction kernels of the speaker-embedding sub-network of the SSNN models (models.py:800-842):
//   inp = add_delta_features(audio_features, 1, 2)  [B,T,2F]  ->  3 dense layers (leaky ReLU 0.3 after the first two)
//   ->  * mask[:, :, 0]  ->  sum over frames / (sum of the mask + 1)  =  speaker embedding [B,200]
// The dense layers run on the tcgen05 GEMM (avsi_gemm_f16); these kernels are what sits between them, forward and
// backward.  Rows are batch-major (r = b * T + t) like the [B,T,.] tensors they come from.
#include "common.cuh"

namespace avsi {

// f32 [rows, cols] (row pitch ld_src) -> f16 [rows, ld_dst], columns >= cols zero: GEMM operand rows need a pitch % 8 == 0
__global__ void __launch_bounds__(256)
cast_pad_f16_kernel(const float* __restrict__ src, int ld_src, int cols, uint16_t* __restrict__ dst, int ld_dst, long long rows) {
  const long long n = rows * ld_dst;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    dst[i] = __half_as_ushort(__float2half_rn(c < cols ? src[r * ld_src + c] : 0.f));
  }
}

__global__ void __launch_bounds__(256)
leaky_relu_kernel(const float* __restrict__ z, long long n, float alpha, uint16_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = z[i];
    out[i] = __half_as_ushort(__float2half_rn(v > 0.f ? v : alpha * v));
  }
}

// dz = da * (z > 0 ? 1 : alpha)  (tf.nn.leaky_relu: max(alpha x, x); derivative alpha at x <= 0)
__global__ void __launch_bounds__(256)
leaky_relu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ da, long long n, float alpha, uint16_t* __restrict__ dz) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dz[i] = __half_as_ushort(__float2half_rn(da[i] * (z[i] > 0.f ? 1.f : alpha)));
}

// out[b,n] = sum_t x[b,t,n] m[b,t] / (sum_t m[b,t] + 1) ; inv_den[b] = 1 / (sum_t m[b,t] + 1).  One block per utterance.
__global__ void __launch_bounds__(256)
masked_time_mean_kernel(const float* __restrict__ x, const float* __restrict__ mask, int mask_ld, int T, int N,
                        float* __restrict__ out, float* __restrict__ inv_den) {
  const int b = blockIdx.x;
  __shared__ float den_sh;
  float cnt = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) cnt += mask[((long long)b * T + t) * mask_ld];
  cnt = warp_sum(cnt);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    den_sh = 1.f / (s + 1.f);
    inv_den[b] = den_sh;
  }
  __syncthreads();
  const float id = den_sh;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(x[((long long)b * T + t) * N + n], mask[((long long)b * T + t) * mask_ld], acc);
    out[(long long)b * N + n] = acc * id;
  }
}

// dx[b,t,n] = d_out[b,n] * m[b,t] * inv_den[b]   (f16: the A operand of the dense layers' backward GEMMs)
__global__ void __launch_bounds__(256)
masked_time_mean_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ mask, int mask_ld,
                            const float* __restrict__ inv_den, int B, int T, int N, uint16_t* __restrict__ dx) {
  const long long n_el = (long long)B * T * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;                                   // b * T + t
    const int n = (int)(i - r * N), b = (int)(r / T);
    dx[i] = __half_as_ushort(__float2half_rn(d_out[(long long)b * N + n] * mask[r * mask_ld] * inv_den[b]));
  }
}

// out[b,n] = sum_t x[(t * B + b) * ld + n]   (time-major rows: the gradient of a vector that was replicated over the frames)
__global__ void __launch_bounds__(256)
time_sum_kernel(const float* __restrict__ x, int ld, int T, int B, int N, float* __restrict__ out) {
  const long long n_el = (long long)B * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / N), n = (int)(i - (long long)b * N);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += x[((long long)t * B + b) * ld + n];
    out[i] = acc;
  }
}

static inline int grid_for(long long n) { return (int)min((n + 255) / 256, (long long)num_sms() * 16); }

}  // namespace avsi

extern "C" int avsi_cast_pad_f16(const float* src, int ld_src, int cols, uint16_t* dst, int ld_dst, int64_t rows, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "args");
  cast_pad_f16_kernel<<<grid_for(rows * ld_dst), 256, 0, (cudaStream_t)stream>>>(src, ld_src, cols, dst, ld_dst, (long long)rows);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_leaky_relu(const float* z, int64_t n, float alpha, uint16_t* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(z && out && n > 0, "args");
  leaky_relu_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(z, (long long)n, alpha, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_leaky_relu_bwd(const float* z, const float* da, int64_t n, float alpha, uint16_t* dz, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(z && da && dz && n > 0, "args");
  leaky_relu_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(z, da, (long long)n, alpha, dz);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_masked_time_mean(const float* x, const float* mask, int mask_ld, int B, int T, int N, float* out,
                                     float* inv_den, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(x && mask && out && inv_den && B > 0 && T > 0 && N > 0 && mask_ld > 0, "args");
  masked_time_mean_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, mask, mask_ld, T, N, out, inv_den);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_masked_time_mean_bwd(const float* d_out, const float* mask, int mask_ld, const float* inv_den, int B,
                                         int T, int N, uint16_t* dx, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(d_out && mask && inv_den && dx && B > 0 && T > 0 && N > 0 && mask_ld > 0, "args");
  masked_time_mean_bwd_kernel<<<grid_for((long long)B * T * N), 256, 0, (cudaStream_t)stream>>>(d_out, mask, mask_ld, inv_den, B, T, N, dx);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_time_sum(const float* x, int ld, int T, int B, int N, float* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(x && out && T > 0 && B > 0 && N > 0 && ld >= N, "args");
  time_sum_kernel<<<grid_for((long long)B * N), 256, 0, (cudaStream_t)stream>>>(x, ld, T, B, N, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
