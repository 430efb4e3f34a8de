// Masked-L1 loss + gradient (one pass) and fp16 column sums (bias gradients).
//
// avsi_masked_l1 replaces prediction / loss of models.py:127-159 (SI) and
// models.py:1920-1963 (MTL) together with their autodiff mirror (about ten element-wise
// and three reduction kernels in the TF graph).
#include "common.cuh"

namespace avsi {

constexpr int L1_THREADS = 256;
constexpr int L1_UNROLL = 4;

__global__ void __launch_bounds__(L1_THREADS)
masked_l1_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ target,
                 const float* __restrict__ mask, const int32_t* __restrict__ seq_len, int B, int T, int F,
                 int mode, float grad_scale_h, const float* __restrict__ grad_scale_dev, double* __restrict__ sums, float* __restrict__ prediction,
                 uint16_t* __restrict__ dlogits, int ldd) {
  // one warp per (b,t) row, grid-stride
  const int lane = threadIdx.x & 31;
  const int warps_per_block = L1_THREADS / 32;
  const float grad_scale = grad_scale_h * (grad_scale_dev ? *grad_scale_dev : 1.f);
  const long long rows = (long long)B * T;
  float s_hole = 0.f, n_hole = 0.f, s_valid = 0.f, n_valid = 0.f, s_all = 0.f;
  for (long long r = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(r / T);
    const int t = (int)(r - (long long)b * T);
    const float sm = (t < seq_len[b]) ? 1.f : 0.f;
    const float* lg = logits + ((long long)t * B + b) * ldl;
    const float* tg = target + r * F;
    const float* mk = mask + r * F;
    uint16_t* dl = dlogits ? dlogits + ((long long)t * B + b) * ldd : nullptr;
    for (int k0 = lane; k0 < F; k0 += 32 * L1_UNROLL) {
     // the loads of L1_UNROLL column groups are requested before the first is used (12 lines in flight per warp)
     float xv[L1_UNROLL], yv[L1_UNROLL], mv[L1_UNROLL];
#pragma unroll
     for (int u = 0; u < L1_UNROLL; ++u) {
       const int k = k0 + 32 * u;
       const bool ok = k < F;
       xv[u] = ok ? __ldg(lg + k) : 0.f;
       yv[u] = ok ? __ldg(tg + k) : 0.f;
       mv[u] = ok ? __ldg(mk + k) : 0.f;
     }
#pragma unroll
     for (int u = 0; u < L1_UNROLL; ++u) {
      const int k = k0 + 32 * u;
      if (k >= F) break;
      const float x = xv[u], y = yv[u], m = mv[u];
      float pred = (mode == 0) ? x : (y * m + x * (1.f - m));
      pred *= sm;
      const float d = y - pred;
      const float ad = fabsf(d);
      s_hole += ad * (1.f - m);
      n_hole += (1.f - m);
      s_valid += ad * m;
      n_valid += m;
      s_all += ad;
      if (prediction) prediction[r * F + k] = pred;
      if (dl) {
        // d|y - pred| / dx = -sign(y - pred) * dpred/dx
        float sg = (d > 0.f) ? -1.f : ((d < 0.f) ? 1.f : 0.f);
        float w = (mode == 0) ? sm : sm * (1.f - m);
        dl[k] = __half_as_ushort(__float2half_rn(grad_scale * sg * w));
      }
     }
    }
  }
  __shared__ double red[5][L1_THREADS / 32];
  float vals[5] = {s_hole, n_hole, s_valid, n_valid, s_all};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double v = warp_sum_d((double)vals[i]);
    if (lane == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int w = 0; w < warps_per_block; ++w) v += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, v);
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sums + 5, (double)rows * F);
}

__global__ void __launch_bounds__(256)
colsum_f16_kernel(const uint16_t* __restrict__ X, int ldx, int rows, int col0, int ncols,
                  float* __restrict__ out) {
  // block handles 32 columns x a slab of rows; threads (32 cols x 8 row lanes)
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < ncols) {
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8)
      acc += __half2float(__ushort_as_half(X[r * ldx + col0 + c]));
  }
  __shared__ float red[8][33];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < ncols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x & 31];
    atomicAdd(out + c, v);
  }
}

}  // namespace avsi

extern "C" int avsi_masked_l1(const float* logits, int ldl, const float* target, const float* mask,
                              const int32_t* seq_len, int B, int T, int F, int mode, float grad_scale,
                              const float* grad_scale_dev, double* sums, float* prediction, uint16_t* dlogits,
                              int ldd, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(logits && target && mask && seq_len && sums, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && F > 0 && ldl >= F, "sizes");
  AVSI_REQUIRE(mode == 0 || mode == 1, "mode");
  AVSI_REQUIRE(!dlogits || ldd >= F, "ldd");
  long long rows = (long long)B * T;
  int blocks = (int)min((rows + 7) / 8, (long long)num_sms() * 8);
  masked_l1_kernel<<<blocks, L1_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, target, mask, seq_len, B, T, F,
                                                                   mode, grad_scale, grad_scale_dev, sums, prediction, dlogits, ldd);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_colsum_f16(const uint16_t* X, int ldx, int rows, int col0, int ncols, float* out,
                               void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(X && out, "null pointer");
  AVSI_REQUIRE(rows > 0 && ncols > 0 && ldx >= col0 + ncols, "sizes");
  dim3 grid((ncols + 31) / 32, (unsigned)min((long long)(rows + 255) / 256, (long long)64));
  if (grid.y == 0) grid.y = 1;
  colsum_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, rows, col0, ncols, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
