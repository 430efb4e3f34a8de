// Optimiser update (TF1 AdamOptimizer semantics) and fp32 -> fp16 weight copies.
//
// avsi_adam_tf replaces the per-variable ApplyAdam kernels behind
// tf.train.AdamOptimizer(...).minimize (models.py:168,178): ONE launch over the flat
// parameter buffer.  epsilon is added to the uncorrected sqrt(v) ("epsilon hat").
#include "common.cuh"

namespace avsi {

__global__ void __launch_bounds__(256)
adam_tf_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
               float* __restrict__ v, long long n, double lr, int step, double b1d, double b2d, float eps,
               float unscale, const float* __restrict__ unscale_dev, float l2, const int32_t* __restrict__ guard) {
  if (guard && guard[0] != 0) return;            // non-finite gradient (avsi_grad_guard): the step is skipped as a whole
  // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) in double; t counts the updates actually applied (skipped steps excluded)
  // step == 0: the call count lives on the device (guard[3], bumped by grad_guard_update) -- what a captured CUDA graph replays
  const int t = (step > 0 ? step : guard[3] + 1) - (guard ? guard[1] : 0);
  const float lr_t = (float)(lr * sqrt(1.0 - pow(b2d, (double)t)) / (1.0 - pow(b1d, (double)t)));
  const float b1 = (float)b1d, omb1 = (float)(1.0 - b1d), b2 = (float)b2d, omb2 = (float)(1.0 - b2d);
  const float us = unscale * (unscale_dev ? *unscale_dev : 1.f) * (guard ? __int_as_float(guard[5]) : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float th = theta[i];
    float gi = fmaf(g[i], us, l2 * th);
    float mi = b1 * m[i] + omb1 * gi;
    float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    theta[i] = th - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// tf.train.GradientDescentOptimizer / MomentumOptimizer(momentum) (models.py:169-173):
//   sgd:       theta -= lr * g
//   momentum:  accum = momentum * accum + g ; theta -= lr * accum          (TF ApplyMomentum, use_nesterov = False)
__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ accum, long long n,
                    float lr, float momentum, float unscale, const float* __restrict__ unscale_dev, float l2,
                    const int32_t* __restrict__ guard) {
  if (guard && guard[0] != 0) return;
  const float us = unscale * (unscale_dev ? *unscale_dev : 1.f) * (guard ? __int_as_float(guard[5]) : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float th = theta[i];
    float gi = fmaf(g[i], us, l2 * th);
    if (accum) {
      gi = fmaf(momentum, accum[i], gi);
      accum[i] = gi;
    }
    theta[i] = th - lr * gi;
  }
}

// Overflow guard of the fp16 gradient path.  The activations' gradients (dlogits, dY, dG) are fp16 and flow loss-scaled;
// a run-away dh saturates at 65504 -> inf -> NaN in the weight gradients.  guard = 8 words:
//   [0] i32 non-finite flag of THIS step   [1] i32 steps skipped so far   [2] i32 finite steps since the last change
//   [3] i32 optimiser calls completed (the device-resident step count of graph replays)
//   [4] f32 dynamic scale s (multiplies the loss gradient fed to the backward pass, a power of two <= 1)
//   [5] f32 1 / s (folded into the optimiser's unscale)
// grad_guard_check sets [0]; the optimiser kernels return without touching theta / m / v when it is set;
// grad_guard_update (after the optimiser) halves s on a skipped step and doubles it back (up to 1) after
// `growth_interval` finite steps -- all on the device, no host synchronisation in the step.
__global__ void __launch_bounds__(256) grad_guard_check_kernel(const float* __restrict__ g, long long n, int32_t* __restrict__ guard) {
  bool bad = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = g[i];
    bad |= !(fabsf(x) <= 3.0e38f);               // inf or NaN
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(guard, 1);
}
__global__ void grad_guard_update_kernel(int32_t* __restrict__ guard, int growth_interval) {
  float s = __int_as_float(guard[4]);
  if (guard[0] != 0) {
    guard[1] += 1;
    guard[2] = 0;
    s = fmaxf(s * 0.5f, 5.9604645e-08f);         // 2^-24
  } else if (++guard[2] >= growth_interval && s < 1.f) {
    guard[2] = 0;
    s *= 2.f;
  }
  guard[3] += 1;
  guard[4] = __float_as_int(s);
  guard[5] = __float_as_int(1.f / s);
}

__global__ void __launch_bounds__(256)
cast_weights_kernel(const float* __restrict__ w, int R, int C, uint16_t* __restrict__ w16,
                    uint16_t* __restrict__ w16t, int halve_sigmoid_rows) {
  // 32x32 tile transpose through shared memory; also writes the straight copy
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty + 8 * i, c = c0 + tx;
    float x = (r < R && c < C) ? w[(long long)r * C + c] : 0.f;
    tile[ty + 8 * i][tx] = x;
    // forward operand copies of the gate matrices carry the 1/2 of sigma(z) = 1/2 tanh(z/2) + 1/2: rows (gate
    // columns) i, f, o are halved (exact in fp16), row g (index 1 of [i,g,f,o]) is not; the transposed copy that
    // the backward GEMMs read stays unscaled
    const float sc = (halve_sigmoid_rows && (r & 3) != 1) ? 0.5f : 1.f;
    if (w16 && r < R && c < C) w16[(long long)r * C + c] = __half_as_ushort(__float2half_rn(x * sc));
  }
  if (!w16t) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < R && c < C) w16t[(long long)c * R + r] = __half_as_ushort(__float2half_rn(tile[tx][ty + 8 * i]));
  }
}

// Host-side batches may arrive in their storage dtypes (int16 samples as in target.wav, int32 as dataset_reader.py:78
// yields them, uint8 {0,1} masks): widened to the fp32 tensors of the feed contract on the device, so the PCIe copy
// moves 2 / 1 bytes per element instead of 4.
template <typename S>
__global__ void __launch_bounds__(256) cast_to_f32_kernel(const S* __restrict__ src, long long n, float* __restrict__ dst) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      float4 v = make_float4((float)src[i], (float)src[i + 1], (float)src[i + 2], (float)src[i + 3]);
      if (((uintptr_t)(dst + i) & 15) == 0) *reinterpret_cast<float4*>(dst + i) = v;
      else {
        dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
      }
    } else {
      for (long long k = i; k < n; ++k) dst[k] = (float)src[k];
    }
  }
}

__global__ void gate_bias_prescale_kernel(const float* __restrict__ b, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = b[i] * (((i & 3) != 1) ? 0.5f : 1.f);
}

__global__ void mtl_scales_kernel(const double* __restrict__ hole_count, int B, float ctc_w, float* __restrict__ out,
                                  const int32_t* __restrict__ guard) {
  // out[0] = S  (scale of the L1 dlogits), out[1] = S * (w/B) * holes (scale of the CTC dlogits),
  // out[2] = 1 / (S * holes)  (optimiser unscale), out[3] = holes
  float holes = fmaxf((float)*hole_count, 1.f);
  float rel = ctc_w / (float)B * holes;         // CTC gradient relative to the +-1 L1 gradient
  float S = 1.f;
  while (rel * S > 64.f) S *= 0.5f;             // keep fp16 dlogits far from 65504
  const float dyn = guard ? __int_as_float(guard[4]) : 1.f;     // dynamic scale of the overflow guard (its inverse is
  out[0] = S * dyn;                                             // applied by the optimiser kernel from guard[5])
  out[1] = S * rel * dyn;
  out[2] = 1.f / (S * holes);
  out[3] = holes;
}

}  // namespace avsi

extern "C" int avsi_adam_tf(float* theta, const float* g, float* m, float* v, int64_t n, double lr, double b1,
                            double b2, double eps, int step, float grad_unscale, const float* grad_unscale_dev,
                            float l2, const int32_t* guard, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(theta && g && m && v, "null pointer");
  AVSI_REQUIRE(n > 0 && (step >= 1 || (step == 0 && guard)), "n > 0, step >= 1 (or 0 with a guard: device-resident count)");
  int blocks = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 8);
  adam_tf_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, g, m, v, (long long)n, lr, step, b1, b2, (float)eps,
                                                          grad_unscale, grad_unscale_dev, l2, guard);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_sgd_momentum(float* theta, const float* g, float* accum, int64_t n, double lr, double momentum,
                                 float grad_unscale, const float* grad_unscale_dev, float l2, const int32_t* guard,
                                 void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(theta && g, "null pointer");
  AVSI_REQUIRE(n > 0, "n > 0");
  int blocks = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 8);
  sgd_momentum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, g, accum, (long long)n, (float)lr, (float)momentum,
                                                               grad_unscale, grad_unscale_dev, l2, guard);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_cast_to_f32(const void* src, int src_type, int64_t n, float* dst, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(src && dst && n > 0, "args");
  AVSI_REQUIRE(src_type >= 0 && src_type <= 2, "src_type: 0 = int16, 1 = uint8, 2 = int32");
  int blocks = (int)min((long long)(n / 4 + 255) / 256 + 1, (long long)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_type == 0) cast_to_f32_kernel<int16_t><<<blocks, 256, 0, st>>>((const int16_t*)src, (long long)n, dst);
  else if (src_type == 1) cast_to_f32_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)src, (long long)n, dst);
  else cast_to_f32_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)src, (long long)n, dst);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_cast_weights(const float* w, int R, int C, uint16_t* w16, uint16_t* w16t, int halve_sigmoid_rows,
                                 void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(w && (w16 || w16t), "null pointer");
  AVSI_REQUIRE(R > 0 && C > 0, "sizes");
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  cast_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, R, C, w16, w16t, halve_sigmoid_rows);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_gate_bias_prescale(const float* bias, int n, float* out, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(bias && out && n > 0, "args");
  gate_bias_prescale_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bias, n, out);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_grad_guard_init(int32_t* guard, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(guard, "null pointer");
  const float one = 1.f;
  int32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  memcpy(&h[4], &one, 4);
  memcpy(&h[5], &one, 4);
  AVSI_CUDA(cudaMemcpyAsync(guard, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  AVSI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));       // h is a stack buffer
  return AVSI_OK;
}

extern "C" int avsi_grad_guard_check(const float* g, int64_t n, int32_t* guard, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(g && guard && n > 0, "args");
  AVSI_CUDA(cudaMemsetAsync(guard, 0, 4, (cudaStream_t)stream));
  int blocks = (int)min((long long)(n + 255) / 256, (long long)num_sms() * 8);
  grad_guard_check_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, (long long)n, guard);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_grad_guard_update(int32_t* guard, int growth_interval, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(guard && growth_interval > 0, "args");
  grad_guard_update_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(guard, growth_interval);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_mtl_scales(const double* hole_count, int B, float ctc_weight, float* out, const int32_t* guard,
                               void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(hole_count && out && B > 0, "args");
  mtl_scales_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hole_count, B, ctc_weight, out, guard);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
