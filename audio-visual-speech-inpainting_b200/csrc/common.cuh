// Shared helpers for the avsi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/avsi_b200.h"

namespace avsi {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;
extern std::atomic<int> g_env_gen;      // bumped by avsi_reload_env(): the cached AVSI_* tuning switches are re-read

// `static` cache of an integer derived from the environment, re-evaluated after avsi_reload_env()
#define AVSI_ENV_CACHE(var, expr)                        \
  static int var##_gen_ = -1;                            \
  static int var = 0;                                    \
  if (var##_gen_ != avsi::g_env_gen.load()) {            \
    var = (expr);                                        \
    var##_gen_ = avsi::g_env_gen.load();                 \
  }
inline int env_is(const char* name, const char* value) {
  const char* e = getenv(name);
  return (e && !strcmp(e, value)) ? 1 : 0;
}
inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

inline int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
  return code;
}

#define AVSI_REQUIRE(cond, msg)                                                     \
  do {                                                                              \
    if (!(cond)) return avsi::set_error(AVSI_ERR_INVALID, "%s: requirement failed: %s", __func__, msg); \
  } while (0)

#define AVSI_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return avsi::set_error(AVSI_ERR_CUDA, "%s: CUDA error: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

#define AVSI_LAUNCH_CHECK()                                                         \
  do {                                                                              \
    avsi::g_launch_count.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess)                                                          \
      return avsi::set_error(AVSI_ERR_CUDA, "%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

// Scratch of the ordered reductions (avsi_set_reduce_scratch): [0, REDUCE_COUNTER_BYTES) tickets of the last-block
// reductions (zero between launches), the rest is the partial-sum area of whichever reduction is running -- all users
// are ordered on ONE stream.  ptr == nullptr: nothing registered, the reductions use floating-point atomics.
constexpr long long REDUCE_COUNTER_BYTES = 4096;
constexpr int REDUCE_SLOT_L1 = 0, REDUCE_SLOT_COLSUM = 8;      // colsum: one ticket per column block (<= 64)
struct ReduceScratch {
  unsigned* counters;
  unsigned char* area;
  long long area_bytes;
};
ReduceScratch reduce_scratch();

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float sigmoidf_fast(float x) {
  // sigma(x) = 0.5 * tanh(0.5 x) + 0.5  (one MUFU.TANH)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float tanhf_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// fast activations for the recurrence: ex2.approx + rcp.approx (2 MUFU, ~2 ulp each), no IEEE fix-up code
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast2(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
// tanh(x) = 2 sigmoid(2x) - 1, saturates cleanly for |x| large
__device__ __forceinline__ float tanh_fast2(float x) { return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f); }
// accurate versions (expf based, ~1e-7 rel) used where the 2e-3 budget is tight
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_acc(float x) {
  float e = __expf(-2.0f * fabsf(x));
  float t = (1.0f - e) / (1.0f + e);
  return copysignf(t, x);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
  __half2 h = *reinterpret_cast<__half2*>(&u);
  return __half22float2(h);
}

// Interleaved ("IL") activation layout: a [R x C] matrix stored as [R/32][C/chunk][32 rows][chunk] with
// 16-byte chunks (8 halves / 4 floats).  A warp whose lanes own 32 consecutive rows reads or writes one
// 512-byte contiguous run per 16-byte column chunk (4 cache lines per instruction instead of 32), and a
// [rows x 8 cols] sub-block is exactly the SWIZZLE_NONE core-matrix layout tcgen05 descriptors address.
// R is padded to a multiple of 32 by the allocator; C must be a multiple of the chunk.
__host__ __device__ __forceinline__ long long il16(long long r, int c, int C) {
  return (((r >> 5) * (long long)(C >> 3) + (c >> 3)) << 8) + ((r & 31) << 3) + (c & 7);
}
__host__ __device__ __forceinline__ long long il32(long long r, int c, int C) {
  return (((r >> 5) * (long long)(C >> 2) + (c >> 2)) << 7) + ((r & 31) << 2) + (c & 3);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace avsi
