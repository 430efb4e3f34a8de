// Library-level state of the C ABI (include/avsi_b200.h): last-error text, version, launch counter.
#include "common.cuh"

namespace avsi {
thread_local char g_last_error[512] = {0};
std::atomic<long long> g_launch_count{0};
std::atomic<int> g_env_gen{0};

// per-device scratch of the ordered (bit-reproducible) reductions, registered by avsi_set_reduce_scratch
static ReduceScratch g_reduce_scratch[64];
ReduceScratch reduce_scratch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return ReduceScratch{nullptr, nullptr, 0};
  return g_reduce_scratch[dev];
}
}  // namespace avsi

extern "C" int avsi_set_reduce_scratch(void* scratch, int64_t bytes, void* stream) {
  using namespace avsi;
  int dev = 0;
  AVSI_CUDA(cudaGetDevice(&dev));
  AVSI_REQUIRE(dev >= 0 && dev < 64, "device index");
  if (!scratch) {
    g_reduce_scratch[dev] = ReduceScratch{nullptr, nullptr, 0};
    return AVSI_OK;
  }
  AVSI_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
  AVSI_REQUIRE(bytes >= avsi_reduce_scratch_min_bytes(), "scratch smaller than avsi_reduce_scratch_min_bytes()");
  // the tickets of the last-block reductions start at zero and reset themselves; the zeroing is complete when this call
  // returns, whichever stream the reductions then run on (not callable inside a stream capture)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  AVSI_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
  AVSI_REQUIRE(cap == cudaStreamCaptureStatusNone, "not inside a CUDA-graph capture");
  AVSI_CUDA(cudaMemsetAsync(scratch, 0, REDUCE_COUNTER_BYTES, (cudaStream_t)stream));
  AVSI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  g_reduce_scratch[dev] = ReduceScratch{reinterpret_cast<unsigned*>(scratch),
                                        reinterpret_cast<unsigned char*>(scratch) + REDUCE_COUNTER_BYTES,
                                        (long long)bytes - REDUCE_COUNTER_BYTES};
  return AVSI_OK;
}

// counters + the partial sums of avsi_masked_l1 / avsi_colsum_f16 at their largest grids (the split-K GEMMs ask for more:
// avsi_gemm_f16_scratch_bytes)
extern "C" int64_t avsi_reduce_scratch_min_bytes(void) { return avsi::REDUCE_COUNTER_BYTES + (4LL << 20); }

extern "C" const char* avsi_last_error(void) { return avsi::g_last_error; }

extern "C" const char* avsi_version(void) {
  return "avsi_b200 0.1 (sm_100a; tcgen05 GEMM, cluster-persistent LSTM, fused STFT front end)";
}

extern "C" int64_t avsi_launch_count(void) { return (int64_t)avsi::g_launch_count.load(); }

// AVSI_* tuning / test switches (kernel selection, activation variant, phase timers) are read from the environment
// once and cached; call this after changing them inside a running process (the parity tests do)
extern "C" void avsi_reload_env(void) { avsi::g_env_gen.fetch_add(1); }

// ABI self-check for the ctypes mirror of the argument structs
extern "C" int avsi_sizeof_frontend_args(void) { return (int)sizeof(avsi_frontend_args); }
extern "C" int avsi_sizeof_istft_args(void) { return (int)sizeof(avsi_istft_args); }
