// Library-level state of the C ABI (include/avsi_b200.h): last-error text, version, launch counter.
#include "common.cuh"

namespace avsi {
thread_local char g_last_error[512] = {0};
std::atomic<long long> g_launch_count{0};
std::atomic<int> g_env_gen{0};
}  // namespace avsi

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
