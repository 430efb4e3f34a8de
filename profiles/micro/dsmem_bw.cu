// Micro-benchmark: per-SM DSMEM exchange bandwidth inside a 4-CTA cluster, the pattern of the recurrence kernels
// (every CTA sends 16 B per thread to each of its 3 peers per round = 24 KB out + 24 KB in at 512 threads).
//   mode 0: cp.async.bulk shared::cta -> shared::cluster, 3 x 8 KB per round issued by one thread
//   mode 1: st.async.v4 (16 B per thread and peer) with mbarrier complete_tx
//   mode 2: mode 0 with the 8 KB split into 4 x 2 KB copies
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bw dsmem_bw.cu ; run: ./dsmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t r) {
  uint32_t o;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
  return o;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(512, 1) k(int rounds, unsigned long long* out, int bytes_per_peer) {
  extern __shared__ __align__(128) unsigned char sm[];
  // [0, 32K) send buffer, [32K, 32K + 3 * 8K) receive slots, then barrier
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 96 * 1024);
  const uint32_t bar_s = smem_u32(bar), send_s = smem_u32(sm), recv_s = smem_u32(sm + 32 * 1024);
  const int tid = threadIdx.x;
  uint32_t j;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(j));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 8192; i += 512) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t total = 3u * (uint32_t)bytes_per_peer;
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  uint32_t dst[3], dbar[3];
  for (int c = 0; c < 3; ++c) {
    const uint32_t peer = (j + 1 + c) & 3;
    dst[c] = map_to_cta(recv_s + (uint32_t)(2 - c) * 8192u, peer);
    dbar[c] = map_to_cta(bar_s, peer);
  }
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (MODE == 0 || MODE == 2) {
      if (tid == 0) {
        const int pieces = (MODE == 2) ? 4 : 1;
        const uint32_t pb = (uint32_t)bytes_per_peer / pieces;
        for (int c = 0; c < 3; ++c)
          for (int q = 0; q < pieces; ++q)
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst[c] + q * pb), "r"(send_s + q * pb), "r"(pb), "r"(dbar[c]) : "memory");
      }
    } else {
      if (tid * 16 < bytes_per_peer) {
        const uint4 v = reinterpret_cast<const uint4*>(sm)[tid];
        for (int c = 0; c < 3; ++c)
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                       ::"r"(dst[c] + tid * 16), "r"(v.x + r), "r"(v.y), "r"(v.z), "r"(v.w), "r"(dbar[c]) : "memory");
      }
    }
    mbar_wait(bar_s, (uint32_t)(r & 1));                 // my 3 incoming pieces of this round have landed
    __syncthreads();
    if (tid == 0 && r + 1 < rounds) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(total) : "memory");
    // everybody re-armed before anybody sends again
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  const int smem = 96 * 1024 + 64, rounds = 2000;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {4, 128}) {
    for (int bytes : {8192, 4096, 2048}) {
      for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) k<0><<<grid, 512, smem>>>(rounds, d, bytes);
          if (mode == 1) k<1><<<grid, 512, smem>>>(rounds, d, bytes);
          if (mode == 2) k<2><<<grid, 512, smem>>>(rounds, d, bytes);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long cyc;
        cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double per_round = (double)cyc / rounds;
        printf("{\"grid\": %d, \"mode\": %d, \"bytes_per_peer\": %d, \"cycles_per_round\": %.0f, \"out_bytes_per_clk\": %.1f}\n", grid, mode, bytes,
               per_round, 3.0 * bytes / per_round);
      }
    }
  }
  return 0;
}
