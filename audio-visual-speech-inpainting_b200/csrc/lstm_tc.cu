// Persistent bidirectional LSTM recurrence on tcgen05 (5th-gen tensor cores), 128-row batch tiles.
//
// Same contract, cluster decomposition and DSMEM exchange as lstm.cu (8 CTAs per (direction,
// batch tile), CTA j owns hidden units [32j, 32j+32) with all four gates), but the recurrent
// product h_{t-1} . W_hh runs as tcgen05.mma (M = 128 batch rows, N = 128 gate columns, K = 256)
// with the CTA's W_hh slice RESIDENT IN SHARED MEMORY for the whole sequence (64 KB, canonical
// K-major SWIZZLE_128B) and the accumulator in TENSOR MEMORY.  ncu / microbenchmarks of the
// mma.sync version showed the legacy HMMA path (about 32 cycles per m16n8k16 per SM
// sub-partition on this part) bounding the step at ~0.08 us per batch row; one tcgen05 step is
// 16 MMAs of 64 cycles for 128 rows.
//
// Step: wait (mbarrier) for the 64 KB h_{t-1} tile pushed by the 8 CTAs -> one thread issues
// 16 tcgen05.mma + commit -> every thread (one batch row x 16 units) tcgen05.ld's its
// accumulators, adds the prefetched x.W_ih pre-activations and bias, runs the cell update with
// c_t in fp32 registers, stores gates / c / h, and pushes its 16 h values (two 16-byte
// chunk) into ITS block of the A-operand tile (K-chunks of 32 units = one source CTA each,
// SWIZZLE_64B) and one thread ships the 8 KB block to the 7 peers with cp.async.bulk DSMEM copies
// (complete_tx on the peers' mbarriers).  Per-thread remote stores (st.async) were measured to be
// packet-rate bound at ~3.5 cycles per store per SM; one CTA barrier per step, no cluster barrier.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avsi {

constexpr int TC_BT = 128;          // batch rows per cluster (= UMMA M)
constexpr int TC_CL = 8;
constexpr int TC_THREADS = 512;     // 16 warps: TMEM lane quarter (w & 3) x 32-column group (w >> 2)
constexpr int TC_HP = 256;
constexpr int TC_G = 1024;
constexpr int TC_BLK = 128 * 64;    // bytes of one K-chunk block: 128 rows x 32 halves (SWIZZLE_64B)

struct LstmTcFwdSmem {
  unsigned char w[8][TC_BLK];       // W_hh slice: [k-chunk of 32][gate column n][32 halves], SW64
  unsigned char h[2][8][TC_BLK];    // h_{t-1} tile, double buffered: [source CTA = k-chunk][batch row][32 halves], SW64
  float bias[128];
  unsigned long long hfull[2];
  unsigned long long mma_done;
  uint32_t tmem_slot;
};

// phase timers (cycles summed over steps) of thread 0 / thread 511 of CTA 0; debug only
__device__ unsigned long long g_tc_timing[16];
#define TC_TICK(i)                                         \
  do {                                                     \
    if (timing) {                                          \
      const long long now_ = clock64();                    \
      tacc[i] += (unsigned long long)(now_ - tprev);       \
      tprev = now_;                                        \
    }                                                      \
  } while (0)

// byte offset of the 16-byte chunk c (0..3) of row `row` inside one K-major SWIZZLE_64B block
__device__ __forceinline__ uint32_t sw64_chunk(int row, int c) {
  return (uint32_t)(row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
}

__global__ void __cluster_dims__(TC_CL, 1, 1) __launch_bounds__(TC_THREADS, 1)
lstm_fwd_tc_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whh, const float* __restrict__ bias,
                   uint16_t* __restrict__ y, float* __restrict__ cst, int T, int B) {
  extern __shared__ unsigned char tc_smem_raw[];
  const uint32_t base_s = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  LstmTcFwdSmem& sm = *reinterpret_cast<LstmTcFwdSmem*>(tc_smem_raw + (base_s - smem_u32(tc_smem_raw)));
  constexpr uint32_t PHASE_BYTES = (TC_CL - 1) * TC_BLK;  // 7 peers x 8 KB per step

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int q = w & 3, cg = w >> 2;              // TMEM lane quarter / 32-column group of this warp
  const int r = q * 32 + lane;                   // batch row inside the tile (= TMEM lane)
  const int cid = blockIdx.x / TC_CL, j = blockIdx.x % TC_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * TC_BT;
  const int row = b0 + r;
  const bool row_ok = row < B;
  const int ul0 = cg * 8;                        // first local unit of this thread (8 units x 4 gates = 32 columns)
  const int ug0 = j * 32 + ul0;                  // first global unit

  const uint32_t w_s = smem_u32(&sm.w[0][0]);
  const uint32_t h_s = smem_u32(&sm.h[0][0][0]);
  const uint32_t hfull_s = smem_u32(&sm.hfull[0]);
  const uint32_t done_s = smem_u32(&sm.mma_done);

  if (tid == 0) {
    mbar_init(hfull_s, 1);
    mbar_init(hfull_s + 8, 1);
    mbar_init(done_s, 1);
    fence_barrier_init();
    if (T > 1) mbar_expect_tx(hfull_s, PHASE_BYTES);
    if (T > 2) mbar_expect_tx(hfull_s + 8, PHASE_BYTES);
  }
  __syncwarp();
  if (w == 0) tmem_alloc(smem_u32(&sm.tmem_slot), 128);
  // W_hh slice -> smem (SW64 blocks of 32 k), bias -> smem
  for (int idx = tid; idx < 128 * 32; idx += TC_THREADS) {
    const int n = idx >> 5, c = idx & 31;        // gate column, 16-byte chunk of the 256-long K row
    const uint4 v = *reinterpret_cast<const uint4*>(whh + ((long long)(dir * TC_G + j * 128 + n)) * TC_HP + c * 8);
    *reinterpret_cast<uint4*>(&sm.w[c >> 2][sw64_chunk(n, c & 3)]) = v;
  }
  if (tid < 128) sm.bias[tid] = bias[dir * TC_G + j * 128 + tid];
  fence_proxy_async();                           // generic-proxy smem writes -> visible to tcgen05.mma
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm.tmem_slot);
  const uint32_t idesc = make_idesc(128, 128, 0, 0);
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);

  float c_state[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c_state[i] = 0.f;
  float4 bq[8];                                  // bias of this thread's 8 units (same for every row)
#pragma unroll
  for (int i = 0; i < 8; ++i) bq[i] = *reinterpret_cast<const float4*>(&sm.bias[(ul0 + i) * 4]);

  // this thread's 16-byte chunk (row r, units ug0..ug0+7) inside block j of the h tile
  unsigned char* my_chunk = &sm.h[0][j][sw64_chunk(r, cg)];

  cluster_sync_all();                            // all CTAs' mbarriers are initialised and armed

  const bool timing = (blockIdx.x == 0) && (tid == 0 || tid == TC_THREADS - 1);
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    const long long grow = (long long)t * B + row;
    // ---- prefetch the 32 pre-activations (8 units x 4 gates = 64 B) of this row -----------
    uint4 pre[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pre[i] = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok) {
      const uint4* src = reinterpret_cast<const uint4*>(gates + grow * (2 * TC_G) + dir * TC_G + ug0 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) pre[i] = src[i];
    }
    uint32_t acc[32];
    TC_TICK(0);                                  // prefetch issue
    if (s > 0) {
      const int pb = (s - 1) & 1;
      if (tid == 0) {
        mbar_wait(hfull_s + 8 * pb, ((s - 1) >> 1) & 1);         // the 7 peer blocks of h_{s-1} have landed
        TC_TICK(1);                              // wait for peers' h
        if (s + 1 < T - 1) mbar_expect_tx(hfull_s + 8 * pb, PHASE_BYTES);
        tc_fence_after();
        const uint32_t a0 = h_s + (uint32_t)pb * 8u * TC_BLK;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint32_t off = (uint32_t)(ks >> 1) * TC_BLK + (uint32_t)(ks & 1) * 32u;
          umma_f16(tmem_base, make_smem_desc(a0 + off, 0u, 512u, 4u), make_smem_desc(w_s + off, 0u, 512u, 4u), idesc,
                   ks > 0 ? 1u : 0u);
        }
        umma_commit(done_s);
        TC_TICK(2);                              // MMA issue
      }
      __syncwarp();
      mbar_wait(done_s, (uint32_t)((s - 1) & 1));
      TC_TICK(3);                                // wait for MMA completion
      tc_fence_after();
      tmem_ld32(taddr, acc);
      tmem_ld_wait();
      tc_fence_before();                         // our tcgen05.ld precedes the next step's MMA (ordered by the barrier below)
      TC_TICK(4);                                // TMEM load
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0u;
    }
    uint4 gout[4];
    float cout[8], hout[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 pv = pre[i >> 1];
      const float2 ig = unpack_half2((i & 1) ? pv.z : pv.x), fo = unpack_half2((i & 1) ? pv.w : pv.y);
      const float gi = sigmoid_fast2(__uint_as_float(acc[4 * i + 0]) + ig.x + bq[i].x);
      const float gg = tanh_fast2(__uint_as_float(acc[4 * i + 1]) + ig.y + bq[i].y);
      const float gf = sigmoid_fast2(__uint_as_float(acc[4 * i + 2]) + fo.x + bq[i].z);
      const float go = sigmoid_fast2(__uint_as_float(acc[4 * i + 3]) + fo.y + bq[i].w);
      const float cc = fmaf(gf, c_state[i], gi * gg);
      c_state[i] = cc;
      cout[i] = cc;
      hout[i] = go * tanh_fast2(cc);
      const uint32_t g0 = pack_half2(gi, gg), g1 = pack_half2(gf, go);
      if (i & 1) {
        gout[i >> 1].z = g0;
        gout[i >> 1].w = g1;
      } else {
        gout[i >> 1].x = g0;
        gout[i >> 1].y = g1;
      }
    }
    const uint4 hv = make_uint4(pack_half2(hout[0], hout[1]), pack_half2(hout[2], hout[3]),
                                pack_half2(hout[4], hout[5]), pack_half2(hout[6], hout[7]));
    TC_TICK(5);                                  // cell update
    if (s + 1 < T) {
      // stage this CTA's h_t slice in its own block of the (s & 1) tile, then ship the whole 8 KB block
      // to the 7 peers with one bulk DSMEM copy each (per-thread remote stores are packet-rate bound)
      *reinterpret_cast<uint4*>(my_chunk + (size_t)(s & 1) * 8 * TC_BLK) = hv;
      fence_proxy_async();
      __syncthreads();
      TC_TICK(6);                                // stage + CTA barrier
      if (tid == 0) {
        const uint32_t src = h_s + (uint32_t)((s & 1) * 8 + j) * TC_BLK;
#pragma unroll
        for (int d = 0; d < TC_CL; ++d)
          if (d != j) bulk_copy_to_cta(map_to_cta(src, (uint32_t)d), src, TC_BLK, map_to_cta(hfull_s + 8u * (s & 1), (uint32_t)d));
      }
    }
    if (row_ok) {
      uint4* gdst = reinterpret_cast<uint4*>(gates + grow * (2 * TC_G) + dir * TC_G + ug0 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) gdst[i] = gout[i];
      float4* cdst = reinterpret_cast<float4*>(cst + grow * (2 * TC_HP) + dir * TC_HP + ug0);
      cdst[0] = make_float4(cout[0], cout[1], cout[2], cout[3]);
      cdst[1] = make_float4(cout[4], cout[5], cout[6], cout[7]);
      *reinterpret_cast<uint4*>(y + grow * (2 * TC_HP) + dir * TC_HP + ug0) = hv;
    }
    TC_TICK(7);                                  // bulk-copy issue + global stores
  }
  if (timing) {
    const int o = (tid == 0) ? 0 : 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) g_tc_timing[o + i] = tacc[i];
  }
  tc_fence_before();
  cluster_sync_all();                            // no CTA exits while peers may still address its smem
  if (w == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

int launch_lstm_fwd_tc(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst, int T, int B,
                       cudaStream_t st) {
  static bool attr_done = false;
  const int smem = (int)sizeof(LstmTcFwdSmem) + 1024;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(lstm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int grid = 2 * ((B + TC_BT - 1) / TC_BT) * TC_CL;
  lstm_fwd_tc_kernel<<<grid, TC_THREADS, smem, st>>>(gates, whh, bias, y, cst, T, B);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

}  // namespace avsi

// debug: cycles per phase summed over the steps of the last lstm_fwd_tc launch (thread 0: [0,8), thread 511: [8,16))
extern "C" int avsi_debug_lstm_tc_timing(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, avsi::g_tc_timing, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}
