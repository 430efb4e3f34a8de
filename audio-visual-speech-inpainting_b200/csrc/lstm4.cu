// Persistent bidirectional LSTM recurrence on tcgen05, cluster of 4 CTAs per (direction, 128-row
// batch tile).  Forward kernel.
//
// Same contract as lstm.cu (avsi_lstm_fwd).  Decomposition: CTA j of the cluster owns hidden units
// [64j, 64j+64) with all four gates = 256 gate columns.  Its W_hh slice [256 gate columns x 256]
// (128 KB fp16) is RESIDENT IN SHARED MEMORY for the whole sequence, the recurrent product
// h_{t-1} . W_hh^T runs on tcgen05.mma (M = 128 batch rows, N = 256, K = 256, accumulator DOUBLE
// BUFFERED in 2 x 256 TMEM columns), the cell update runs on 16 warps straight out of TMEM.
//
// Why 4 CTAs: the per-step all-gather of h_t through distributed shared memory is the scarce
// resource (measured 12..22 B/clk per SM for bulk DSMEM copies, profiles/README.md).  With 8 CTAs per
// 128 rows every SM ships 56 KB per step and uses 1/8 of its tensor/MUFU throughput; with 4 CTAs
// it ships 48 KB for twice the work, and 33 clusters (132 SMs) are co-resident instead of 15 (120).
//
// The step is software-pipelined in two halves of the hidden units (pass p = units 16cg + 8p .. +8 of each
// warp's 16 units; K is ordered [pass][source CTA][warp] so that a CTA's half is one contiguous 8 KB
// block of the A tile):
//   compute warps    prefetch x.W_ih pre-activations -> wait `done` (this step's MMA chain retired) ->
//                    pass 0: tcgen05.ld 32 columns, gates (the pre-activations already hold the bias and the 1/2 of the
//                    sigmoid gates: sigma(z) = 1/2 tanh(z/2) + 1/2 is one MUFU + one FFMA) / cell update (c_t in fp32 registers), stores,
//                    wait `afree` (every CTA's chain retired: the peers have consumed the previous push),
//                    write the h_t chunk into the LOCAL A tile, arrive staged[0] -> pass 1 likewise.
//   control thread   staged[0] -> 3 bulk DSMEM copies of the 8 KB half into the peers' A tiles (complete_tx
//                    on their hfull[0]); staged[1] -> second half; hfull[0] -> first K-half of the NEXT
//                    step's chain (8 tcgen05.mma into the other accumulator buffer, overlapping the second
//                    half's transfer); hfull[1] -> second K-half; tcgen05.commit -> `done` (local) and,
//                    multicast, `afree` of all 4 CTAs.
// The A tile is K-major SWIZZLE_NONE: [k-chunk of 8 units][row][16 B]; a warp's stores are 512 contiguous
// bytes (conflict free).  Gate pre-activations / activated gates and the c_t stash are INTERLEAVED in
// global memory (common.cuh il16 / il32): every warp access is a run of 512 contiguous bytes.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avsi {

constexpr int L4_BT = 128;            // batch rows per cluster (UMMA M)
constexpr int L4_CL = 4;              // CTAs per cluster
constexpr int L4_NC = 256;            // gate columns per CTA (UMMA N) = 64 units x 4 gates
constexpr int L4_HP = 256;            // padded hidden size (UMMA K total)
constexpr int L4_G = 1024;            // gate columns per direction
constexpr int L4_CWARPS = 16;         // compute warps
constexpr int L4_THREADS = (L4_CWARPS + 1) * 32;
constexpr uint32_t L4_W_BYTES = L4_NC * L4_HP * 2;      // 131072
constexpr uint32_t L4_A_BYTES = L4_BT * L4_HP * 2;      // 65536
constexpr uint32_t L4_HALF_BYTES = L4_BT * 32 * 2;      // 8192: one CTA's half (32 units) of the A tile
constexpr uint32_t L4_W_LBO = 4096, L4_W_SBO = 128;     // W: [k-chunk][gate column][16 B]
constexpr uint32_t L4_A_LBO = 2048, L4_A_SBO = 128;     // A: [k-chunk][row][16 B]

struct Lstm4Smem {
  unsigned char w[L4_W_BYTES];
  unsigned char a[L4_A_BYTES];
  unsigned long long hfull[2];        // tx barriers: peers' halves of h have landed
  unsigned long long done;            // local MMA chain retired
  unsigned long long afree[2];        // all 4 CTAs' chains of the step retired (multicast commit), by step parity
  unsigned long long staged[2];       // the 16 compute warps have written half p of h_t into the local A tile
  uint32_t tmem_slot;
  // BULK variant: per-warp staging of one pass's activated gates [4 column chunks][32 rows][16 B] = the 2 KB run they
  // occupy in the interleaved global layout, written to HBM by ONE cp.async.bulk per warp and pass
  alignas(128) unsigned char gstage[L4_CWARPS][2048];
};

// ACT = 0: tanh.approx.f32 (1 MUFU per activation, 2^-11 relative -- the precision h_t is stored in)
// ACT = 1: ex2.approx + rcp.approx (2 MUFU, ~2 ulp)
// argument = z / 2 (the forward weight copies and the projection bias are pre-halved for the sigmoid gates)
template <int ACT>
__device__ __forceinline__ float act_sigmoid_half(float xh) {
  if (ACT == 0) return fmaf(0.5f, tanhf_fast(xh), 0.5f);
  return sigmoid_fast2(2.f * xh);
}
template <int ACT>
__device__ __forceinline__ float act_tanh(float x) {
  if (ACT == 0) return tanhf_fast(x);
  return tanh_fast2(x);
}

// phase timers (cycles summed over steps) of CTA 0: control thread [0,8), compute thread 0 [8,16); debug only
__device__ unsigned long long g_l4_timing[16];
#define L4_TICK(i)                                         \
  do {                                                     \
    if (TIMING && timing) {                                \
      const long long now_ = clock64();                    \
      tacc[i] += (unsigned long long)(now_ - tprev);       \
      tprev = now_;                                        \
    }                                                      \
  } while (0)

// BULK (B % 32 == 0: a warp's 32 rows are one row block of the interleaved layout): the activated gates leave through
// shared memory + cp.async.bulk (async proxy) instead of 4 STG.128 per thread and pass
// CFENCE: the generic -> async proxy fence is executed on the consumer side (the control thread, after its acquire of
//   the hand-over barrier, with no memory operations of its own in flight) instead of by each of the 512 writers, where
//   MEMBAR.ALL.CTA waits for the thread's global loads and stores.
// BPF (B % 32 == 0): the next steps' pre-activation lines are pulled into L2 by four 16 KB cp.async.bulk.prefetch of the
//   control thread instead of eight CCTL per compute thread (LSU queue slots).
template <int ACT, bool TIMING, bool BULK, bool CFENCE, bool BPF>
__global__ void __cluster_dims__(L4_CL, 1, 1) __launch_bounds__(L4_THREADS, 1)
lstm4_fwd_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whh, const float* __restrict__ bias,
                 uint16_t* __restrict__ y, float* __restrict__ cst, int T, int B, int PREFETCH, int y_il) {
  extern __shared__ unsigned char l4_smem_raw[];
  const uint32_t raw_s = smem_u32(l4_smem_raw);
  const uint32_t base_s = (raw_s + 127u) & ~127u;
  Lstm4Smem& sm = *reinterpret_cast<Lstm4Smem*>(l4_smem_raw + (base_s - raw_s));

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int cid = blockIdx.x / L4_CL, j = blockIdx.x % L4_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * L4_BT;

  const uint32_t w_s = smem_u32(&sm.w[0]);
  const uint32_t a_s = smem_u32(&sm.a[0]);
  const uint32_t hfull_s = smem_u32(&sm.hfull[0]);
  const uint32_t done_s = smem_u32(&sm.done);
  const uint32_t afree_s = smem_u32(&sm.afree[0]);
  const uint32_t staged_s = smem_u32(&sm.staged[0]);
  constexpr uint32_t PUSH_BYTES = (L4_CL - 1) * L4_HALF_BYTES;

  if (tid == 0) {
    mbar_init(hfull_s, 1);
    mbar_init(hfull_s + 8, 1);
    mbar_init(done_s, 1);
    mbar_init(afree_s, L4_CL);
    mbar_init(afree_s + 8, L4_CL);
    mbar_init(staged_s, L4_CWARPS);
    mbar_init(staged_s + 8, L4_CWARPS);
    fence_barrier_init();
    if (T > 1) {                                              // push round 0 (h_0)
      mbar_expect_tx(hfull_s, PUSH_BYTES);
      mbar_expect_tx(hfull_s + 8, PUSH_BYTES);
    }
  }
  if (w == L4_CWARPS) tmem_alloc(smem_u32(&sm.tmem_slot), 512);
  // W_hh slice -> smem.  K position kpos = [pass p][source CTA j'][warp cg] <-> unit chunk 8j' + 2cg + p;
  // element (gate column n, kpos, e) at kpos*4096 + n*16 + e*2
  for (int idx = tid; idx < L4_NC * 32; idx += L4_THREADS) {
    const int n = idx & (L4_NC - 1), kpos = idx >> 8;
    const int c = 8 * ((kpos >> 2) & 3) + 2 * (kpos & 3) + (kpos >> 4);
    const uint4 v = *reinterpret_cast<const uint4*>(whh + ((long long)(dir * L4_G + j * L4_NC + n)) * L4_HP + c * 8);
    *reinterpret_cast<uint4*>(&sm.w[(uint32_t)kpos * L4_W_LBO + (uint32_t)n * 16u]) = v;
  }
  // The bias rides on the MMA: hidden unit 255 is padding (H = 250), its slot of the A tile is held at 1.0 and the
  // k = 255 column of the shared-memory copy of W_hh holds the (prescaled) bias of every gate column, so
  // h_{t-1} . W_hh^T + b comes out of the tensor core.  Unit 255 = k-chunk position 31, element 7.
  __syncthreads();
  if (tid < L4_NC)
    *reinterpret_cast<uint16_t*>(&sm.w[31u * L4_W_LBO + (uint32_t)tid * 16u + 14u]) =
        __half_as_ushort(__float2half_rn(bias[dir * L4_G + j * L4_NC + tid]));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm.tmem_slot);
  cluster_sync_all();                               // every CTA's barriers are initialised and armed
  const bool timing = TIMING && (blockIdx.x == 0) && (tid == 0 || tid == L4_CWARPS * 32);
  unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();

  if (w == L4_CWARPS) {
    // ===================================================================== control warp
    if (lane == 0) {
      const uint32_t idesc = make_idesc(L4_BT, L4_NC, 0, 0);
      uint32_t dst_a[L4_CL], dst_bar[L4_CL];
#pragma unroll
      for (int d = 0; d < L4_CL; ++d) {
        dst_a[d] = map_to_cta(a_s, (uint32_t)d);
        dst_bar[d] = map_to_cta(hfull_s, (uint32_t)d);
      }
      // 16 KB runs of this CTA's 256 gate columns (interleaved layout: one run per 32-row block) of chain step sp
      auto prefetch_step = [&](int sp) {
        const int tt = dir ? (T - 1 - sp) : sp;
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
          if (b0 + 32 * rb < B)
            bulk_prefetch_l2(gates + il16((long long)tt * B + b0 + 32 * rb, dir * L4_G + j * L4_NC, 2 * L4_G), 16384u);
      };
      if (BPF)
        for (int sp = 1; sp < PREFETCH && sp < T; ++sp) prefetch_step(sp);
      for (int s = 0; s + 1 < T; ++s) {
        if (BPF && s + PREFETCH < T) prefetch_step(s + PREFETCH);      // PREFETCH = distance in steps
        const uint32_t td = tmem_base + (uint32_t)(((s + 1) & 1) * L4_NC);
        auto push_half = [&](int p) {          // ship half p of h_s as soon as the warps have staged it
          mbar_wait(staged_s + 8u * p, (uint32_t)(s & 1));
          if (CFENCE) fence_proxy_async();
          L4_TICK(p);
          const uint32_t off = (uint32_t)(16 * p + 4 * j) * L4_A_LBO;
#pragma unroll
          for (int d = 0; d < L4_CL; ++d)
            if (d != j) bulk_copy_to_cta(dst_a[d] + off, a_s + off, L4_HALF_BYTES, dst_bar[d] + 8u * p);
        };
        auto mma_half = [&](int p) {           // K-half p of the chain of step s+1, once the peers' halves have landed
          mbar_wait(hfull_s + 8u * p, (uint32_t)(s & 1));
          if (s + 2 < T) mbar_expect_tx(hfull_s + 8u * p, PUSH_BYTES);    // re-arm for round s+1
          L4_TICK(2 + 2 * p);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t kc = (uint32_t)(16 * p + 2 * ks);
            umma_f16(td, make_smem_desc(a_s + kc * L4_A_LBO, L4_A_LBO, L4_A_SBO, 0u),
                     make_smem_desc(w_s + kc * L4_W_LBO, L4_W_LBO, L4_W_SBO, 0u), idesc, (p > 0 || ks > 0) ? 1u : 0u);
          }
          L4_TICK(3 + 2 * p);
        };
        // (issuing the first K-half before the second push, i.e. under the second cell-update pass, was measured 13 % SLOWER:
        // the chain's TMEM / shared-memory traffic slows that pass down by more than the overlap gains, profiles/README.md)
        push_half(0);
        push_half(1);
        mma_half(0);
        mma_half(1);
        umma_commit(done_s);
        umma_commit_mc(afree_s + 8u * (s & 1), (uint16_t)0xF);
      }
    }
    if (timing) {
#pragma unroll
      for (int i = 0; i < 8; ++i) g_l4_timing[i] = tacc[i];
    }
    __syncwarp();
  } else {
    // ===================================================================== compute warps
    const int q = w & 3, cg = w >> 2;               // TMEM lane quarter / column group of this warp
    const int r = q * 32 + lane;                    // batch row inside the tile (= TMEM lane)
    const int row = b0 + r;
    const bool row_ok = row < B;
    float c_state[2][8];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) c_state[p][i] = 0.f;

    for (int s = 0; s < T; ++s) {
      const int t = dir ? (T - 1 - s) : s;
      const long long grow = (long long)t * B + row;
      // ---- prefetch the pre-activations of both passes (2 x 8 units x 4 gates x 2 B = 128 B) --------------
      uint4 pre[2][4];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i) pre[p][i] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int col0 = dir * L4_G + (j * 64 + cg * 16 + p * 8) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pre[p][i] = *reinterpret_cast<const uint4*>(gates + il16(grow, col0 + 8 * i, 2 * L4_G));
        }
      }
      // the lines of the NEXT step are pulled into L2 now (a whole step ahead), so that its loads -- issued only
      // when this step's epilogue is done -- are L2 hits instead of DRAM reads queued behind the stores
      if (!BPF && PREFETCH && row_ok && s + 1 < T && (lane & 7) == 0) {          // one lane per 128-byte line
        const long long gn = (long long)(dir ? (t - 1) : (t + 1)) * B + row;
        const int col0 = dir * L4_G + (j * 64 + cg * 16) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) prefetch_l2(gates + il16(gn, col0 + 8 * i, 2 * L4_G));
      }
      L4_TICK(0);
      if (s > 0) {
        mbar_wait(done_s, (uint32_t)((s - 1) & 1));
        tc_fence_after();
      }
      L4_TICK(1);
      const uint32_t td = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((s & 1) * L4_NC);
      uint4 hv[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int ul0 = cg * 16 + p * 8;            // first local unit of this pass (a warp owns 16 contiguous units)
        uint4 gout[4];
        float hout[8];
        // the 32 accumulator columns of a pass are read in two halves of 4 units: 16 live accumulator registers instead
        // of 32 (with 17 warps per CTA the budget is 96 registers per thread; the 32-column read spilled, and the
        // spill reloads queue behind the step's global stores in the LSU)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t acc[16];
          if (s > 0) {
            tmem_ld16(td + (uint32_t)((ul0 + 4 * hh) * 4), acc);
            tmem_ld_wait();
          } else {                                  // h_{-1} = 0: no chain ran, the bias comes from global memory once
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias + dir * L4_G + (j * 64 + ul0 + 4 * hh + i) * 4);
              acc[4 * i + 0] = __float_as_uint(b4.x);
              acc[4 * i + 1] = __float_as_uint(b4.y);
              acc[4 * i + 2] = __float_as_uint(b4.z);
              acc[4 * i + 3] = __float_as_uint(b4.w);
            }
          }
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int i = 4 * hh + ii;
            const uint4 pv = pre[p][i >> 1];
            const float2 ig = unpack_half2((i & 1) ? pv.z : pv.x), fo = unpack_half2((i & 1) ? pv.w : pv.y);
            const float gi = act_sigmoid_half<ACT>(__uint_as_float(acc[4 * ii + 0]) + ig.x);
            const float gg = act_tanh<ACT>(__uint_as_float(acc[4 * ii + 1]) + ig.y);
            const float gf = act_sigmoid_half<ACT>(__uint_as_float(acc[4 * ii + 2]) + fo.x);
            const float go = act_sigmoid_half<ACT>(__uint_as_float(acc[4 * ii + 3]) + fo.y);
            const float cc = fmaf(gf, c_state[p][i], gi * gg);
            c_state[p][i] = cc;
            hout[i] = go * act_tanh<ACT>(cc);
            const uint32_t g0 = pack_half2(gi, gg), g1 = pack_half2(gf, go);
            if (i & 1) {
              gout[i >> 1].z = g0;
              gout[i >> 1].w = g1;
            } else {
              gout[i >> 1].x = g0;
              gout[i >> 1].y = g1;
            }
          }
        }
        hv[p] = make_uint4(pack_half2(hout[0], hout[1]), pack_half2(hout[2], hout[3]),
                           pack_half2(hout[4], hout[5]), pack_half2(hout[6], hout[7]));
        if (s + 1 < T) {
          // every CTA's chain of this step has retired: the A tiles are free and the peers have consumed the
          // previous push (whose source is the block overwritten here)
          if (p == 0 && s > 0) mbar_wait(afree_s + 8u * ((s - 1) & 1), (uint32_t)(((s - 1) >> 1) & 1));
          uint4 ha = hv[p];
          if (p == 1 && cg == 3 && j == L4_CL - 1) ha.w = (ha.w & 0xFFFFu) | 0x3C000000u;   // unit 255 := 1.0 (bias slot)
          *reinterpret_cast<uint4*>(&sm.a[(uint32_t)(16 * p + 4 * j + cg) * L4_A_LBO + (uint32_t)r * 16u]) = ha;
          if (!CFENCE) fence_proxy_async();          // generic-proxy writes -> visible to UMMA / bulk copies
          tc_fence_before();                         // our tcgen05.ld's precede the chain that reuses this buffer
          __syncwarp();
          if (lane == 0) mbar_arrive_local(staged_s + 8u * p);
        }
        if (BULK) {
          // the previous pass's bulk store of this warp has read the staging block (it was issued a whole pass ago)
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          unsigned char* stg = &sm.gstage[w][0];
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + i * 512 + lane * 16) = gout[i];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row_ok) {               // lane 0 = first row of the block; B % 32 == 0: the warp is all valid or all out
            bulk_store_s2g(gates + il16(grow, dir * L4_G + (j * 64 + ul0) * 4, 2 * L4_G), smem_u32(stg), 2048u);
            bulk_commit();
          }
        }
        if (row_ok) {
          const int ug0 = j * 64 + ul0;
          if (!BULK) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(gates + il16(grow, dir * L4_G + ug0 * 4 + 8 * i, 2 * L4_G)) = gout[i];
          }
          *reinterpret_cast<float4*>(cst + il32(grow, dir * L4_HP + ug0, 2 * L4_HP)) =
              make_float4(c_state[p][0], c_state[p][1], c_state[p][2], c_state[p][3]);
          *reinterpret_cast<float4*>(cst + il32(grow, dir * L4_HP + ug0 + 4, 2 * L4_HP)) =
              make_float4(c_state[p][4], c_state[p][5], c_state[p][6], c_state[p][7]);
        }
        L4_TICK(2 + p);
      }
      if (row_ok) {
        const int u0 = dir * L4_HP + j * 64 + cg * 16;
        if (y_il) {                                  // interleaved y: two 512-byte runs per warp instead of 32 scattered sectors
          *reinterpret_cast<uint4*>(y + il16(grow, u0, 2 * L4_HP)) = hv[0];
          *reinterpret_cast<uint4*>(y + il16(grow, u0 + 8, 2 * L4_HP)) = hv[1];
        } else {                                     // h_t of the warp's 16 units: one 32-byte store per row
          st_global_v8(y + grow * (2 * L4_HP) + u0, hv[0], hv[1]);
        }
      }
      L4_TICK(4);
    }
    if (timing) {
#pragma unroll
      for (int i = 0; i < 8; ++i) g_l4_timing[8 + i] = tacc[i];
    }
    if (BULK && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  cluster_sync_all();                                // no CTA exits while peers may still address its smem
  if (w == L4_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_lstm4_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst, int T, int B,
                     int y_il, cudaStream_t st) {
  // AVSI_LSTM_ACT=exact: ex2/rcp activations; AVSI_L4_TIMING=1: in-kernel phase timers (profiles/bench_lstm.py)
  AVSI_ENV_CACHE(mode, env_is("AVSI_LSTM_ACT", "exact") | (env_is("AVSI_L4_TIMING", "1") << 1));
  AVSI_ENV_CACHE(pf, env_int("AVSI_L4_PREFETCH", 1));            // L2 prefetch distance in steps (bulk variant; the per-thread variant always looks one step ahead); 0: none
  AVSI_ENV_CACHE(bulk_env, env_int("AVSI_L4_BULK", 0));           // AVSI_L4_BULK=1: gates through shared memory + cp.async.bulk (measured: no gain, profiles/README.md)
  AVSI_ENV_CACHE(cfence, env_int("AVSI_L4_CFENCE", 1));           // 0: proxy fence in the 512 writers (round-1 form, A/B runs)
  AVSI_ENV_CACHE(bpf_env, env_int("AVSI_L4_BPF", 1));             // 0: per-thread L2 prefetch
  const int bulk = (bulk_env && B % 32 == 0) ? 1 : 0;
  const int bpf = (bpf_env && pf && B % 32 == 0) ? 1 : 0;
  const int smem = (int)sizeof(Lstm4Smem) + 128;
  const int grid = 2 * ((B + L4_BT - 1) / L4_BT) * L4_CL;
  auto launch = [&](auto kern) -> int {
    static const void* prepared[16];                               // kernels whose shared-memory limit is already raised
    static int n_prepared = 0;
    bool seen = false;
    for (int i = 0; i < n_prepared; ++i) seen |= (prepared[i] == (const void*)kern);
    if (!seen) {
      AVSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      if (n_prepared < 16) prepared[n_prepared++] = (const void*)kern;
    }
    kern<<<grid, L4_THREADS, smem, st>>>(gates, whh, bias, y, cst, T, B, pf, y_il);
    return AVSI_OK;
  };
  const int act = mode & 1;
  int rc;
  if (mode & 2) {                                                  // phase timers: default variant only
    rc = act ? launch(lstm4_fwd_kernel<1, true, false, true, false>) : launch(lstm4_fwd_kernel<0, true, false, true, false>);
  } else if (bulk) {
    rc = act ? launch(lstm4_fwd_kernel<1, false, true, false, false>) : launch(lstm4_fwd_kernel<0, false, true, false, false>);
  } else if (act) {
    rc = bpf ? launch(lstm4_fwd_kernel<1, false, false, true, true>) : launch(lstm4_fwd_kernel<1, false, false, true, false>);
  } else if (!cfence) {
    rc = launch(lstm4_fwd_kernel<0, false, false, false, false>);
  } else {
    rc = bpf ? launch(lstm4_fwd_kernel<0, false, false, true, true>) : launch(lstm4_fwd_kernel<0, false, false, true, false>);
  }
  if (rc != AVSI_OK) return rc;
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

}  // namespace avsi

// debug: cycles per phase summed over the steps of the last lstm4 forward launch (control thread [0,8), compute thread 0 [8,16))
extern "C" int avsi_debug_lstm4_timing(unsigned long long* out16) {
  return cudaMemcpyFromSymbol(out16, avsi::g_l4_timing, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -2;
}
