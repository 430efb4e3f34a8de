// BPTT of the bidirectional LSTM recurrence on tcgen05, cluster of 4 CTAs per (direction, 128-row
// batch tile).  Same contract as the mma.sync kernel in lstm.cu (avsi_lstm_bwd): the activated gates
// stashed by the forward kernel are overwritten in place with the pre-activation gradients dG.
//
// CTA j owns hidden units [64j, 64j+64) = 256 gate columns.  Per step (reverse chain order):
//   dh      = dy_t + sum of the 4 CTAs' partial  dG_{t+1} . W_hh  restricted to the own 64 units
//             (own partial straight from TMEM in fp32, 3 peers' partials as fp16 from shared-memory slots)
//   dG_t    = cell backward (c_t, dc carried in fp32)  -> the A operand tile in shared memory (K-major, K = own gate
//             columns) and, out of that tile, to global memory by TMA tensor stores of the control thread
//   partial = dG_t[128 x 256 own gate cols] . W_hh^T slice [256 gate cols x 256 h_in]   (tcgen05, TMEM)
//             issued as two K-halves so that the second half of the cell math overlaps the first MMA chain
//             (accumulator columns are ordered [pass][owner][warp][8 units])
//   reduce-scatter: every compute thread converts the peers' columns of its row to fp16 straight out of TMEM and ships
//             them with st.async (16 B per thread, owner and pass; complete_tx on the owner's slotfull barrier) -- no
//             shared-memory staging, no hand-over to the control thread.
// The bias gradient (column sums of dG over rows and time) costs nothing extra: it is a second, tiny MMA
// chain  dG^T(view of the same A tile, MN-major) . ones  accumulating in 32 spare TMEM columns over all steps.
//
// Per-row state that must survive a step (c_t and the running dc of the 64 own units) is parked in 128 spare
// TMEM columns (tcgen05.st / tcgen05.ld), which keeps 16 compute warps under 96 registers.
//
// Shared memory: W_hh^T slice 128 KB (resident for the whole sequence) + 32 KB A-half + 48 KB partial slots = 208 KB.
// Hand-shakes (all mbarriers, no cluster barrier in the loop):
//   slotfull[p] tx barriers, peers' partials of the previous step for the units of pass p have landed
//   consumed   3 remote arrives: every peer has READ its slots       -> my threads may write into them again
//   stagedA[h] 16 warps wrote K-half h of the A tile ; freeA / done0 / done : tcgen05.commit of the first chain / the
//              second chain / the second chain + its bias chain (freeA and done also count the TMA stores' read of the A-half)
//   slotread   16 warps have read their slots (the control thread then tells the peers: consumed)
#include <string.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avsi {

constexpr int B4_BT = 128;
constexpr int B4_CL = 4;
constexpr int B4_HP = 256;
constexpr int B4_G = 1024;
constexpr int B4_CWARPS = 16;
constexpr int B4_THREADS = (B4_CWARPS + 1) * 32;
constexpr uint32_t B4_W_BYTES = 256 * 256 * 2;
constexpr uint32_t B4_SLICE = B4_BT * 64 * 2;          // 16384: partial of one owner (64 units), fp16
constexpr uint32_t B4_AH_BYTES = 2 * B4_SLICE;          // A-half: [16 k-chunks][128 rows][16 B]
// TMEM columns: [0,256) partial dh accumulator (unit 64c + 16cg + 8p + i at column 128p + 32c + 8cg + i), [256,288) bias-gradient accumulators, [288,352) c_t, [352,416) dc
constexpr uint32_t B4_TM_BIAS = 256, B4_TM_C = 288, B4_TM_DC = 352;

struct Lstm4BwdSmem {
  unsigned char wt[B4_W_BYTES];      // [k-chunk position 0..31][h_in n 0..255][16 B]
  unsigned char ah[B4_AH_BYTES];     // A-half [16 k-chunks][128 rows][16 B]
  unsigned char slots[3 * B4_SLICE]; // [2 passes][3 sources][4 warps][128 rows][16 B]
  unsigned char ones[128];           // 8 x 8 halves of 1.0 (B operand of the bias MMA, strides 0)
  unsigned long long slotfull[2], consumed, stagedA[2], freeA, done0, done, slotread;
  uint32_t tmem_slot;
};

// packed fp32x2 arithmetic (sm_100: one issue slot for two lanes)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// phase timers (cycles summed over steps) of CTA 0: control thread [0,12), compute thread 0 [12,24); debug only
__device__ unsigned long long g_b4_timing[24];
#define B4_TICK(i)                                         \
  do {                                                     \
    if (TIMING && timing) {                                \
      const long long now_ = clock64();                    \
      tacc[i] += (unsigned long long)(now_ - tprev);       \
      tprev = now_;                                        \
    }                                                      \
  } while (0)

// CFENCE: the generic -> async proxy fence is executed by the consumer (the control thread, after its acquire of the
//   hand-over barrier) instead of by each of the 512 writers, whose fence (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC under nvcc
//   12.9) would wait for their loads in flight.
// BPF (B % 32 == 0): the next steps' gate / cell-state / dL/dy lines are pulled into L2 by cp.async.bulk.prefetch of the
//   control thread (contiguous 16 / 8 / 4 KB runs of the interleaved layouts) instead of 12 CCTL per compute thread.
// STMA (B % 128 == 0, needs CFENCE): dG leaves through the A-half it is staged in anyway -- four TMA tensor stores of the
//   control thread per half ([4 column chunks][4 row blocks][512 B] boxes of the interleaved layout) instead of four
//   STG.128 per compute thread; freeA / done then also wait for the stores to have READ the A-half.
template <bool TIMING, bool CFENCE, bool BPF, bool STMA>
__global__ void __cluster_dims__(B4_CL, 1, 1) __launch_bounds__(B4_THREADS, 1)
lstm4_bwd_kernel(const __grid_constant__ CUtensorMap tmG, uint16_t* __restrict__ gates, const uint16_t* __restrict__ whhT, const float* __restrict__ cst,
                 const uint16_t* __restrict__ dy, float* __restrict__ dbias, int dbias_tile_stride, int T, int B, int PFD) {
  extern __shared__ unsigned char b4_smem_raw[];
  const uint32_t raw_s = smem_u32(b4_smem_raw);
  const uint32_t base_s = (raw_s + 127u) & ~127u;
  Lstm4BwdSmem& sm = *reinterpret_cast<Lstm4BwdSmem*>(b4_smem_raw + (base_s - raw_s));

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int cid = blockIdx.x / B4_CL, j = blockIdx.x % B4_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * B4_BT;

  const uint32_t wt_s = smem_u32(&sm.wt[0]), ah_s = smem_u32(&sm.ah[0]), slots_s = smem_u32(&sm.slots[0]);
  const uint32_t ones_s = smem_u32(&sm.ones[0]);
  const uint32_t slotfull_s = smem_u32(&sm.slotfull[0]);
  const uint32_t consumed_s = smem_u32(&sm.consumed), stagedA_s = smem_u32(&sm.stagedA[0]);
  const uint32_t freeA_s = smem_u32(&sm.freeA), done0_s = smem_u32(&sm.done0), done_s = smem_u32(&sm.done);
  const uint32_t slotread_s = smem_u32(&sm.slotread);
  constexpr uint32_t B4_HALF = B4_SLICE / 2;            // 8192: one owner's partial for the units of one pass
  constexpr uint32_t PUSH_BYTES = 3 * B4_HALF;          // per pass

  if (tid == 0) {
    mbar_init(slotfull_s, 1);
    mbar_init(slotfull_s + 8, 1);
    mbar_init(consumed_s, 3);
    mbar_init(stagedA_s, B4_CWARPS);
    mbar_init(stagedA_s + 8, B4_CWARPS);
    mbar_init(freeA_s, STMA ? 2 : 1);
    mbar_init(done0_s, 1);
    mbar_init(done_s, STMA ? 2 : 1);
    mbar_init(slotread_s, B4_CWARPS);
    fence_barrier_init();
    if (T > 1) {
      mbar_expect_tx(slotfull_s, PUSH_BYTES);
      mbar_expect_tx(slotfull_s + 8, PUSH_BYTES);
    }
  }
  if (w == B4_CWARPS) tmem_alloc(smem_u32(&sm.tmem_slot), 512);
  // W_hh^T slice -> smem.  K position kpos = [half h][cg][i] <-> gate-column chunk 32j + 8cg + 4h + i
  // (the order in which the compute warps fill the two A-halves).
  for (int idx = tid; idx < 256 * 32; idx += B4_THREADS) {
    const int n = idx & 255, kpos = idx >> 8;
    const int h = kpos >> 4, cgk = (kpos >> 2) & 3, i = kpos & 3;
    const int gc = 32 * j + 8 * cgk + 4 * h + i;
    // accumulator column n = 128 pass + 32 owner + 8 warp + e  <->  hidden unit 64 owner + 16 warp + 8 pass + e
    const int unit = 64 * ((n >> 5) & 3) + 16 * ((n >> 3) & 3) + 8 * (n >> 7) + (n & 7);
    const uint4 v = *reinterpret_cast<const uint4*>(whhT + (long long)unit * (2 * B4_G) + dir * B4_G + gc * 8);
    *reinterpret_cast<uint4*>(&sm.wt[(uint32_t)kpos * 4096u + (uint32_t)n * 16u]) = v;
  }
  if (tid < 8) *reinterpret_cast<uint4*>(&sm.ones[tid * 16]) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&sm.tmem_slot);
  cluster_sync_all();
  const bool timing = TIMING && (blockIdx.x == 0) && (tid == 0 || tid == B4_CWARPS * 32);
  unsigned long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();

  if (w == B4_CWARPS) {
    // ===================================================================== control warp
    if (lane == 0) {
      const uint32_t idesc_main = make_idesc(B4_BT, 256, 0, 0);
      const uint32_t idesc_bias = make_idesc(128, 16, 1, 0);          // A = MN-major view of the A-half (dG^T)
      uint32_t peer_consumed[3];
#pragma unroll
      for (int cp = 0; cp < 3; ++cp) peer_consumed[cp] = map_to_cta(consumed_s, (uint32_t)(cp + (cp >= j ? 1 : 0)));
      // chain step sp: gates / dL/dy of its frame, cell state of the frame before it in chain order
      auto prefetch_step = [&](int sp) {
        const int tt = dir ? sp : (T - 1 - sp), tq = dir ? (tt + 1) : (tt - 1);
#pragma unroll
        for (int rb = 0; rb < 4; ++rb)
          if (b0 + 32 * rb < B) {
            const long long g = (long long)tt * B + b0 + 32 * rb;
            bulk_prefetch_l2(gates + il16(g, dir * B4_G + 256 * j, 2 * B4_G), 16384u);
            bulk_prefetch_l2(dy + il16(g, dir * B4_HP + 64 * j, 2 * B4_HP), 4096u);
            if (sp + 1 < T)
              bulk_prefetch_l2(cst + il32((long long)tq * B + b0 + 32 * rb, dir * B4_HP + 64 * j, 2 * B4_HP), 8192u);
          }
      };
      if (BPF)
        for (int sp = 1; sp < PFD && sp < T; ++sp) prefetch_step(sp);
      for (int s = 0; s < T; ++s) {
        if (BPF && s + PFD < T) prefetch_step(s + PFD);               // PFD = prefetch distance in steps
        if (s > 0 && s + 1 < T) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {                             // peers' partials of step s-1 are here: re-arm for step s
            mbar_wait(slotfull_s + 8u * ph, (uint32_t)((s - 1) & 1));
            mbar_expect_tx(slotfull_s + 8u * ph, PUSH_BYTES);
          }
        }
        B4_TICK(0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(stagedA_s + 8u * h, (uint32_t)(s & 1));
          if (CFENCE) fence_proxy_async();
          B4_TICK(1 + 2 * h);
          tc_fence_after();
          if (STMA) {                                                  // issued ahead of the chains: the reads overlap their issue
            const int tt = dir ? s : (T - 1 - s);
            const int rb0 = (int)(((long long)tt * B + b0) >> 5), cc0 = (dir * B4_G + 256 * j) / 8 + 4 * h;
#pragma unroll
            for (int cgk = 0; cgk < 4; ++cgk) tma_store_3d(&tmG, ah_s + (uint32_t)cgk * 8192u, 0, rb0, cc0 + 8 * cgk);
            bulk_commit();
          }
          if (h == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_f16(tmem_base, make_smem_desc(ah_s + (uint32_t)ks * 4096u, 2048u, 128u, 0u),
                       make_smem_desc(wt_s + (uint32_t)(2 * ks) * 4096u, 4096u, 128u, 0u), idesc_main, ks > 0 ? 1u : 0u);
          } else {
            // second K-half; its commit does not wait for the bias chain.  (Issuing it as two N = 128 halves with a commit
            // after the first -- the columns the receivers need first -- was measured 4 % slower: profiles/README.md)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_f16(tmem_base, make_smem_desc(ah_s + (uint32_t)ks * 4096u, 2048u, 128u, 0u),
                       make_smem_desc(wt_s + (uint32_t)(16 + 2 * ks) * 4096u, 4096u, 128u, 0u), idesc_main, 1u);
            umma_commit(done0_s);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_f16(tmem_base + B4_TM_BIAS + 16u * h, make_smem_desc(ah_s + (uint32_t)ks * 256u, 128u, 2048u, 0u),
                     make_smem_desc(ones_s, 0u, 0u, 0u), idesc_bias, (s > 0 || ks > 0) ? 1u : 0u);
          umma_commit(h == 0 ? freeA_s : done_s);                      // both chains read the A-half
          if (STMA) {
            bulk_wait_read0();                                         // the A-half may be overwritten once the chain retired
            mbar_arrive_local(h == 0 ? freeA_s : done_s);
          }
          B4_TICK(2 + 2 * h);
          if (h == 0 && s > 0) {
            // my warps read their slots at the start of the second half: tell the senders as early as possible
            mbar_wait(slotread_s, (uint32_t)((s - 1) & 1));
#pragma unroll
            for (int cp = 0; cp < 3; ++cp) mbar_arrive_remote_relaxed(peer_consumed[cp]);
            B4_TICK(5);
          }
        }
      }
    }
    if (timing) {
#pragma unroll
      for (int i = 0; i < 12; ++i) g_b4_timing[i] = tacc[i];
    }
    if (STMA && lane == 0) bulk_wait0();
    __syncwarp();
  } else {
    // ===================================================================== compute warps
    const int q = w & 3, cg = w >> 2;               // TMEM lane quarter / unit group (16 contiguous units)
    const int r = q * 32 + lane;
    const int row = b0 + r;
    const bool row_ok = row < B;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t peer_slot[3], peer_full[3];            // this thread's 16 bytes in my slot at owner c, and c's slotfull barriers
#pragma unroll
    for (int cp = 0; cp < 3; ++cp) {
      const int c = cp + (cp >= j ? 1 : 0);
      const int src = (j < c) ? j : j - 1;
      peer_slot[cp] = map_to_cta(slots_s + (uint32_t)src * B4_HALF + (uint32_t)cg * 2048u + (uint32_t)r * 16u, (uint32_t)c);
      peer_full[cp] = map_to_cta(slotfull_s, (uint32_t)c);
    }

    for (int s = 0; s < T; ++s) {
      const int t = dir ? s : (T - 1 - s);          // reverse of the forward chain order
      const int tp = dir ? (t + 1) : (t - 1);       // chain predecessor
      const bool has_prev = (s + 1 < T);
      const long long grow = (long long)t * B + row, gprev = (long long)tp * B + row;
      auto load_inputs = [&](int h, uint4 (&G4)[4], float (&CP)[8], uint4& DY) {
        const int ug0 = 64 * j + 16 * cg + 8 * h;
#pragma unroll
        for (int i = 0; i < 4; ++i) G4[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) CP[i] = 0.f;
        DY = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            G4[i] = *reinterpret_cast<const uint4*>(gates + il16(grow, dir * B4_G + ug0 * 4 + 8 * i, 2 * B4_G));
          DY = *reinterpret_cast<const uint4*>(dy + il16(grow, dir * B4_HP + ug0, 2 * B4_HP));   // interleaved dL/dy
          if (has_prev) {
            const float4 a = *reinterpret_cast<const float4*>(cst + il32(gprev, dir * B4_HP + ug0, 2 * B4_HP));
            const float4 b = *reinterpret_cast<const float4*>(cst + il32(gprev, dir * B4_HP + ug0 + 4, 2 * B4_HP));
            CP[0] = a.x; CP[1] = a.y; CP[2] = a.z; CP[3] = a.w;
            CP[4] = b.x; CP[5] = b.y; CP[6] = b.z; CP[7] = b.w;
          }
        }
      };
      // the inputs of the first half do not depend on the peers: request them before waiting; the lines of the
      // NEXT step are pulled into L2 now so that its loads are L2 hits
      uint4 g4[4], dyv;
      float cprev[8];
      load_inputs(0, g4, cprev, dyv);
      if (!BPF && row_ok && has_prev && (lane & 7) == 0) {     // one lane per 128-byte line of the interleaved runs
        const int tn = dir ? (t + 1) : (t - 1), tnp = dir ? (t + 2) : (t - 2);
        const long long gn = (long long)tn * B + row, gnp = (long long)tnp * B + row;
        const int ug = 64 * j + 16 * cg;
#pragma unroll
        for (int i = 0; i < 8; ++i) prefetch_l2(gates + il16(gn, dir * B4_G + ug * 4 + 8 * i, 2 * B4_G));
        if (s + 2 < T) {
#pragma unroll
          for (int i = 0; i < 4; ++i) prefetch_l2(cst + il32(gnp, dir * B4_HP + ug + 4 * i, 2 * B4_HP));
        }
      }
      B4_TICK(9);
      uint32_t own1[8];                             // own partial of the second half's units (read before the MMA chain overwrites it)
      if (s > 0) {
        mbar_wait(slotfull_s, (uint32_t)((s - 1) & 1));
        tc_fence_after();
        tmem_ld8(trow + (uint32_t)(128 + 32 * j + 8 * cg), own1);
        tmem_ld_wait();
      }
      B4_TICK(0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ul0 = 16 * cg + 8 * h;            // first local unit of the pass
        const int ug0 = 64 * j + ul0;
        // ---- inputs ------------------------------------------------------------------------------------------
        float c_cur[8], dc_st[8], dhrec[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c_cur[i] = dc_st[i] = dhrec[i] = 0.f;
        if (row_ok) {
          if (s == 0) {
            const float4 a = *reinterpret_cast<const float4*>(cst + il32(grow, dir * B4_HP + ug0, 2 * B4_HP));
            const float4 b = *reinterpret_cast<const float4*>(cst + il32(grow, dir * B4_HP + ug0 + 4, 2 * B4_HP));
            c_cur[0] = a.x; c_cur[1] = a.y; c_cur[2] = a.z; c_cur[3] = a.w;
            c_cur[4] = b.x; c_cur[5] = b.y; c_cur[6] = b.z; c_cur[7] = b.w;
          }
        }
        if (s > 0) {
          uint32_t cc[8], dd[8], own[8];
          tmem_ld8(trow + B4_TM_C + (uint32_t)ul0, cc);
          tmem_ld8(trow + B4_TM_DC + (uint32_t)ul0, dd);
          if (h == 0) tmem_ld8(trow + (uint32_t)(32 * j + 8 * cg), own);
          tmem_ld_wait();
          if (h == 1) mbar_wait(slotfull_s + 8u, (uint32_t)((s - 1) & 1));      // second half of the peers' partials
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            c_cur[i] = __uint_as_float(cc[i]);
            dc_st[i] = __uint_as_float(dd[i]);
            dhrec[i] = __uint_as_float(h == 0 ? own[i] : own1[i]);
          }
#pragma unroll
          for (int src = 0; src < 3; ++src) {
            const uint4 v = *reinterpret_cast<const uint4*>(&sm.slots[(uint32_t)h * PUSH_BYTES + (uint32_t)src * B4_HALF +
                                                                      (uint32_t)cg * 2048u + (uint32_t)r * 16u]);
            const float2 a = unpack_half2(v.x), b = unpack_half2(v.y), c = unpack_half2(v.z), d = unpack_half2(v.w);
            dhrec[0] += a.x; dhrec[1] += a.y; dhrec[2] += b.x; dhrec[3] += b.y;
            dhrec[4] += c.x; dhrec[5] += c.y; dhrec[6] += d.x; dhrec[7] += d.y;
          }
          if (h == 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive_local(slotread_s);
          }
        }
        // ---- cell backward -------------------------------------------------------------------------------------
        // two units at a time in packed fp32x2 (FADD2 / FMUL2 / FFMA2: half the issue slots of the scalar form -- the cell
        // arithmetic is 40 % of this kernel's issue slots); unit pair k = units 2k, 2k+1 = the four words of g4[k]
        uint4 pk[4];
        uint32_t cnew[8], dnew[8];
        const uint32_t dyw[4] = {dyv.x, dyv.y, dyv.z, dyv.w};
        const f2 one = pk2(1.f, 1.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint4 gv = g4[k];
          const float2 ig0 = unpack_half2(gv.x), fo0 = unpack_half2(gv.y), ig1 = unpack_half2(gv.z), fo1 = unpack_half2(gv.w);
          const f2 gi = pk2(ig0.x, ig1.x), gg = pk2(ig0.y, ig1.y), gf = pk2(fo0.x, fo1.x), go = pk2(fo0.y, fo1.y);
          const float2 dy2 = unpack_half2(dyw[k]);
          const f2 dh = add2(pk2(dy2.x, dy2.y), pk2(dhrec[2 * k], dhrec[2 * k + 1]));
          const f2 tc = pk2(tanhf_fast(c_cur[2 * k]), tanhf_fast(c_cur[2 * k + 1]));
          const f2 cp = pk2(cprev[2 * k], cprev[2 * k + 1]);
          const f2 d_o = mul2(mul2(dh, tc), mul2(go, sub2(one, go)));
          const f2 dc = fma2(mul2(dh, go), sub2(one, mul2(tc, tc)), pk2(dc_st[2 * k], dc_st[2 * k + 1]));
          const f2 d_i = mul2(mul2(dc, gg), mul2(gi, sub2(one, gi)));
          const f2 d_g = mul2(mul2(dc, gi), sub2(one, mul2(gg, gg)));
          const f2 d_f = mul2(mul2(dc, cp), mul2(gf, sub2(one, gf)));
          const f2 dn = mul2(dc, gf);
          float a0, a1, b0, b1, c0, c1, e0, e1, n0, n1;
          upk2(d_i, a0, a1);
          upk2(d_g, b0, b1);
          upk2(d_f, c0, c1);
          upk2(d_o, e0, e1);
          upk2(dn, n0, n1);
          dnew[2 * k] = __float_as_uint(n0);
          dnew[2 * k + 1] = __float_as_uint(n1);
          cnew[2 * k] = __float_as_uint(cprev[2 * k]);
          cnew[2 * k + 1] = __float_as_uint(cprev[2 * k + 1]);
          pk[k] = make_uint4(pack_half2(a0, b0), pack_half2(c0, e0), pack_half2(a1, b1), pack_half2(c1, e1));
        }
        if (h == 0) load_inputs(1, g4, cprev, dyv);   // second half's inputs: in flight during the A-half hand-over
        if (has_prev) {                               // park c_{t_prev} and dc for the next step
          tmem_st8(trow + B4_TM_C + (uint32_t)ul0, cnew);
          tmem_st8(trow + B4_TM_DC + (uint32_t)ul0, dnew);
        }
        if (!STMA && row_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(gates + il16(grow, dir * B4_G + ug0 * 4 + 8 * i, 2 * B4_G)) = pk[i];
        }
        B4_TICK(1 + 2 * h);
        // A-half: the chain that read it (and the TMA store out of it) must have retired -- the previous step's second
        // chain was waited for at the end of that step
        if (h == 1) mbar_wait(freeA_s, (uint32_t)(s & 1));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(&sm.ah[(uint32_t)(cg * 4 + i) * 2048u + (uint32_t)r * 16u]) = pk[i];
        if (has_prev) tmem_st_wait();
        if (!CFENCE) fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(stagedA_s + 8u * h);
        B4_TICK(2 + 2 * h);
      }
      // ---- the partial of this step: the other owners' columns of this thread's row -> fp16 -> st.async into their slots ----
      if (has_prev) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          mbar_wait(ph == 0 ? done0_s : done_s, (uint32_t)(s & 1));
          tc_fence_after();
          if (ph == 0) B4_TICK(7);
          uint32_t acc[3][8];                         // the three owners' columns in flight together
#pragma unroll
          for (int cp = 0; cp < 3; ++cp) {
            const int c = cp + (cp >= j ? 1 : 0);
            tmem_ld8(trow + (uint32_t)(128 * ph + 32 * c + 8 * cg), acc[cp]);
          }
          tmem_ld_wait();
          if (ph == 0 && s > 0) mbar_wait(consumed_s, (uint32_t)((s - 1) & 1));     // the peers have read their slots of the last step
#pragma unroll
          for (int cp = 0; cp < 3; ++cp)
            st_async_v4(peer_slot[cp] + (uint32_t)ph * PUSH_BYTES,
                        pack_half2(__uint_as_float(acc[cp][0]), __uint_as_float(acc[cp][1])),
                        pack_half2(__uint_as_float(acc[cp][2]), __uint_as_float(acc[cp][3])),
                        pack_half2(__uint_as_float(acc[cp][4]), __uint_as_float(acc[cp][5])),
                        pack_half2(__uint_as_float(acc[cp][6]), __uint_as_float(acc[cp][7])), peer_full[cp] + 8u * ph);
        }
        tc_fence_before();                            // the loads above precede the next step's first chain (accumulate = 0)
        B4_TICK(8);
      } else {
        mbar_wait(done_s, (uint32_t)(s & 1));         // last step: all chains retired before the bias columns are read
        tc_fence_after();
      }
    }
    if (timing) {
#pragma unroll
      for (int i = 0; i < 12; ++i) g_b4_timing[12 + i] = tacc[i];
    }
    // ---- bias gradient: TMEM lane m of half h = sum over rows and steps of local gate column m ----------------------
    if (cg == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[8];
        tmem_ld8(trow + B4_TM_BIAS + 16u * h, v);
        tmem_ld_wait();
        const int m = q * 32 + lane, kc = m >> 3, e = m & 7;
        const int cgk = kc >> 2, i = kc & 3;
        const int col = (64 * j + 16 * cgk + 8 * h) * 4 + 8 * i + e;
        // per-tile row of the scratch matrix (ordered reduction afterwards) or atomics straight into dbias
        if (dbias_tile_stride > 0) dbias[(long long)(cid >> 1) * dbias_tile_stride + dir * B4_G + col] = __uint_as_float(v[0]);
        else atomicAdd(dbias + dir * B4_G + col, __uint_as_float(v[0]));
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (w == B4_CWARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int get_tmap_il(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t brb, uint32_t bcc, CUtensorMap* out);

int launch_lstm4_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy, float* dbias,
                     int dbias_tile_stride, int T, int B, cudaStream_t st) {
  AVSI_ENV_CACHE(timing, env_is("AVSI_B4_TIMING", "1"));   // in-kernel phase timers (profiles/bench_lstm.py)
  const int smem = (int)sizeof(Lstm4BwdSmem) + 128;
  AVSI_ENV_CACHE(cfence, env_int("AVSI_B4_CFENCE", 1));    // 0: proxy fence in the 512 writers (round-1 form, A/B runs)
  AVSI_ENV_CACHE(bpf_env, env_int("AVSI_B4_BPF", 1));      // 0: per-thread L2 prefetch
  AVSI_ENV_CACHE(pfd, env_int("AVSI_B4_PFD", 1));          // its distance in steps
  const int bpf = (bpf_env && pfd > 0 && B % 32 == 0) ? 1 : 0;
  const int grid = 2 * ((B + B4_BT - 1) / B4_BT) * B4_CL;
  AVSI_ENV_CACHE(stma_env, env_int("AVSI_B4_STMA", 1));    // 0: dG through per-thread STG.128
  const int stma = (stma_env && cfence && B % 128 == 0) ? 1 : 0;
  CUtensorMap tmG;
  memset(&tmG, 0, sizeof(tmG));
  if (stma) {
    const int rc_map = get_tmap_il(gates, (uint64_t)T * B, 2 * B4_G, 2 * B4_G, 4, 4, &tmG);
    if (rc_map != AVSI_OK) return rc_map;
  }
  auto launch = [&](auto kern) -> int {
    static const void* prepared[8];                        // kernels whose shared-memory limit is already raised
    static int n_prepared = 0;
    bool seen = false;
    for (int i = 0; i < n_prepared; ++i) seen |= (prepared[i] == (const void*)kern);
    if (!seen) {
      AVSI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      if (n_prepared < 8) prepared[n_prepared++] = (const void*)kern;
    }
    kern<<<grid, B4_THREADS, smem, st>>>(tmG, gates, whhT, cst, dy, dbias, dbias_tile_stride, T, B, pfd);
    return AVSI_OK;
  };
  int rc;
  if (timing) rc = launch(lstm4_bwd_kernel<true, true, false, false>);
  else if (!cfence) rc = launch(lstm4_bwd_kernel<false, false, false, false>);
  else if (stma) rc = bpf ? launch(lstm4_bwd_kernel<false, true, true, true>) : launch(lstm4_bwd_kernel<false, true, false, true>);
  else rc = bpf ? launch(lstm4_bwd_kernel<false, true, true, false>) : launch(lstm4_bwd_kernel<false, true, false, false>);
  if (rc != AVSI_OK) return rc;
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

}  // namespace avsi

// debug: cycles per phase summed over the steps of the last lstm4 backward launch (control [0,12), compute thread 0 [12,24))
extern "C" int avsi_debug_lstm4_bwd_timing(unsigned long long* out24) {
  return cudaMemcpyFromSymbol(out24, avsi::g_b4_timing, sizeof(unsigned long long) * 24) == cudaSuccess ? 0 : -2;
}
