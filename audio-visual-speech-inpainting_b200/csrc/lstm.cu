// Persistent bidirectional LSTM recurrence, forward and BPTT.
//
// Replaces cudnnRNNForwardTraining / cudnnRNNBackwardData behind
// tf.contrib.cudnn_rnn.CudnnLSTM (models.py:95-104) and the CudnnCompatibleLSTMCell
// while-loop (models.py:106-115).  Cell (TF LSTMBlockCell, forget_bias 0):
//   z = x.W_ih + h.W_hh + b ; c = sig(f) c + sig(i) tanh(g) ; h = sig(o) tanh(c)
//
// One launch runs both directions of a layer over all T steps.  A thread-block cluster of
// 8 CTAs owns one (direction, batch tile of BT rows); CTA j owns hidden units [32j, 32j+32)
// with all four gates, so the cell update is CTA-local.  Its 256 x 128 slice of W_hh lives in
// REGISTERS as mma.sync B-fragments for the whole sequence (64 regs/thread); the recurrent
// product runs on mma.sync.m16n8k16 (fp16 in, fp32 accumulate); c_t stays in fp32 registers.
//
// Per-step exchange (the critical path of a 250..1667-step dependent chain) goes through
// DISTRIBUTED SHARED MEMORY: every thread pushes its packed h_t values straight into the
// double-buffered h tile of all 8 CTAs with st.async ... mbarrier::complete_tx::bytes, and
// every CTA waits on its own mbarrier for the BT x 512 bytes of the step.  No __syncthreads,
// no cluster barrier, no fence and no L2 round trip inside the time loop (ncu of the first
// version -- exchange through L2 + barrier.cluster -- showed 83 % of issue slots idle on
// barrier / membar / long-scoreboard stalls, profiles/r01a_lstm_fwd_kernel_summary.txt).
// x.W_ih pre-activations of the step are prefetched before the wait.  Gate columns are laid
// out [dir][unit][i,g,f,o]: a thread's four gates are one 8-byte load and the activated gates
// overwrite the pre-activations in place (stash for BPTT); BPTT overwrites them with dgates.
// The gate tensor and the cell-state stash are stored INTERLEAVED (common.cuh il16 / il32).
//
// BPTT: CTA j multiplies ITS 128 dgate columns with its W_hh^T slice (K-split, registers);
// warp w's partial dh[BT, 32] belongs to CTA w and is pushed there as fp16 through the same
// st.async / mbarrier mechanism (reduce-scatter in DSMEM), the owner sums the 8 slots.
// Bias gradients accumulate in registers over t.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avsi {

constexpr int LS_HP = 256;        // padded hidden size
constexpr int LS_G = 1024;        // gate columns per direction
constexpr int LS_CL = 8;          // CTAs per cluster
constexpr int LS_THREADS = 256;
constexpr int LS_HSTRIDE = LS_HP + 8;     // halves per smem row of h (528 B, conflict-free ldmatrix)
constexpr int LS_DSTRIDE = 128 + 8;       // halves per smem row of dgates

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// fast, accurate-enough activations: ex2.approx + rcp.approx (2 MUFU each), ~2 ulp
__device__ __forceinline__ float fsigmoid(float x) { return sigmoid_fast2(x); }
__device__ __forceinline__ float ftanh(float x) { return tanh_fast2(x); }

template <int BT>
struct LstmFwdSmem {
  uint16_t hbuf[2][BT * LS_HSTRIDE];
  unsigned long long mbar[2];
};

template <int BT>
__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
lstm_fwd_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whh, const float* __restrict__ bias,
                uint16_t* __restrict__ y, float* __restrict__ cst, int T, int B, int y_il) {
  constexpr int MT = BT / 16;
  constexpr uint32_t PHASE_BYTES = BT * LS_HP * 2;
  extern __shared__ __align__(16) unsigned char ls_smem[];
  LstmFwdSmem<BT>& sm = *reinterpret_cast<LstmFwdSmem<BT>*>(ls_smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, uu = lane & 3;
  const int cid = blockIdx.x / LS_CL, j = blockIdx.x % LS_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int u_base = j * 32 + w * 4;
  const int u = u_base + uu;                       // hidden unit whose cell this thread updates

  const uint32_t hbuf_s = (uint32_t)__cvta_generic_to_shared(&sm.hbuf[0][0]);
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&sm.mbar[0]);
  if (tid == 0) {
    mbar_init(mbar_s, 1);
    mbar_init(mbar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (T > 1) mbar_expect_tx(mbar_s, PHASE_BYTES);          // h_0 lands in buffer 0
    if (T > 2) mbar_expect_tx(mbar_s + 8, PHASE_BYTES);      // h_1 lands in buffer 1
  }

  // W_hh fragments: mma column c of n-tile nt <-> unit (c/2), gate (c%2) + 2 nt
  uint32_t wf[16][2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = (u_base + (g >> 1)) * 4 + (g & 1) + 2 * nt;
    const uint16_t* wrow = whh + ((long long)(dir * LS_G + col)) * LS_HP;
#pragma unroll
    for (int kt = 0; kt < 16; ++kt) {
      wf[kt][nt][0] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16 + uu * 2);
      wf[kt][nt][1] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16 + uu * 2 + 8);
    }
  }

  float bq[4];                                     // prescaled bias (i, f, o halved like the pre-activations)
#pragma unroll
  for (int q = 0; q < 4; ++q) bq[q] = bias[dir * LS_G + u * 4 + q];

  float c_state[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) c_state[mt][0] = c_state[mt][1] = 0.f;

  // destinations of this lane's pushes: CTAs 2*uu and 2*uu+1 (h tile and mbarrier base there)
  const uint32_t dst_h0 = map_to_cta(hbuf_s, 2 * uu), dst_h1 = map_to_cta(hbuf_s, 2 * uu + 1);
  const uint32_t dst_m0 = map_to_cta(mbar_s, 2 * uu), dst_m1 = map_to_cta(mbar_s, 2 * uu + 1);

  cluster_sync_all();                              // every CTA's mbarriers are initialised and armed

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    // ---- prefetch x.W_ih pre-activations of this step (independent of h) ----------------
    uint2 pre[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int row = b0 + mt * 16 + g + rh * 8;
        pre[mt][rh] = make_uint2(0u, 0u);
        if (row < B)
          pre[mt][rh] = *reinterpret_cast<const uint2*>(gates + il16((long long)t * B + row, dir * LS_G + u * 4, 2 * LS_G));
      }
    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    if (s > 0) {
      const int pb = (s - 1) & 1;
      mbar_wait(mbar_s + 8 * pb, ((s - 1) >> 1) & 1);          // all of h_{s-1} has landed here
      if (tid == 0 && s + 1 < T - 1) mbar_expect_tx(mbar_s + 8 * pb, PHASE_BYTES);   // re-arm for h_{s+1}
      const uint32_t hb = hbuf_s + (uint32_t)(pb * BT * LS_HSTRIDE) * 2u;
#pragma unroll
      for (int kt = 0; kt < 16; ++kt) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(hb + (uint32_t)((mt * 16 + (lane & 15)) * LS_HSTRIDE + kt * 16 + (lane >> 4) * 8) * 2u,
                      a0, a1, a2, a3);
          mma16816(acc[mt][0], a0, a1, a2, a3, wf[kt][0][0], wf[kt][0][1]);
          mma16816(acc[mt][1], a0, a1, a2, a3, wf[kt][1][0], wf[kt][1][1]);
        }
      }
    }
    // ---- gates, cell update, stores, push ----------------------------------------------------
    const bool push = (s + 1 < T);
    const uint32_t boff = (uint32_t)((s & 1) * BT * LS_HSTRIDE) * 2u;
    const uint32_t moff = 8u * (s & 1);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int rl = mt * 16 + g + rh * 8;
        const int row = b0 + rl;
        const float2 p_ig = unpack_half2(pre[mt][rh].x), p_fo = unpack_half2(pre[mt][rh].y);
        // the i, f, o columns of the pre-activations, the bias and the W_hh rows arrive pre-halved
        const float gi = fsigmoid(2.f * (acc[mt][0][rh * 2 + 0] + p_ig.x + bq[0]));
        const float gg = ftanh(acc[mt][0][rh * 2 + 1] + p_ig.y + bq[1]);
        const float gf = fsigmoid(2.f * (acc[mt][1][rh * 2 + 0] + p_fo.x + bq[2]));
        const float go = fsigmoid(2.f * (acc[mt][1][rh * 2 + 1] + p_fo.y + bq[3]));
        const float c = gf * c_state[mt][rh] + gi * gg;
        c_state[mt][rh] = c;
        const float h = go * ftanh(c);
        // quad all-gather of the 4 units of this row: lo = units (0,1), hi = units (2,3)
        const float hx = __shfl_xor_sync(0xffffffffu, h, 1);
        const uint32_t pr = (uu & 1) ? pack_half2(hx, h) : pack_half2(h, hx);
        const uint32_t po = __shfl_xor_sync(0xffffffffu, pr, 2);
        const uint32_t lo = (uu & 2) ? po : pr, hi = (uu & 2) ? pr : po;
        if (push) {
          const uint32_t off = boff + (uint32_t)(rl * LS_HSTRIDE + u_base) * 2u;
          st_async_v2(dst_h0 + off, lo, hi, dst_m0 + moff);
          st_async_v2(dst_h1 + off, lo, hi, dst_m1 + moff);
        }
        if (row < B) {
          const long long r = (long long)t * B + row;
          *reinterpret_cast<uint2*>(gates + il16(r, dir * LS_G + u * 4, 2 * LS_G)) =
              make_uint2(pack_half2(gi, gg), pack_half2(gf, go));
          cst[il32(r, dir * LS_HP + u, 2 * LS_HP)] = c;
          if (uu == 0)
            *reinterpret_cast<uint2*>(y + (y_il ? il16(r, dir * LS_HP + u_base, 2 * LS_HP) : r * (2 * LS_HP) + dir * LS_HP + u_base)) =
                make_uint2(lo, hi);
        }
      }
  }
  cluster_sync_all();                              // no CTA exits while peers may still address its smem
}

template <int BT>
struct LstmBwdSmem {
  uint16_t dgbuf[2][BT * LS_DSTRIDE];              // this CTA's dgates (MMA A operand), double buffered
  uint16_t pbuf[2][LS_CL][BT * 32];                // partial dh slots [parity][source CTA][row][unit]
  unsigned long long mbar[2];
};

template <int BT>
__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
lstm_bwd_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whhT, const float* __restrict__ cst,
                const uint16_t* __restrict__ dy, float* __restrict__ dbias, int dbias_tile_stride, int T, int B) {
  constexpr int MT = BT / 16;
  constexpr uint32_t PHASE_BYTES = LS_CL * BT * 32 * 2;
  extern __shared__ __align__(16) unsigned char ls_smem[];
  LstmBwdSmem<BT>& sm = *reinterpret_cast<LstmBwdSmem<BT>*>(ls_smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, uu = lane & 3;
  const int cid = blockIdx.x / LS_CL, j = blockIdx.x % LS_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int ul = w * 4 + uu;                        // unit local to this CTA (0..31)
  const int u = j * 32 + ul;

  const uint32_t dg_s = (uint32_t)__cvta_generic_to_shared(&sm.dgbuf[0][0]);
  const uint32_t pb_s = (uint32_t)__cvta_generic_to_shared(&sm.pbuf[0][0][0]);
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&sm.mbar[0]);
  if (tid == 0) {
    mbar_init(mbar_s, 1);
    mbar_init(mbar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (T > 1) mbar_expect_tx(mbar_s, PHASE_BYTES);
    if (T > 2) mbar_expect_tx(mbar_s + 8, PHASE_BYTES);
  }

  // W_hh^T fragments for partial dh[BT, 256] = dgates[BT, own 128 cols] . W_hh[own cols, 256]:
  // n = h_in index w*32 + nt*8 + g ; k = local gate column kt*16 + uu*2 (+8)
  uint32_t wf[8][4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const uint16_t* wrow = whhT + ((long long)(w * 32 + nt * 8 + g)) * (2 * LS_G) + dir * LS_G + j * 128;
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      wf[kt][nt][0] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16 + uu * 2);
      wf[kt][nt][1] = *reinterpret_cast<const uint32_t*>(wrow + kt * 16 + uu * 2 + 8);
    }
  }
  float dc_state[MT][2], c_cur[MT][2], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) dc_state[mt][0] = dc_state[mt][1] = c_cur[mt][0] = c_cur[mt][1] = 0.f;

  // warp w's partial columns [32w, 32w+32) belong to CTA w: slot [source = j] there
  const uint32_t dst_p = map_to_cta(pb_s, (uint32_t)w) + (uint32_t)(j * BT * 32) * 2u;
  const uint32_t dst_m = map_to_cta(mbar_s, (uint32_t)w);

  cluster_sync_all();

  for (int s = 0; s < T; ++s) {
    const int t = dir ? s : (T - 1 - s);            // reverse of the forward chain order
    const int tp = dir ? (t + 1) : (t - 1);         // chain predecessor (forward-time h_{prev})
    const bool has_prev = (s + 1 < T);
    uint2 gt[MT][2];
    float dyv[MT][2], cprev[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int row = b0 + mt * 16 + g + rh * 8;
        gt[mt][rh] = make_uint2(0u, 0u);
        dyv[mt][rh] = 0.f;
        cprev[mt][rh] = 0.f;
        if (row < B) {
          const long long r = (long long)t * B + row;
          gt[mt][rh] = *reinterpret_cast<const uint2*>(gates + il16(r, dir * LS_G + u * 4, 2 * LS_G));
          dyv[mt][rh] = __half2float(__ushort_as_half(dy[il16(r, dir * LS_HP + u, 2 * LS_HP)]));   // interleaved dL/dy
          if (has_prev) cprev[mt][rh] = cst[il32((long long)tp * B + row, dir * LS_HP + u, 2 * LS_HP)];
          if (s == 0) c_cur[mt][rh] = cst[il32(r, dir * LS_HP + u, 2 * LS_HP)];
        }
      }
    float dhr[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dhr[mt][0] = dhr[mt][1] = 0.f;
    if (s > 0) {
      const int pb = (s - 1) & 1;
      mbar_wait(mbar_s + 8 * pb, ((s - 1) >> 1) & 1);          // the 8 partial slots of step s-1 are here
      if (tid == 0 && s + 1 < T - 1) mbar_expect_tx(mbar_s + 8 * pb, PHASE_BYTES);
      const uint16_t* slots = &sm.pbuf[pb][0][0];
#pragma unroll
      for (int src = 0; src < LS_CL; ++src)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh)
            dhr[mt][rh] += __half2float(__ushort_as_half(slots[(src * BT + mt * 16 + g + rh * 8) * 32 + ul]));
    }
    uint16_t* dgb = &sm.dgbuf[s & 1][0];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int rl = mt * 16 + g + rh * 8;
        const float2 ig = unpack_half2(gt[mt][rh].x), fo = unpack_half2(gt[mt][rh].y);
        const float gi = ig.x, gg = ig.y, gf = fo.x, go = fo.y;
        const float dh = dyv[mt][rh] + dhr[mt][rh];
        const float tc = ftanh(c_cur[mt][rh]);
        const float d_o = dh * tc * go * (1.f - go);
        const float dc = dc_state[mt][rh] + dh * go * (1.f - tc * tc);
        const float d_i = dc * gg * gi * (1.f - gi);
        const float d_g = dc * gi * (1.f - gg * gg);
        const float d_f = dc * cprev[mt][rh] * gf * (1.f - gf);
        dc_state[mt][rh] = dc * gf;
        c_cur[mt][rh] = cprev[mt][rh];
        const uint2 pk = make_uint2(pack_half2(d_i, d_g), pack_half2(d_f, d_o));
        *reinterpret_cast<uint2*>(dgb + rl * LS_DSTRIDE + ul * 4) = pk;
        if (b0 + rl < B) {
          *reinterpret_cast<uint2*>(gates + il16((long long)t * B + b0 + rl, dir * LS_G + u * 4, 2 * LS_G)) = pk;
          db[0] += d_i;
          db[1] += d_g;
          db[2] += d_f;
          db[3] += d_o;
        }
      }
    if (has_prev) {
      __syncthreads();                              // the CTA's dgates tile is complete (double buffered)
      float acc[MT][4][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
      const uint32_t db_s = dg_s + (uint32_t)((s & 1) * BT * LS_DSTRIDE) * 2u;
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(db_s + (uint32_t)((mt * 16 + (lane & 15)) * LS_DSTRIDE + kt * 16 + (lane >> 4) * 8) * 2u,
                      a0, a1, a2, a3);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816(acc[mt][nt], a0, a1, a2, a3, wf[kt][nt][0], wf[kt][nt][1]);
        }
      }
      // push warp w's partial [BT, 32] (fp16) to its owner CTA w: 4x4 quad transpose so that lane uu
      // sends the 16-byte chunk of n-tile uu
      const uint32_t poff = (uint32_t)((s & 1) * LS_CL * BT * 32) * 2u;
      const uint32_t moff = 8u * (s & 1);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          uint32_t x[4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) x[nt] = pack_half2(acc[mt][nt][rh * 2], acc[mt][nt][rh * 2 + 1]);
          const bool o1 = (uu & 1), o2 = (uu & 2);
          const uint32_t r0 = __shfl_xor_sync(0xffffffffu, o1 ? x[0] : x[1], 1);
          const uint32_t r1 = __shfl_xor_sync(0xffffffffu, o1 ? x[2] : x[3], 1);
          const uint32_t k0 = o1 ? x[1] : x[0], k1 = o1 ? x[3] : x[2];
          const uint32_t a0lo = o1 ? r0 : k0, a0hi = o1 ? k0 : r0;     // tile (uu&1), lanes (2m, 2m+1)
          const uint32_t a1lo = o1 ? r1 : k1, a1hi = o1 ? k1 : r1;     // tile (uu&1)+2
          const uint32_t rl_ = __shfl_xor_sync(0xffffffffu, o2 ? a0lo : a1lo, 2);
          const uint32_t rh_ = __shfl_xor_sync(0xffffffffu, o2 ? a0hi : a1hi, 2);
          const uint32_t klo = o2 ? a1lo : a0lo, khi = o2 ? a1hi : a0hi;
          const uint32_t y0 = o2 ? rl_ : klo, y1 = o2 ? rh_ : khi, y2 = o2 ? klo : rl_, y3 = o2 ? khi : rh_;
          const int rl = mt * 16 + g + rh * 8;
          st_async_v4(dst_p + poff + (uint32_t)(rl * 32 + uu * 8) * 2u, y0, y1, y2, y3, dst_m + moff);
        }
    }
  }
  // bias gradient: sum over the 8 row-lanes (g) that share this unit, one atomic per (unit, gate, CTA)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v = db[q];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    // dbias_tile_stride > 0: this row tile's sums go to their own row of a scratch matrix (plain stores; reduced in
    // tile order by dbias_reduce_kernel: run-to-run identical bits), else straight into dbias with atomics
    if (g == 0) {
      if (dbias_tile_stride > 0) dbias[(long long)(cid >> 1) * dbias_tile_stride + dir * LS_G + u * 4 + q] = v;
      else atomicAdd(dbias + dir * LS_G + u * 4 + q, v);
    }
  }
  cluster_sync_all();
}

static int cluster_slots() {
  // co-resident 8-CTA clusters (15 on a 148-SM B200: GPC packing)
  static int slots = 0;
  if (slots == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LS_CL * 64);
    cfg.blockDim = dim3(LS_THREADS);
    cfg.dynamicSmemBytes = sizeof(LstmFwdSmem<64>);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LS_CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    cudaFuncSetAttribute(lstm_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LstmFwdSmem<64>));
    if (cudaOccupancyMaxActiveClusters(&n, lstm_fwd_kernel<64>, &cfg) != cudaSuccess || n <= 0) n = 15;
    cudaGetLastError();
    slots = n;
  }
  return slots;
}

static int pick_bt(int B) {
  // smallest batch tile whose 2*ceil(B/BT) clusters are all co-resident: more, smaller tiles
  // shorten the per-step critical path; otherwise the largest tile (fewest waves).
  const int slots = cluster_slots();
  const int cand[3] = {16, 32, 64};
  for (int i = 0; i < 3; ++i)
    if (2 * ((B + cand[i] - 1) / cand[i]) <= slots) return cand[i];
  return 64;
}

template <int BT>
static int launch_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst, int T, int B, int y_il,
                      cudaStream_t st) {
  static bool attr_done = false;
  const int smem = (int)sizeof(LstmFwdSmem<BT>);
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int grid = 2 * ((B + BT - 1) / BT) * LS_CL;
  lstm_fwd_kernel<BT><<<grid, LS_THREADS, smem, st>>>(gates, whh, bias, y, cst, T, B, y_il);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

template <int BT>
static int launch_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy, float* dbias,
                      int dbias_tile_stride, int T, int B, cudaStream_t st) {
  static bool attr_done = false;
  const int smem = (int)sizeof(LstmBwdSmem<BT>);
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int grid = 2 * ((B + BT - 1) / BT) * LS_CL;
  lstm_bwd_kernel<BT><<<grid, LS_THREADS, smem, st>>>(gates, whhT, cst, dy, dbias, dbias_tile_stride, T, B);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

int launch_lstm4_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy, float* dbias,
                     int dbias_tile_stride, int T, int B, cudaStream_t st);   // lstm4_bwd.cu

// dbias[c] += sum over row tiles (in tile order) of part[tile][c]: the deterministic tail of the bias gradient
__global__ void dbias_reduce_kernel(const float* __restrict__ part, int tiles, int n, float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float acc = dbias[c];
  for (int t = 0; t < tiles; ++t) acc += __ldcg(part + (long long)t * n + c);
  dbias[c] = acc;
}
int launch_lstm4_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst, int T, int B,
                     int y_il, cudaStream_t st);     // lstm4.cu

// which forward kernel: tcgen05 (128-row tiles) once the batch no longer fits 16-row mma.sync tiles
// in one wave of clusters.  AVSI_LSTM_FWD=mma|tc overrides (A/B measurements only).
// returns 0 = mma.sync register-resident kernel (small batches: 16-row tiles, shortest step), 2 = tcgen05 4-CTA
// kernel (lstm4.cu).  AVSI_LSTM_FWD=mma|l4 overrides (A/B measurements and tests).
static int fwd_kernel_choice(int B) {
  AVSI_ENV_CACHE(mode, env_is("AVSI_LSTM_FWD", "mma") ? 1 : (env_is("AVSI_LSTM_FWD", "l4") ? 3 : 0));
  if (mode == 1) return 0;
  if (mode == 3) return 2;
  return pick_bt(B) > 32 ? 2 : 0;      // measured crossover: 16/32-row mma.sync tiles win while they fit one wave (B <= 224)
}

}  // namespace avsi

extern "C" int avsi_lstm_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst,
                             int T, int B, int y_il, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(gates && whh && bias && y && cst, "null pointer");
  AVSI_REQUIRE(T > 0 && B > 0, "T,B > 0");
  const int bt = pick_bt(B);
  cudaStream_t st = (cudaStream_t)stream;
  const int kc = fwd_kernel_choice(B);
  if (kc == 2) return launch_lstm4_fwd(gates, whh, bias, y, cst, T, B, y_il, st);
  if (bt == 16) return launch_fwd<16>(gates, whh, bias, y, cst, T, B, y_il, st);
  if (bt == 32) return launch_fwd<32>(gates, whh, bias, y, cst, T, B, y_il, st);
  return launch_fwd<64>(gates, whh, bias, y, cst, T, B, y_il, st);
}

extern "C" int64_t avsi_lstm_bwd_scratch_bytes(int B) {
  // one row of 2 x 1024 bias-gradient sums per row tile (at most ceil(B / 16) tiles, whichever kernel runs); the
  // reduce-scatter of partial dh lives in distributed shared memory
  if (B <= 0) return 0;
  return (int64_t)((B + 15) / 16) * 2 * avsi::LS_G * (int64_t)sizeof(float);
}

extern "C" int avsi_lstm_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy,
                             float* dbias, void* scratch, int T, int B, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(gates && whhT && cst && dy && dbias, "null pointer");
  AVSI_REQUIRE(T > 0 && B > 0, "T,B > 0");
  AVSI_REQUIRE(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned");
  const int bt = pick_bt(B);
  cudaStream_t st = (cudaStream_t)stream;
  // AVSI_LSTM_BWD=mma|l4 overrides (A/B measurements and the parity tests of the tcgen05 path at small batches)
  AVSI_ENV_CACHE(bmode, env_is("AVSI_LSTM_BWD", "mma") ? 1 : (env_is("AVSI_LSTM_BWD", "l4") ? 2 : 0));
  // with scratch: per-tile bias-gradient sums + an ordered reduction (bit-reproducible); without: fp32 atomics into dbias
  float* part = reinterpret_cast<float*>(scratch);
  const int stride = part ? 2 * LS_G : 0;
  float* dst = part ? part : dbias;
  int rc, tiles;
  if (bmode == 2 || (bmode == 0 && bt > 32)) {
    tiles = (B + 127) / 128;
    rc = launch_lstm4_bwd(gates, whhT, cst, dy, dst, stride, T, B, st);
  } else {
    tiles = (B + bt - 1) / bt;
    if (bt == 16) rc = launch_bwd<16>(gates, whhT, cst, dy, dst, stride, T, B, st);
    else if (bt == 32) rc = launch_bwd<32>(gates, whhT, cst, dy, dst, stride, T, B, st);
    else rc = launch_bwd<64>(gates, whhT, cst, dy, dst, stride, T, B, st);
  }
  if (rc != AVSI_OK || !part) return rc;
  dbias_reduce_kernel<<<(2 * LS_G + 255) / 256, 256, 0, st>>>(part, tiles, 2 * LS_G, dbias);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
