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
// REGISTERS as mma.sync B-fragments for the whole sequence (64 regs/thread), the recurrent
// product runs on mma.sync.m16n8k16 (fp16 in, fp32 accumulate), c_t stays in fp32 registers.
// Per step the CTAs exchange h_t through the layer-output buffer in L2 (it must be written
// anyway) and one barrier.cluster release/acquire; x.W_ih pre-activations for the step are
// prefetched before the barrier wait.  Gate columns are laid out [dir][unit][i,g,f,o] so a
// thread's four gates are one 8-byte load and the activated gates overwrite the
// pre-activations in place (stash for BPTT); BPTT overwrites them again with dgates.
//
// BPTT: CTA j multiplies ITS 128 dgate columns with its W_hh^T slice (K-split, registers),
// partial dh[BT,256] go through an L2 scratch and are summed by the owners after the
// cluster barrier (reduce-scatter).  Bias gradients accumulate in registers over t.
#include "common.cuh"

namespace avsi {

constexpr int LS_HP = 256;        // padded hidden size
constexpr int LS_G = 1024;        // gate columns per direction
constexpr int LS_CL = 8;          // CTAs per cluster
constexpr int LS_THREADS = 256;
constexpr int LS_HSTRIDE = LS_HP + 8;     // halves per smem row of h (528 B, conflict-free ldmatrix)
constexpr int LS_DSTRIDE = 128 + 8;       // halves per smem row of dgates

__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int BT>
__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
lstm_fwd_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whh, const float* __restrict__ bias,
                uint16_t* __restrict__ y, float* __restrict__ cst, int T, int B) {
  constexpr int MT = BT / 16;
  __shared__ __align__(16) uint16_t hbuf[BT * LS_HSTRIDE];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, uu = lane & 3;
  const int cid = blockIdx.x / LS_CL, j = blockIdx.x % LS_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int u_base = j * 32 + w * 4;
  const int u = u_base + uu;                       // hidden unit whose cell this thread updates

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
  float bq[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bq[q] = bias[dir * LS_G + u * 4 + q];

  float c_state[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) c_state[mt][0] = c_state[mt][1] = 0.f;

  const uint32_t hbuf_s = (uint32_t)__cvta_generic_to_shared(hbuf);

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    const int tp = dir ? (t + 1) : (t - 1);
    // ---- prefetch x.W_ih pre-activations of this step (independent of h) ----------------
    uint2 pre[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int row = b0 + mt * 16 + g + rh * 8;
        pre[mt][rh] = make_uint2(0u, 0u);
        if (row < B)
          pre[mt][rh] = *reinterpret_cast<const uint2*>(gates + ((long long)t * B + row) * (2 * LS_G) + dir * LS_G + u * 4);
      }
    float acc[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    if (s > 0) {
      cluster_wait_acquire();                      // every CTA's h_{t-1} slice is in L2
      for (int idx = tid; idx < BT * 32; idx += LS_THREADS) {
        const int r = idx >> 5, ch = idx & 31;
        const uint32_t dst = hbuf_s + (uint32_t)(r * LS_HSTRIDE + ch * 8) * 2u;
        if (b0 + r < B)
          cp_async16(dst, y + ((long long)tp * B + b0 + r) * (2 * LS_HP) + dir * LS_HP + ch * 8);
        else
          *reinterpret_cast<uint4*>(hbuf + r * LS_HSTRIDE + ch * 8) = make_uint4(0, 0, 0, 0);
      }
      cp_async_wait_all();
      __syncthreads();
#pragma unroll
      for (int kt = 0; kt < 16; ++kt) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(hbuf_s + (uint32_t)((mt * 16 + (lane & 15)) * LS_HSTRIDE + kt * 16 + (lane >> 4) * 8) * 2u,
                      a0, a1, a2, a3);
          mma16816(acc[mt][0], a0, a1, a2, a3, wf[kt][0][0], wf[kt][0][1]);
          mma16816(acc[mt][1], a0, a1, a2, a3, wf[kt][1][0], wf[kt][1][1]);
        }
      }
    }
    // ---- gates, cell update, stores --------------------------------------------------------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int row = b0 + mt * 16 + g + rh * 8;
        const float2 p_ig = unpack_half2(pre[mt][rh].x), p_fo = unpack_half2(pre[mt][rh].y);
        const float gi = sigmoidf_acc(acc[mt][0][rh * 2 + 0] + p_ig.x + bq[0]);
        const float gg = tanhf_acc(acc[mt][0][rh * 2 + 1] + p_ig.y + bq[1]);
        const float gf = sigmoidf_acc(acc[mt][1][rh * 2 + 0] + p_fo.x + bq[2]);
        const float go = sigmoidf_acc(acc[mt][1][rh * 2 + 1] + p_fo.y + bq[3]);
        const float c = gf * c_state[mt][rh] + gi * gg;
        c_state[mt][rh] = c;
        const float h = go * tanhf_acc(c);
        const float hn = __shfl_down_sync(0xffffffffu, h, 1);
        if (row < B) {
          const long long r = (long long)t * B + row;
          *reinterpret_cast<uint2*>(gates + r * (2 * LS_G) + dir * LS_G + u * 4) =
              make_uint2(pack_half2(gi, gg), pack_half2(gf, go));
          cst[r * (2 * LS_HP) + dir * LS_HP + u] = c;
          if ((uu & 1) == 0) *reinterpret_cast<uint32_t*>(y + r * (2 * LS_HP) + dir * LS_HP + u) = pack_half2(h, hn);
        }
      }
    if (s + 1 < T) cluster_arrive_release();       // publishes this CTA's h_t stores to the cluster
  }
}

template <int BT>
__global__ void __cluster_dims__(LS_CL, 1, 1) __launch_bounds__(LS_THREADS, 1)
lstm_bwd_kernel(uint16_t* __restrict__ gates, const uint16_t* __restrict__ whhT, const float* __restrict__ cst,
                const uint16_t* __restrict__ dy, float* __restrict__ dbias, float* __restrict__ scratch, int T, int B) {
  constexpr int MT = BT / 16;
  __shared__ __align__(16) uint16_t dgbuf[BT * LS_DSTRIDE];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, uu = lane & 3;
  const int cid = blockIdx.x / LS_CL, j = blockIdx.x % LS_CL;
  const int dir = cid & 1, b0 = (cid >> 1) * BT;
  const int ul = w * 4 + uu;                        // unit local to this CTA (0..31)
  const int u = j * 32 + ul;

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
  float* part = scratch + (long long)cid * (2LL * LS_CL * BT * LS_HP);
  float dc_state[MT][2], c_cur[MT][2], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) dc_state[mt][0] = dc_state[mt][1] = 0.f;
  const uint32_t dg_s = (uint32_t)__cvta_generic_to_shared(dgbuf);

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
          gt[mt][rh] = *reinterpret_cast<const uint2*>(gates + r * (2 * LS_G) + dir * LS_G + u * 4);
          dyv[mt][rh] = __half2float(__ushort_as_half(dy[r * (2 * LS_HP) + dir * LS_HP + u]));
          if (has_prev) cprev[mt][rh] = cst[((long long)tp * B + row) * (2 * LS_HP) + dir * LS_HP + u];
          if (s == 0) c_cur[mt][rh] = cst[r * (2 * LS_HP) + dir * LS_HP + u];
        } else if (s == 0) {
          c_cur[mt][rh] = 0.f;
        }
      }
    float dhr[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) dhr[mt][0] = dhr[mt][1] = 0.f;
    if (s > 0) {
      cluster_wait_acquire();                       // all partial dh of the previous step are in L2
      const float* pb = part + (long long)((s - 1) & 1) * (LS_CL * BT * LS_HP);
#pragma unroll
      for (int src = 0; src < LS_CL; ++src)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int rh = 0; rh < 2; ++rh)
            dhr[mt][rh] += __ldcg(pb + ((long long)src * BT + mt * 16 + g + rh * 8) * LS_HP + u);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        const int rl = mt * 16 + g + rh * 8;
        const float2 ig = unpack_half2(gt[mt][rh].x), fo = unpack_half2(gt[mt][rh].y);
        const float gi = ig.x, gg = ig.y, gf = fo.x, go = fo.y;
        const float dh = dyv[mt][rh] + dhr[mt][rh];
        const float tc = tanhf_acc(c_cur[mt][rh]);
        const float d_o = dh * tc * go * (1.f - go);
        const float dc = dc_state[mt][rh] + dh * go * (1.f - tc * tc);
        const float d_i = dc * gg * gi * (1.f - gi);
        const float d_g = dc * gi * (1.f - gg * gg);
        const float d_f = dc * cprev[mt][rh] * gf * (1.f - gf);
        dc_state[mt][rh] = dc * gf;
        c_cur[mt][rh] = cprev[mt][rh];
        const uint2 pk = make_uint2(pack_half2(d_i, d_g), pack_half2(d_f, d_o));
        *reinterpret_cast<uint2*>(dgbuf + rl * LS_DSTRIDE + ul * 4) = pk;
        if (b0 + rl < B) {
          *reinterpret_cast<uint2*>(gates + ((long long)t * B + b0 + rl) * (2 * LS_G) + dir * LS_G + u * 4) = pk;
          db[0] += d_i;
          db[1] += d_g;
          db[2] += d_f;
          db[3] += d_o;
        }
      }
    if (has_prev) {
      __syncthreads();
      float acc[MT][4][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(dg_s + (uint32_t)((mt * 16 + (lane & 15)) * LS_DSTRIDE + kt * 16 + (lane >> 4) * 8) * 2u,
                      a0, a1, a2, a3);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816(acc[mt][nt], a0, a1, a2, a3, wf[kt][nt][0], wf[kt][nt][1]);
        }
      }
      float* pw = part + (long long)(s & 1) * (LS_CL * BT * LS_HP) + (long long)j * BT * LS_HP;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int n = w * 32 + nt * 8 + uu * 2;
          __stcg(reinterpret_cast<float2*>(pw + (long long)(mt * 16 + g) * LS_HP + n), make_float2(acc[mt][nt][0], acc[mt][nt][1]));
          __stcg(reinterpret_cast<float2*>(pw + (long long)(mt * 16 + g + 8) * LS_HP + n), make_float2(acc[mt][nt][2], acc[mt][nt][3]));
        }
      cluster_arrive_release();
    }
  }
  // bias gradient: sum over the 8 row-lanes (g) that share this unit, one atomic per (unit, gate, CTA)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v = db[q];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (g == 0) atomicAdd(dbias + dir * LS_G + u * 4 + q, v);
  }
}

static int pick_bt(int B) {
  // smallest batch tile whose 2*ceil(B/BT) clusters are all co-resident (about 16 on 148 SMs):
  // more, smaller tiles shorten the per-step critical path.
  static int slots = 0;
  if (slots == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LS_CL * 64);
    cfg.blockDim = dim3(LS_THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LS_CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_fwd_kernel<64>, &cfg) != cudaSuccess || n <= 0) n = 16;
    cudaGetLastError();
    slots = n;
  }
  const int cand[3] = {16, 32, 64};
  for (int i = 0; i < 3; ++i)
    if (2 * ((B + cand[i] - 1) / cand[i]) <= slots) return cand[i];
  return 64;
}

}  // namespace avsi

extern "C" int avsi_lstm_fwd(uint16_t* gates, const uint16_t* whh, const float* bias, uint16_t* y, float* cst,
                             int T, int B, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(gates && whh && bias && y && cst, "null pointer");
  AVSI_REQUIRE(T > 0 && B > 0, "T,B > 0");
  const int bt = pick_bt(B);
  const int grid = 2 * ((B + bt - 1) / bt) * LS_CL;
  cudaStream_t st = (cudaStream_t)stream;
  if (bt == 16) lstm_fwd_kernel<16><<<grid, LS_THREADS, 0, st>>>(gates, whh, bias, y, cst, T, B);
  else if (bt == 32) lstm_fwd_kernel<32><<<grid, LS_THREADS, 0, st>>>(gates, whh, bias, y, cst, T, B);
  else lstm_fwd_kernel<64><<<grid, LS_THREADS, 0, st>>>(gates, whh, bias, y, cst, T, B);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int64_t avsi_lstm_bwd_scratch_bytes(int B) {
  using namespace avsi;
  if (B <= 0) return 0;
  const int bt = pick_bt(B);
  const long long clusters = 2LL * ((B + bt - 1) / bt);
  return clusters * 2LL * LS_CL * bt * LS_HP * (long long)sizeof(float);
}

extern "C" int avsi_lstm_bwd(uint16_t* gates, const uint16_t* whhT, const float* cst, const uint16_t* dy,
                             float* dbias, void* scratch, int T, int B, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(gates && whhT && cst && dy && dbias && scratch, "null pointer");
  AVSI_REQUIRE(T > 0 && B > 0, "T,B > 0");
  const int bt = pick_bt(B);
  const int grid = 2 * ((B + bt - 1) / bt) * LS_CL;
  cudaStream_t st = (cudaStream_t)stream;
  float* sc = reinterpret_cast<float*>(scratch);
  if (bt == 16) lstm_bwd_kernel<16><<<grid, LS_THREADS, 0, st>>>(gates, whhT, cst, dy, dbias, sc, T, B);
  else if (bt == 32) lstm_bwd_kernel<32><<<grid, LS_THREADS, 0, st>>>(gates, whhT, cst, dy, dbias, sc, T, B);
  else lstm_bwd_kernel<64><<<grid, LS_THREADS, 0, st>>>(gates, whhT, cst, dy, dbias, sc, T, B);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
