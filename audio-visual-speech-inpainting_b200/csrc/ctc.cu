// CTC negative log-likelihood + gradient on the GPU.
//
// Replaces tf.nn.ctc_loss (a CPU-only op in TF1, forcing a device->host->device hop every
// training step) as called at models.py:1950-1953 / models_asr.py:146-148: unnormalised
// time-major logits, blank = C-1, ctc_merge_repeated=True, frames t >= seq_len ignored.
// One 128-thread CTA per utterance: log-softmax rows -> alpha (forward in t) -> beta and the
// gradient softmax - posterior (backward in t), in fp32 log space RELATIVE TO A RUNNING SHIFT:
// log alpha_t falls by ~log C per frame (-5800 after the 1667 frames of a 20 s utterance, where one
// fp32 ulp is 5e-4 -- measured: 6e-3 relative L2 on the gradient), so every CTC_RENORM frames the
// state vector is re-centred on its maximum and the shift accumulates in a double; the posterior
// exponent combines the two small fp32 residuals with the double-precision shifts.  alpha and the
// log-softmax live in an L2-resident workspace; the recursion state sits in shared memory.
#include "common.cuh"

namespace avsi {

constexpr int CTC_THREADS = 128;
constexpr int CTC_MAX_S = 1024;   // 2*Lmax+1 limit (shared-memory state)
constexpr int CTC_MAX_C = 128;
constexpr int CTC_RENORM = 8;     // frames between re-centrings of the alpha / beta vectors
#define CTC_NEG (-1e30f)

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m <= CTC_NEG) return CTC_NEG;
  // expf, not __expf: the fast exponential's error (~1e-6 absolute on the sum, with a bias) adds up along the chain --
  // measured 2e-3 relative on the gradient after the 1667 frames of a 20 s utterance; log1pf / expf keep it below 1e-4
  return m + log1pf(expf(fminf(a, b) - m));
}

// SPT = states per thread: ceil((2 Lmax + 1) / 128).  The reference pads labels to 50 (tfrecord_utils.py:101): 101 states,
// SPT = 1 -- the per-state loops then have one iteration instead of eight predicated ones, and the state vectors take
// SPT * 128 entries of shared memory instead of 1024 (more utterances per SM).
template <int SPT>
__global__ void __launch_bounds__(CTC_THREADS)
ctc_kernel(const float* __restrict__ logits, int ldl, int col0, int C, const int32_t* __restrict__ labels, int Lmax,
           const int32_t* __restrict__ lab_len, const int32_t* __restrict__ seq_len, int B, int T, float grad_scale,
           const float* __restrict__ grad_scale_dev, float* __restrict__ nll, uint16_t* __restrict__ dlogits,
           int ldd, int dcol0, float* __restrict__ ws_logp, float* __restrict__ ws_alpha, double* __restrict__ ws_shift) {
  constexpr int NS = SPT * CTC_THREADS;             // states this instantiation can hold
  __shared__ int ext[NS];
  __shared__ unsigned char skip[NS];
  __shared__ float st[2][NS];
  __shared__ float gam[NS];                         // posterior mass of every state at the current frame
  __shared__ unsigned short cls_states[NS];         // states grouped by class, ascending inside a class ...
  __shared__ int cls_start[CTC_MAX_C + 1];          // ... class c owns cls_states[cls_start[c] .. cls_start[c+1])
  __shared__ float ll_sh;
  __shared__ float red[CTC_THREADS / 32];
  __shared__ int bad_label;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blank = C - 1;
  const int L = min(max(lab_len[b], 0), Lmax);
  const int Tb = min(max(seq_len[b], 0), T);
  const int S = 2 * L + 1;
  const int Smax = 2 * Lmax + 1;
  const float scale = grad_scale * (grad_scale_dev ? *grad_scale_dev : 1.f);
  float* logp = ws_logp + (long long)b * T * C;
  float* alpha = ws_alpha + (long long)b * T * Smax;
  double* ashift = ws_shift + (long long)b * T;      // alpha[t][s] holds log alpha_t(s) - ashift[t]

  if (tid == 0) bad_label = 0;
  __syncthreads();
  for (int s = tid; s < S; s += CTC_THREADS) {
    int e = blank;
    if (s & 1) {
      e = labels[b * Lmax + (s >> 1)];
      // tf.nn.ctc_loss raises InvalidArgument for a label outside [0, num_classes - 1); here the utterance is
      // reported infeasible (nll = +inf, zero gradient) and the index never reaches memory
      if (e < 0 || e >= blank) {
        bad_label = 1;
        e = 0;
      }
    }
    ext[s] = e;
  }
  __syncthreads();
  for (int s = tid; s < S; s += CTC_THREADS) skip[s] = (s >= 2 && ext[s] != blank && ext[s] != ext[s - 2]) ? 1 : 0;
  // states of every class in ascending order (a counting sort, once per utterance): the per-frame posterior of a class is
  // then summed by ONE thread in a fixed order -- bit-reproducible, unlike a scatter of the states through atomics
  if (dlogits) {
    for (int c = tid; c < C; c += CTC_THREADS) {
      int n = 0;
      for (int s = 0; s < S; ++s) n += (ext[s] == c);
      cls_start[c + 1] = n;
    }
    __syncthreads();
    if (tid == 0) {
      cls_start[0] = 0;
      for (int c = 0; c < C; ++c) cls_start[c + 1] += cls_start[c];
    }
    __syncthreads();
    for (int c = tid; c < C; c += CTC_THREADS) {
      int k = cls_start[c];
      for (int s = 0; s < S; ++s)
        if (ext[s] == c) cls_states[k++] = (unsigned short)s;
    }
  }

  // ---- log-softmax, one warp per frame ---------------------------------------------------
  for (int t = warp; t < Tb; t += CTC_THREADS / 32) {
    const float* row = logits + ((long long)t * B + b) * ldl + col0;
    float mx = CTC_NEG;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
    sum = warp_sum(sum);
    const float lz = mx + logf(sum);
    for (int c = lane; c < C; c += 32) logp[t * C + c] = row[c] - lz;
  }
  // frames beyond the sequence get zero gradient
  if (dlogits)
    for (int i = tid; i < (T - Tb) * C; i += CTC_THREADS) {
      const int t = Tb + i / C, c = i % C;
      dlogits[((long long)t * B + b) * ldd + dcol0 + c] = 0;
    }
  __syncthreads();
  if (Tb == 0) {
    if (tid == 0) nll[b] = (L == 0) ? 0.f : INFINITY;
    return;
  }

  // ---- alpha -------------------------------------------------------------------------------
  for (int s = tid; s < S; s += CTC_THREADS) {
    float a = CTC_NEG;
    if (s == 0) a = logp[blank];
    else if (s == 1) a = logp[ext[1]];
    st[0][s] = a;
    alpha[s] = a;
  }
  if (tid == 0) ashift[0] = 0.0;
  double a_shift = 0.0;                               // identical in every thread
  __syncthreads();
  // The recursions are a dependent chain of Tb steps: the global (L2) operands of step t+1 are requested during
  // step t so that only shared-memory latency sits on the chain.  S <= 8 * CTC_THREADS (Lmax limit) -> <= 8 states per thread.
  float lp_next[SPT];
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = tid + i * CTC_THREADS;
    lp_next[i] = (s < S && Tb > 1) ? logp[C + ext[s]] : 0.f;
  }
  for (int t = 1; t < Tb; ++t) {
    const float* prev = st[(t - 1) & 1];
    float* cur = st[t & 1];
    const bool renorm = (t % CTC_RENORM) == 0;
    float areg[SPT], amax = CTC_NEG;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = tid + i * CTC_THREADS;
      areg[i] = CTC_NEG;
      if (s >= S) continue;
      const float lp = lp_next[i];
      if (t + 1 < Tb) lp_next[i] = logp[(t + 1) * C + ext[s]];
      float a = prev[s];
      if (s >= 1) a = lse2(a, prev[s - 1]);
      if (skip[s]) a = lse2(a, prev[s - 2]);
      a = (a <= CTC_NEG) ? CTC_NEG : a + lp;
      areg[i] = a;
      amax = fmaxf(amax, a);
    }
    if (renorm) {                                    // re-centre on the maximum: one extra block reduction every CTC_RENORM frames
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) red[warp] = amax;
      __syncthreads();
      float m = red[0];
#pragma unroll
      for (int k = 1; k < CTC_THREADS / 32; ++k) m = fmaxf(m, red[k]);
      if (m > CTC_NEG) {
        a_shift += (double)m;
#pragma unroll
        for (int i = 0; i < SPT; ++i)
          if (areg[i] > CTC_NEG) areg[i] -= m;
      }
    }
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = tid + i * CTC_THREADS;
      if (s >= S) break;
      cur[s] = areg[i];
      alpha[(long long)t * Smax + s] = areg[i];
    }
    if (tid == 0) ashift[t] = a_shift;
    __syncthreads();
  }
  if (tid == 0) {
    const float* last = st[(Tb - 1) & 1];
    ll_sh = (S == 1) ? last[0] : lse2(last[S - 1], last[S - 2]);
  }
  __syncthreads();
  const float ll_rel = ll_sh;                         // log p(l|x) - a_shift
  const bool feasible = ll_rel > CTC_NEG && !bad_label;
  const double ll = (double)ll_rel + a_shift;
  if (tid == 0) nll[b] = feasible ? (float)(-ll) : INFINITY;
  if (!dlogits) return;

  // ---- beta + gradient -------------------------------------------------------------------
  float al_next[SPT], lpc_next = 0.f;               // operands of the next (earlier) frame, requested one step ahead
#pragma unroll
  for (int i = 0; i < SPT; ++i) {
    const int s = tid + i * CTC_THREADS;
    lp_next[i] = (s < S) ? logp[(Tb - 1) * C + ext[s]] : 0.f;
    al_next[i] = (s < S) ? alpha[(long long)(Tb - 1) * Smax + s] : 0.f;
  }
  if (tid < C) lpc_next = logp[(Tb - 1) * C + tid];
  double b_shift = 0.0;                               // st[.][s] holds log beta_t(s) - b_shift
  for (int t = Tb - 1; t >= 0; --t) {
    float* cur = st[t & 1];
    const float* nxt = st[(t + 1) & 1];
    const float lpc = lpc_next;
    if (t > 0 && tid < C) lpc_next = logp[(t - 1) * C + tid];
    const bool renorm = ((Tb - 1 - t) % CTC_RENORM) == 0 && t != Tb - 1;
    float breg[SPT], lpreg[SPT], alreg[SPT], bmax = CTC_NEG;
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = tid + i * CTC_THREADS;
      breg[i] = CTC_NEG;
      lpreg[i] = alreg[i] = 0.f;
      if (s >= S) continue;
      const float lp = lp_next[i], al = al_next[i];
      if (t > 0) {
        lp_next[i] = logp[(t - 1) * C + ext[s]];
        al_next[i] = alpha[(long long)(t - 1) * Smax + s];
      }
      float bt;
      if (t == Tb - 1) {
        bt = (s == S - 1 || s == S - 2) ? 0.f : CTC_NEG;
      } else {
        bt = nxt[s];
        if (s + 1 < S) bt = lse2(bt, nxt[s + 1]);
        if (s + 2 < S && skip[s + 2]) bt = lse2(bt, nxt[s + 2]);
      }
      bt = (bt <= CTC_NEG) ? CTC_NEG : bt + lp;
      breg[i] = bt;
      lpreg[i] = lp;
      alreg[i] = al;
      bmax = fmaxf(bmax, bt);
    }
    if (renorm) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
      if (lane == 0) red[warp] = bmax;
      __syncthreads();
      float m = red[0];
#pragma unroll
      for (int k = 1; k < CTC_THREADS / 32; ++k) m = fmaxf(m, red[k]);
      if (m > CTC_NEG) {
        b_shift += (double)m;
#pragma unroll
        for (int i = 0; i < SPT; ++i)
          if (breg[i] > CTC_NEG) breg[i] -= m;
      }
    }
    // log alpha_t + log beta_t - log p(l|x) = (alpha residual + beta residual) + (ashift[t] + b_shift - ll): the bracket is a
    // difference of large numbers, taken in double
    const float shift = (float)(ashift[t] + b_shift - ll);
#pragma unroll
    for (int i = 0; i < SPT; ++i) {
      const int s = tid + i * CTC_THREADS;
      if (s >= S) break;
      cur[s] = breg[i];
      // posterior mass of state s: alpha * beta / (y_t(ext_s) * p(l|x)); alpha*beta counts y_t twice -> subtract lp once.
      // It is a probability (<= 1): no max-shift needed.
      const float e = alreg[i] + breg[i] - lpreg[i] + shift;
      gam[s] = (feasible && alreg[i] > CTC_NEG && breg[i] > CTC_NEG) ? expf(fminf(e, 0.f)) : 0.f;
    }
    __syncthreads();
    // gradient of the NLL wrt the logits: softmax - posterior; the posterior of class c = its states' masses, added in
    // ascending state order on four interleaved partial sums (the blank owns L + 1 states)
    for (int c = tid; c < C; c += CTC_THREADS) {
      const float lp = (c == tid) ? lpc : logp[t * C + c];
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
      int k = cls_start[c];
      const int kz = cls_start[c + 1];
      for (; k + 4 <= kz; k += 4) {
        p0 += gam[cls_states[k]];
        p1 += gam[cls_states[k + 1]];
        p2 += gam[cls_states[k + 2]];
        p3 += gam[cls_states[k + 3]];
      }
      for (; k < kz; ++k) p0 += gam[cls_states[k]];
      const float grad = feasible ? (expf(lp) - ((p0 + p1) + (p2 + p3))) : 0.f;
      dlogits[((long long)t * B + b) * ldd + dcol0 + c] = __half_as_ushort(__float2half_rn(scale * grad));
    }
    __syncthreads();
  }
}

}  // namespace avsi

extern "C" int64_t avsi_ctc_workspace_bytes(int B, int T, int Lmax) {
  if (B <= 0 || T <= 0 || Lmax < 0) return 0;
  // log-softmax [B,T,128] f32 + alpha [B,T,2 Lmax + 1] f32 (padded to an even count) + alpha shifts [B,T] f64
  const int64_t smax = (2LL * Lmax + 1 + 1) & ~1LL;
  return (int64_t)B * T * (avsi::CTC_MAX_C + smax) * (int64_t)sizeof(float) + (int64_t)B * T * (int64_t)sizeof(double);
}

extern "C" int avsi_ctc_loss(const float* logits, int ldl, int col0, int C, const int32_t* labels, int Lmax,
                             const int32_t* lab_len, const int32_t* seq_len, int B, int T, float grad_scale,
                             const float* grad_scale_dev, float* nll, uint16_t* dlogits, int ldd, int dcol0,
                             void* workspace, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(logits && labels && lab_len && seq_len && nll && workspace, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && C > 1 && C <= CTC_MAX_C, "sizes");
  AVSI_REQUIRE(Lmax >= 1 && 2 * Lmax + 1 <= CTC_MAX_S, "Lmax");
  AVSI_REQUIRE(ldl >= col0 + C, "ldl");
  AVSI_REQUIRE(!dlogits || ldd >= dcol0 + C, "ldd");
  float* ws_logp = reinterpret_cast<float*>(workspace);
  AVSI_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
  float* ws_alpha = ws_logp + (long long)B * T * CTC_MAX_C;
  const long long smax = (2LL * Lmax + 1 + 1) & ~1LL;
  double* ws_shift = reinterpret_cast<double*>(ws_alpha + (long long)B * T * smax);
  const int spt = (2 * Lmax + 1 + CTC_THREADS - 1) / CTC_THREADS;
#define CTC_LAUNCH(SPT_)                                                                                              \
  ctc_kernel<SPT_><<<B, CTC_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, col0, C, labels, Lmax, lab_len, seq_len, B, T, \
                                                               grad_scale, grad_scale_dev, nll, dlogits, ldd, dcol0,  \
                                                               ws_logp, ws_alpha, ws_shift)
  if (spt <= 1) CTC_LAUNCH(1);
  else if (spt <= 2) CTC_LAUNCH(2);
  else if (spt <= 4) CTC_LAUNCH(4);
  else CTC_LAUNCH(8);
#undef CTC_LAUNCH
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
