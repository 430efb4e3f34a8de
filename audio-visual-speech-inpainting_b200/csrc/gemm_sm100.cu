// fp16 x fp16 -> fp32 GEMMs on the 5th-gen tensor cores (tcgen05.mma kind::f16, accumulators in TMEM, operands
// staged by TMA).
//
// Replaces the cuBLAS/cuDNN contractions behind CudnnLSTM's input projections
// (models.py:95-104), the tf.matmul heads (models.py:117-123, 1902-1912) and their gradients.
//
// Two kernels share the operand / epilogue code:
//   gemm_f16_2sm_kernel  (below, the hot one: projections, dX, dW) persistent CTA pairs, cta_group::2, 256 x 256 tiles
//   gemm_f16_kernel      (first) small or narrow problems (head GEMM, tests): one 128 x BN tile per CTA, cta_group::1,
//                        2 CTAs co-resident per SM so one CTA's epilogue overlaps the other's main loop:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) into a STAGES-deep ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer, tcgen05.commit -> mbarriers
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns) -> registers -> global
// Operand layouts:
//   trans == 0  A [M,K], B [N,K] row-major  -> K-major smem tiles   (box 64(k) x rows)
//   trans == 1  A [K,M], B [K,N] row-major  -> MN-major smem tiles  (boxes 64(mn) x 64(k))
//   layout & 1  A is stored INTERLEAVED (common.cuh il16: [rows/32][cols/8][32][8]); one 3-D TMA box per
//               stage lands it as SWIZZLE_NONE core matrices: K-major [k-chunk][128 rows][16 B] (trans 0,
//               LBO 2048 / SBO 128) or MN-major [m-chunk][64 k-rows][16 B] (trans 1, LBO 128 / SBO 1024)
//   layout & 4  B is stored interleaved likewise (trans = 1 only: the [K, N] operand of the dW contractions, e.g. the
//               layer outputs Y); tile = [n-chunk][64 k-rows][16 B], the mirror image of the interleaved A^T tile
//   layout & 2  the f16 output (out_mode 0) is written interleaved: a warp's 32 rows x 8 columns are one
//               contiguous 512-byte store
// Shared-memory matrix descriptors follow the canonical SWIZZLE_128B layouts
// (K-major: SBO = 1024 B; MN-major: SBO = 1024 B, LBO = 8192 B), version = 1 (sm_100).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avsi {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

// Masked-L1 loss + gradient fused into the epilogue of the head GEMM (avsi_head_l1): the arguments of avsi_masked_l1
// (loss.cu).  The fp32 logits then never reach HBM (655 MB written and read back per step at B = 2048), except the
// columns >= logits_from_col that a second loss (CTC) still reads.
struct L1Fuse {
  const float* target;          // [B, T, F] f32
  const float* mask;            // [B, T, F] f32
  const int32_t* seq_len;       // [B]
  const float* grad_scale_dev;  // optional device factor of the gradient
  double* partial;              // [L1_PARTS][8] partial sums (zeroed by the launcher, reduced by l1_partials_kernel)
  uint16_t* dlogits;            // [M, ldd] f16, columns < F written
  int B, T, F, mode, ldd, logits_from_col;
  float grad_scale;
};
constexpr int L1_PARTS = 64;

struct GemmParams {
  void* C;
  const float* bias;
  int ldc, M, N, K;
  int trans, out_mode, split_k;
  int a_il, c_il, b_il;
  int fuse_l1;
  L1Fuse l1;
  float* part;                  // out_mode 2, ordered split-K: partial products [k split][M][part_ld] f32 (else nullptr: atomics)
  int part_ld;
};

// one element of the masked-L1 loss (same arithmetic as masked_l1_kernel, loss.cu): x = logit, y = target, m = mask,
// sm = 1 inside the utterance; returns the fp16 gradient word and adds to the five running sums
__device__ __forceinline__ uint16_t l1_element(float x, float y, float m, float sm, int mode, float grad_scale, float (&acc)[5]) {
  float pred = (mode == 0) ? x : (y * m + x * (1.f - m));
  pred *= sm;
  const float d = y - pred;
  const float ad = fabsf(d);
  acc[0] += ad * (1.f - m);
  acc[1] += (1.f - m);
  acc[2] += ad * m;
  acc[3] += m;
  acc[4] += ad;
  const float sg = (d > 0.f) ? -1.f : ((d < 0.f) ? 1.f : 0.f);
  const float w = (mode == 0) ? sm : sm * (1.f - m);
  return __half_as_ushort(__float2half_rn(grad_scale * sg * w));
}

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;    // barriers + alignment slack
};

// one warp-lane's 32 consecutive accumulator columns [n0, n0+32) of output row `row`
// (ks = index of the k split that produced them: selects the slice of the ordered split-K partials)
__device__ __forceinline__ void epilogue_store(const GemmParams& p, int row, int n0, bool full, const uint32_t (&r)[32], int ks = 0) {
  if (p.out_mode == 0 && p.c_il) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n0 + 8 * j >= p.N) break;
      uint4 v;
      v.x = pack_half2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
      v.y = pack_half2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
      v.z = pack_half2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
      v.w = pack_half2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.C) + il16(row, n0 + 8 * j, p.ldc)) = v;
    }
  } else if (p.out_mode == 0) {
    uint16_t* dst = reinterpret_cast<uint16_t*>(p.C) + (long long)row * p.ldc + n0;
    if (full && (p.ldc % 8 == 0)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        v.x = pack_half2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
        v.y = pack_half2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
        v.z = pack_half2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
        v.w = pack_half2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
        reinterpret_cast<uint4*>(dst)[j] = v;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) dst[j] = __half_as_ushort(__float2half_rn(__uint_as_float(r[j])));
    }
  } else if (p.out_mode == 1) {
    float* dst = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + n0;
    if (full && (p.ldc % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v;
        v.x = __uint_as_float(r[4 * j + 0]);
        v.y = __uint_as_float(r[4 * j + 1]);
        v.z = __uint_as_float(r[4 * j + 2]);
        v.w = __uint_as_float(r[4 * j + 3]);
        if (p.bias) {
          v.x += __ldg(p.bias + n0 + 4 * j + 0);
          v.y += __ldg(p.bias + n0 + 4 * j + 1);
          v.z += __ldg(p.bias + n0 + 4 * j + 2);
          v.w += __ldg(p.bias + n0 + 4 * j + 3);
        }
        reinterpret_cast<float4*>(dst)[j] = v;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) dst[j] = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + n0 + j) : 0.f);
    }
  } else if (p.part) {
    // ordered split-K: this split's product goes to its own slice with plain stores; splitk_reduce_kernel adds the slices
    // to C in split order (bit-reproducible, unlike the atomics below)
    float* dst = p.part + ((long long)ks * p.M + row) * p.part_ld + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1]),
                                                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    } else {
      for (int j = 0; j < 32; ++j)
        if (n0 + j < p.N) dst[j] = __uint_as_float(r[j]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + n0;
    for (int j = 0; j < 32; ++j)
      if (n0 + j < p.N) atomicAdd(dst + j, __uint_as_float(r[j]));
  }
}

// C[M, N] += sum over the k splits, in split order, of part[split][M][part_ld]
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, int part_ld, float* __restrict__ C, int ldc) {
  const int vpr = part_ld >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * vpr) return;
  const int row = (int)(idx / vpr), c = (int)(idx - (long long)row * vpr) * 4;
  if (c >= N) return;
  float* dst = C + (long long)row * ldc + c;
  const bool vec = (c + 4 <= N) && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) {
    acc = *reinterpret_cast<const float4*>(dst);
  } else {
    acc.x = dst[0];
    if (c + 1 < N) acc.y = dst[1];
    if (c + 2 < N) acc.z = dst[2];
    if (c + 3 < N) acc.w = dst[3];
  }
  const float* src = part + (long long)row * part_ld + c;
  const long long slice = (long long)M * part_ld;
#pragma unroll 4
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + s * slice));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (vec) {
    *reinterpret_cast<float4*>(dst) = acc;
  } else {
    dst[0] = acc.x;
    if (c + 1 < N) dst[1] = acc.y;
    if (c + 2 < N) dst[2] = acc.z;
    if (c + 3 < N) dst[3] = acc.w;
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using S = GemmSmem<BN, STAGES>;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_blk = blockIdx.x / n_tiles;
  const int n_blk = blockIdx.x % n_tiles;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int chunk = (kb_total + p.split_k - 1) / p.split_k;
  const int kb0 = blockIdx.y * chunk;
  const int kb1 = min(kb_total, kb0 + chunk);
  const int nkb = kb1 - kb0;
  if (nkb <= 0) return;   // uniform for the whole CTA

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        mbar_expect_tx(full_bar(s), S::STAGE_BYTES);
        const int k0 = (kb0 + i) * GEMM_BK;
        if (p.trans == 0) {
          if (p.a_il) tma_load_3d(sa, &tmA, full_bar(s), 0, m_blk * (GEMM_BM / 32), k0 >> 3);
          else tma_load_2d(sa, &tmA, full_bar(s), k0, m_blk * GEMM_BM);
          tma_load_2d(sb, &tmB, full_bar(s), k0, n_blk * BN);
        } else {
          if (p.a_il) tma_load_3d(sa, &tmA, full_bar(s), 0, k0 >> 5, m_blk * (GEMM_BM / 8));
          else {
#pragma unroll
            for (int a = 0; a < GEMM_BM / 64; ++a) tma_load_2d(sa + a * 8192, &tmA, full_bar(s), m_blk * GEMM_BM + 64 * a, k0);
          }
          if (p.b_il) tma_load_3d(sb, &tmB, full_bar(s), 0, k0 >> 5, n_blk * (BN / 8));
          else {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) tma_load_2d(sb + b * 8192, &tmB, full_bar(s), n_blk * BN + 64 * b, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(GEMM_BM, BN, p.trans, p.trans);
      const uint32_t lbo = p.trans ? 8192u : 0u;
      const uint32_t kstep = p.trans ? 2048u : 32u;   // bytes per UMMA_K = 16 advance
      // interleaved A: SWIZZLE_NONE core matrices (see header comment)
      const uint32_t a_lbo = p.trans ? 128u : 2048u, a_sbo = p.trans ? 1024u : 128u;
      const uint32_t a_kstep = p.trans ? 256u : 4096u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          const uint64_t da = p.a_il ? make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, 0u)
                                     : make_smem_desc(sa + k * kstep, lbo, 1024u);
          const uint64_t db = p.b_il ? make_smem_desc(sb + k * a_kstep, a_lbo, a_sbo, 0u)
                                     : make_smem_desc(sb + k * kstep, lbo, 1024u);
          umma_f16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));               // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);                // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                // TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m_blk * GEMM_BM + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    // fp32 row-major output (the head's logits): a thread owns a ROW of the accumulator, so a direct store instruction
    // scatters 32 x 16 bytes over 32 rows.  The pipeline's shared memory is idle once the accumulator is complete: each
    // warp turns its 32 x 32 block through a padded tile there and stores 4 full 128-byte row segments per instruction.
    const bool via_smem = (p.out_mode == 1) && (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                          (S::BAR_OFFSET >= 4 * 32 * 36 * 4);
    float* tile = reinterpret_cast<float*>(smem_dyn + (smem_base - smem_u32(smem_dyn))) + (warp - 2) * (32 * 36);
    float l1acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const float l1gs = p.fuse_l1 ? p.l1.grad_scale * (p.l1.grad_scale_dev ? *p.l1.grad_scale_dev : 1.f) : 0.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      const int n0 = n_blk * BN + c * 32;
      if (n0 >= p.N) continue;                       // uniform for the warp
      const bool full = (n0 + 32 <= p.N);
      if (via_smem && full) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(tile + lane * 36 + 4 * j) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int cl = (lane & 7) * 4;               // this lane's 4 columns of the block
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias)                                  // (the bias is a view into the flat parameter buffer: no alignment promise)
          bv = make_float4(__ldg(p.bias + n0 + cl), __ldg(p.bias + n0 + cl + 1), __ldg(p.bias + n0 + cl + 2),
                           __ldg(p.bias + n0 + cl + 3));
        if (p.fuse_l1) {
          // masked-L1 on the fly: this lane holds 4 consecutive logits of 8 rows; target / mask rows are [b, t]-major
          const L1Fuse& q = p.l1;
          const int k0 = n0 + cl;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float yv[4][4], mv[4][4], smv[4];
            int grows[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {              // the loads of four rows are in flight together
              const int rl = 4 * (4 * half + i) + (lane >> 3);
              const int grow = m_blk * GEMM_BM + quarter * 32 + rl;
              grows[i] = grow;
              const int gr = grow < p.M ? grow : p.M - 1;
              const int t = gr / q.B, b = gr - t * q.B;
              smv[i] = (t < __ldg(q.seq_len + b)) ? 1.f : 0.f;
              const long long base = ((long long)b * q.T + t) * q.F + k0;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const bool ok = k0 + e < q.F;
                yv[i][e] = ok ? __ldg(q.target + base + e) : 0.f;
                mv[i][e] = ok ? __ldg(q.mask + base + e) : 0.f;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 4 * (4 * half + i) + (lane >> 3);
              const int grow = grows[i];
              float4 v = *reinterpret_cast<const float4*>(tile + rl * 36 + cl);
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
              if (grow >= p.M) continue;
              const float xs[4] = {v.x, v.y, v.z, v.w};
              uint16_t g[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) g[e] = (k0 + e < q.F) ? l1_element(xs[e], yv[i][e], mv[i][e], smv[i], q.mode, l1gs, l1acc) : (uint16_t)0;
              uint16_t* dl = q.dlogits + (long long)grow * q.ldd + k0;
              if (k0 + 4 <= q.F) {
                *reinterpret_cast<uint2*>(dl) = make_uint2((uint32_t)g[0] | ((uint32_t)g[1] << 16), (uint32_t)g[2] | ((uint32_t)g[3] << 16));
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (k0 + e < q.F) dl[e] = g[e];
              }
              if (k0 + 4 > q.logits_from_col)          // columns another loss still reads as logits
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (long long)grow * p.ldc + k0) = v;
            }
          }
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + (lane >> 3);        // row of the block
          const int grow = m_blk * GEMM_BM + quarter * 32 + rl;
          float4 v = *reinterpret_cast<const float4*>(tile + rl * 36 + cl);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if (grow < p.M) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (long long)grow * p.ldc + n0 + cl) = v;
        }
        __syncwarp();
        continue;
      }
      if (!row_ok) continue;
      if (p.fuse_l1) {
        // partial column block (the last columns of the head): this lane owns a row
        const L1Fuse& q = p.l1;
        const int t = row / q.B, b = row - t * q.B;
        const float sm = (t < __ldg(q.seq_len + b)) ? 1.f : 0.f;
        const long long base = ((long long)b * q.T + t) * q.F;
        for (int jj = 0; jj < 32 && n0 + jj < p.N; ++jj) {
          const int k = n0 + jj;
          const float x = __uint_as_float(r[jj]) + (p.bias ? __ldg(p.bias + k) : 0.f);
          if (k < q.F)
            q.dlogits[(long long)row * q.ldd + k] = l1_element(x, __ldg(q.target + base + k), __ldg(q.mask + base + k), sm, q.mode, l1gs, l1acc);
          if (k >= q.logits_from_col) reinterpret_cast<float*>(p.C)[(long long)row * p.ldc + k] = x;
        }
        continue;
      }
      epilogue_store(p, row, n0, full, r, (int)blockIdx.y);
    }
    if (p.fuse_l1) {
      // five sums: warp reduction in double, one atomic per warp and sum into one of L1_PARTS partial rows
      double* dst = p.l1.partial + (size_t)(blockIdx.x % L1_PARTS) * 8;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const double v = warp_sum_d((double)l1acc[i]);
        if (lane == 0 && v != 0.0) atomicAdd(dst + i, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// folds the partial rows of a fused head + masked-L1 launch into sums[0..5] (sums[5] = number of loss elements)
__global__ void l1_partials_kernel(const double* __restrict__ partial, double* __restrict__ sums, double count) {
  const int i = threadIdx.x;
  if (i < 5) {
    double v = 0.0;
    for (int pth = 0; pth < L1_PARTS; ++pth) v += partial[pth * 8 + i];
    sums[i] += v;
  }
  if (i == 5) sums[5] += count;
}

// ---------------------------------------------------------------- persistent CTA-pair kernel
// The one-tile-per-CTA kernel above re-reads its operands from L2 at 128 B per tensor-core clock; the L2 delivers
// ~42 B/clk per SM (measured LTS cap, B300_MICROARCH.md), which is why it saturates near 0.8 PFLOP/s whatever the
// pipeline depth (profiles/README.md).  This kernel halves the operand traffic per flop twice over:
//   * cta_group::2: two CTAs of a cluster form one 256 x BN tile; each stages its own 128 rows of A and HALF of the
//     B tile, the tensor cores of both SMs read both halves (tcgen05.mma.cta_group::2 issued by the leader);
//   * BN = 256 columns per tile.
// and it is persistent (one cluster per SM pair walks the tile list), with the accumulator double buffered in TMEM so
// the epilogue of tile i (tcgen05.ld -> registers -> global, 4 warps per CTA) overlaps the main loop of tile i+1.
//   warp 0  TMA producer (both CTAs; completion is signalled on the LEADER's `full` barrier)
//   warp 1  leader CTA only: tcgen05.mma issue; tcgen05.commit multicast frees the smem slot in BOTH CTAs
//   warps 2..5  epilogue of the CTA's own 128 accumulator rows
constexpr int G2_EPI_WARPS = 4;                  // one warp per TMEM lane quarter (8 warps were measured slower on the dW / dX shapes)
constexpr int G2_THREADS = (2 + G2_EPI_WARPS) * 32;

// WIDE: one 256 x 512 tile per CTA pair (two N = 256 MMAs per k-step into one 512-column accumulator) instead of
// 256 x 256 with a double-buffered accumulator.  The operand bytes per flop drop by a quarter -- the CTA-pair kernel is
// bound by the ~6300 B/clk the L2 feeds the SMs' TMA units chip-wide (tensor pipe 57 % active with 256 x 256 tiles) --
// at the price of an epilogue that no longer overlaps the next tile's main loop: chosen for the long-K shapes (dW).
template <int BN, bool WIDE>
struct Gemm2Smem {
  static constexpr int TN = WIDE ? 2 * BN : BN;                  // tile width
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // 16 KB: this CTA's 128 rows
  static constexpr int BH_BYTES = (BN / 2) * GEMM_BK * 2;        // this CTA's half of one N = BN operand
  static constexpr int B_BYTES = (TN / BN) * BH_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = WIDE ? 4 : ((BN == 256) ? 6 : 8);
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

template <int BN, bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_f16_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using S = Gemm2Smem<BN, WIDE>;
  constexpr int STAGES = S::STAGES;
  constexpr int TN = S::TN, NH = TN / BN;                  // tile width, N = BN operands per k-step
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader CTA
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };       // per CTA
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };  // used in the leader CTA
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (p.N + TN - 1) / TN;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int chunk = (kb_total + p.split_k - 1) / p.split_k;
  const int n_work = m_tiles * n_tiles * p.split_k;        // work item = (k split, m tile, n tile), n fastest

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);                            // one arrive.expect_tx per CTA of the pair
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * G2_EPI_WARPS);           // epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // (An L2 prefetch cursor running 4..32 k-blocks ahead of these loads -- cp.async.bulk.prefetch.tensor -- was built and
    // measured: 10-50 % SLOWER on every shape of the step, profiles/README.md; the ring is not starved by DRAM latency.)
    if (lane == 0) {
      int it = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters) {
        const int n_blk = w % n_tiles, m_blk = (w / n_tiles) % m_tiles, ks = w / (n_tiles * m_tiles);
        const int kb0 = ks * chunk, kb1 = min(kb_total, kb0 + chunk);
        const int m0 = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM;       // this CTA's A rows
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * S::STAGE_BYTES;
          const uint32_t lead_full = map_to_cta(full_bar(s), 0);
          mbar_expect_tx_cluster(lead_full, S::STAGE_BYTES);
          const int k0 = kb * GEMM_BK;
          if (p.trans == 0) {
            if (p.a_il) tma_load_3d_2sm(sa, &tmA, lead_full, 0, m0 / 32, k0 >> 3);
            else tma_load_2d_2sm(sa, &tmA, lead_full, k0, m0);
          } else {
            if (p.a_il) tma_load_3d_2sm(sa, &tmA, lead_full, 0, k0 >> 5, m0 / 8);
            else {
#pragma unroll
              for (int a = 0; a < GEMM_BM / 64; ++a) tma_load_2d_2sm(sa + a * 8192, &tmA, lead_full, m0 + 64 * a, k0);
            }
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {                                 // this CTA's half of each N = BN operand
            const uint32_t sb = sa + S::A_BYTES + h * S::BH_BYTES;
            const int n0 = n_blk * TN + h * BN + (int)rank * (BN / 2);
            if (p.trans == 0) {
              tma_load_2d_2sm(sb, &tmB, lead_full, k0, n0);
            } else if (p.b_il) {
              tma_load_3d_2sm(sb, &tmB, lead_full, 0, k0 >> 5, n0 / 8);
            } else {
#pragma unroll
              for (int b = 0; b < BN / 128; ++b) tma_load_2d_2sm(sb + b * 8192, &tmB, lead_full, n0 + 64 * b, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(2 * GEMM_BM, BN, p.trans, p.trans);
      const uint32_t lbo = p.trans ? 8192u : 0u;
      const uint32_t kstep = p.trans ? 2048u : 32u;
      const uint32_t a_lbo = p.trans ? 128u : 2048u, a_sbo = p.trans ? 1024u : 128u;
      const uint32_t a_kstep = p.trans ? 256u : 4096u;
      int it = 0, tile = 0;
      for (int w = cluster_id; w < n_work; w += n_clusters, ++tile) {
        const int ks = w / (n_tiles * m_tiles);
        const int kb0 = ks * chunk, kb1 = min(kb_total, kb0 + chunk);
        const int acc = WIDE ? 0 : (tile & 1);                           // WIDE: one 512-column accumulator, no double buffer
        mbar_wait(tempty_bar(acc), ((WIDE ? tile : (tile >> 1)) & 1) ^ 1);   // both CTAs' epilogues drained this buffer
        tc_fence_after();
        const uint32_t td = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * S::STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = p.a_il ? make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, 0u)
                                       : make_smem_desc(sa + k * kstep, lbo, 1024u);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const uint32_t sb = sa + S::A_BYTES + h * S::BH_BYTES;
              const uint64_t db = p.b_il ? make_smem_desc(sb + k * a_kstep, a_lbo, a_sbo, 0u)
                                         : make_smem_desc(sb + k * kstep, lbo, 1024u);
              umma_f16_2sm(td + (uint32_t)(h * BN), da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit_2sm(empty_bar(s), (uint16_t)3);                    // frees the slot in both CTAs
        }
        umma_commit_2sm(tfull_bar(acc), (uint16_t)3);                    // accumulator complete (both CTAs)
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;                       // TMEM lane quarter a warp may read = warp id mod 4
    constexpr int NSPLIT = G2_EPI_WARPS / 4;            // warps per lane quarter
    const int chalf = (warp - 2) >> 2;                  // which share of the tile's columns this warp converts
    const uint32_t lead_tempty0 = map_to_cta(tempty_bar(0), 0);
    int tile = 0;
    for (int w = cluster_id; w < n_work; w += n_clusters, ++tile) {
      const int n_blk = w % n_tiles, m_blk = (w / n_tiles) % m_tiles, ks = w / (n_tiles * m_tiles);
      const int kb0 = ks * chunk, kb1 = min(kb_total, kb0 + chunk);
      const int acc = WIDE ? 0 : (tile & 1);
      mbar_wait(tfull_bar(acc), (WIDE ? tile : (tile >> 1)) & 1);
      tc_fence_after();
      const int row = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M && kb1 > kb0;
#pragma unroll 1
      for (int c = chalf * (TN / 32 / NSPLIT); c < (chalf + 1) * (TN / 32 / NSPLIT); ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        tmem_ld_wait();
        const int n0 = n_blk * TN + c * 32;
        if (!row_ok || n0 >= p.N) continue;
        epilogue_store(p, row, n0, n0 + 32 <= p.N, r, ks);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(lead_tempty0 + 8u * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------- A-stationary CTA-pair kernel (K <= 512, trans = 0)
// The gate projections are [M, K<=512] x [2048, K]^T: with 256 x 256 tiles every tile loads 256 KB of A and 256 KB of B for
// 67 MFLOP, and the L2 -> SM path (~6300 B/clk chip-wide) keeps the tensor pipe at 57 %.  Here a CTA pair keeps its 256
// rows of A (the whole K, 8 k-slices of 16 KB per CTA) in shared memory while it sweeps ALL n tiles of that row block:
// only B streams (4-stage ring), A costs 1/n_tiles of what it did -- 44 % less operand traffic at N = 2048.
// Work item = one 256-row block; per item n_tiles output tiles, accumulators double-buffered in TMEM as before.
// A k-slice is released (a_empty) by the LAST n tile's MMAs of that slice, so the next row block's slice reloads under
// the rest of that tile's main loop.
struct GemmAStatSmem {
  static constexpr int KS_MAX = 8;                               // k-slices of 64: K <= 512
  static constexpr int A_SLICE = GEMM_BM * GEMM_BK * 2;          // 16 KB: this CTA's 128 rows x 64 k
  static constexpr int B_BYTES = 128 * GEMM_BK * 2;              // this CTA's half of the 256-wide B tile
  static constexpr int STAGES = 4;
  static constexpr int B_OFFSET = KS_MAX * A_SLICE;              // 128 KB
  static constexpr int BAR_OFFSET = B_OFFSET + STAGES * B_BYTES; // + 64 KB
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_f16_2sm_astat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using S = GemmAStatSmem;
  constexpr int STAGES = S::STAGES, BN = 256, KS = S::KS_MAX;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  auto bfull_bar = [&](int s) { return bar_base + 8u * s; };                          // leader CTA
  auto bempty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };              // per CTA
  auto afull_bar = [&](int k) { return bar_base + 8u * (2 * STAGES + k); };           // leader CTA
  auto aempty_bar = [&](int k) { return bar_base + 8u * (2 * STAGES + KS + k); };     // per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 * KS + a); };  // per CTA
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 * KS + 2 + a); };   // leader CTA
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * KS + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kbt = (p.K + GEMM_BK - 1) / GEMM_BK;           // <= KS

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bfull_bar(s), 2);
      mbar_init(bempty_bar(s), 1);
    }
    for (int k = 0; k < KS; ++k) {
      mbar_init(afull_bar(k), 2);
      mbar_init(aempty_bar(k), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * G2_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int it = 0, mi = 0;
      for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters, ++mi) {
        const int m0 = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM;
        for (int n_blk = 0; n_blk < n_tiles; ++n_blk) {
          const int n0 = n_blk * BN + (int)rank * (BN / 2);
          for (int kb = 0; kb < kbt; ++kb, ++it) {
            const int k0 = kb * GEMM_BK;
            if (n_blk == 0) {                                            // this row block's A slice kb
              mbar_wait(aempty_bar(kb), (uint32_t)((mi & 1) ^ 1));
              const uint32_t lead_afull = map_to_cta(afull_bar(kb), 0);
              mbar_expect_tx_cluster(lead_afull, S::A_SLICE);
              const uint32_t sa = smem_base + kb * S::A_SLICE;
              if (p.a_il) tma_load_3d_2sm(sa, &tmA, lead_afull, 0, m0 / 32, k0 >> 3);
              else tma_load_2d_2sm(sa, &tmA, lead_afull, k0, m0);
            }
            const int s = it % STAGES;
            mbar_wait(bempty_bar(s), (uint32_t)(((it / STAGES) & 1) ^ 1));
            const uint32_t lead_bfull = map_to_cta(bfull_bar(s), 0);
            mbar_expect_tx_cluster(lead_bfull, S::B_BYTES);
            tma_load_2d_2sm(smem_base + S::B_OFFSET + s * S::B_BYTES, &tmB, lead_bfull, k0, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(2 * GEMM_BM, BN, 0, 0);
      int it = 0, mi = 0, tile = 0;
      for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters, ++mi) {
        for (int n_blk = 0; n_blk < n_tiles; ++n_blk, ++tile) {
          const int acc = tile & 1;
          mbar_wait(tempty_bar(acc), (uint32_t)(((tile >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t td = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < kbt; ++kb, ++it) {
            if (n_blk == 0) mbar_wait(afull_bar(kb), (uint32_t)(mi & 1));
            const int s = it % STAGES;
            mbar_wait(bfull_bar(s), (uint32_t)((it / STAGES) & 1));
            tc_fence_after();
            const uint32_t sa = smem_base + kb * S::A_SLICE;
            const uint32_t sb = smem_base + S::B_OFFSET + s * S::B_BYTES;
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              const uint64_t da = p.a_il ? make_smem_desc(sa + k * 4096u, 2048u, 128u, 0u) : make_smem_desc(sa + k * 32u, 0u, 1024u);
              const uint64_t db = make_smem_desc(sb + k * 32u, 0u, 1024u);
              umma_f16_2sm(td, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(bempty_bar(s), (uint16_t)3);
            if (n_blk == n_tiles - 1) umma_commit_2sm(aempty_bar(kb), (uint16_t)3);   // slice free for the next row block
          }
          umma_commit_2sm(tfull_bar(acc), (uint16_t)3);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    constexpr int NSPLIT = G2_EPI_WARPS / 4;
    const int chalf = (warp - 2) >> 2;
    const uint32_t lead_tempty0 = map_to_cta(tempty_bar(0), 0);
    int tile = 0;
    for (int m_blk = cluster_id; m_blk < m_tiles; m_blk += n_clusters) {
      const int row = m_blk * 2 * GEMM_BM + (int)rank * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      for (int n_blk = 0; n_blk < n_tiles; ++n_blk, ++tile) {
        const int acc = tile & 1;
        mbar_wait(tfull_bar(acc), (uint32_t)((tile >> 1) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf * (BN / 32 / NSPLIT); c < (chalf + 1) * (BN / 32 / NSPLIT); ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
          tmem_ld_wait();
          const int n0 = n_blk * BN + c * 32;
          if (!row_ok || n0 >= p.N) continue;
          epilogue_store(p, row, n0, n0 + 32 <= p.N, r);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(lead_tempty0 + 8u * acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------- debug / validation GEMM (CUDA cores)
#ifdef AVSI_DEBUG_KERNELS   // CUDA-core bisecting kernel: compiled only into debug builds (nvcc -DAVSI_DEBUG_KERNELS), never into the product library
// Same contract as the tensor-core kernel; selected only by AVSI_GEMM_DEBUG_SIMT=1 to bisect
// failures.  Never used silently.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const uint16_t* __restrict__ A, int lda, const uint16_t* __restrict__ B, int ldb, GemmParams p) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    // sa[ty][tx] = A(row, k0+tx) ; sb[ty][tx] = B(col_of_ty.., k0+tx)
    int ar = blockIdx.y * 16 + ty, ak = k0 + tx;
    int bn = blockIdx.x * 16 + ty, bk = k0 + tx;
    float av = 0.f, bv = 0.f;
    if (ar < p.M && ak < p.K)
      av = __half2float(__ushort_as_half(p.trans ? A[(long long)ak * lda + ar] : A[(long long)ar * lda + ak]));
    if (bn < p.N && bk < p.K)
      bv = __half2float(__ushort_as_half(p.trans ? B[(long long)bk * ldb + bn] : B[(long long)bn * ldb + bk]));
    sa[ty][tx] = av;
    sb[ty][tx] = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (row < p.M && col < p.N) {
    if (p.out_mode == 0)
      reinterpret_cast<uint16_t*>(p.C)[(long long)row * p.ldc + col] = __half_as_ushort(__float2half_rn(acc));
    else if (p.out_mode == 1)
      reinterpret_cast<float*>(p.C)[(long long)row * p.ldc + col] = acc + (p.bias ? p.bias[col] : 0.f);
    else
      atomicAdd(reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col, acc);
  }
}

// ---------------------------------------------------------------- host side
#endif  // AVSI_DEBUG_KERNELS

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<uint64_t>()(k.d0 * 31 + k.d1);
    h = h * 1000003u ^ std::hash<uint64_t>()(k.ld * 131 + k.b0 * 7 + k.b1);
    return h;
  }
};

// 2-D fp16 tensor map: dims {d0 (contiguous), d1}, row pitch ld elements, box {b0, b1}, 128B swizzle.
static int get_tmap(const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, d0, d1, ld, b0, b1};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return AVSI_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(AVSI_ERR_CUDA, "%s: cuTensorMapEncodeTiled entry point unavailable%s", "get_tmap");
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return set_error(AVSI_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed, CUresult %s", "get_tmap", buf);
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return AVSI_OK;
}

// 3-D tensor map over an interleaved fp16 matrix [rows x ld]: dims {256 (one 32-row x 8-col block, contiguous),
// row blocks, column chunks}; box {256, brb, bcc}; no swizzle.  Rows / columns beyond the extents read as zero.
int get_tmap_il(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t brb, uint32_t bcc,
                CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, brb, bcc};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return AVSI_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(AVSI_ERR_CUDA, "%s: cuTensorMapEncodeTiled entry point unavailable%s", "get_tmap_il");
  cuuint64_t dims[3] = {256, (rows + 31) / 32, (cols + 7) / 8};
  cuuint64_t strides[2] = {(ld / 8) * 512, 512};
  cuuint32_t box[3] = {256, brb, bcc};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return set_error(AVSI_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed, CUresult %s", "get_tmap_il", buf);
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return AVSI_OK;
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using S = GemmSmem<BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(gemm_f16_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_done = true;
  }
  dim3 grid(((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN), p.split_k);
  gemm_f16_kernel<BN, STAGES><<<grid, GEMM_THREADS, S::TOTAL, st>>>(ta, tb, p);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

template <int BN, bool WIDE>
static int launch_gemm_2sm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using S = Gemm2Smem<BN, WIDE>;
  static bool attr_done = false;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(gemm_f16_2sm_kernel<BN, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_done = true;
  }
  const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM), n_tiles = (p.N + S::TN - 1) / S::TN;
  const long long n_work = (long long)m_tiles * n_tiles * p.split_k;
  long long clusters = num_sms() / 2;
  if (clusters > n_work) clusters = n_work;
  gemm_f16_2sm_kernel<BN, WIDE><<<(unsigned)(2 * clusters), G2_THREADS, S::TOTAL, st>>>(ta, tb, p);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

static int launch_gemm_2sm_astat(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using S = GemmAStatSmem;
  static bool attr_done = false;
  if (!attr_done) {
    AVSI_CUDA(cudaFuncSetAttribute(gemm_f16_2sm_astat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_done = true;
  }
  const long long m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  long long clusters = num_sms() / 2;
  if (clusters > m_tiles) clusters = m_tiles;
  gemm_f16_2sm_astat_kernel<<<(unsigned)(2 * clusters), G2_THREADS, S::TOTAL, st>>>(ta, tb, p);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

}  // namespace avsi

// Which kernel serves a problem, and with how many k splits: shared by avsi_gemm_f16 and avsi_gemm_f16_scratch_bytes.
namespace avsi {
struct GemmPlan {
  int kind;      // 0 one-tile kernel (bn), 1 CTA pairs 256 x 256, 2 CTA pairs 256 x 512, 3 A-stationary CTA pairs
  int bn, split_k;
};
static GemmPlan plan_gemm(int M, int N, int K, int trans, int out_mode, int split_k) {
  GemmPlan g{0, 128, split_k};
  // tile width: 128 (3 stages, 2 CTAs/SM) by default; 256 (4 stages, 1 CTA/SM) for long-K problems
  // where the main loop dominates; 64 for narrow outputs.  AVSI_GEMM_BN overrides (tuning only).
  AVSI_ENV_CACHE(bn_env, env_int("AVSI_GEMM_BN", 0));
  // large problems: persistent CTA-pair kernel (256 x 256 tiles, cta_group::2).  AVSI_GEMM_2SM=0 disables it.
  AVSI_ENV_CACHE(use_2sm, env_is("AVSI_GEMM_2SM", "0") ? 0 : (env_is("AVSI_GEMM_2SM", "2") ? 2 : 1));
  if (use_2sm && N >= 256 && (use_2sm == 2 || ((N + 255) / 256) * 256 * 3 <= N * 4) && (long long)M * N * K >= (1LL << 29)) {
    // 256 x 512 tiles (one accumulator, a quarter less operand traffic per flop) where the main loop is long enough to
    // carry the un-overlapped epilogue and the output is at least 3/4 of a 512-wide tile: the split-K dW shapes, and
    // fixed-K shapes with K >= 2048 (dX).  AVSI_GEMM_WIDE = 0 never, 2 whenever N allows (A/B runs).
    AVSI_ENV_CACHE(wide_env, env_int("AVSI_GEMM_WIDE", 1));
    const bool n_fits = (N % 512 == 0) || (N % 512 >= 384);
    const bool wide = wide_env && N >= 384 && n_fits && (wide_env == 2 || out_mode == 2 || K >= 2048);
    // split-K only as far as needed to give every SM pair a work item
    const int tn = wide ? 512 : 256;
    const int m_t = (M + 255) / 256, n_t = (N + tn - 1) / tn, kbt = (K + GEMM_BK - 1) / GEMM_BK;
    if (out_mode == 2) {
      // split K so that the work items fill whole waves of SM pairs (80 items on 74 pairs would take two waves)
      const int pairs = num_sms() / 2, tiles = m_t * n_t;
      int best = 1;
      double best_eff = 0.0;
      for (int sk = 1; sk <= kbt && sk * tiles <= 4 * pairs; ++sk) {
        const int work = sk * tiles, waves = (work + pairs - 1) / pairs;
        const double eff = (double)work / ((double)waves * pairs);
        if (eff > best_eff + 1e-9) {
          best_eff = eff;
          best = sk;
        }
      }
      g.split_k = best;
    }
    // A-stationary sweep for the projection shape: trans = 0, the whole K in 8 k-slices, several n tiles to share A,
    // enough row blocks to fill the chip.  AVSI_GEMM_ASTAT=0 disables it (A/B runs).
    AVSI_ENV_CACHE(astat_env, env_int("AVSI_GEMM_ASTAT", 1));
    if (astat_env && trans == 0 && out_mode != 2 && K <= GemmAStatSmem::KS_MAX * GEMM_BK && n_t >= 4 && m_t >= num_sms()) g.kind = 3;
    else g.kind = wide ? 2 : 1;
    g.bn = 256;
    return g;
  }
  int bn = 128;
  if (N > 128 && K >= 4096) bn = 256;
  if (N <= 64) bn = 64;
  if (bn_env == 64 || bn_env == 128 || bn_env == 256) bn = bn_env;
  g.bn = bn;
  return g;
}
// k splits that get at least one k block (the others leave their slice of the ordered partials unwritten)
static int live_splits(int K, int split_k) {
  const int kbt = (K + GEMM_BK - 1) / GEMM_BK, chunk = (kbt + split_k - 1) / split_k;
  return (kbt + chunk - 1) / chunk;
}
}  // namespace avsi

// Bytes of reduction scratch (avsi_set_reduce_scratch) with which this out_mode-2 problem runs its split-K sum in a fixed
// order; 0 when it does not split K.
extern "C" int64_t avsi_gemm_f16_scratch_bytes(int M, int N, int K, int trans, int out_mode, int split_k) {
  using namespace avsi;
  if (M <= 0 || N <= 0 || K <= 0 || out_mode != 2 || split_k < 1) return 0;
  const GemmPlan g = plan_gemm(M, N, K, trans, out_mode, split_k);
  if (g.split_k <= 1) return 0;
  return REDUCE_COUNTER_BYTES + (int64_t)live_splits(K, g.split_k) * M * ((N + 3) & ~3) * (int64_t)sizeof(float);
}

extern "C" int avsi_gemm_f16(const uint16_t* A, int lda, const uint16_t* B, int ldb, void* C, int ldc,
                             const float* bias, int M, int N, int K, int trans, int out_mode, int split_k,
                             int layout, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(A && B && C, "null pointer");
  AVSI_REQUIRE(M > 0 && N > 0 && K > 0, "M,N,K > 0");
  AVSI_REQUIRE(trans == 0 || trans == 1, "trans");
  AVSI_REQUIRE(out_mode >= 0 && out_mode <= 2, "out_mode");
  AVSI_REQUIRE(split_k >= 1 && (split_k == 1 || out_mode == 2), "split_k > 1 needs out_mode 2");
  AVSI_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "lda/ldb multiples of 8");
  AVSI_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "A/B 16-byte aligned");
  AVSI_REQUIRE(ldc >= N, "ldc >= N");
  AVSI_REQUIRE(layout >= 0 && layout <= 7, "layout");
  AVSI_REQUIRE(!(layout & 2) || (out_mode == 0 && ldc % 8 == 0), "interleaved output needs out_mode 0 and ldc % 8 == 0");
  AVSI_REQUIRE(!(layout & 4) || (trans == 1 && ldb % 8 == 0), "interleaved B needs trans = 1 and ldb % 8 == 0");
  GemmParams p{C, bias, ldc, M, N, K, trans, out_mode, split_k, layout & 1, (layout >> 1) & 1, (layout >> 2) & 1};
  cudaStream_t st = (cudaStream_t)stream;

#ifdef AVSI_DEBUG_KERNELS
  AVSI_ENV_CACHE(debug_simt, env_is("AVSI_GEMM_DEBUG_SIMT", "1"));
  if (debug_simt && layout == 0) {
    GemmParams q = p;
    q.split_k = 1;
    dim3 grid((N + 15) / 16, (M + 15) / 16);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, q);
    AVSI_LAUNCH_CHECK();
    return AVSI_OK;
  }
#endif

  const GemmPlan g = plan_gemm(M, N, K, trans, out_mode, split_k);
  p.split_k = g.split_k;
  // split-K sums in a fixed order when the reduction scratch is registered (avsi_set_reduce_scratch): every split
  // writes its product to its own slice, splitk_reduce_kernel adds the slices to C.  Without scratch: fp32 atomics.
  int splits_live = 0;
  if (out_mode == 2 && p.split_k > 1) {
    const ReduceScratch rs = reduce_scratch();
    if (rs.counters) {
      splits_live = live_splits(K, p.split_k);
      p.part_ld = (N + 3) & ~3;
      AVSI_REQUIRE((long long)splits_live * M * p.part_ld * (long long)sizeof(float) <= rs.area_bytes,
                   "reduce scratch too small (avsi_gemm_f16_scratch_bytes)");
      p.part = reinterpret_cast<float*>(rs.area);
    }
  }
  auto finish = [&](int rc) -> int {
    if (rc != AVSI_OK || !p.part) return rc;
    const long long vecs = (long long)M * (p.part_ld >> 2);
    splitk_reduce_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, st>>>(p.part, splits_live, M, N, p.part_ld,
                                                                        reinterpret_cast<float*>(C), ldc);
    AVSI_LAUNCH_CHECK();
    return AVSI_OK;
  };

  if (g.kind != 0) {
    CUtensorMap ta2, tb2;
    int rc2;
    if (trans == 0) {
      rc2 = (layout & 1) ? get_tmap_il(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM / 32, GEMM_BK / 8, &ta2)
                         : get_tmap(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM, &ta2);
      if (rc2) return rc2;
      rc2 = get_tmap(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK, 128, &tb2);
      if (rc2) return rc2;
    } else {
      rc2 = (layout & 1) ? get_tmap_il(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK / 32, GEMM_BM / 8, &ta2)
                         : get_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BK, &ta2);
      if (rc2) return rc2;
      rc2 = (layout & 4) ? get_tmap_il(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK / 32, 128 / 8, &tb2)
                         : get_tmap(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BK, &tb2);
      if (rc2) return rc2;
    }
    if (g.kind == 3) return launch_gemm_2sm_astat(ta2, tb2, p, st);
    if (g.kind == 2) return finish(launch_gemm_2sm<256, true>(ta2, tb2, p, st));
    return finish(launch_gemm_2sm<256, false>(ta2, tb2, p, st));
  }
  const int bn = g.bn;
  CUtensorMap ta, tb;
  int rc;
  if (trans == 0) {
    rc = (layout & 1) ? get_tmap_il(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM / 32, GEMM_BK / 8, &ta)
                      : get_tmap(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM, &ta);
    if (rc) return rc;
    rc = get_tmap(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK, (uint32_t)bn, &tb);
    if (rc) return rc;
  } else {
    rc = (layout & 1) ? get_tmap_il(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK / 32, GEMM_BM / 8, &ta)
                      : get_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, GEMM_BK, &ta);
    if (rc) return rc;
    rc = (layout & 4) ? get_tmap_il(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK / 32, (uint32_t)bn / 8, &tb)
                      : get_tmap(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, GEMM_BK, &tb);
    if (rc) return rc;
  }
  if (bn == 256) return finish(launch_gemm<256, 4>(ta, tb, p, st));
  if (bn == 128) return finish(launch_gemm<128, 3>(ta, tb, p, st));
  return finish(launch_gemm<64, 4>(ta, tb, p, st));
}

// Head GEMM with the masked-L1 loss and its gradient in the epilogue (training step): logits = A[M,K] . W[N,K]^T + bias are
// consumed where they are produced -- prediction, |target - prediction| sums and the fp16 gradient of avsi_masked_l1 -- and
// only the columns >= logits_from_col (the phone-recognition head of the MTL models, read by avsi_ctc_loss) are written.
extern "C" int avsi_head_l1(const uint16_t* A, int lda, int a_layout, const uint16_t* W, int ldw, const float* bias, int M, int N,
                            int K, float* logits, int ldc, int logits_from_col, const float* target, const float* mask,
                            const int32_t* seq_len, int B, int T, int F, int mode, float grad_scale,
                            const float* grad_scale_dev, double* sums, double* partial_ws, uint16_t* dlogits, int ldd,
                            void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(A && W && logits && target && mask && seq_len && sums && partial_ws && dlogits, "null pointer");
  AVSI_REQUIRE(M > 0 && N > 0 && K > 0 && B > 0 && T > 0 && (long long)B * T == M, "M = B * T");
  AVSI_REQUIRE(F > 0 && F <= N && ldd >= F && ldd % 4 == 0, "F <= N, ldd >= F, ldd % 4 == 0");
  AVSI_REQUIRE(mode == 0 || mode == 1, "mode");
  AVSI_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "operand alignment");
  AVSI_REQUIRE(ldc >= N && ldc % 4 == 0 && ((uintptr_t)logits % 16 == 0) && ((uintptr_t)dlogits % 8 == 0), "output alignment");
  AVSI_REQUIRE(a_layout == 0 || a_layout == 1, "a_layout");
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p{logits, bias, ldc, M, N, K, 0, 1, 1, a_layout, 0, 0};
  p.fuse_l1 = 1;
  p.l1 = L1Fuse{target, mask, seq_len, grad_scale_dev, partial_ws, dlogits, B, T, F, mode, ldd, logits_from_col, grad_scale};
  AVSI_CUDA(cudaMemsetAsync(partial_ws, 0, sizeof(double) * L1_PARTS * 8, st));
  CUtensorMap ta, tb;
  int rc = a_layout ? get_tmap_il(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM / 32, GEMM_BK / 8, &ta)
                    : get_tmap(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, GEMM_BK, 128, &tb);
  if (rc) return rc;
  rc = launch_gemm<128, 3>(ta, tb, p, st);
  if (rc) return rc;
  l1_partials_kernel<<<1, 32, 0, st>>>(partial_ws, sums, (double)M * (double)F);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_head_l1_workspace_bytes(void) { return (int)(sizeof(double) * avsi::L1_PARTS * 8); }
