// Masked-L1 loss + gradient (one pass) and fp16 column sums (bias gradients).
//
// avsi_masked_l1 replaces prediction / loss of models.py:127-159 (SI) and
// models.py:1920-1963 (MTL) together with their autodiff mirror (about ten element-wise
// and three reduction kernels in the TF graph).
#include "common.cuh"

namespace avsi {

constexpr int L1_THREADS = 256;
constexpr int L1_UNROLL = 4;

__global__ void __launch_bounds__(L1_THREADS)
masked_l1_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ target,
                 const float* __restrict__ mask, const int32_t* __restrict__ seq_len, int B, int T, int F,
                 int mode, float grad_scale_h, const float* __restrict__ grad_scale_dev, double* __restrict__ sums, float* __restrict__ prediction,
                 uint16_t* __restrict__ dlogits, int ldd, double* __restrict__ partials, unsigned* __restrict__ ticket) {
  // one warp per (b,t) row, grid-stride
  const int lane = threadIdx.x & 31;
  const int warps_per_block = L1_THREADS / 32;
  const float grad_scale = grad_scale_h * (grad_scale_dev ? *grad_scale_dev : 1.f);
  const long long rows = (long long)B * T;
  float s_hole = 0.f, n_hole = 0.f, s_valid = 0.f, n_valid = 0.f, s_all = 0.f;
  for (long long r = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(r / T);
    const int t = (int)(r - (long long)b * T);
    const float sm = (t < seq_len[b]) ? 1.f : 0.f;
    const float* lg = logits + ((long long)t * B + b) * ldl;
    const float* tg = target + r * F;
    const float* mk = mask + r * F;
    uint16_t* dl = dlogits ? dlogits + ((long long)t * B + b) * ldd : nullptr;
    for (int k0 = lane; k0 < F; k0 += 32 * L1_UNROLL) {
     // the loads of L1_UNROLL column groups are requested before the first is used (12 lines in flight per warp)
     float xv[L1_UNROLL], yv[L1_UNROLL], mv[L1_UNROLL];
#pragma unroll
     for (int u = 0; u < L1_UNROLL; ++u) {
       const int k = k0 + 32 * u;
       const bool ok = k < F;
       xv[u] = ok ? __ldg(lg + k) : 0.f;
       yv[u] = ok ? __ldg(tg + k) : 0.f;
       mv[u] = ok ? __ldg(mk + k) : 0.f;
     }
#pragma unroll
     for (int u = 0; u < L1_UNROLL; ++u) {
      const int k = k0 + 32 * u;
      if (k >= F) break;
      const float x = xv[u], y = yv[u], m = mv[u];
      float pred = (mode == 0) ? x : (y * m + x * (1.f - m));
      pred *= sm;
      const float d = y - pred;
      const float ad = fabsf(d);
      s_hole += ad * (1.f - m);
      n_hole += (1.f - m);
      s_valid += ad * m;
      n_valid += m;
      s_all += ad;
      if (prediction) prediction[r * F + k] = pred;
      if (dl) {
        // d|y - pred| / dx = -sign(y - pred) * dpred/dx
        float sg = (d > 0.f) ? -1.f : ((d < 0.f) ? 1.f : 0.f);
        float w = (mode == 0) ? sm : sm * (1.f - m);
        dl[k] = __half_as_ushort(__float2half_rn(grad_scale * sg * w));
      }
     }
    }
  }
  __shared__ double red[5][L1_THREADS / 32];
  float vals[5] = {s_hole, n_hole, s_valid, n_valid, s_all};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double v = warp_sum_d((double)vals[i]);
    if (lane == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int w = 0; w < warps_per_block; ++w) v += red[threadIdx.x][w];
    if (partials) partials[(long long)blockIdx.x * 5 + threadIdx.x] = v;
    else atomicAdd(sums + threadIdx.x, v);
  }
  if (!partials) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(sums + 5, (double)rows * F);
    return;
  }
  // ordered tail (bit-reproducible sums): the block that draws the last ticket adds all blocks' partials in a fixed order
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double v = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += L1_THREADS) v += __ldcg(partials + (long long)b * 5 + i);
    v = warp_sum_d(v);
    if (lane == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int w = 0; w < warps_per_block; ++w) v += red[threadIdx.x][w];
    sums[threadIdx.x] += v;
  }
  if (threadIdx.x == 0) {
    sums[5] += (double)rows * F;
    *ticket = 0u;
  }
}

// Column sums of an f16 matrix (bias gradients): a thread owns 8 consecutive columns (one 16-byte load per row) and every
// 8th ... row of its block's slab, so a warp reads whole 512-byte row segments; fp32 partials meet in shared memory and
// leave as one atomicAdd per column and block.  HBM-bound: the matrix is read once.
constexpr int CS_VECS = 64;                       // column vectors (x 8 columns) per block
constexpr int CS_ROWS = 4;                        // row lanes per block (256 threads)

__global__ void __launch_bounds__(CS_VECS * CS_ROWS)
colsum_f16_kernel(const uint16_t* __restrict__ X, int ldx, int rows, int col0, int ncols,
                  float* __restrict__ out, float* __restrict__ partials, unsigned* __restrict__ tickets) {
  const int v = threadIdx.x % CS_VECS, rl = threadIdx.x / CS_VECS;
  const int c = (blockIdx.x * CS_VECS + v) * 8;                     // first of this thread's 8 columns (relative to col0)
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const long long rows_per = ((long long)rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rows_per, r1 = min((long long)rows, r0 + rows_per);
  if (c < ncols) {
    const bool vec = (c + 8 <= ncols) && (((col0 + c) & 7) == 0) && ((ldx & 7) == 0);
    for (long long r = r0 + rl; r < r1; r += CS_ROWS) {
      const uint16_t* src = X + r * ldx + col0 + c;
      if (vec) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
        const float2 a = unpack_half2(q.x), b2 = unpack_half2(q.y), c2 = unpack_half2(q.z), d = unpack_half2(q.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b2.x; acc[3] += b2.y;
        acc[4] += c2.x; acc[5] += c2.y; acc[6] += d.x; acc[7] += d.y;
      } else {
        for (int i = 0; i < 8 && c + i < ncols; ++i) acc[i] += __half2float(__ushort_as_half(__ldg(src + i)));
      }
    }
  }
  __shared__ float red[CS_ROWS][CS_VECS][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][v][i] = acc[i];
  __syncthreads();
  if (!partials) {
    if (rl == 0 && c < ncols && r1 > r0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (c + i < ncols) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < CS_ROWS; ++k) t += red[k][v][i];
          atomicAdd(out + c + i, t);
        }
      }
    }
    return;
  }
  // ordered tail (bit-reproducible): row slab blockIdx.y's sums go to row blockIdx.y of the partial matrix; the block of
  // this column block that draws the last ticket adds the slabs in slab order
  const int pcols = gridDim.x * CS_VECS * 8;
  if (rl == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < CS_ROWS; ++k) t += red[k][v][i];
      partials[(long long)blockIdx.y * pcols + blockIdx.x * CS_VECS * 8 + v * 8 + i] = t;
    }
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int y = rl; y < (int)gridDim.y; y += CS_ROWS) {
    const float4* src = reinterpret_cast<const float4*>(partials + (long long)y * pcols + blockIdx.x * CS_VECS * 8 + v * 8);
    const float4 a = __ldcg(src), b = __ldcg(src + 1);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][v][i] = acc[i];
  __syncthreads();
  if (rl == 0 && c < ncols) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c + i < ncols) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < CS_ROWS; ++k) t += red[k][v][i];
        out[c + i] += t;
      }
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
}

}  // namespace avsi

extern "C" int avsi_masked_l1(const float* logits, int ldl, const float* target, const float* mask,
                              const int32_t* seq_len, int B, int T, int F, int mode, float grad_scale,
                              const float* grad_scale_dev, double* sums, float* prediction, uint16_t* dlogits,
                              int ldd, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(logits && target && mask && seq_len && sums, "null pointer");
  AVSI_REQUIRE(B > 0 && T > 0 && F > 0 && ldl >= F, "sizes");
  AVSI_REQUIRE(mode == 0 || mode == 1, "mode");
  AVSI_REQUIRE(!dlogits || ldd >= F, "ldd");
  long long rows = (long long)B * T;
  int blocks = (int)min((rows + 7) / 8, (long long)num_sms() * 8);
  // ordered sums when the reduction scratch is registered (avsi_set_reduce_scratch), atomics otherwise
  const ReduceScratch rs = reduce_scratch();
  double* partials = nullptr;
  unsigned* ticket = nullptr;
  if (rs.counters) {
    AVSI_REQUIRE((long long)blocks * 5 * (long long)sizeof(double) <= rs.area_bytes, "reduce scratch too small");
    partials = reinterpret_cast<double*>(rs.area);
    ticket = rs.counters + REDUCE_SLOT_L1;
  }
  masked_l1_kernel<<<blocks, L1_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, target, mask, seq_len, B, T, F,
                                                                   mode, grad_scale, grad_scale_dev, sums, prediction, dlogits, ldd,
                                                                   partials, ticket);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}

extern "C" int avsi_colsum_f16(const uint16_t* X, int ldx, int rows, int col0, int ncols, float* out,
                               void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(X && out, "null pointer");
  AVSI_REQUIRE(rows > 0 && ncols > 0 && ldx >= col0 + ncols, "sizes");
  const unsigned gx = (unsigned)((ncols + CS_VECS * 8 - 1) / (CS_VECS * 8));
  long long gy = ((long long)num_sms() * 8 + gx - 1) / gx;                       // ~8 blocks per SM in total
  if (gy > (rows + 63) / 64) gy = (rows + 63) / 64;
  if (gy < 1) gy = 1;
  dim3 grid(gx, (unsigned)gy);
  const ReduceScratch rs = reduce_scratch();
  float* partials = nullptr;
  unsigned* tickets = nullptr;
  if (rs.counters && gx <= 64) {
    AVSI_REQUIRE((long long)gy * gx * CS_VECS * 8 * (long long)sizeof(float) <= rs.area_bytes, "reduce scratch too small");
    partials = reinterpret_cast<float*>(rs.area);
    tickets = rs.counters + REDUCE_SLOT_COLSUM;
  }
  colsum_f16_kernel<<<grid, CS_VECS * CS_ROWS, 0, (cudaStream_t)stream>>>(X, ldx, rows, col0, ncols, out, partials, tickets);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
