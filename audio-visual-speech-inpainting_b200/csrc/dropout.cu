// Inverted dropout on the BLSTM outputs ahead of the linear head(s): tf.nn.dropout(rnn_outputs, rate) of
// models.py:117 (SI), :1901 (MTL) and models_asr.py:120 -- y * keep / (1 - rate), keep = (u >= rate).
// The mask is never stored: it is a pure function of (seed, offset, element index) through Philox-4x32-10, so the
// backward pass regenerates it on dY with the same (seed, offset).  HBM-bound: one 16-byte chunk (8 halves, two
// Philox blocks) per thread, coalesced; `keep_out` (tests, parity against the oracle with the same mask) is optional.
#include "common.cuh"

namespace avsi {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__global__ void __launch_bounds__(256)
dropout_f16_kernel(const uint16_t* __restrict__ src, int ld_src, uint16_t* __restrict__ dst, int ld_dst, long long rows,
                   int chunks_per_row, uint32_t thresh, float scale, uint2 key, unsigned long long offset,
                   uint8_t* __restrict__ keep_out, int layout) {
  const long long total = rows * chunks_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks_per_row;
    const int ch = (int)(i - r * chunks_per_row);
    // layout bit 0 / 1: src / dst is INTERLEAVED (common.cuh il16, ld = logical row length); the mask is a function of
    // the logical (row, chunk) either way
    const long long so = (layout & 1) ? il16(r, ch * 8, ld_src) : r * ld_src + ch * 8;
    const long long dofs = (layout & 2) ? il16(r, ch * 8, ld_dst) : r * ld_dst + ch * 8;
    const uint4 in = *reinterpret_cast<const uint4*>(src + so);
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint32_t rnd[8];
    {
      const uint4 a = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32) << 1), key);
      const uint4 b = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, ((uint32_t)(offset >> 32) << 1) | 1u), key);
      rnd[0] = a.x; rnd[1] = a.y; rnd[2] = a.z; rnd[3] = a.w;
      rnd[4] = b.x; rnd[5] = b.y; rnd[6] = b.z; rnd[7] = b.w;
    }
    uint32_t o[4];
    uint32_t kbits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool k0 = rnd[2 * j] >= thresh, k1 = rnd[2 * j + 1] >= thresh;
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      const __half2 h = __floats2half2_rn(k0 ? f.x * scale : 0.f, k1 ? f.y * scale : 0.f);
      o[j] = *reinterpret_cast<const uint32_t*>(&h);
      kbits |= (k0 ? 1u : 0u) << (2 * j) | (k1 ? 1u : 0u) << (2 * j + 1);
    }
    *reinterpret_cast<uint4*>(dst + dofs) = make_uint4(o[0], o[1], o[2], o[3]);
    if (keep_out) {
      uint2 kb;
      kb.x = (kbits & 1u) | ((kbits >> 1 & 1u) << 8) | ((kbits >> 2 & 1u) << 16) | ((kbits >> 3 & 1u) << 24);
      kb.y = (kbits >> 4 & 1u) | ((kbits >> 5 & 1u) << 8) | ((kbits >> 6 & 1u) << 16) | ((kbits >> 7 & 1u) << 24);
      *reinterpret_cast<uint2*>(keep_out + (r * chunks_per_row + ch) * 8) = kb;
    }
  }
}

}  // namespace avsi

extern "C" int avsi_dropout_f16(const void* src, int ld_src, void* dst, int ld_dst, int64_t rows, int cols, float rate,
                                uint64_t seed, uint64_t offset, void* keep_out, int layout, void* stream) {
  using namespace avsi;
  AVSI_REQUIRE(src && dst, "null pointer");
  AVSI_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld_src >= cols && ld_dst >= cols, "sizes (cols multiple of 8)");
  AVSI_REQUIRE(ld_src % 8 == 0 && ld_dst % 8 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, "16-byte rows");
  AVSI_REQUIRE(rate >= 0.f && rate < 1.f, "rate in [0, 1)");
  AVSI_REQUIRE(!(layout & 1) || ld_src == cols, "interleaved src: ld = cols");
  AVSI_REQUIRE(!(layout & 2) || ld_dst == cols, "interleaved dst: ld = cols");
  AVSI_REQUIRE(!keep_out || (uintptr_t)keep_out % 8 == 0, "keep_out alignment");
  // keep = (u >= rate) with u uniform on [0, 1) in 2^-32 steps
  const double t = (double)rate * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  const long long total = rows * (long long)(cols / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  dropout_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)src, ld_src, (uint16_t*)dst, ld_dst, rows, cols / 8, thresh, 1.0f / (1.0f - rate),
      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), offset, (uint8_t*)keep_out, layout);
  AVSI_LAUNCH_CHECK();
  return AVSI_OK;
}
