// CTC prefix beam search on the HOST: the reference's `decoding` op, tf.nn.ctc_beam_search_decoder(tm_logits,
// sequence_lengths, beam_width=20) at models.py:1627 / :2027 and models_asr.py:139 -- a CPU op in TensorFlow as
// well, run at logging / validation steps, never inside the training step.  Plain prefix beam
// search over the log-softmax of the logits, blank = C - 1, every label tried at every frame (TF's
// label_selection_size = 0), top path only; `merge_repeated` reproduces TF-1's post-processing of the emitted
// path (consecutive equal labels of the OUTPUT collapsed, the default of tf.nn.ctc_beam_search_decoder).
// Utterances are independent: a small thread pool walks the batch.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <map>
#include <thread>
#include <vector>

#include "common.cuh"

namespace avsi {
namespace {

constexpr float NEG_INF = -INFINITY;

inline float log_add(float a, float b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  const float m = a > b ? a : b;
  return m + log1pf(expf(-fabsf(a - b)));
}

struct Prob {
  float pb = NEG_INF, pnb = NEG_INF;   // log P(prefix, ends in blank) / (ends in its last label)
  float total() const { return log_add(pb, pnb); }
};

typedef std::map<std::vector<int>, Prob> Beam;

void decode_one(const float* logits, int T, int stride, int C, int beam_width, bool merge_repeated, int max_out, int* out,
                int* out_len, float* log_prob) {
  const int blank = C - 1;
  Beam beam;
  beam[std::vector<int>()].pb = 0.f;
  std::vector<float> lp(C);
  std::vector<std::pair<float, const std::vector<int>*>> order;
  for (int t = 0; t < T; ++t) {
    const float* x = logits + (size_t)t * stride;
    float mx = x[0];
    for (int c = 1; c < C; ++c) mx = std::max(mx, x[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(x[c] - mx);
    const float lse = mx + logf(se);
    for (int c = 0; c < C; ++c) lp[c] = x[c] - lse;
    Beam next;
    for (const auto& kv : beam) {
      const std::vector<int>& y = kv.first;
      const Prob& p = kv.second;
      const float tot = p.total();
      Prob& same = next[y];
      same.pb = log_add(same.pb, tot + lp[blank]);
      if (!y.empty()) same.pnb = log_add(same.pnb, p.pnb + lp[y.back()]);
      std::vector<int> ext(y);
      ext.push_back(0);
      for (int c = 0; c < C; ++c) {
        if (c == blank) continue;
        ext.back() = c;
        const float from = (!y.empty() && y.back() == c) ? p.pb : tot;   // a repeat needs a blank in between
        if (from == NEG_INF) continue;
        Prob& e = next[ext];
        e.pnb = log_add(e.pnb, from + lp[c]);
      }
    }
    if ((int)next.size() > beam_width) {
      order.clear();
      for (const auto& kv : next) order.emplace_back(kv.second.total(), &kv.first);
      std::nth_element(order.begin(), order.begin() + beam_width, order.end(),
                       [](const std::pair<float, const std::vector<int>*>& a, const std::pair<float, const std::vector<int>*>& b) {
                         return a.first > b.first || (a.first == b.first && *a.second < *b.second);
                       });
      Beam pruned;
      for (int i = 0; i < beam_width; ++i) pruned[*order[i].second] = next[*order[i].second];
      beam.swap(pruned);
    } else {
      beam.swap(next);
    }
  }
  const std::vector<int>* best = nullptr;
  float best_lp = NEG_INF;
  for (const auto& kv : beam) {
    const float tot = kv.second.total();
    if (best == nullptr || tot > best_lp) {
      best = &kv.first;
      best_lp = tot;
    }
  }
  int n = 0;
  for (size_t i = 0; best && i < best->size(); ++i) {
    if (merge_repeated && i > 0 && (*best)[i] == (*best)[i - 1]) continue;
    if (n < max_out) out[n] = (*best)[i];
    ++n;
  }
  *out_len = n;
  if (log_prob) *log_prob = best_lp;
}

}  // namespace
}  // namespace avsi

extern "C" int avsi_ctc_beam_search_host(const float* logits, int T, int B, int ldl, int col0, int C, const int* seq_len,
                                         int beam_width, int merge_repeated, int max_out, int* out, int* out_len,
                                         float* log_prob, int n_threads) {
  using namespace avsi;
  AVSI_REQUIRE(logits && seq_len && out && out_len, "null pointer");
  AVSI_REQUIRE(T > 0 && B > 0 && C >= 2 && col0 >= 0 && ldl >= col0 + C, "sizes");
  AVSI_REQUIRE(beam_width >= 1 && max_out >= 1, "beam_width, max_out >= 1");
  for (long long i = 0; i < (long long)B * max_out; ++i) out[i] = -1;
  std::atomic<int> next_b(0);
  auto work = [&]() {
    for (int b = next_b.fetch_add(1); b < B; b = next_b.fetch_add(1)) {
      const int len = std::max(0, std::min(T, seq_len[b]));
      // time-major rows t*B + b
      decode_one(logits + (size_t)b * ldl + col0, len, B * ldl, C, beam_width, merge_repeated != 0, max_out,
                 out + (size_t)b * max_out, out_len + b, log_prob ? log_prob + b : nullptr);
    }
  };
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, B));
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int i = 0; i < nt; ++i) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  return AVSI_OK;
}

// CRC-32C (Castagnoli) over host memory, slicing-by-8: TFRecord frames (tfrecord_utils.py via tf.python_io) and the
// tensor-bundle checkpoints of tf.train.Saver (training.py:114,267,335) checksum every record / tensor with it.
extern "C" uint32_t avsi_crc32c_host(const void* data, uint64_t n, uint32_t crc) {
  static uint32_t tab[8][256];
  static std::atomic<int> ready(0);
  if (!ready.load(std::memory_order_acquire)) {
    uint32_t t[8][256];
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    for (int s = 0; s < 8; ++s)
      for (int i = 0; i < 256; ++i) tab[s][i] = t[s][i];
    ready.store(1, std::memory_order_release);
  }
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = ~crc;
  while (n >= 8) {
    const uint32_t lo = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
    const uint32_t hi = (uint32_t)p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
    const uint32_t x = c ^ lo;
    c = tab[7][x & 0xFF] ^ tab[6][(x >> 8) & 0xFF] ^ tab[5][(x >> 16) & 0xFF] ^ tab[4][x >> 24] ^
        tab[3][hi & 0xFF] ^ tab[2][(hi >> 8) & 0xFF] ^ tab[1][(hi >> 16) & 0xFF] ^ tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return ~c;
}
