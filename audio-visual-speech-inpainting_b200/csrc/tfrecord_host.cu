// Host-side parser of the reference's TFRecord payload: one serialized tf.train.SequenceExample written by
// tfrecord_utils.py:19-41 ('fixed' mode) and read by dataset_reader.py:62-79 through tf.parse_single_sequence_example:
//   context        sequence_length, labels_length (int64), target_audio_wav (floats), sample_path (bytes)
//   feature_lists  mask [T x F floats], video_features [T x V floats], labels [L x 1 float]
// straight into the caller's (batch) buffers.  The pure-Python parser of tfrecord_io.py walks every one of the ~550
// nested Feature messages of a record in the interpreter (300 records/s); this one is a single pass over the bytes, and
// ctypes releases the GIL around it so that a thread pool scales it across host cores.  No CUDA in this file.
#include <cstring>

#include "common.cuh"

namespace avsi {
namespace {

struct Span {
  const unsigned char* p;
  const unsigned char* end;
};

inline bool rd_varint(Span& s, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
    const unsigned char b = *s.p++;
    v |= (uint64_t)(b & 0x7F) << shift;
    if (b < 0x80) return true;
  }
  return false;
}
inline bool rd_len(Span& s, Span& out) {
  uint64_t n;
  if (!rd_varint(s, n) || n > (uint64_t)(s.end - s.p)) return false;
  out.p = s.p;
  out.end = s.p + n;
  s.p += n;
  return true;
}
inline bool skip(Span& s, int wt) {
  uint64_t v;
  Span t;
  switch (wt) {
    case 0: return rd_varint(s, v);
    case 1: if (s.end - s.p < 8) return false; s.p += 8; return true;
    case 2: return rd_len(s, t);
    case 5: if (s.end - s.p < 4) return false; s.p += 4; return true;
    default: return false;
  }
}
inline bool key_is(const Span& k, const char* name) {
  const size_t n = strlen(name);
  return (size_t)(k.end - k.p) == n && memcmp(k.p, name, n) == 0;
}

// Feature{float_list = 2 {value = 1, packed or not}} -> dst[*n ...]; returns false on malformed input / overflow
bool feature_floats(Span f, float* dst, long long cap, long long& n) {
  while (f.p < f.end) {
    uint64_t tag;
    if (!rd_varint(f, tag)) return false;
    if ((tag >> 3) == 2 && (tag & 7) == 2) {
      Span fl;
      if (!rd_len(f, fl)) return false;
      while (fl.p < fl.end) {
        uint64_t t2;
        if (!rd_varint(fl, t2)) return false;
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) {               // packed
          Span pk;
          if (!rd_len(fl, pk)) return false;
          const long long cnt = (pk.end - pk.p) / 4;
          if ((pk.end - pk.p) % 4 || n + cnt > cap) return false;
          memcpy(dst + n, pk.p, (size_t)cnt * 4);
          n += cnt;
        } else if ((t2 >> 3) == 1 && (t2 & 7) == 5) {        // one value
          if (fl.end - fl.p < 4 || n + 1 > cap) return false;
          memcpy(dst + n, fl.p, 4);
          fl.p += 4;
          ++n;
        } else if (!skip(fl, (int)(t2 & 7))) {
          return false;
        }
      }
    } else if (!skip(f, (int)(tag & 7))) {
      return false;
    }
  }
  return true;
}

bool feature_int64_first(Span f, int64_t& out) {
  while (f.p < f.end) {
    uint64_t tag;
    if (!rd_varint(f, tag)) return false;
    if ((tag >> 3) == 3 && (tag & 7) == 2) {
      Span il;
      if (!rd_len(f, il)) return false;
      while (il.p < il.end) {
        uint64_t t2, v;
        if (!rd_varint(il, t2)) return false;
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) {
          Span pk;
          if (!rd_len(il, pk)) return false;
          if (pk.p < pk.end) {
            if (!rd_varint(pk, v)) return false;
            out = (int64_t)v;
            return true;
          }
        } else if ((t2 >> 3) == 1 && (t2 & 7) == 0) {
          if (!rd_varint(il, v)) return false;
          out = (int64_t)v;
          return true;
        } else if (!skip(il, (int)(t2 & 7))) {
          return false;
        }
      }
    } else if (!skip(f, (int)(tag & 7))) {
      return false;
    }
  }
  return true;
}

bool feature_bytes_first(Span f, char* dst, int cap, int64_t& n) {
  while (f.p < f.end) {
    uint64_t tag;
    if (!rd_varint(f, tag)) return false;
    if ((tag >> 3) == 1 && (tag & 7) == 2) {
      Span bl;
      if (!rd_len(f, bl)) return false;
      while (bl.p < bl.end) {
        uint64_t t2;
        if (!rd_varint(bl, t2)) return false;
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) {
          Span v;
          if (!rd_len(bl, v)) return false;
          n = v.end - v.p;
          if (dst && cap > 0) memcpy(dst, v.p, (size_t)std::min<int64_t>(n, cap));
          return true;
        } else if (!skip(bl, (int)(t2 & 7))) {
          return false;
        }
      }
    } else if (!skip(f, (int)(tag & 7))) {
      return false;
    }
  }
  return true;
}

// map entry {1: key, 2: value}
bool map_entry(Span e, Span& key, Span& val) {
  key = Span{nullptr, nullptr};
  val = Span{nullptr, nullptr};
  while (e.p < e.end) {
    uint64_t tag;
    if (!rd_varint(e, tag)) return false;
    if ((tag & 7) == 2 && ((tag >> 3) == 1 || (tag >> 3) == 2)) {
      Span v;
      if (!rd_len(e, v)) return false;
      if ((tag >> 3) == 1) key = v;
      else val = v;
    } else if (!skip(e, (int)(tag & 7))) {
      return false;
    }
  }
  return key.p != nullptr;
}

// FeatureList{1: repeated Feature} of float rows -> dst row after row; every row must have the same length
bool feature_list_rows(Span fl, float* dst, long long cap, int64_t& rows, int64_t& cols) {
  long long n = 0;
  rows = 0;
  cols = -1;
  while (fl.p < fl.end) {
    uint64_t tag;
    if (!rd_varint(fl, tag)) return false;
    if ((tag >> 3) == 1 && (tag & 7) == 2) {
      Span f;
      if (!rd_len(fl, f)) return false;
      const long long before = n;
      if (!feature_floats(f, dst, cap, n)) return false;
      if (cols < 0) cols = n - before;
      else if (n - before != cols) return false;
      ++rows;
    } else if (!skip(fl, (int)(tag & 7))) {
      return false;
    }
  }
  if (cols < 0) cols = 0;
  return true;
}

}  // namespace
}  // namespace avsi

extern "C" int avsi_parse_av_sample_host(const void* rec, uint64_t len, float* wav, int64_t wav_cap, float* mask,
                                         int64_t mask_cap, float* video, int64_t video_cap, float* labels,
                                         int64_t labels_cap, char* path, int path_cap, int64_t* meta) {
  using namespace avsi;
  AVSI_REQUIRE(rec && wav && mask && video && labels && meta, "null pointer");
  for (int i = 0; i < 9; ++i) meta[i] = 0;
  Span s{static_cast<const unsigned char*>(rec), static_cast<const unsigned char*>(rec) + len};
  bool ok = true;
  while (ok && s.p < s.end) {
    uint64_t tag;
    if (!(ok = rd_varint(s, tag))) break;
    const int field = (int)(tag >> 3), wt = (int)(tag & 7);
    if (wt == 2 && (field == 1 || field == 2)) {
      Span body;
      if (!(ok = rd_len(s, body))) break;
      while (ok && body.p < body.end) {                       // Features / FeatureLists: repeated map entries (field 1)
        uint64_t t2;
        if (!(ok = rd_varint(body, t2))) break;
        if ((t2 >> 3) == 1 && (t2 & 7) == 2) {
          Span entry, key, val;
          if (!(ok = rd_len(body, entry) && map_entry(entry, key, val))) break;
          if (!val.p) continue;
          if (field == 1) {
            if (key_is(key, "sequence_length")) ok = feature_int64_first(val, meta[0]);
            else if (key_is(key, "labels_length")) ok = feature_int64_first(val, meta[1]);
            else if (key_is(key, "target_audio_wav")) {
              long long n = 0;
              ok = feature_floats(val, wav, wav_cap, n);
              meta[2] = n;
            } else if (key_is(key, "sample_path")) ok = feature_bytes_first(val, path, path_cap, meta[8]);
          } else {
            if (key_is(key, "mask")) ok = feature_list_rows(val, mask, mask_cap, meta[3], meta[4]);
            else if (key_is(key, "video_features")) ok = feature_list_rows(val, video, video_cap, meta[5], meta[6]);
            else if (key_is(key, "labels")) {
              int64_t cols;
              ok = feature_list_rows(val, labels, labels_cap, meta[7], cols) && (cols == 1 || meta[7] == 0);
            }
          }
        } else {
          ok = skip(body, (int)(t2 & 7));
        }
      }
    } else {
      ok = skip(s, wt);
    }
  }
  if (!ok) return set_error(AVSI_ERR_INVALID, "%s: malformed SequenceExample or a buffer too small%s", "avsi_parse_av_sample_host");
  return AVSI_OK;
}
