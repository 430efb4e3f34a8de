"""TFRecord files of tf.train.SequenceExample without TensorFlow.

The reference stores every utterance as one `data_XXXXX.tfrecord` holding a SequenceExample with the keys of
tfrecord_utils.py:19-41 ('fixed' mode) and parses it in dataset_reader.py:62-79.  Both formats are stable and
small, so they are handled here directly:

  TFRecord framing   uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)
  SequenceExample    context = 1 (Features), feature_lists = 2 (FeatureLists)
  Features           map<string, Feature> feature = 1        FeatureLists  map<string, FeatureList> feature_list = 1
  FeatureList        repeated Feature feature = 1
  Feature            oneof { BytesList = 1, FloatList = 2, Int64List = 3 }, each `repeated value = 1` (packed for numbers)

Pure Python / numpy; host-side data plumbing only (nothing here is on the GPU hot path).
"""
import struct

import numpy as np

_CRC_TABLE = None
_NATIVE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78                                   # CRC-32C (Castagnoli), reflected
        tab = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = [int(x) for x in tab]
    return _CRC_TABLE


def _native():
    """The library's host routines (CRC-32C, SequenceExample parser), or None when it cannot be loaded."""
    global _NATIVE
    if _NATIVE is None:
        try:
            from . import _lib
            _NATIVE = _lib.load()
        except Exception:                                  # noqa: BLE001 -- the pure-Python code below is complete
            _NATIVE = False
    return _NATIVE or None


def crc32c(data):
    if len(data) >= 1024:
        lib = _native()
        if lib is not None:
            raw = bytes(data) if not isinstance(data, bytes) else data
            return int(lib.avsi_crc32c_host(raw, len(raw), 0))
    tab = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire format ---------------------------------------------------------------------------------
def _varint(buf, pos):
    r = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if b < 0x80:
            return r, pos
        shift += 7


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _fields(buf):
    """Yield (field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield f, wt, v


def _parse_feature(buf):
    """Feature -> numpy array (float32 / int64) or list of bytes."""
    for f, wt, v in _fields(buf):
        if f == 1:                                            # BytesList
            return [bytes(x) for ff, _, x in _fields(v) if ff == 1]
        if f == 2:                                            # FloatList
            parts = []
            for ff, w2, x in _fields(v):
                if ff == 1:
                    parts.append(np.frombuffer(x, '<f4'))
            return np.concatenate(parts) if len(parts) != 1 else parts[0]
        if f == 3:                                            # Int64List
            vals = []
            for ff, w2, x in _fields(v):
                if ff != 1:
                    continue
                if w2 == 0:
                    vals.append(x)
                else:
                    p = 0
                    while p < len(x):
                        val, p = _varint(x, p)
                        vals.append(val)
            a = np.array(vals, np.uint64).astype(np.int64)
            return a
    return np.zeros(0, np.float32)


def _parse_map(buf, value_parser):
    out = {}
    for f, _, entry in _fields(buf):
        if f != 1:
            continue
        key, val = None, None
        for ff, _, x in _fields(entry):
            if ff == 1:
                key = bytes(x).decode()
            elif ff == 2:
                val = value_parser(x)
        out[key] = val
    return out


def parse_sequence_example(data):
    """bytes -> (context {name: array | [bytes]}, feature_lists {name: [array, ...]})."""
    buf = memoryview(data)
    context, lists = {}, {}
    for f, _, v in _fields(buf):
        if f == 1:
            context = _parse_map(v, _parse_feature)
        elif f == 2:
            lists = _parse_map(v, lambda fl: [_parse_feature(x) for ff, _, x in _fields(fl) if ff == 1])
    return context, lists


def _ld(field, payload):
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def _enc_feature(value):
    if isinstance(value, (bytes, str)):
        value = [value]
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], (bytes, str)):
        body = b''.join(_ld(1, v.encode() if isinstance(v, str) else v) for v in value)
        return _ld(1, body)
    a = np.asarray(value)
    if a.dtype.kind in 'iu':
        packed = b''.join(_enc_varint(int(x)) for x in a.ravel())
        return _ld(3, _ld(1, packed))
    return _ld(2, _ld(1, a.astype('<f4').ravel().tobytes()))


def _enc_map(d, enc_value):
    return b''.join(_ld(1, _ld(1, k.encode()) + _ld(2, enc_value(v))) for k, v in d.items())


def build_sequence_example(context, feature_lists):
    """Inverse of parse_sequence_example (context values: ints / floats arrays / bytes; lists: rows)."""
    ctx = _enc_map(context, _enc_feature)
    fls = _enc_map(feature_lists, lambda rows: b''.join(_ld(1, _enc_feature(r)) for r in rows))
    return _ld(1, ctx) + _ld(2, fls)


# ---- record files -----------------------------------------------------------------------------------------
def read_records(path, verify=False):
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError('truncated TFRecord header in %s' % path)
            (n,), (hcrc,) = struct.unpack('<Q', head[:8]), struct.unpack('<I', head[8:])
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise IOError('truncated TFRecord in %s' % path)
            if verify and (masked_crc(head[:8]) != hcrc or masked_crc(data) != struct.unpack('<I', tail)[0]):
                raise IOError('TFRecord CRC mismatch in %s' % path)
            yield data


def write_records(path, records):
    with open(path, 'wb') as f:
        for data in records:
            ln = struct.pack('<Q', len(data))
            f.write(ln + struct.pack('<I', masked_crc(ln)) + data + struct.pack('<I', masked_crc(data)))


def serialize_sample_fixed(seq_len, lab_len, target_audio_wav, video_features, mask, labels, sample_path, embedding=None):
    """Same keys and value kinds as tfrecord_utils.py:19-41 (the 'fixed' mode, the only working one -- SURVEY.md 2.4);
    with `embedding`, the context of tfrecord_emb_utils.py:19-43 (one more float feature)."""
    ctx = {'sequence_length': np.array([seq_len], np.int64), 'labels_length': np.array([lab_len], np.int64),
           'target_audio_wav': np.asarray(target_audio_wav, np.float32), 'sample_path': sample_path.encode()}
    if embedding is not None:
        ctx['embedding'] = np.asarray(embedding, np.float32)
    return build_sequence_example(
        ctx,
        {'mask': [np.asarray(r, np.float32) for r in mask],
         'video_features': [np.asarray(r, np.float32) for r in video_features],
         'labels': [np.array([l], np.float32) for l in labels]})


def parse_av_sample(data, num_audio_samples=48000, audio_feat_size=257, video_feat_size=136, max_frames=None,
                    max_labels=64):
    """One record of the reference's TFRecords -> (seq_len, lab_len, wav f32 [N], sample_path bytes, labels f32 [L],
    video f32 [T, V], mask f32 [T, F]) through the library's host parser (avsi_parse_av_sample_host); None when the
    library is unavailable (callers fall back to parse_sequence_example)."""
    import ctypes
    lib = _native()
    if lib is None:
        return None
    if max_frames is None:
        max_frames = max(1, len(data) // (4 * (audio_feat_size + 1)))      # a mask row costs > 4 F bytes
    wav = np.empty(num_audio_samples, np.float32)
    mask = np.empty(max_frames * audio_feat_size, np.float32)
    video = np.empty(max_frames * max(video_feat_size, 1), np.float32)
    labels = np.empty(max_labels, np.float32)
    path = ctypes.create_string_buffer(512)
    meta = np.zeros(9, np.int64)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    raw = bytes(data) if not isinstance(data, bytes) else data
    rc = lib.avsi_parse_av_sample_host(raw, len(raw), vp(wav), wav.size, vp(mask), mask.size, vp(video), video.size,
                                       vp(labels), labels.size, path, 512, vp(meta))
    if rc != 0:
        raise ValueError('malformed TFRecord payload: %s' % lib.avsi_last_error().decode(errors='replace'))
    seq_len, lab_len, n_wav, mr, mc, vr, vc, nl, pl = (int(x) for x in meta)
    return (seq_len, lab_len, wav[:n_wav], path.raw[:min(pl, 512)], labels[:nl].copy(),
            video[:vr * vc].reshape(vr, vc) if vr else np.zeros((0, video_feat_size), np.float32),
            mask[:mr * mc].reshape(mr, mc) if mr else np.zeros((0, audio_feat_size), np.float32))


def parse_av_batch(records, num_audio_samples=48000, audio_feat_size=257, video_feat_size=136, pool=None):
    """A list of records -> the DataManager's 7-tuple, every record parsed by the library straight into its row of the
    batch arrays (no per-sample arrays, no stacking); `pool` (a ThreadPoolExecutor) spreads the rows over host cores,
    the C call runs without the GIL.  None when the library is unavailable.  All records must share T and the padded
    label count (true of the reference's 'fixed' TFRecords)."""
    import ctypes
    lib = _native()
    if lib is None:
        return None
    first = parse_av_sample(records[0], num_audio_samples, audio_feat_size, video_feat_size)
    T, L = first[6].shape[0], first[4].shape[0]
    V = first[5].shape[1] if first[5].size else video_feat_size
    B = len(records)
    wav = np.empty((B, num_audio_samples), np.float32)
    mask = np.empty((B, T, audio_feat_size), np.float32)
    video = np.empty((B, T, V), np.float32)
    labels = np.zeros((B, max(L, 1)), np.float32)
    meta = np.zeros((B, 9), np.int64)
    paths = [ctypes.create_string_buffer(512) for _ in range(B)]
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)

    def one(i):
        raw = records[i] if isinstance(records[i], bytes) else bytes(records[i])
        rc = lib.avsi_parse_av_sample_host(raw, len(raw), vp(wav[i]), wav.shape[1], vp(mask[i]), mask[i].size, vp(video[i]),
                                           video[i].size, vp(labels[i]), labels.shape[1], paths[i], 512, vp(meta[i]))
        if rc != 0:
            raise ValueError('malformed TFRecord payload in record %d of the batch' % i)
        m = meta[i]
        if m[2] != num_audio_samples:
            raise ValueError('target_audio_wav has %d samples, expected %d' % (m[2], num_audio_samples))
        if m[3] != T or m[4] != audio_feat_size or m[5] != T or m[6] != V or m[7] != L:
            raise ValueError('records of one batch differ in shape (T, F, V, labels): %s' % (m[3:8].tolist(),))
    if pool is not None and B > 1:
        list(pool.map(one, range(B)))
    else:
        for i in range(B):
            one(i)
    return (meta[:, 0].astype(np.int32), meta[:, 1].astype(np.int32), wav.astype(np.int32),
            [paths[i].raw[:min(int(meta[i, 8]), 512)] for i in range(B)], labels[:, :L], video, mask)
