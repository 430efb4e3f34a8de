"""TFRecord / SequenceExample I/O without TensorFlow (host data plumbing of the drop-in jobs)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_crc32c_known_answers():
    from avsi_b200 import tfrecord_io as tio
    assert tio.crc32c(b'123456789') == 0xE3069283              # RFC 3720 check value
    assert tio.crc32c(b'') == 0
    assert tio.crc32c(bytes(32)) == 0x8A9136AA                  # 32 zero bytes (RFC 3720 B.4)
    # TFRecord length header of an 8-byte little-endian 0: masking is ((crc >> 15 | crc << 17) + 0xa282ead8)
    c = tio.crc32c(bytes(8))
    assert tio.masked_crc(bytes(8)) == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_sequence_example_roundtrip_and_datamanager(tmp_path):
    from avsi_b200 import tfrecord_io as tio
    from avsi_b200.dataset_reader import DataManager
    rng = np.random.default_rng(0)
    T, N = 25, 4800
    files, truth = [], []
    for i in range(5):
        wav = np.round(rng.normal(0, 3000, N)).astype(np.int16)
        video = rng.standard_normal((T, 136)).astype(np.float32)
        mask = np.ones((T, 257), np.float32)
        mask[3 + i:8 + i] = 0
        lab_len = 12 + i
        labels = np.zeros(50, np.int64)
        labels[:lab_len] = rng.integers(0, 33, lab_len)
        rec = tio.serialize_sample_fixed(T, lab_len, wav, video, mask, labels, 's%02d' % i)
        f = str(tmp_path / ('data_%05d.tfrecord' % (i + 1)))
        tio.write_records(f, [rec])
        files.append(f)
        truth.append((wav, video, mask, labels, lab_len))
    # raw parse
    ctx, seq = tio.parse_sequence_example(next(tio.read_records(files[2], verify=True)))
    assert int(ctx['sequence_length'][0]) == T and int(ctx['labels_length'][0]) == 14
    assert ctx['sample_path'] == [b's02']
    assert np.array_equal(ctx['target_audio_wav'], truth[2][0].astype(np.float32))
    assert len(seq['mask']) == T and np.array_equal(np.stack(seq['mask']), truth[2][2])
    assert np.array_equal(np.concatenate(seq['labels']), truth[2][3].astype(np.float32))
    # corrupted payload is caught when verification is on
    raw = bytearray(open(files[0], 'rb').read())
    raw[40] ^= 0xFF
    bad = str(tmp_path / 'bad.tfrecord')
    open(bad, 'wb').write(bytes(raw))
    with pytest.raises(IOError):
        list(tio.read_records(bad, verify=True))
    # batches in the reference's tuple order, remainder kept, sharding disjoint and complete
    dm = DataManager(num_audio_samples=N)
    _, it = dm.get_iterator(dm.get_dataset(files, shuffle=False), batch_size=2, n_epochs=1)
    batches = list(it)
    assert [len(b[0]) for b in batches] == [2, 2, 1]
    seq_len, lab_len, wav, paths, labels, video, mask = batches[0]
    assert wav.dtype == np.int32 and wav.shape == (2, N) and video.shape == (2, T, 136) and mask.shape == (2, T, 257)
    assert paths == [b's00', b's01'] and list(lab_len) == [12, 13] and labels.shape == (2, 50)
    assert np.array_equal(video[1], truth[1][1])
    seen = []
    for r in range(2):
        d = DataManager(num_audio_samples=N, rank=r, world=2)
        _, it = d.get_iterator(d.get_dataset(files, shuffle=True, seed=1), batch_size=8, n_epochs=1)
        seen.append(sorted(p for b in it for p in b[3]))
    assert sorted(seen[0] + seen[1]) == [b's%02d' % i for i in range(5)] and not set(seen[0]) & set(seen[1])


def _records(rng, n, T=25, N=4800, L=50):
    from avsi_b200 import tfrecord_io as tio
    out = []
    for i in range(n):
        wav = np.round(rng.normal(0, 3000, N)).astype(np.int16)
        video = rng.standard_normal((T, 136)).astype(np.float32)
        mask = (rng.random((T, 257)) > 0.2).astype(np.float32)
        labels = np.zeros(L, np.int64)
        labels[:10 + i % 7] = rng.integers(0, 33, 10 + i % 7)
        out.append(tio.serialize_sample_fixed(T - i % 3, 10 + i % 7, wav, video, mask, labels, 'dir/s%03d' % i))
    return out


def test_native_parser_equals_python_parser(tmp_path):
    """avsi_parse_av_sample_host (the host C routine the DataManager uses) against the pure-Python wire-format parser:
    same values for packed and un-packed float lists, clean errors on truncated payloads and small buffers."""
    from avsi_b200 import tfrecord_io as tio
    rng = np.random.default_rng(3)
    recs = _records(rng, 6)
    for r in recs:
        fast = tio.parse_av_sample(r, num_audio_samples=4800)
        assert fast is not None, 'library not built'
        ctx, seq = tio.parse_sequence_example(r)
        assert fast[0] == int(ctx['sequence_length'][0]) and fast[1] == int(ctx['labels_length'][0])
        assert np.array_equal(fast[2], ctx['target_audio_wav']) and fast[3] == ctx['sample_path'][0]
        assert np.array_equal(fast[4], np.concatenate(seq['labels']))
        assert np.array_equal(fast[5], np.stack(seq['video_features'])) and np.array_equal(fast[6], np.stack(seq['mask']))
    # un-packed floats (one fixed32 per value: legal wire format) and an unknown context key are accepted
    def feature_unpacked(vals):
        body = b''.join(b'\x0d' + np.float32(v).tobytes() for v in vals)            # field 1, wire type 5
        return tio._ld(2, body)
    ctx = tio._enc_map({'sequence_length': np.array([3], np.int64), 'labels_length': np.array([1], np.int64),
                        'extra': np.array([7], np.int64)}, tio._enc_feature)
    ctx += tio._ld(1, tio._ld(1, b'target_audio_wav') + tio._ld(2, feature_unpacked([1.5, -2.0, 3.25])))
    fls = tio._enc_map({'mask': [np.ones(4, np.float32)] * 3, 'video_features': [np.zeros(2, np.float32)] * 3,
                        'labels': [np.array([5.0], np.float32)]}, lambda rows: b''.join(tio._ld(1, tio._enc_feature(r)) for r in rows))
    rec = tio._ld(1, ctx) + tio._ld(2, fls)
    got = tio.parse_av_sample(rec, num_audio_samples=3, audio_feat_size=4, video_feat_size=2)
    assert got[0] == 3 and got[1] == 1 and got[2].tolist() == [1.5, -2.0, 3.25] and got[4].tolist() == [5.0]
    assert got[5].shape == (3, 2) and got[6].shape == (3, 4) and got[3] == b''
    # malformed / too small
    with pytest.raises(ValueError):
        tio.parse_av_sample(recs[0][:len(recs[0]) // 2], num_audio_samples=4800)
    with pytest.raises(ValueError):
        tio.parse_av_sample(recs[0], num_audio_samples=100)                        # wav does not fit
    ragged = tio.build_sequence_example({'sequence_length': np.array([2], np.int64)},
                                        {'mask': [np.ones(4, np.float32), np.ones(3, np.float32)]})
    with pytest.raises(ValueError):
        tio.parse_av_sample(ragged, num_audio_samples=8, audio_feat_size=4)


def test_datamanager_native_batches_equal_python_batches(tmp_path):
    """The batch-level native path (thread pool, one batch of prefetch) yields exactly the batches of the per-sample
    Python path, shuffled or not, remainder included; a batch of mixed shapes is an error."""
    from avsi_b200 import tfrecord_io as tio
    from avsi_b200.dataset_reader import DataManager
    rng = np.random.default_rng(4)
    recs = _records(rng, 11)
    files = []
    for k in range(3):
        f = str(tmp_path / ('data_%05d.tfrecord' % k))
        tio.write_records(f, recs[4 * k:4 * k + 4])
        files.append(f)
    dm = DataManager(num_audio_samples=4800, buffer_size=5, num_parallel_calls=3)
    for shuffle in (False, True):
        ds = dm.get_dataset(files, shuffle=shuffle, seed=9)
        _, it = dm.get_iterator(ds, batch_size=4, n_epochs=2)
        fast = list(it)
        slow = []
        for _ in range(2):
            cur = []
            for s in dm.get_dataset(files, shuffle=shuffle, seed=9).samples():
                cur.append(s)
                if len(cur) == 4:
                    slow.append(cur)
                    cur = []
            if cur:
                slow.append(cur)
        assert [len(b[0]) for b in fast] == [len(c) for c in slow] == [4, 4, 3, 4, 4, 3]
        for fb, sb in zip(fast, slow):
            for i, smp in enumerate(sb):
                assert fb[0][i] == smp[0] and fb[1][i] == smp[1] and fb[3][i] == smp[3]
                assert np.array_equal(fb[2][i], smp[2]) and fb[2].dtype == np.int32
                assert np.array_equal(fb[4][i], smp[4]) and np.array_equal(fb[5][i], smp[5]) and np.array_equal(fb[6][i], smp[6])
    odd = str(tmp_path / 'odd.tfrecord')
    tio.write_records(odd, [recs[0], _records(rng, 1, T=30)[0]])
    _, it = dm.get_iterator(dm.get_dataset([odd], shuffle=False), batch_size=2, n_epochs=1)
    with pytest.raises(ValueError):
        list(it)
