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
