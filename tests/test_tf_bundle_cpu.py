"""TensorFlow tensor-bundle checkpoints without TensorFlow (tf_bundle.py).  Unpinned by a genuine TF file (none ships
with the reference, TF is not installable here): the reader is exercised on tables the writer emits with restart
points and prefix compression, on a hand-assembled table, and on corruption."""
import os
import struct

import numpy as np
import pytest

from avsi_b200 import tf_bundle


def _vars(rng):
    pre = 'av-blstm/cudnn_lstm/stack_bidirectional_rnn/cell_%d/bidirectional_rnn/%s/cudnn_compatible_lstm_cell/%s'
    v = {}
    for l in range(3):
        for d in ('fw', 'bw'):
            v[pre % (l, d, 'kernel')] = rng.standard_normal((40 + 3 * l, 16)).astype(np.float32)
            v[pre % (l, d, 'bias')] = rng.standard_normal(16).astype(np.float32)
            v[pre % (l, d, 'kernel') + '/Adam'] = rng.standard_normal((40 + 3 * l, 16)).astype(np.float32)
    v['av-blstm/logits/weights'] = rng.standard_normal((8, 257)).astype(np.float32)
    v['av-blstm/Variable'] = np.asarray(1234, np.int32)                  # the global step: a scalar
    v['beta1_power'] = np.asarray(0.9 ** 7, np.float32)
    v['empty'] = np.zeros((0, 3), np.float64)
    v['wide'] = rng.integers(-5, 5, (3, 2, 2)).astype(np.int64)
    return v


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    v = _vars(rng)
    prefix = str(tmp_path / 'netmodel' / 'sinet')
    tf_bundle.write_bundle(prefix, v)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    got = tf_bundle.read_bundle(prefix)
    assert sorted(got) == sorted(v)
    for k in v:
        assert got[k].dtype == v[k].dtype and got[k].shape == v[k].shape and np.array_equal(got[k], v[k]), k
    assert tf_bundle.latest_checkpoint(str(tmp_path / 'netmodel')) == prefix
    # the data file is the tensors back to back in key order
    order = sorted(v, key=lambda s: s.encode())
    assert os.path.getsize(prefix + '.data-00000-of-00001') == sum(v[k].nbytes for k in order)


def test_many_keys_span_blocks_and_restarts(tmp_path):
    items = [(('layer/%04d/w' % i).encode(), struct.pack('<I', i) * (1 + i % 5)) for i in range(700)]
    path = str(tmp_path / 't.index')
    tf_bundle.write_table(path, items, block_size=512)
    assert tf_bundle.read_table(path) == items
    with pytest.raises(ValueError):
        tf_bundle.write_table(path, [(b'b', b''), (b'a', b'')])


def test_hand_assembled_table(tmp_path):
    """One data block written byte by byte from the format description: keys 'ab', 'abc' (shares 2), 'b' (restart)."""
    def entry(shared, suffix, value):
        return bytes([shared, len(suffix), len(value)]) + suffix + value
    body = entry(0, b'ab', b'1') + entry(2, b'c', b'22')
    r1 = len(body)
    body += entry(0, b'b', b'')
    block = body + struct.pack('<III', 0, r1, 2)

    def framed(b):
        return b + b'\x00' + struct.pack('<I', tf_bundle._mask(tf_bundle.crc32c(b + b'\x00')))
    meta = struct.pack('<II', 0, 1)
    off_meta = len(framed(block))
    index_body = bytes([0, 1, 2]) + b'b' + bytes([0, len(block)])
    index = index_body + struct.pack('<II', 0, 1)
    off_index = off_meta + len(framed(meta))
    footer = bytes([off_meta, len(meta), off_index, len(index)])
    data = framed(block) + framed(meta) + framed(index) + footer + b'\x00' * (40 - len(footer)) + \
        struct.pack('<Q', 0xdb4775248b80fb57)
    path = str(tmp_path / 'h.index')
    with open(path, 'wb') as f:
        f.write(data)
    assert tf_bundle.read_table(path) == [(b'ab', b'1'), (b'abc', b'22'), (b'b', b'')]


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(1)
    prefix = str(tmp_path / 'ckpt')
    tf_bundle.write_bundle(prefix, _vars(rng))
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[100] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        tf_bundle.read_bundle(prefix)
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[10] ^= 1
    open(prefix + '.index', 'wb').write(bytes(idx))
    with pytest.raises(ValueError):
        tf_bundle.read_bundle(prefix)
    with pytest.raises(ValueError, match='not a valid checkpoint'):
        tf_bundle.read_bundle(str(tmp_path / 'missing'))


def test_masked_crc_known_answer():
    assert tf_bundle.crc32c(b'123456789') == 0xE3069283
    big = bytes(range(256)) * 64                                           # takes the library routine when it is built
    from avsi_b200 import tfrecord_io
    assert tf_bundle.crc32c(big) == tfrecord_io.crc32c(big)


def test_training_driver_checkpoint_formats(tmp_path):
    """training._save honours `checkpoint_format` (npz | tf | both); both containers hold the same variables."""
    from avsi_b200 import checkpoint, training

    class FakeModel(object):
        optimizer_choice = 'sgd'                               # no Adam slots to export
        all_vars = {'av-blstm/logits/weights': np.arange(12, dtype=np.float32).reshape(3, 4),
                    'av-blstm/Variable': np.asarray(7, np.int32)}
    for fmt, files in (('npz', ['ck.npz']), ('tf', ['ck.index', 'ck.data-00000-of-00001']),
                       ('both', ['ck.npz', 'ck.index', 'ck.data-00000-of-00001'])):
        d = tmp_path / fmt
        d.mkdir()
        training._save(FakeModel(), str(d / 'ck'), {'checkpoint_format': fmt})
        assert sorted(os.listdir(str(d))) == sorted(files + (['checkpoint'] if fmt != 'npz' else []))
        got = checkpoint.load(str(d / 'ck'))
        assert set(got) == set(FakeModel.all_vars)
        for k, v in FakeModel.all_vars.items():
            assert np.array_equal(got[k], v) and got[k].dtype == v.dtype
    with pytest.raises(SystemExit):
        training._save(FakeModel(), str(tmp_path / 'x'), {'checkpoint_format': 'hdf5'})
