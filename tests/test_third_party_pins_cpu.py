"""Pins of the on-disk formats against INDEPENDENT implementations that exist in this image (TensorFlow itself does
not, and cannot be installed: no network, no wheel in /opt/wheelhouse).

* TFRecord framing (length, masked CRC-32C of the length, payload, masked CRC-32C of the payload): tensorboard's
  RecordWriter / PyRecordReader -- TensorFlow-project code that TensorBoard ships so that it can read and write event
  files without TensorFlow -- writes what tfrecord_io.read_records must read, and reads what write_records writes.
* The wire format of tf.train.SequenceExample (tensorflow/core/example/{feature,example}.proto) and of the tensor
  bundle's BundleHeaderProto / BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto): the message schemas
  are rebuilt from their published .proto definitions with the protobuf runtime (google.protobuf descriptors), and the
  official encoder / decoder is held against the hand-written ones of tfrecord_io.py / tf_bundle.py in both directions,
  on a sample laid out exactly like tfrecord_utils.py:19-41 builds it.

What stays unpinned: the sorted-string-table container of the checkpoint `.index` file (no LevelDB / TF reader in the
image) -- covered by round trips and hand-assembled blocks in tests/test_tf_bundle_cpu.py only."""
import os
import struct

import numpy as np
import pytest


_POOL = {}


def _pool():
    """tf.train.SequenceExample and the tensor-bundle protos as dynamic protobuf messages (built once)."""
    if not _POOL:
        _POOL.update(_build_pool())
    return _POOL


def _build_pool():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    f = descriptor_pb2.FileDescriptorProto()
    f.name = 'avsi_test/tf_formats.proto'
    f.package = 'tensorflow'
    f.syntax = 'proto3'
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name):
        m = f.message_type.add()
        m.name = name
        return m

    def field(m, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        fd = m.field.add()
        fd.name, fd.number, fd.type, fd.label = name, number, ftype, label
        if type_name:
            fd.type_name = type_name
        if packed is not None:
            fd.options.packed = packed
        if oneof is not None:
            fd.oneof_index = oneof
        return fd

    def map_field(m, name, number, value_type_name):
        entry = m.nested_type.add()
        entry.name = ''.join(p.capitalize() for p in name.split('_')) + 'Entry'
        entry.options.map_entry = True
        field(entry, 'key', 1, T.TYPE_STRING)
        field(entry, 'value', 2, T.TYPE_MESSAGE, type_name=value_type_name)
        field(m, name, number, T.TYPE_MESSAGE, T.LABEL_REPEATED, '.tensorflow.%s.%s' % (m.name, entry.name))
    # tensorflow/core/example/feature.proto
    field(msg('BytesList'), 'value', 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(msg('FloatList'), 'value', 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(msg('Int64List'), 'value', 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    feat = msg('Feature')
    feat.oneof_decl.add().name = 'kind'
    field(feat, 'bytes_list', 1, T.TYPE_MESSAGE, type_name='.tensorflow.BytesList', oneof=0)
    field(feat, 'float_list', 2, T.TYPE_MESSAGE, type_name='.tensorflow.FloatList', oneof=0)
    field(feat, 'int64_list', 3, T.TYPE_MESSAGE, type_name='.tensorflow.Int64List', oneof=0)
    map_field(msg('Features'), 'feature', 1, '.tensorflow.Feature')
    field(msg('FeatureList'), 'feature', 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, '.tensorflow.Feature')
    map_field(msg('FeatureLists'), 'feature_list', 1, '.tensorflow.FeatureList')
    # tensorflow/core/example/example.proto
    se = msg('SequenceExample')
    field(se, 'context', 1, T.TYPE_MESSAGE, type_name='.tensorflow.Features')
    field(se, 'feature_lists', 2, T.TYPE_MESSAGE, type_name='.tensorflow.FeatureLists')
    # tensorflow/core/framework/tensor_shape.proto, versions.proto, protobuf/tensor_bundle.proto
    shp = msg('TensorShapeProto')
    dim = shp.nested_type.add()
    dim.name = 'Dim'
    field(dim, 'size', 1, T.TYPE_INT64)
    field(dim, 'name', 2, T.TYPE_STRING)
    field(shp, 'dim', 2, T.TYPE_MESSAGE, T.LABEL_REPEATED, '.tensorflow.TensorShapeProto.Dim')
    field(shp, 'unknown_rank', 3, T.TYPE_BOOL)
    ver = msg('VersionDef')
    field(ver, 'producer', 1, T.TYPE_INT32)
    field(ver, 'min_consumer', 2, T.TYPE_INT32)
    hdr = msg('BundleHeaderProto')
    field(hdr, 'num_shards', 1, T.TYPE_INT32)
    field(hdr, 'endianness', 2, T.TYPE_INT32)            # enum Endianness {LITTLE = 0; BIG = 1}: same wire type
    field(hdr, 'version', 3, T.TYPE_MESSAGE, type_name='.tensorflow.VersionDef')
    ent = msg('BundleEntryProto')
    field(ent, 'dtype', 1, T.TYPE_INT32)                 # enum DataType
    field(ent, 'shape', 2, T.TYPE_MESSAGE, type_name='.tensorflow.TensorShapeProto')
    field(ent, 'shard_id', 3, T.TYPE_INT32)
    field(ent, 'offset', 4, T.TYPE_INT64)
    field(ent, 'size', 5, T.TYPE_INT64)
    field(ent, 'crc32c', 6, T.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    get = getattr(message_factory, 'GetMessageClass', None)
    if get is None:                                     # older protobuf
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return {n: get(pool.FindMessageTypeByName('tensorflow.' + n))
            for n in ('SequenceExample', 'BundleHeaderProto', 'BundleEntryProto')}


def _sample(rng, T=7, n_audio=40, L=4):
    return dict(seq_len=T, lab_len=L, wav=np.round(rng.normal(0, 3000, n_audio)).astype(np.float32),
                video=rng.standard_normal((T, 136)).astype(np.float32), mask=(rng.uniform(size=(T, 257)) > 0.3).astype(np.float32),
                labels=np.array([3, 0, 32, 7] + [0] * (50 - L), np.float32), path='s1_train/bbaf2n')


def _reference_style_example(cls, s):
    """tfrecord_utils.py:19-41 (serialize_sample_fixed), statement for statement, on the protobuf-runtime message."""
    example = cls()
    example.context.feature['sequence_length'].int64_list.value.append(s['seq_len'])
    example.context.feature['labels_length'].int64_list.value.append(s['lab_len'])
    example.context.feature['target_audio_wav'].float_list.value.extend(s['wav'].tolist())
    example.context.feature['sample_path'].bytes_list.value.append(s['path'].encode())
    fl_mask = example.feature_lists.feature_list['mask']
    fl_video = example.feature_lists.feature_list['video_features']
    fl_labels = example.feature_lists.feature_list['labels']
    for v in s['video']:
        fl_video.feature.add().float_list.value.extend(v.tolist())
    for m in s['mask']:
        fl_mask.feature.add().float_list.value.extend(m.tolist())
    for l in s['labels']:
        fl_labels.feature.add().float_list.value.append(float(l))
    return example


def test_sequence_example_written_by_the_protobuf_runtime_is_parsed(tmp_path):
    from avsi_b200 import tfrecord_io as tio
    s = _sample(np.random.default_rng(0))
    data = _reference_style_example(_pool()['SequenceExample'], s).SerializeToString()
    ctx, seq = tio.parse_sequence_example(data)
    assert int(ctx['sequence_length'][0]) == s['seq_len'] and int(ctx['labels_length'][0]) == s['lab_len']
    assert np.array_equal(np.asarray(ctx['target_audio_wav'], np.float32), s['wav'])
    assert ctx['sample_path'][0] == s['path'].encode()
    assert np.array_equal(np.stack(seq['video_features']).astype(np.float32), s['video'])
    assert np.array_equal(np.stack(seq['mask']).astype(np.float32), s['mask'])
    assert np.array_equal(np.concatenate(seq['labels']).astype(np.float32), s['labels'])
    fast = tio.parse_av_sample(data, len(s['wav']), 257, 136)          # the native parser of the loader, when built
    if fast is not None:
        seq_len, lab_len, wav, path, labels, video, mask = fast
        assert (int(seq_len), int(lab_len), path) == (s['seq_len'], s['lab_len'], s['path'].encode())
        assert np.array_equal(wav.astype(np.float32), s['wav']) and np.array_equal(video, s['video'])
        assert np.array_equal(mask, s['mask']) and np.array_equal(labels, s['labels'])
    # through a TFRecord file written by TensorBoard's writer
    from tensorboard.summary.writer.record_writer import RecordWriter
    path = str(tmp_path / 'tb_written.tfrecord')
    with open(path, 'wb') as fh:
        w = RecordWriter(fh)
        w.write(data)
        w.write(b'')
        w.write(b'x' * 70000)
        w.flush()
    recs = [bytes(r) for r in tio.read_records(path, verify=True)]
    assert recs == [data, b'', b'x' * 70000]


def test_sequence_example_we_write_is_read_by_protobuf_and_tensorboard(tmp_path):
    from avsi_b200 import tfrecord_io as tio
    s = _sample(np.random.default_rng(1))
    data = tio.serialize_sample_fixed(s['seq_len'], s['lab_len'], s['wav'], s['video'], s['mask'], s['labels'], s['path'])
    ex = _pool()['SequenceExample']()
    ex.ParseFromString(bytes(data))
    ref = _reference_style_example(_pool()['SequenceExample'], s)
    assert list(ex.context.feature['sequence_length'].int64_list.value) == [s['seq_len']]
    assert list(ex.context.feature['sample_path'].bytes_list.value) == [s['path'].encode()]
    assert np.array_equal(np.asarray(ex.context.feature['target_audio_wav'].float_list.value, np.float32), s['wav'])
    for name, want in (('video_features', s['video']), ('mask', s['mask'])):
        got = np.stack([np.asarray(f.float_list.value, np.float32) for f in ex.feature_lists.feature_list[name].feature])
        assert np.array_equal(got, want), name
    assert [f.float_list.value[0] for f in ex.feature_lists.feature_list['labels'].feature] == s['labels'].tolist()
    # same message as the reference-style construction (map entries may be ordered differently on the wire)
    assert ex == ref
    path = str(tmp_path / 'ours.tfrecord')
    tio.write_records(path, [data, b'', b'abc'])
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import PyRecordReader_New
    rd = PyRecordReader_New(path)
    got = []
    while True:
        try:
            rd.GetNext()
        except Exception:                                     # OutOfRangeError at EOF; a CRC mismatch would raise DataLossError earlier
            break
        got.append(bytes(rd.record()))
    assert got == [bytes(data), b'', b'abc']


def test_bundle_protos_against_the_protobuf_runtime(tmp_path):
    from avsi_b200 import tf_bundle
    P = _pool()
    # an entry as TensorFlow would write it -> our parser
    e = P['BundleEntryProto']()
    e.dtype, e.shard_id, e.offset, e.size, e.crc32c = 1, 0, 123456789012, 4000, 0xDEADBEEF
    for d in (643, 1000):
        e.shape.dim.add().size = d
    got = tf_bundle._parse_entry(e.SerializeToString())
    assert (got['dtype'], got['shape'], got['shard_id'], got['offset'], got['size'], got['crc32c'], got['sliced']) == \
        (1, [643, 1000], 0, 123456789012, 4000, 0xDEADBEEF, False)
    scalar = P['BundleEntryProto']()                         # rank-0 int32 (the global step `Variable`): empty shape message
    scalar.dtype, scalar.size, scalar.crc32c = 3, 4, 7
    scalar.shape.SetInParent()
    got = tf_bundle._parse_entry(scalar.SerializeToString())
    assert got['shape'] == [] and got['dtype'] == 3 and got['offset'] == 0
    # a bundle we write -> the official decoder on every index value
    rng = np.random.default_rng(2)
    variables = {'av-blstm/logits/weights': rng.standard_normal((500, 257)).astype(np.float32),
                 'av-blstm/logits/biases': rng.standard_normal(257).astype(np.float32),
                 'av-blstm/Variable': np.asarray(41, np.int32)}
    prefix = str(tmp_path / 'sinet')
    tf_bundle.write_bundle(prefix, variables)
    table = dict(tf_bundle.read_table(prefix + '.index'))
    hdr = P['BundleHeaderProto']()
    hdr.ParseFromString(bytes(table[b'']))                  # the header lives under the empty key
    assert hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer >= 1
    data = open(prefix + '.data-00000-of-00001', 'rb').read()
    from tensorboard.compat.proto import types_pb2
    for name, arr in variables.items():
        ent = P['BundleEntryProto']()
        ent.ParseFromString(bytes(table[name.encode()]))
        assert ent.dtype == {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.int32): types_pb2.DT_INT32}[arr.dtype]
        assert [d.size for d in ent.shape.dim] == list(arr.shape)
        raw = data[ent.offset:ent.offset + ent.size]
        assert np.array_equal(np.frombuffer(raw, arr.dtype).reshape(arr.shape), arr)
        # crc32c field = masked CRC-32C of the tensor bytes (tensor_bundle.cc); CRC known answer in test_tf_bundle_cpu.py
        c = tf_bundle.crc32c(raw)
        assert ent.crc32c == ((((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff)


def test_mel_matrix_against_transformers_mel_filter_bank():
    """tf.signal.linear_to_mel_weight_matrix(80, 257, 16000, 125, 7600) (audio_processing.py:63-65): HTK mel scale,
    triangles linear IN MEL (not in Hz, unlike librosa / torchaudio defaults), DC row zero.  transformers.audio_utils
    implements the same definition independently (`mel_scale='htk', triangularize_in_mel_space=True`)."""
    transformers = pytest.importorskip('transformers')
    from transformers.audio_utils import mel_filter_bank
    from avsi_b200 import audio_processing as ap
    from oracle import stft as ostft
    for n_mel, lo, hi in ((80, 125.0, 7600.0), (40, 20.0, 4000.0)):
        ref = mel_filter_bank(num_frequency_bins=257, num_mel_filters=n_mel, min_frequency=lo, max_frequency=hi,
                              sampling_rate=16000, norm=None, mel_scale='htk', triangularize_in_mel_space=True)
        for m in (ostft.linear_to_mel_weight_matrix(n_mel, 257, 16000, lo, hi),
                  ap.linear_to_mel_weight_matrix(n_mel, 257, 16000, lo, hi)):
            assert m.shape == (257, n_mel) and np.abs(np.asarray(m, np.float64) - ref).max() < 1e-6
    # and it is NOT the Hz-domain triangle bank
    hz = mel_filter_bank(num_frequency_bins=257, num_mel_filters=80, min_frequency=125.0, max_frequency=7600.0,
                         sampling_rate=16000, norm=None, mel_scale='htk', triangularize_in_mel_space=False)
    assert np.abs(hz - ostft.linear_to_mel_weight_matrix(80, 257, 16000, 125.0, 7600.0)).max() > 1e-3


def test_mfcc_against_scipy_dct():
    """tf.signal.mfccs_from_log_mel_spectrograms = DCT-II (unnormalised, factor 2) * rsqrt(2 * num_mel_bins), first
    num_mfcc coefficients kept (audio_processing.py:75-82)."""
    from scipy.fft import dct
    from oracle import stft as ostft
    rng = np.random.default_rng(5)
    logmel = rng.standard_normal((3, 11, 80))
    ref = dct(logmel, type=2, norm=None, axis=-1) / np.sqrt(2.0 * 80)
    got = ostft.get_mfcc(logmel, 13)
    assert np.abs(got - ref[..., :13]).max() < 1e-12


def test_resample_oracle_against_scipy_signal_resample():
    """downsampling (audio_processing.py:9-16) delegates to scipy.signal.resample; scipy is in the image, so the oracle's
    restatement of its real-input branch (bins kept, Nyquist doubling / halving, num / n_in) is held against scipy itself
    on even and odd lengths in both directions, and at the GRID lengths (150 000 samples at 50 kHz -> 48 000)."""
    from scipy import signal
    from oracle import resample as ores
    rng = np.random.default_rng(17)
    for nx, num in ((150000, 48000), (1000, 320), (1001, 320), (1000, 321), (999, 333), (320, 1000), (321, 1000),
                    (320, 1001), (7, 3), (64, 64), (5, 8), (6, 4), (2, 1), (1, 5)):
        x = np.round(rng.normal(0, 3000, nx))
        ref = signal.resample(x, num)
        got = ores.resample(x, num)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (nx, num)
    wav = np.round(rng.normal(0, 3000, 150000)).astype(np.int16)
    ref = signal.resample(wav, int(16000 * (len(wav) / 50000.0)))
    assert np.abs(ores.downsampling(wav, 50000, 16000) - ref).max() < 1e-8
    assert ores.downsampling(wav, 16000, 16000) is wav


def test_delta_preemphasis_edit_distance_against_torchaudio():
    """Three more independent implementations that ship in the image (torchaudio): `delta` (audio_processing.py:84-93: SYMMETRIC
    padding one frame at a time = edge replication, sum_i i (x[t+i] - x[t-i]) / (2 sum i^2)) = compute_deltas(win_length
    2N + 1, mode 'replicate'); `preemphasis` (:19-22); the Levenshtein distance behind `per` (tf.edit_distance,
    models.py:1718)."""
    torchaudio = pytest.importorskip('torchaudio')
    import torch
    from avsi_b200 import models
    from oracle import stft as ostft
    rng = np.random.default_rng(29)
    x = rng.standard_normal((3, 41, 13))
    for N in (1, 2, 3):
        ref = torchaudio.functional.compute_deltas(torch.from_numpy(x).permute(0, 2, 1), win_length=2 * N + 1,
                                                   mode='replicate').permute(0, 2, 1).numpy()
        assert np.abs(ostft.delta(x, N) - ref).max() < 1e-12
    wav = rng.standard_normal((2, 500))
    ref = torchaudio.functional.preemphasis(torch.from_numpy(wav), coeff=0.95).numpy()
    assert np.abs(ostft.preemphasis(wav, 0.95) - ref).max() < 1e-12
    for _ in range(50):
        a = rng.integers(0, 5, rng.integers(0, 12)).tolist()
        b = rng.integers(0, 5, rng.integers(0, 12)).tolist()
        assert models.edit_distance(a, b) == torchaudio.functional.edit_distance(a, b)
