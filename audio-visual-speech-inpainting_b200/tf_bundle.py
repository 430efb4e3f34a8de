"""TensorFlow checkpoint bundles without TensorFlow.

`tf.train.Saver().save(sess, '<dir>/sinet')` (training.py:114,267,335) writes `sinet.index` +
`sinet.data-00000-of-00001` (+ a `checkpoint` state file).  The formats, restated from their published definitions:

  *.index   a leveldb-style sorted string table: data blocks of prefix-compressed (key, value) entries
            [shared varint32 | non_shared varint32 | value_len varint32 | key suffix | value], a restart array
            (uint32 offsets + count) closing every block, a 5-byte trailer per block (compression type, masked CRC-32C
            of block + type), a meta-index block, an index block whose values are BlockHandles (varint64 offset, size),
            and a 48-byte footer (two handles padded to 40 bytes + magic 0xdb4775248b80fb57).  The Saver writes it
            uncompressed.  Key "" holds BundleHeaderProto {num_shards=1, endianness=2, version=3}; every other key is
            a variable name holding BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32,
            masked CRC-32C of the tensor bytes)}.
  *.data-*  the raw little-endian tensor bytes, back to back in key order.

No TF bundle ships with the reference and TF is not installable here, so this reader is NOT pinned to a genuine
file: it is checked by round trips through the writer below (which emits restart points and prefix compression the
way the table builder does), by the CRCs, and by hand-built blocks in tests/test_tf_bundle_cpu.py."""
import os
import struct

import numpy as np

from . import tfrecord_io as _io

MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items()}


def crc32c(data):
    """Masked-CRC building block; the C routine of the library when it is built (17 MB tensors), pure Python otherwise."""
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    if len(data) >= 4096:
        try:
            from . import _lib
            return int(_lib.load().avsi_crc32c_host(data, len(data), 0))
        except Exception:
            pass
    return _io.crc32c(data)


def _mask(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- table reading ------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify=True):
    body = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored, = struct.unpack_from('<I', buf, offset + size + 1)
    if verify and _mask(crc32c(bytes(body) + bytes([ctype]))) != stored:
        raise ValueError('checkpoint index: block checksum mismatch at offset %d' % offset)
    if ctype != 0:
        raise ValueError('checkpoint index: compressed block (type %d) not supported; tf.train.Saver writes none' % ctype)
    return body


def _block_entries(block):
    """Yield (key, value) of one block, undoing the prefix compression."""
    n_restarts, = struct.unpack_from('<I', block, len(block) - 4)
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _io._varint(block, pos)
        non_shared, pos = _io._varint(block, pos)
        vlen, pos = _io._varint(block, pos)
        if shared > len(key):
            raise ValueError('checkpoint index: corrupt key prefix')
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _handle(buf, pos=0):
    off, pos = _io._varint(buf, pos)
    size, pos = _io._varint(buf, pos)
    return off, size, pos


def read_table(path, verify=True):
    """All (key, value) pairs of a sorted string table file, in key order."""
    with open(path, 'rb') as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError('%s is not a checkpoint index (bad magic)' % path)
    footer = buf[len(buf) - 48:]
    _, _, pos = _handle(footer)                               # meta-index handle (unused)
    ioff, isize, _ = _handle(footer, pos)
    out = []
    for _, hv in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, size, _ = _handle(hv)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


def _parse_entry(value):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, wt, v in _io._fields(memoryview(value)):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, v2 in _io._fields(v):
                if f2 == 2:                                    # TensorShapeProto.dim
                    size = 0
                    for f3, _, v3 in _io._fields(v2):
                        if f3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = struct.unpack('<I', bytes(v))[0]
        elif f == 7:
            e['sliced'] = True
    return e


def read_bundle(prefix, verify=True):
    """`prefix` as given to saver.save / saver.restore -> {variable name: numpy array}."""
    index = prefix + '.index'
    if not os.path.exists(index):
        raise ValueError('%s is not a valid checkpoint' % prefix)
    entries = read_table(index, verify)
    if not entries or entries[0][0] != b'':
        raise ValueError('%s: missing bundle header' % index)
    num_shards, endian = 1, 0
    for f, _, v in _io._fields(memoryview(entries[0][1])):
        if f == 1:
            num_shards = v
        elif f == 2:
            endian = v
    if endian != 0:
        raise ValueError('big-endian bundles are not supported')
    shards = {}
    out = {}
    for key, value in entries[1:]:
        e = _parse_entry(value)
        name = key.decode()
        if e['sliced']:
            raise ValueError('%s: partitioned variable %s not supported' % (index, name))
        if e['dtype'] not in DTYPES:
            raise ValueError('%s: dtype %d of %s not supported' % (index, e['dtype'], name))
        if e['shard_id'] not in shards:
            with open('%s.data-%05d-of-%05d' % (prefix, e['shard_id'], num_shards), 'rb') as f:
                shards[e['shard_id']] = f.read()
        raw = shards[e['shard_id']][e['offset']:e['offset'] + e['size']]
        dt = np.dtype(DTYPES[e['dtype']])
        if len(raw) != e['size'] or e['size'] != int(np.prod(e['shape'], dtype=np.int64)) * dt.itemsize:
            raise ValueError('%s: size mismatch for %s' % (index, name))
        if verify and e['crc32c'] is not None and _mask(crc32c(raw)) != e['crc32c']:
            raise ValueError('%s: tensor checksum mismatch for %s' % (index, name))
        out[name] = np.frombuffer(raw, dt).reshape(e['shape']).copy()
    return out


# ---- table writing ------------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b''
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _io._enc_varint(shared) + _io._enc_varint(len(key) - shared) + _io._enc_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _ld(field, payload):
    return _io._enc_varint((field << 3) | 2) + _io._enc_varint(len(payload)) + payload


def _vi(field, value):
    return _io._enc_varint(field << 3) + _io._enc_varint(value)


def write_table(path, items, block_size=4096):
    """items: sorted [(key bytes, value bytes)] -> sorted string table file."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        out.extend(b'\x00' + struct.pack('<I', _mask(crc32c(block_bytes + b'\x00'))))
        return _io._enc_varint(off) + _io._enc_varint(len(block_bytes))
    blk = _BlockBuilder()
    for i, (k, v) in enumerate(items):
        if i and k <= items[i - 1][0]:
            raise ValueError('table keys must be strictly increasing')
        blk.add(k, v)
        if len(blk.buf) >= block_size:
            index.add(k, emit(blk.finish()))
            blk = _BlockBuilder()
    if blk.buf or not items:
        index.add(items[-1][0] if items else b'', emit(blk.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    out.extend(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC))
    with open(path, 'wb') as f:
        f.write(bytes(out))


def write_bundle(prefix, variables):
    """{name: array} -> `<prefix>.index`, `<prefix>.data-00000-of-00001` and the `checkpoint` state file of its
    directory, readable by tf.train.Saver.restore / tf.train.load_checkpoint."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _vi(1, 1) + _vi(2, 0) + _ld(3, _vi(1, 1))         # num_shards = 1, little endian, version.producer = 1
    items = [(b'', header)]
    offset = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in sorted(variables, key=lambda s: s.encode()):
            a = np.asarray(variables[name])                  # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in DTYPE_ENUM:
                raise ValueError('dtype %s of %s has no TensorFlow counterpart here' % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
            shape = b''.join(_ld(2, _vi(1, int(d))) for d in a.shape)
            entry = _vi(1, DTYPE_ENUM[a.dtype]) + _ld(2, shape)
            if offset:
                entry += _vi(4, offset)
            entry += _vi(5, len(raw)) + _io._enc_varint((6 << 3) | 5) + struct.pack('<I', _mask(crc32c(raw)))
            items.append((name.encode(), entry))
            f.write(raw)
            offset += len(raw)
    write_table(prefix + '.index', items)
    base = os.path.basename(prefix)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def latest_checkpoint(folder):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file, or None."""
    state = os.path.join(folder, 'checkpoint')
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith('model_checkpoint_path:'):
                name = line.split(':', 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(folder, name)
    return None
