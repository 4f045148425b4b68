"""Reader (and a minimal writer) for TensorFlow's V2 checkpoint format -- what `tf.train.Saver` writes and the reference
restores with `saver.restore(sess, path)` (sunrgbd_detection/test_semisup.py:158-159, train_semisup_adv.py:224-237, 450-467) --
without TensorFlow: `load_checkpoint(prefix)` returns the `{variable name: ndarray}` dict `runtime.VariableStore` takes
(SURVEY 8f rank 5).

Format (restated from TensorFlow's published tensor_bundle / table sources; no TensorFlow-written file exists in this
environment, so the reader is pinned by (a) tests/golden/tf_ckpt_handmade.*, a fixture assembled byte by byte by an
independent script in TensorFlow's own layout -- restart interval 16, shortened separator keys in the index block, a
DT_STRING entry, bit-serial CRC --, (b) round trips through the writer below, and (c) independent implementations that ARE in the
image: TensorBoard's bundled copy of TensorFlow's generated protos (types.proto's DataType enum, TensorShapeProto) and its CRC32C,
with every BundleEntryProto also parsed by google.protobuf (tests/test_tf_checkpoint_cpu.py)):

* `<prefix>.index` is a LevelDB-style sorted table.  Footer = last 48 bytes: metaindex BlockHandle, index BlockHandle
  (each two varint64: offset, size), zero padding to 40 bytes, 8-byte little-endian magic 0xdb4775248b80fb57.
  A block = entries `varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value`, then the restart
  offsets (uint32 each) and their count (uint32); on disk every block is followed by a 5-byte trailer (compression type,
  masked crc32c).  TensorBundle writes uncompressed blocks (type 0); snappy blocks (type 1) are rejected.
  The index block maps separator keys to the BlockHandles of the data blocks.
* key "" -> BundleHeaderProto (num_shards = 1, endianness = 2 (0 = little), version = 3);
  key <tensor name> -> BundleEntryProto (dtype = 1, shape = 2 {dim = 2 {size = 1}}, shard_id = 3, offset = 4, size = 5,
  crc32c = 6 (fixed32), slices = 7).
* `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian row-major bytes of every tensor at [offset, offset + size).

`remap_scope` covers the reference's restore maps: pretrained F-PointNet variables saved without a scope are loaded into
`class_agnostic/...`, BoxPC variables into `D_boxpc_branch/...`.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------------------------------------ primitives
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_CRC_TABLE = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of the table trailers and of BundleEntryProto.crc32c."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = crc ^ 0xffffffff
    for b in bytes(data):
        c = _CRC_TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def masked_crc(data):
    """leveldb / TF crc mask: rotate right by 15 bits and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


def _parse_proto(buf):
    """Flat protobuf field scan -> {field number: [values]} (varint -> int, fixed32/64 -> int, length-delimited -> bytes)."""
    out, pos, n = {}, 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.setdefault(field, []).append(v)
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# ------------------------------------------------------------------------------------------------ table reader
def _read_block(buf, offset, size, verify=True):
    block = buf[offset:offset + size]
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError('compressed table block (type %d): TensorBundle index files are written uncompressed' % ctype)
    if verify:
        want = struct.unpack_from('<I', buf, offset + size + 1)[0]
        if masked_crc(buf[offset:offset + size + 1]) != want:
            raise ValueError('index block checksum mismatch at offset %d' % offset)
    return block


def _block_entries(block):
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix, verify=True):
    """-> (header dict, {name: entry dict(dtype, shape, shard_id, offset, size, crc32c)})"""
    if not os.path.exists(prefix + '.index') and os.path.isfile(prefix):
        raise ValueError('%s is a single-file (V1, tf.train.SaverDef.V1) checkpoint; only the V2 format (<prefix>.index + '
                         '<prefix>.data-*) is read here -- re-save it with tf.train.Saver(write_version=V2)' % prefix)
    buf = open(prefix + '.index', 'rb').read()
    if len(buf) < 48 or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError('%s.index is not a TensorFlow V2 checkpoint index (bad magic)' % prefix)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex offset
    _, pos = _varint(footer, pos)          # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(buf, off, size, verify)):
            f = _parse_proto(value)
            if key == b'':
                header = dict(num_shards=f.get(1, [1])[0], endianness=f.get(2, [0])[0])
                continue
            shape = []
            if 2 in f:
                sp = _parse_proto(f[2][0])
                shape = [_signed64(_parse_proto(d).get(1, [0])[0]) for d in sp.get(2, [])]
            if 7 in f:
                raise ValueError('%s: partitioned (sliced) variables are not supported' % key.decode())
            entries[key.decode()] = dict(dtype=f.get(1, [0])[0], shape=shape, shard_id=f.get(3, [0])[0], offset=f.get(4, [0])[0],
                                         size=f.get(5, [0])[0], crc32c=f.get(6, [None])[0])
    if header is None:
        raise ValueError('%s.index has no bundle header entry' % prefix)
    if header['endianness'] != 0:
        raise ValueError('big-endian checkpoints are not supported')
    return header, entries


def list_variables(prefix):
    """[(name, shape)] like tf.train.list_variables."""
    _, entries = read_index(prefix)
    return [(k, list(v['shape'])) for k, v in sorted(entries.items())]


def load_checkpoint(prefix, names=None, verify_data=False):
    """{name: ndarray} for every (or the named) variable of the checkpoint `prefix` (the path given to saver.save / restore)."""
    header, entries = read_index(prefix)
    shards = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e['dtype'] not in DTYPES:
            continue                                  # strings (e.g. Saver bookkeeping) and exotic types are skipped
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap('%s.data-%05d-of-%05d' % (prefix, sid, header['num_shards']), dtype=np.uint8, mode='r')
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if verify_data and e['crc32c'] is not None and masked_crc(raw.tobytes()) != e['crc32c']:
            raise ValueError('%s: tensor data checksum mismatch' % name)
        arr = np.frombuffer(raw.tobytes(), dtype=DTYPES[e['dtype']])
        count = int(np.prod(e['shape'])) if e['shape'] else 1
        if arr.size != count:
            raise ValueError('%s: %d bytes do not match shape %s' % (name, e['size'], e['shape']))
        out[name] = arr.reshape(e['shape']).copy()
    if names is not None:
        missing = [n for n in names if n not in out]
        if missing:
            raise KeyError('not in checkpoint: %s' % missing)
    return out


def remap_scope(variables, prefix, only=None):
    """The reference restores un-scoped pretrained variables into a scoped graph (train_semisup_adv.py:224-237: '' ->
    'class_agnostic/'; :450-467: '' -> 'D_boxpc_branch/').  Returns {prefix + name: array} for the names accepted by
    `only(name)` (all when None); Saver bookkeeping / optimizer slots ('beta1_power', '.../Adam', '.../Adam_1') are dropped."""
    out = {}
    for k, v in variables.items():
        if k in ('beta1_power', 'beta2_power', 'global_step') or k.endswith('/Adam') or k.endswith('/Adam_1'):
            continue
        if only is not None and not only(k):
            continue
        out[prefix + k] = v
    return out


# ------------------------------------------------------------------------------------------------ writer (tests, export)
def _put_proto_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v & 0xffffffffffffffff)


def _put_proto_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b'', restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def save_checkpoint(prefix, variables, block_size=4096):
    """Writes {name: ndarray} as a single-shard V2 checkpoint (`prefix.index`, `prefix.data-00000-of-00001`)."""
    d = os.path.dirname(prefix)
    if d and not os.path.isdir(d):
        os.makedirs(d)
    items = []
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        off = 0
        for name in sorted(variables):
            a = np.asarray(variables[name], order='C')        # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in DTYPE_IDS:
                raise ValueError('%s: dtype %s has no TensorFlow type id here' % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            # canonical proto3, as TensorFlow's C++ serialiser writes it: scalar fields equal to zero are omitted (the first
            # tensor's offset, an empty tensor's size, a zero dimension); the shape message is always present
            nz = lambda field, v: _put_proto_varint(field, v) if v else b''
            shape = b''.join(_put_proto_bytes(2, nz(1, int(s))) for s in a.shape)
            crc = masked_crc(raw)
            entry = _put_proto_varint(1, DTYPE_IDS[a.dtype]) + _put_proto_bytes(2, shape) + nz(4, off) + nz(5, len(raw)) + \
                ((_put_varint((6 << 3) | 5) + struct.pack('<I', crc)) if crc else b'')
            items.append((name.encode(), entry))
            off += len(raw)
    header = _put_proto_varint(1, 1) + _put_proto_bytes(3, _put_proto_varint(1, 1))       # num_shards = 1, version.producer = 1
    items = [(b'', header)] + items                                                        # "" sorts first
    out, index = bytearray(), _BlockBuilder(restart_interval=1)

    def flush(bb):
        block = bb.finish()
        handle = _put_varint(len(out)) + _put_varint(len(block))
        out.extend(block + b'\x00' + struct.pack('<I', masked_crc(block + b'\x00')))
        return handle
    bb = _BlockBuilder()
    for key, value in items:
        bb.add(key, value)
        if len(bb.buf) >= block_size:
            index.add(key, flush(bb))            # the last key of the block is a valid separator
            bb = _BlockBuilder()
    if bb.count:
        index.add(bb.last, flush(bb))
    meta_handle = flush(_BlockBuilder())
    index_handle = flush(index)
    footer = meta_handle + index_handle
    out.extend(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC))
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
