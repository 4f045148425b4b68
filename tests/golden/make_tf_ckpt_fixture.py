"""Hand-assembles a small TensorFlow V2 checkpoint (tests/golden/tf_ckpt_handmade.{index,data-00000-of-00001}) WITHOUT
importing transferable3d_b200.tf_checkpoint: every structure is spelled out here byte by byte from the published formats
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{format,block_builder,table_builder}.cc,
tensorflow/core/protobuf/tensor_bundle.proto, tensorflow/core/framework/{types,tensor_shape,versions}.proto), in the way
TensorFlow's own writer lays them out -- which differs from what `save_checkpoint` in this repo emits:

  * restart interval 16 inside the data blocks, with a second restart point (20 entries in block 0), so that entries with
    shared = 0 appear mid-block and the restart array has two offsets;
  * prefix-compressed keys whose shared length crosses a '/' ("a/conv1/bn/beta" -> "a/conv1/bn/gamma": shared 11);
  * index-block keys are SHORTENED SEPARATORS (leveldb FindShortestSeparator), not the last key of the block:
    block 0 ends with "a/conv9/weights", block 1 starts with "c/fc1/biases" -> separator "b" (first differing byte + 1,
    truncated there); the last block's key is the short successor "t" of its last key "shape_probe";
  * BundleHeaderProto with version {producer: 1} and the endianness field omitted (default little);
  * BundleEntryProto with fields in TensorFlow's order; the scalar int64 has an EMPTY shape message; offsets grow in key order;
  * a DT_STRING entry (Saver bookkeeping style) that a float reader must skip;
  * two shards' worth of naming is NOT used: num_shards = 1.

The CRC here is a bit-serial CRC-32C, independent of the table-driven one in the product.
Run:  python tests/golden/make_tf_ckpt_fixture.py      (rewrites the two fixture files and prints their sha256)
"""
import hashlib
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def crc32c_bitwise(data):
    c = 0xffffffff
    for byte in data:
        c ^= byte
        for _ in range(8):
            c = (c >> 1) ^ (0x82f63b78 & -(c & 1))
    return c ^ 0xffffffff


def mask(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


def vint(v):
    out = b''
    while v >= 0x80:
        out += bytes([(v & 0x7f) | 0x80])
        v >>= 7
    return out + bytes([v])


DT_FLOAT, DT_INT32, DT_STRING, DT_INT64 = 1, 3, 7, 9


def entry_proto(dtype, shape, offset, size, crc):
    p = b'\x08' + vint(dtype)                                        # field 1 varint: dtype
    sh = b''.join(b'\x12' + vint(len(b'\x08' + vint(d))) + b'\x08' + vint(d) for d in shape)    # TensorShapeProto.dim (field 2) {size (1)}
    p += b'\x12' + vint(len(sh)) + sh                                # field 2: shape (present, possibly empty)
    if offset:
        p += b'\x20' + vint(offset)                                  # field 4: offset (proto3: zero omitted)
    p += b'\x28' + vint(size)                                        # field 5: size
    p += b'\x35' + struct.pack('<I', crc)                            # field 6: fixed32 masked crc32c of the bytes
    return p


def build():
    rng = np.random.RandomState(20260117)
    tensors = []                                                     # (name, dtype id, array or raw bytes)
    for i in range(1, 4):
        cin, cout = (3, 8) if i == 1 else (8, 8)
        tensors.append(('a/conv%d/biases' % i, DT_FLOAT, rng.randn(cout).astype('<f4')))
        tensors.append(('a/conv%d/bn/beta' % i, DT_FLOAT, rng.randn(cout).astype('<f4')))
        tensors.append(('a/conv%d/bn/gamma' % i, DT_FLOAT, rng.randn(cout).astype('<f4')))
        tensors.append(('a/conv%d/bn/moving_mean' % i, DT_FLOAT, rng.randn(cout).astype('<f4')))
        tensors.append(('a/conv%d/bn/moving_variance' % i, DT_FLOAT, rng.rand(cout).astype('<f4')))
        tensors.append(('a/conv%d/weights' % i, DT_FLOAT, rng.randn(1, 1, cin, cout).astype('<f4')))
    tensors.append(('a/conv9/weights', DT_FLOAT, rng.randn(1, 1, 2, 5).astype('<f4')))
    tensors.append(('c/fc1/biases', DT_FLOAT, rng.randn(67).astype('<f4')))
    tensors.append(('c/fc1/weights', DT_FLOAT, rng.randn(12, 67).astype('<f4')))
    tensors.append(('global_step', DT_INT64, np.array(4242, dtype='<i8')))
    tensors.append(('save_counter_names', DT_STRING, b'\x05hello'))  # string tensor bytes (length-prefixed), skipped by readers of floats
    tensors.append(('shape_probe', DT_INT32, np.arange(24, dtype='<i4').reshape(2, 3, 4)))
    tensors.sort(key=lambda t: t[0].encode())
    data = b''
    items = [(b'', b'\x08\x01' + b'\x1a\x02\x08\x01')]               # BundleHeaderProto: num_shards = 1, version {producer = 1}
    arrays = {}
    for name, dt, a in tensors:
        raw = a if isinstance(a, bytes) else a.tobytes()
        shape = [] if isinstance(a, bytes) else list(a.shape)
        items.append((name.encode(), entry_proto(dt, shape, len(data), len(raw), mask(crc32c_bitwise(raw)))))
        data += raw
        if not isinstance(a, bytes):
            arrays[name] = a
    # ---- data blocks: block 0 = header + the 19 'a/...' keys (20 entries: restarts at entries 0 and 16), block 1 = the rest
    split = 1 + sum(1 for t in tensors if t[0].startswith('a/'))

    def block(entries):
        buf, restarts, last = b'', [], b''
        for n, (k, v) in enumerate(entries):
            if n % 16 == 0:
                restarts.append(len(buf))
                shared = 0
            else:
                shared = 0
                while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                    shared += 1
            buf += vint(shared) + vint(len(k) - shared) + vint(len(v)) + k[shared:] + v
            last = k
        return buf + b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))

    out = b''

    def emit(b):
        nonlocal out
        handle = vint(len(out)) + vint(len(b))
        out += b + b'\x00' + struct.pack('<I', mask(crc32c_bitwise(b + b'\x00')))
        return handle
    h0 = emit(block(items[:split]))
    h1 = emit(block(items[split:]))
    assert items[split - 1][0] == b'a/conv9/weights' and items[split][0] == b'c/fc1/biases'
    # leveldb separators: FindShortestSeparator("a/conv9/weights", "c/fc1/biases") -> first byte differs and 'a' + 1 < 'c',
    # so the separator is "b"; FindShortSuccessor("shape_probe") -> "t"
    index_entries = [(b'b', h0), (b't', h1)]
    ib = b''
    ir = []
    for k, v in index_entries:                                       # index block: restart interval 1
        ir.append(len(ib))
        ib += vint(0) + vint(len(k)) + vint(len(v)) + k + v
    ib += b''.join(struct.pack('<I', r) for r in ir) + struct.pack('<I', len(ir))
    meta = emit(struct.pack('<I', 0) + struct.pack('<I', 1))         # empty metaindex block: one restart (offset 0), count 1
    idxh = emit(ib)
    footer = meta + idxh
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', 0xdb4775248b80fb57)
    return out, data, arrays


if __name__ == '__main__':
    index, data, arrays = build()
    with open(os.path.join(HERE, 'tf_ckpt_handmade.index'), 'wb') as f:
        f.write(index)
    with open(os.path.join(HERE, 'tf_ckpt_handmade.data-00000-of-00001'), 'wb') as f:
        f.write(data)
    print('index', len(index), hashlib.sha256(index).hexdigest())
    print('data', len(data), hashlib.sha256(data).hexdigest())
