"""-m "not gpu": transferable3d_b200.tf_checkpoint -- TensorFlow V2 checkpoint reader without TensorFlow.
Pinned by: the CRC-32C known-answer vector, the byte layout of a hand-assembled BundleEntryProto and footer, and a
round trip through the writer (multi-block index with prefix-compressed keys, several dtypes, scalars).  No
TensorFlow-written file is available here: the on-disk format is restated from the published sources, see the module
docstring."""
import os
import struct

import numpy as np
import pytest

from transferable3d_b200 import tf_checkpoint as ck


def test_crc32c_known_answers():
    assert ck.crc32c(b'123456789') == 0xE3069283                    # the standard CRC-32C check value
    assert ck.crc32c(b'') == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                       # 32 zero bytes (RFC 3720 B.4)
    c = ck.crc32c(b'123456789')
    assert ck.masked_crc(b'123456789') == ((((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff)


def test_entry_and_footer_byte_layout(tmp_path):
    v = {'w': np.arange(6, dtype=np.float32).reshape(2, 3)}
    prefix = os.path.join(str(tmp_path), 'm.ckpt')
    ck.save_checkpoint(prefix, v)
    idx = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', idx[-8:])[0] == 0xdb4775248b80fb57 and len(idx) >= 48
    # BundleEntryProto of 'w': dtype DT_FLOAT (08 01), shape {dim{size 2} dim{size 3}} (12 08 12 02 08 02 12 02 08 03),
    # offset 0 (omitted: proto3 does not serialise a zero scalar), size 24 (28 18), crc32c fixed32 (35 ....)
    want = bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03, 0x28, 0x18, 0x35])
    assert want in idx
    raw = open(prefix + '.data-00000-of-00001', 'rb').read()
    assert raw == v['w'].tobytes()
    header, entries = ck.read_index(prefix)
    assert header == dict(num_shards=1, endianness=0)
    assert entries['w']['shape'] == [2, 3] and entries['w']['crc32c'] == ck.masked_crc(raw)


def test_round_trip_many_variables(tmp_path):
    rng = np.random.RandomState(0)
    v = {}
    for net in ('class_agnostic/inst_seg', 'class_agnostic/tnet', 'class_agnostic/box_est', 'class_dependent/box_refine'):
        for i in range(1, 12):
            cin, cout = rng.randint(3, 40), rng.randint(3, 40)
            v['%s/conv%d/weights' % (net, i)] = rng.randn(1, 1, cin, cout).astype(np.float32)
            v['%s/conv%d/biases' % (net, i)] = rng.randn(cout).astype(np.float32)
            for s in ('beta', 'gamma', 'moving_mean', 'moving_variance'):
                v['%s/conv%d/bn/%s' % (net, i, s)] = rng.randn(cout).astype(np.float32)
    v['global_step'] = np.array(1234, dtype=np.int64)               # scalar
    v['flags'] = np.array([True, False, True])
    v['half'] = rng.randn(5).astype(np.float16)
    v['beta1_power'] = np.array(0.9, dtype=np.float32)
    v['class_agnostic/tnet/conv1/weights/Adam'] = np.zeros((1, 1, 3, 4), np.float32)
    prefix = os.path.join(str(tmp_path), 'sub', 'model.ckpt')
    ck.save_checkpoint(prefix, v, block_size=512)                   # many data blocks
    back = ck.load_checkpoint(prefix, verify_data=True)
    assert set(back) == set(v)
    for k in v:
        assert back[k].dtype == np.asarray(v[k]).dtype and back[k].shape == np.asarray(v[k]).shape and np.array_equal(back[k], v[k]), k
    names = ck.list_variables(prefix)
    assert [n for n, _ in names] == sorted(v) and dict(names)['global_step'] == []
    some = ck.load_checkpoint(prefix, names=['flags', 'half'])
    assert set(some) == {'flags', 'half'}
    with pytest.raises(KeyError):
        ck.load_checkpoint(prefix, names=['nope'])
    # scope remap of the reference's restore maps; optimizer slots dropped
    r = ck.remap_scope({k: a for k, a in back.items() if not k.startswith('class_dependent')}, 'D_boxpc_branch/',
                       only=lambda n: 'tnet' in n)
    assert all(k.startswith('D_boxpc_branch/class_agnostic/tnet') for k in r) and not any(k.endswith('/Adam') for k in r)
    assert 'D_boxpc_branch/class_agnostic/tnet/conv1/weights' in r


def test_corruption_is_detected(tmp_path):
    prefix = os.path.join(str(tmp_path), 'm.ckpt')
    ck.save_checkpoint(prefix, {'a': np.arange(10, dtype=np.float32), 'b': np.ones((3, 3), np.float32)})
    idx = bytearray(open(prefix + '.index', 'rb').read())
    bad = bytearray(idx)
    bad[-1] ^= 0xff                                                 # magic
    open(prefix + '.index', 'wb').write(bytes(bad))
    with pytest.raises(ValueError):
        ck.load_checkpoint(prefix)
    bad = bytearray(idx)
    bad[3] ^= 0x01                                                  # a byte inside the first data block
    open(prefix + '.index', 'wb').write(bytes(bad))
    with pytest.raises(ValueError):
        ck.load_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(bytes(idx))
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[5] ^= 0x10
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    ck.load_checkpoint(prefix)                                      # data checksums are opt-in (pure-python CRC is slow)
    with pytest.raises(ValueError):
        ck.load_checkpoint(prefix, verify_data=True)


def test_checkpoint_feeds_the_variable_store(tmp_path):
    """weights -> checkpoint -> load_checkpoint gives back the dict runtime.VariableStore / the oracle take."""
    from transferable3d_b200 import weights
    v = weights.make_weights_boxpc()
    prefix = os.path.join(str(tmp_path), 'boxpc.ckpt')
    ck.save_checkpoint(prefix, {k: np.asarray(a) for k, a in v.items()})
    back = ck.load_checkpoint(prefix)
    assert set(back) == set(v) and all(np.array_equal(back[k], np.asarray(v[k])) for k in v)


def test_handmade_fixture_not_produced_by_the_writer():
    """tests/golden/tf_ckpt_handmade.* is assembled byte by byte by tests/golden/make_tf_ckpt_fixture.py (which does not
    import the product module) in TensorFlow's own table layout: restart interval 16 with a mid-block restart, shortened
    separator keys ("b", "t") in the index block, a DT_STRING entry, an empty shape message for the scalar, a bit-serial
    CRC.  The reader must return exactly the arrays the script generated."""
    import hashlib
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    prefix = os.path.join(here, 'tf_ckpt_handmade')
    idx = open(prefix + '.index', 'rb').read()
    dat = open(prefix + '.data-00000-of-00001', 'rb').read()
    assert hashlib.sha256(idx).hexdigest() == 'c8dcc03b2e5a34315d52d97b4e02cf0b6b3dad49e9e5188bc29acacf3fcc4580'
    assert hashlib.sha256(dat).hexdigest() == 'a059084545405792ddb8dd83393fcfa165ce93d14737c24eeec24debd6d10a40'
    spec = importlib.util.spec_from_file_location('make_tf_ckpt_fixture', os.path.join(here, 'make_tf_ckpt_fixture.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    index2, data2, arrays = mod.build()
    assert index2 == idx and data2 == dat                           # the committed files are what the script builds
    assert mod.crc32c_bitwise(b'123456789') == 0xE3069283           # the independent CRC agrees with the published check value
    # structure the writer in the product never emits: two restart points in the first data block, separator keys
    header, entries = ck.read_index(prefix)
    assert header == dict(num_shards=1, endianness=0)
    assert set(entries) == set(arrays) | {'save_counter_names'}
    assert entries['global_step']['shape'] == [] and entries['global_step']['dtype'] == 9
    assert entries['a/conv1/biases']['offset'] == 0 and entries['shape_probe']['shape'] == [2, 3, 4]
    back = ck.load_checkpoint(prefix, verify_data=True)
    assert set(back) == set(arrays)                                 # the string entry is skipped
    for k, a in arrays.items():
        assert back[k].dtype == a.dtype and back[k].shape == a.shape and np.array_equal(back[k], a), k
    assert [n for n, _ in ck.list_variables(prefix)] == sorted(entries)
    # a flipped byte in the second data block (reached only through the shortened separator key's handle) is detected
    bad = bytearray(idx)
    pos = idx.index(b'c/fc1/biases')
    bad[pos + 20] ^= 0x40
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p2 = os.path.join(d, 'x')
        open(p2 + '.index', 'wb').write(bytes(bad))
        open(p2 + '.data-00000-of-00001', 'wb').write(dat)
        with pytest.raises(ValueError):
            ck.load_checkpoint(p2)


def test_v1_checkpoint_is_named_as_such(tmp_path):
    p = os.path.join(str(tmp_path), 'model.ckpt')
    open(p, 'wb').write(b'\x00' * 64)
    with pytest.raises(ValueError, match='V1'):
        ck.load_checkpoint(p)


# ---- independent implementations found in the image: TensorBoard ships TensorFlow's generated protos and a CRC32C ---------------
def _bundle_entry_class():
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto) built with google.protobuf on top of TensorFlow's OWN
    generated TensorShapeProto (tensorboard.compat.proto.tensor_shape_pb2); the field numbers are the published ones."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2
    pool = descriptor_pool.Default()
    name = 't3d_test/tensor_bundle_entry.proto'
    try:
        fd = pool.FindFileByName(name)
    except KeyError:
        f = descriptor_pb2.FileDescriptorProto(name=name, package='t3d_test', syntax='proto3',
                                               dependency=[tensor_shape_pb2.DESCRIPTOR.name])
        m = f.message_type.add(name='BundleEntryProto')
        L, T = descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL, descriptor_pb2.FieldDescriptorProto
        m.field.add(name='dtype', number=1, label=L, type=T.TYPE_INT32)
        m.field.add(name='shape', number=2, label=L, type=T.TYPE_MESSAGE, type_name='.' + tensor_shape_pb2.TensorShapeProto.DESCRIPTOR.full_name)
        m.field.add(name='shard_id', number=3, label=L, type=T.TYPE_INT32)
        m.field.add(name='offset', number=4, label=L, type=T.TYPE_INT64)
        m.field.add(name='size', number=5, label=L, type=T.TYPE_INT64)
        m.field.add(name='crc32c', number=6, label=L, type=T.TYPE_FIXED32)
        fd = pool.Add(f) if hasattr(pool, 'Add') and not hasattr(pool, 'AddSerializedFile') else pool.AddSerializedFile(f.SerializeToString())
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('t3d_test.BundleEntryProto'))


def _raw_entries(prefix):
    """(key, raw value bytes) of every data-block entry of <prefix>.index, through the reader's own table walk."""
    import struct
    from transferable3d_b200 import tf_checkpoint as ck
    buf = open(prefix + '.index', 'rb').read()
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = ck._varint(footer, pos)
    _, pos = ck._varint(footer, pos)
    idx_off, pos = ck._varint(footer, pos)
    idx_size, pos = ck._varint(footer, pos)
    out = []
    for _, handle in ck._block_entries(ck._read_block(buf, idx_off, idx_size, True)):
        off, p = ck._varint(handle, 0)
        size, p = ck._varint(handle, p)
        out.extend(ck._block_entries(ck._read_block(buf, off, size, True)))
    return out


def test_crc32c_against_tensorboards_implementation():
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb
    from transferable3d_b200 import tf_checkpoint as ck
    rng = np.random.RandomState(0)
    for n in (0, 1, 2, 7, 8, 9, 63, 64, 1000, 4097):
        data = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        assert ck.crc32c(data) == tb.crc32c(data), n
        assert ck.masked_crc(data) == tb.masked_crc32c(data), n


def test_dtype_table_against_tensorflows_types_proto():
    from tensorboard.compat.proto import types_pb2
    from tensorboard.compat.tensorflow_stub import dtypes
    from transferable3d_b200 import tf_checkpoint as ck
    for enum_id, np_type in ck.DTYPES.items():
        name = types_pb2.DataType.Name(enum_id)                    # raises for an id TensorFlow does not define
        assert dtypes.as_dtype(enum_id).as_numpy_dtype == np_type, (enum_id, name)
    assert types_pb2.DataType.Value('DT_FLOAT') == ck.DTYPE_IDS[np.dtype(np.float32)]
    assert types_pb2.DataType.Value('DT_STRING') not in ck.DTYPES         # skipped by load_checkpoint


@pytest.mark.parametrize('which', ['written', 'handmade'])
def test_index_entries_parse_with_google_protobuf_and_tensorflows_shape_proto(which, tmp_path):
    """Every BundleEntryProto of an index file, parsed by google.protobuf into TensorFlow's own TensorShapeProto, gives the dtype,
    shape, offset, size and checksum the hand-written parser of tf_checkpoint.read_index reports (and nothing is left unparsed)."""
    from transferable3d_b200 import tf_checkpoint as ck
    if which == 'written':
        rng = np.random.RandomState(1)
        prefix = os.path.join(str(tmp_path), 'm.ckpt')
        ck.save_checkpoint(prefix, {'a/weights': rng.randn(1, 6, 64).astype(np.float32), 'a/biases': rng.randn(64).astype(np.float32),
                                    'global_step': np.asarray(12345678901, dtype=np.int64), 'b/x' * 30: rng.randint(0, 9, (3, 0, 2)).astype(np.int32),
                                    'flags': np.asarray([True, False])}, block_size=128)
    else:
        prefix = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tf_ckpt_handmade')
    Entry = _bundle_entry_class()
    _, entries = ck.read_index(prefix)
    seen = 0
    for key, value in _raw_entries(prefix):
        if key == b'':
            continue
        e = Entry()
        e.ParseFromString(bytes(value))
        mine = entries[key.decode()]
        assert (e.dtype, [d.size for d in e.shape.dim], e.shard_id, e.offset, e.size) == \
            (mine['dtype'], mine['shape'], mine['shard_id'], mine['offset'], mine['size']), key
        assert mine['crc32c'] is None or e.crc32c == mine['crc32c']
        # nothing unparsed: re-serialising the parsed message reproduces the entry byte for byte (unknown fields would be kept too,
        # so also check the length of what the known fields alone encode)
        assert e.SerializeToString() == bytes(value), key
        assert len(bytes(value)) == e.ByteSize()
        seen += 1
    assert seen == len(entries) >= 3
