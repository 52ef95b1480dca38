'''
The TensorFlow-checkpoint reader / writer (composer_b200/tf_checkpoint.py).  No file written by real TensorFlow
exists in this image, so the format is pinned from three sides: published check values of the primitives (CRC32C,
LevelDB's mask), a byte-level fixture assembled by hand from the format description (independent of the writer), and
writer -> reader round trips over several table blocks.
'''
import os
import struct
import subprocess
import sys

import numpy as np

from composer_b200 import tf_checkpoint as tfc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_crc32c_check_values():
    assert tfc.crc32c(b'123456789') == 0xE3069283                       # the standard CRC-32C check value
    assert tfc.crc32c(b'\x00' * 32) == 0x8A9136AA                       # RFC 3720 B.4: 32 bytes of zeros
    assert tfc.crc32c(b'\xff' * 32) == 0x62A8AB43                       # RFC 3720 B.4: 32 bytes of ones
    crc = tfc.crc32c(b'foo')
    assert tfc.masked_crc32c(b'foo') == ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def _hand_block(entries):
    '''A table block written entry by entry from the format description, without prefix compression.'''
    out, restarts = b'', []
    for key, value in entries:
        restarts.append(len(out))
        out += bytes([0, len(key), len(value)]) + key + value             # shared 0, single-byte varints
    for r in restarts:
        out += struct.pack('<I', r)
    return out + struct.pack('<I', len(restarts))


def test_reader_on_a_hand_assembled_checkpoint(tmp_path):
    value = np.arange(6, dtype='<f4').reshape(2, 3)
    raw = value.tobytes()
    # BundleEntryProto: dtype DT_FLOAT, shape {dim {size 2} dim {size 3}}, size 24, crc32c (offset 0 omitted)
    entry = b'\x08\x01' + b'\x12\x08' + b'\x12\x02\x08\x02' + b'\x12\x02\x08\x03' + b'\x28\x18' + \
        b'\x35' + struct.pack('<I', tfc.masked_crc32c(raw))
    header = b'\x08\x01\x1a\x02\x08\x01'                                  # num_shards 1, version {producer 1}
    name = b'model/wte/weight/.ATTRIBUTES/VARIABLE_VALUE'
    data_block = _hand_block([(b'', header), (name, entry)])
    trailer = lambda block: b'\x00' + struct.pack('<I', tfc.masked_crc32c(block + b'\x00'))
    meta_block = _hand_block([])
    file = data_block + trailer(data_block)
    meta_offset = len(file)
    file += meta_block + trailer(meta_block)
    index_block = _hand_block([(name, bytes([0, len(data_block)]))])     # handle: offset 0, size
    index_offset = len(file)
    file += index_block + trailer(index_block)
    footer = bytes([meta_offset, len(meta_block), index_offset, len(index_block)])
    file += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', tfc.TABLE_MAGIC)
    prefix = str(tmp_path / 'ckpt-7')
    with open(prefix + '.index', 'wb') as handle:
        handle.write(file)
    with open(prefix + '.data-00000-of-00001', 'wb') as handle:
        handle.write(raw)
    (tmp_path / 'checkpoint').write_text('model_checkpoint_path: "ckpt-7"\nall_model_checkpoint_paths: "ckpt-7"\n')
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    bundle = tfc.read_bundle(prefix)
    assert list(bundle) == [name.decode()]
    np.testing.assert_array_equal(bundle[name.decode()], value)
    # a flipped data byte must be caught by the entry's checksum
    with open(prefix + '.data-00000-of-00001', 'wb') as handle:
        handle.write(b'\x01' + raw[1:])
    try:
        tfc.read_bundle(prefix)
    except ValueError as error:
        assert 'checksum' in str(error)
    else:
        raise AssertionError('corrupted tensor data went unnoticed')


def test_bundle_round_trip_over_several_blocks(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {'model/decoder_blocks/%d/mlp/c_fc/weight%s' % (i, tfc.SUFFIX): rng.standard_normal((5, 7)).astype(np.float32)
               for i in range(120)}                                        # > 4 KB of index entries: several data blocks
    tensors['step' + tfc.SUFFIX] = np.asarray(41, dtype=np.int64)
    tensors['model/wte/weight' + tfc.SUFFIX] = rng.standard_normal((390, 16)).astype(np.float32)
    prefix = str(tmp_path / 'ckpt-1')
    tfc.write_bundle(prefix, tensors)
    back = tfc.read_bundle(prefix)
    assert sorted(back) == sorted(tensors)
    for name, value in tensors.items():
        np.testing.assert_array_equal(back[name], value)
        assert back[name].dtype == value.dtype and back[name].shape == value.shape
    keys = [k for k, _ in tfc.read_table(prefix + '.index')]
    assert keys == sorted(keys) and keys[0] == b''


def test_object_graph_names_follow_the_reference():
    # transformer.py:681-693: blocks are a Python list attribute `decoder_blocks`, Keras names h_1 .. h_L
    assert tfc.object_path('h_1/attn/c_attn/weight') == 'model/decoder_blocks/0/attn/c_attn/weight'
    assert tfc.object_path('h_8/ln_2/gamma') == 'model/decoder_blocks/7/ln_2/gamma'
    assert tfc.object_path('wpe/embeddings') == 'model/wpe/embeddings'
    assert tfc.object_path('ln_f/beta') == 'model/ln_f/beta'


def test_weights_and_slots_survive_the_name_mapping(tmp_path):
    from oracle import transformer_oracle as oracle
    cfg = oracle.OracleConfig(vocab_size=31, embedding_size=16, window_size=8, decoder_layers_count=2, attention_head_count=4)
    weights = oracle.init_parameters(cfg, seed=5)
    rng = np.random.default_rng(1)
    adam_m = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in weights.items()}
    adam_v = {k: rng.random(v.shape).astype(np.float32) for k, v in weights.items()}
    prefix = str(tmp_path / 'ckpt-3')
    tfc.write_bundle(prefix, tfc.from_arrays(weights, adam_m, adam_v, {'step': 9, 'epoch': 2, 'iterations': 8}))
    shapes = {k: v.shape for k, v in weights.items()}
    w2, m2, v2, counters = tfc.to_arrays(tfc.read_bundle(prefix), list(weights), shapes)
    assert counters == {'step': 9, 'epoch': 2, 'iterations': 8}
    for name in weights:
        np.testing.assert_array_equal(w2[name], weights[name])
        np.testing.assert_array_equal(m2[name], adam_m[name])
        np.testing.assert_array_equal(v2[name], adam_v[name])


def test_convert_tool_round_trip(tmp_path):
    from oracle import transformer_oracle as oracle
    cfg = oracle.OracleConfig(vocab_size=31, embedding_size=16, window_size=8, decoder_layers_count=2, attention_head_count=4)
    weights = oracle.init_parameters(cfg, seed=6)
    tf_dir, npz_dir, tf_again = tmp_path / 'tf', tmp_path / 'npz', tmp_path / 'tf2'
    tf_dir.mkdir()
    tfc.write_bundle(str(tf_dir / 'ckpt-4'), tfc.from_arrays(weights, None, None, {'step': 5, 'epoch': 3, 'iterations': 4}))
    (tf_dir / 'checkpoint').write_text('model_checkpoint_path: "ckpt-4"\n')
    tool = os.path.join(ROOT, 'tools', 'convert_checkpoint.py')
    subprocess.run([sys.executable, tool, 'to-npz', str(tf_dir), str(npz_dir)], check=True, capture_output=True)
    data = np.load(npz_dir / 'ckpt-1.npz')
    for name, value in weights.items():
        np.testing.assert_array_equal(data['variables/' + name], value)
    assert int(data['step']) == 5 and int(data['epoch']) == 3
    subprocess.run([sys.executable, tool, 'to-tf', str(npz_dir), str(tf_again)], check=True, capture_output=True)
    back = tfc.read_bundle(tfc.latest_checkpoint(str(tf_again)))
    for name, value in weights.items():
        np.testing.assert_array_equal(back[tfc.object_path(name) + tfc.SUFFIX], value)
