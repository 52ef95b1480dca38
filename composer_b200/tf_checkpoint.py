'''
TensorFlow checkpoints without TensorFlow: a reader / writer of the tensor-bundle format that
``tf.train.Checkpoint`` / ``CheckpointManager`` write (what the reference's ``Transformer.train`` saves,
``composer/models/transformer.py:890-900``, and ``BaseModel.load_from_checkpoint`` restores,
``composer/models/__init__.py:66-90``), and the mapping between the reference's object graph and this
package's flat parameter arena.  Host-side, numpy only.

Format (restated from TensorFlow's sources, none of which is in this image):

* ``<prefix>.index`` is an immutable sorted string table in LevelDB's table format
  (``tensorflow/core/lib/io/table_builder.cc``, ``block_builder.cc``, ``format.cc``): data blocks of
  prefix-compressed entries ``varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value``
  followed by a restart array (uint32 offsets, then their count), each block followed by a 5-byte trailer
  (compression type, masked CRC32C); one index block maps the last key of each data block to its
  ``BlockHandle`` (varint64 offset, varint64 size); a 48-byte footer holds the metaindex and index handles and the
  magic number 0xdb4775248b80fb57.  ``BundleWriter`` disables compression.
* key ``""`` holds a ``BundleHeaderProto`` (1 num_shards, 2 endianness: LITTLE = 0 / BIG = 1, 3 version {producer = 1});
  every other key is a tensor name with a ``BundleEntryProto``: 1 dtype, 2 shape (``TensorShapeProto``: repeated
  dim { 1 size }), 3 shard_id, 4 offset, 5 size, 6 fixed32 crc32c (masked) -- ``tensor_bundle.proto``.
* ``<prefix>.data-00000-of-00001`` is the tensors' raw little-endian bytes at those offsets.
* an object-based checkpoint names a variable by its attribute path from the root:
  ``model/decoder_blocks/0/attn/c_attn/weight/.ATTRIBUTES/VARIABLE_VALUE``; Adam's slots are
  ``<variable path>/.OPTIMIZER_SLOT/optimizer/{m,v}/.ATTRIBUTES/VARIABLE_VALUE``; ``step``, ``epoch`` and
  ``optimizer/iter`` are scalar int64 variables; ``_CHECKPOINTABLE_OBJECT_GRAPH`` holds the serialized
  ``TrackableObjectGraph`` (string tensor), which a name-based reader such as this one does not need.

PARITY STATUS: verified only against its own writer and a hand-assembled fixture (tests/test_tf_checkpoint.py);
no file written by real TensorFlow was available to check it against.
'''

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_INT32, DT_STRING, DT_INT64 = 1, 3, 7, 9
_DTYPES = {DT_FLOAT: np.dtype('<f4'), DT_INT32: np.dtype('<i4'), DT_INT64: np.dtype('<i8')}
SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'


# ---------------------------------------------------------------------------
# CRC32C (Castagnoli), masked as LevelDB / TensorFlow store it
# ---------------------------------------------------------------------------

def _crc_table():
    table = []
    for i in range(256):
        crc = i
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
        table.append(crc)
    return table


_CRC = _crc_table()


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for byte in bytes(data):
        crc = _CRC[(crc ^ byte) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------
# varints and the two protos
# ---------------------------------------------------------------------------

def _put_varint(value):
    out = bytearray()
    while value >= 0x80:
        out.append((value & 0x7F) | 0x80)
        value >>= 7
    out.append(value)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        byte = buf[pos]
        pos += 1
        result |= (byte & 0x7F) << shift
        if byte < 0x80:
            return result, pos
        shift += 7


def _parse_fields(buf):
    '''Protobuf wire format -> list of (field number, wire type, value).'''

    fields, pos = [], 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        number, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _get_varint(buf, pos)
        elif wire == 1:
            value, pos = struct.unpack_from('<Q', buf, pos)[0], pos + 8
        elif wire == 2:
            length, pos = _get_varint(buf, pos)
            value, pos = bytes(buf[pos:pos + length]), pos + length
        elif wire == 5:
            value, pos = struct.unpack_from('<I', buf, pos)[0], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wire)
        fields.append((number, wire, value))
    return fields


def _parse_entry(buf):
    entry = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None}
    for number, _, value in _parse_fields(buf):
        if number == 1:
            entry['dtype'] = value
        elif number == 2:
            for n2, _, dim in _parse_fields(value):
                if n2 == 2:                                       # TensorShapeProto.dim
                    size = 0
                    for n3, _, v3 in _parse_fields(dim):
                        if n3 == 1:
                            size = v3
                    entry['shape'].append(size)
        elif number == 3:
            entry['shard_id'] = value
        elif number == 4:
            entry['offset'] = value
        elif number == 5:
            entry['size'] = value
        elif number == 6:
            entry['crc32c'] = value
    return entry


def _build_entry(dtype, shape, offset, size, crc):
    dims = b''.join(b'\x12' + _put_varint(len(d)) + d for d in (b'\x08' + _put_varint(int(s)) for s in shape))
    out = b'\x08' + _put_varint(dtype) + b'\x12' + _put_varint(len(dims)) + dims
    if offset:
        out += b'\x20' + _put_varint(offset)
    out += b'\x28' + _put_varint(size) + b'\x35' + struct.pack('<I', crc)
    return out


# ---------------------------------------------------------------------------
# LevelDB table: read
# ---------------------------------------------------------------------------

def _read_block(data, offset, size):
    block = data[offset:offset + size]
    if data[offset + size] != 0:
        raise ValueError('compressed table blocks are not supported (BundleWriter never compresses)')
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    entries, pos, key = [], 0, b''
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        value_len, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        entries.append((key, block[pos:pos + value_len]))
        pos += value_len
    return entries


def read_table(path):
    '''All (key, value) pairs of a LevelDB-format table file, in key order.'''

    with open(path, 'rb') as handle:
        data = handle.read()
    if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError('%s is not a TensorFlow checkpoint index (bad table magic)' % path)
    footer = data[-48:]
    _, pos = _get_varint(footer, 0)                 # metaindex handle
    _, pos = _get_varint(footer, pos)
    index_offset, pos = _get_varint(footer, pos)
    index_size, pos = _get_varint(footer, pos)
    pairs = []
    for _, handle_bytes in _read_block(data, index_offset, index_size):
        offset, p = _get_varint(handle_bytes, 0)
        size, _ = _get_varint(handle_bytes, p)
        pairs.extend(_read_block(data, offset, size))
    return pairs


# ---------------------------------------------------------------------------
# LevelDB table: write (one data block per `block_size` bytes, restart interval 16)
# ---------------------------------------------------------------------------

def _build_block(entries, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            limit = min(len(last), len(key))
            while shared < limit and last[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_table(path, pairs, block_size=4096):
    pairs = sorted(pairs)
    out, index_entries = bytearray(), []

    def emit(block):
        offset = len(out)
        out.extend(block)
        out.extend(b'\x00' + struct.pack('<I', masked_crc32c(block + b'\x00')))
        return _put_varint(offset) + _put_varint(len(block))

    pending, pending_bytes = [], 0
    for key, value in pairs:
        pending.append((key, value))
        pending_bytes += len(key) + len(value) + 3
        if pending_bytes >= block_size:
            index_entries.append((pending[-1][0], emit(_build_block(pending))))
            pending, pending_bytes = [], 0
    if pending:
        index_entries.append((pending[-1][0], emit(_build_block(pending))))
    metaindex = emit(_build_block([]))
    index = emit(_build_block(index_entries, restart_interval=1))
    footer = metaindex + index
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(path, 'wb') as handle:
        handle.write(bytes(out))


# ---------------------------------------------------------------------------
# Tensor bundle
# ---------------------------------------------------------------------------

def read_bundle(prefix, verify_crc=True):
    '''``{tensor name: numpy array}`` of the checkpoint ``<prefix>.index`` / ``<prefix>.data-*`` (string tensors skipped).'''

    tensors, shards = {}, {}
    for key, value in read_table(prefix + '.index'):
        if key == b'':
            header = {n: v for n, _, v in _parse_fields(value)}
            num_shards = header.get(1, 1)
            if header.get(2, 0) == 1:
                raise ValueError('big-endian checkpoints are not supported')
            continue
        entry = _parse_entry(value)
        if entry['dtype'] not in _DTYPES:
            continue                                  # e.g. the DT_STRING object graph
        shard = entry['shard_id']
        if shard not in shards:
            with open('%s.data-%05d-of-%05d' % (prefix, shard, num_shards), 'rb') as handle:
                shards[shard] = handle.read()
        raw = shards[shard][entry['offset']:entry['offset'] + entry['size']]
        if verify_crc and entry['crc32c'] is not None and masked_crc32c(raw) != entry['crc32c']:
            raise ValueError('checksum mismatch for tensor %s' % key.decode())
        dtype = _DTYPES[entry['dtype']]
        tensors[key.decode()] = np.frombuffer(raw, dtype=dtype).reshape(entry['shape']).copy()
    return tensors


def write_bundle(prefix, tensors):
    '''Writes ``{name: array}`` (float32 / int32 / int64) as a one-shard tensor bundle.'''

    data, pairs = bytearray(), []
    header = b'\x08\x01' + b'\x1a\x02\x08\x01'        # num_shards 1, endianness LITTLE (0, the default), version {producer 1}
    pairs.append((b'', header))
    for name in sorted(tensors):
        array = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
        dtype = {np.dtype('float32'): DT_FLOAT, np.dtype('int32'): DT_INT32, np.dtype('int64'): DT_INT64}[array.dtype]
        raw = array.astype(array.dtype.newbyteorder('<')).tobytes()
        pairs.append((name.encode(), _build_entry(dtype, array.shape, len(data), len(raw), masked_crc32c(raw))))
        data.extend(raw)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.data-00000-of-00001', 'wb') as handle:
        handle.write(bytes(data))
    write_table(prefix + '.index', pairs)


def latest_checkpoint(directory):
    '''``tf.train.latest_checkpoint``: the prefix named by ``model_checkpoint_path`` in the ``checkpoint`` state file.'''

    state = os.path.join(directory, 'checkpoint')
    if not os.path.exists(state):
        return None
    with open(state) as handle:
        for line in handle:
            if line.startswith('model_checkpoint_path:'):
                name = line.split(':', 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(directory, name)
    return None


# ---------------------------------------------------------------------------
# The reference's object graph  <->  this package's variable names
# ---------------------------------------------------------------------------

def object_path(keras_name, root='model'):
    '''
    Checkpoint key of a variable of the reference's ``Transformer`` (attribute path from the ``Checkpoint`` root,
    transformer.py:890: ``model=self``): ``h_3/attn/c_attn/weight`` (the Keras name, block names 1-based,
    transformer.py:692) is ``model/decoder_blocks/2/attn/c_attn/weight`` (list index 0-based, :681-693).
    '''

    parts = keras_name.split('/')
    if parts[0].startswith('h_'):
        parts = ['decoder_blocks', str(int(parts[0][2:]) - 1)] + parts[1:]
    return '/'.join([root] + parts)


def to_arrays(bundle, layout_names, shapes, root='model'):
    '''
    ``{keras name: array}`` for the flat arena plus ``(adam_m, adam_v, counters)`` (dicts, possibly empty) from the
    tensors of a reference checkpoint.  Missing variables raise with the key that was looked for.
    '''

    weights, adam_m, adam_v = {}, {}, {}
    for name in layout_names:
        key = object_path(name, root) + SUFFIX
        if key not in bundle:
            raise KeyError('the checkpoint has no variable %s (looked for %s)' % (name, key))
        weights[name] = bundle[key].astype(np.float32).reshape(shapes[name])
        for slot, target in (('m', adam_m), ('v', adam_v)):
            slot_key = '%s/.OPTIMIZER_SLOT/optimizer/%s%s' % (object_path(name, root), slot, SUFFIX)
            if slot_key in bundle:
                target[name] = bundle[slot_key].astype(np.float32).reshape(shapes[name])
    counters = {}
    for label, key in (('step', 'step'), ('epoch', 'epoch'), ('iterations', 'optimizer/iter')):
        if key + SUFFIX in bundle:
            counters[label] = int(np.asarray(bundle[key + SUFFIX]).reshape(-1)[0])
    return weights, adam_m, adam_v, counters


def from_arrays(weights, adam_m=None, adam_v=None, counters=None, root='model'):
    '''The inverse of :func:`to_arrays`: tensors of a checkpoint the reference's ``Checkpoint.restore`` can match by name.'''

    bundle = {}
    for name, value in weights.items():
        bundle[object_path(name, root) + SUFFIX] = np.asarray(value, dtype=np.float32)
        for slot, source in (('m', adam_m), ('v', adam_v)):
            if source and name in source:
                bundle['%s/.OPTIMIZER_SLOT/optimizer/%s%s' % (object_path(name, root), slot, SUFFIX)] = \
                    np.asarray(source[name], dtype=np.float32)
    for label, key in (('step', 'step'), ('epoch', 'epoch'), ('iterations', 'optimizer/iter')):
        if counters and label in counters:
            bundle[key + SUFFIX] = np.asarray(counters[label], dtype=np.int64)
    return bundle
