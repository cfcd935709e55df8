"""TensorFlow checkpoint (tensor bundle, "V2" format) reader without TensorFlow - SURVEY 8f #2.

The reference restores its generator from `tf.train.Saver` checkpoints (generate.py:88-90:
`Saver(ema.variables_to_restore()).restore(sess, <prefix>)`), i.e. a pair of files

    <prefix>.index                   sorted string table: "" -> BundleHeaderProto, name -> BundleEntryProto
    <prefix>.data-00000-of-0000N     raw little-endian tensor bytes, addressed by (shard_id, offset, size)

TensorFlow is not a dependency of this package (and is not installable here), so the two on-disk formats are
restated from their published definitions:
  * the table format of tensorflow/core/lib/io/table (LevelDB's SSTable): data blocks of prefix-compressed
    entries [shared varint32 | non_shared varint32 | value_len varint32 | key suffix | value], a restart array
    and its length at the end of every block, a 5-byte block trailer (compression type, masked crc32c), an index
    block mapping separator keys to block handles (offset varint64, size varint64) and a 48-byte footer
    (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57);
  * tensorflow/core/protobuf/tensor_bundle.proto: BundleHeaderProto {num_shards=1, endianness=2, version=3},
    BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32), slices=7};
    TensorShapeProto {dim=2 {size=1, name=2}, unknown_rank=3}.
PARITY UNPINNED: no TensorFlow-written checkpoint exists in this environment, so the reader is tested against the
writer below (same restatement) and against hand-assembled byte fixtures, not against TensorFlow's own output.
Checksums are verified when present (crc32c, masked as in lib/hash/crc32c.h); snappy-compressed blocks and
partitioned (sliced) variables are reported as unsupported rather than guessed.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ crc32c
def _crc_table():
    poly = 0x82F63B78
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if (c & 1) else (c >> 1)
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_CRC_TAB = _crc_table()


def _crc_bytes(c, data):
    tab = _CRC_TAB
    for b in bytes(data):
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c


def crc32c(data):
    """CRC-32C (Castagnoli), the checksum of lib/hash/crc32c.h.
    Large buffers are cut into K equal segments whose registers advance together as one NumPy vector (the register
    update is linear over GF(2): the state after a segment from state s is Z(s) ^ R, with R the segment run from state
    0 and Z the advance through as many zero bytes; Z is obtained from 32 extra lanes that start from the basis
    states and read zeros)."""
    buf = np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview, np.ndarray)) else data, dtype=np.uint8)
    n = buf.size
    if n < (1 << 14):
        return _crc_bytes(0xFFFFFFFF, buf.tobytes()) ^ 0xFFFFFFFF
    K = int(min(4096, n // 64))
    seg = n // K
    body = buf[:K * seg].reshape(K, seg)
    state = np.zeros(K + 32, dtype=np.uint32)
    state[K:] = np.uint32(1) << np.arange(32, dtype=np.uint32)
    cols = np.zeros((seg, K + 32), dtype=np.uint32)
    cols[:, :K] = body.T
    tab = _CRC_TAB
    for j in range(seg):
        state = tab[(state ^ cols[j]) & np.uint32(0xFF)] ^ (state >> np.uint32(8))
    zcols = [int(v) for v in state[K:]]                       # Z applied to basis state i
    c = 0xFFFFFFFF
    for k in range(K):
        z = 0
        i = 0
        while c:
            if c & 1:
                z ^= zcols[i]
            c >>= 1
            i += 1
        c = z ^ int(state[k])
    return _crc_bytes(c, buf[K * seg:].tobytes()) ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _read_varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _write_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> list of (field number, wire type, value); value is int (varint / fixed) or bytes (length-delimited)"""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _parse_shape(buf):
    dims = []
    for field, wt, v in _parse_proto(buf):
        if field == 2 and wt == 2:
            size = 0
            for f2, w2, v2 in _parse_proto(v):
                if f2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
        elif field == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


# ------------------------------------------------------------------------------------------------ table reader
def _read_block(data, offset, size, verify=True):
    block = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    ctype = trailer[0]
    if verify:
        want = struct.unpack('<I', trailer[1:5])[0]
        if want != masked_crc(block + trailer[:1]):
            raise ValueError("table block checksum mismatch")
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d): snappy-compressed checkpoints are not supported" % ctype)
    return block


def _block_entries(block):
    """iterate (key, value) of one block in order"""
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        value = bytes(block[pos:pos + vlen])
        pos += vlen
        yield key, value


def read_table(path, verify=True):
    """all (key, value) pairs of an SSTable file, in key order"""
    with open(path, 'rb') as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError("%s: too short for a table footer" % path)
    footer = data[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("%s: bad table magic (not a tensor-bundle index)" % path)
    pos = 0
    _, pos = _read_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p = _read_varint(handle, 0)
        bsize, p = _read_varint(handle, p)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


# ------------------------------------------------------------------------------------------------ bundle reader
class BundleReader(object):
    """`tf.train.load_checkpoint(prefix)`-like access: .keys(), .shape(name), .dtype(name), .get_tensor(name)"""

    def __init__(self, prefix, verify=True):
        self.prefix = prefix
        self.verify = verify
        self.entries = {}
        self.num_shards = 1
        for key, value in read_table(prefix + '.index', verify):
            fields = _parse_proto(value)
            if key == b'':
                for f, wt, v in fields:
                    if f == 1:
                        self.num_shards = v
                    elif f == 2 and v != 0:
                        raise NotImplementedError("big-endian checkpoint")
                continue
            e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'sliced': False}
            for f, wt, v in fields:
                if f == 1:
                    e['dtype'] = v
                elif f == 2:
                    e['shape'] = _parse_shape(v)
                elif f == 3:
                    e['shard_id'] = v
                elif f == 4:
                    e['offset'] = v
                elif f == 5:
                    e['size'] = v
                elif f == 6:
                    e['crc32c'] = v
                elif f == 7:
                    e['sliced'] = True
            self.entries[key.decode('utf-8')] = e

    def keys(self):
        return sorted(self.entries)

    def has_tensor(self, name):
        return name in self.entries

    def shape(self, name):
        return self.entries[name]['shape']

    def dtype(self, name):
        return np.dtype(_DTYPES[self.entries[name]['dtype']])

    def get_tensor(self, name):
        e = self.entries[name]
        if e['sliced']:
            raise NotImplementedError("partitioned variable %s" % name)
        if e['dtype'] not in _DTYPES:
            raise NotImplementedError("dtype id %d of %s" % (e['dtype'], name))
        shard = "%s.data-%05d-of-%05d" % (self.prefix, e['shard_id'], self.num_shards)
        with open(shard, 'rb') as f:
            f.seek(e['offset'])
            raw = f.read(e['size'])
        if len(raw) != e['size']:
            raise ValueError("truncated data shard for %s" % name)
        if self.verify and e['crc32c'] is not None and masked_crc(raw) != e['crc32c']:      # every tensor, whatever its size
            raise ValueError("tensor checksum mismatch for %s" % name)
        arr = np.frombuffer(raw, dtype=_DTYPES[e['dtype']])
        n = int(np.prod(e['shape'])) if e['shape'] else 1
        if arr.size != n:
            raise ValueError("size of %s does not match its shape" % name)
        return arr.reshape(e['shape']).copy()


def is_bundle(prefix):
    return os.path.exists(prefix + '.index')


def generator_weights(prefix, wanted):
    """what `Saver(ema.variables_to_restore()).restore` gives the generator (generate.py:88-90): for every variable the
    graph wants, its ExponentialMovingAverage shadow when the checkpoint has one, else the variable itself.
    Shadows are saved as '<scope>/<var>/ExponentialMovingAverage', possibly under the optimiser's name scope.
    `wanted`: iterable of variable names -> dict name -> ndarray"""
    rd = BundleReader(prefix)
    keys = rd.keys()
    out = {}
    for name in wanted:
        cands = [name + '/ExponentialMovingAverage', 'optimiser/' + name + '/ExponentialMovingAverage', name]
        hit = next((c for c in cands if rd.has_tensor(c)), None)
        if hit is None:      # any other enclosing name scope of the shadow
            suffix = '/' + name + '/ExponentialMovingAverage'
            hit = next((k for k in keys if k.endswith(suffix)), None)
        if hit is not None:
            out[name] = rd.get_tensor(hit)
    return out


# ------------------------------------------------------------------------------------------------ writer (tests, tools)
def _proto_field(field, wt, payload):
    return _write_varint((field << 3) | wt) + payload


def _shape_proto(shape):
    out = b''
    for d in shape:
        dim = _proto_field(1, 0, _write_varint(int(d)))
        out += _proto_field(2, 2, _write_varint(len(dim)) + dim)
    return out


def _build_block(pairs, restart_interval=16):
    buf, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        buf += _write_varint(shared) + _write_varint(len(k) - shared) + _write_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        buf += struct.pack('<I', r)
    buf += struct.pack('<I', len(restarts))
    return bytes(buf)


def write_bundle(prefix, tensors, block_entries=64):
    """write {name: ndarray} as a one-shard tensor bundle in the layout described in the module docstring"""
    names = sorted(tensors)
    data = bytearray()
    pairs = [(b'', _proto_field(1, 0, _write_varint(1)) + _proto_field(3, 2, _write_varint(2) + _proto_field(1, 0, _write_varint(1))))]
    for n in names:
        a = np.asarray(tensors[n])          # (ascontiguousarray would promote scalars to rank 1)
        raw = a.tobytes()                   # C order
        shape = _shape_proto(a.shape)
        entry = _proto_field(1, 0, _write_varint(_DTYPE_IDS[a.dtype]))
        entry += _proto_field(2, 2, _write_varint(len(shape)) + shape)
        if len(data):
            entry += _proto_field(4, 0, _write_varint(len(data)))
        entry += _proto_field(5, 0, _write_varint(len(raw)))
        entry += _proto_field(6, 5, struct.pack('<I', masked_crc(raw)))
        pairs.append((n.encode('utf-8'), entry))
        data += raw
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    out = bytearray()
    index_pairs = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b'\x00' + struct.pack('<I', masked_crc(block + b'\x00')))
        return _write_varint(off) + _write_varint(len(block))

    for i in range(0, len(pairs), block_entries):
        chunk = pairs[i:i + block_entries]
        handle = emit(_build_block(chunk))
        index_pairs.append((chunk[-1][0] + b'\x00' if chunk[-1][0] else b'\x00', handle))      # separator >= last key
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block(index_pairs, restart_interval=1))
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + '.index', 'wb') as f:
        f.write(bytes(out))
