"""TensorFlow-free reader / writer of TF1 ``tf.train.Saver`` checkpoints (the "V2" tensor bundle:
``<prefix>.index`` + ``<prefix>.data-00000-of-00001``) -- what ``AC_IRL(saved_network=...)`` restores
(ac_irl.py:108-111) and ``outerloop`` saves (ac_irl.py:948).

Format, restated from TensorFlow's sources (tensorflow/core/util/tensor_bundle, core/lib/io/table*, the
LevelDB table format it reuses, core/protobuf/tensor_bundle.proto):

  * ``.index`` is an immutable sorted string table: data blocks of prefix-compressed entries
    ``varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value`` followed by the
    restart offsets (uint32 LE each) and their count; every block is followed by a 5-byte trailer
    (compression type, masked CRC32C of block + type).  An index block maps separator keys to block
    handles (``varint64 offset | varint64 size``); the 48-byte footer holds the metaindex and index handles
    and the magic 0xdb4775248b80fb57.
  * key ``""`` -> ``BundleHeaderProto {num_shards=1, endianness=2, version=3}``; every other key is a variable
    name -> ``BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked)}``.
  * ``.data-0000k-of-0000n`` holds the raw little-endian row-major tensor bytes at [offset, offset+size).

Only what checkpoints of dense float / int variables need is implemented: uncompressed blocks (the bundle
writer never compresses), no tensor slices.  The protobuf messages are decoded by a 40-line wire-format parser.
There is no TensorFlow in this environment, so the writer + reader are tested against each other and against
the published CRC32C / varint known answers -- NOT against a file written by TensorFlow itself.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
BLOCK_RESTART_INTERVAL = 16
BLOCK_SIZE = 4096
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT_OF_NP = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32,
             np.dtype("int64"): DT_INT64}


# ----------------------------------------------------------------------------- CRC32C (Castagnoli), masked as in TF
def _crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC = _crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varints and protobuf wire format
def put_varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def parse_proto(buf):
    """{field number: [values]} of one message: varints as int, fixed32 / fixed64 as int, length-delimited as bytes."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = get_varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wire == 2:
            n, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + n]); pos += n
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(field, []).append(v)
    return out


def _field(field, wire, payload):
    return put_varint((field << 3) | wire) + payload


def _shape_proto(shape):
    return b"".join(_field(2, 2, put_varint(len(d)) + d) for d in (_field(1, 0, put_varint(int(s))) for s in shape))


def _entry_proto(dtype, shape, offset, size, crc):
    msg = _field(1, 0, put_varint(dtype))
    sp = _shape_proto(shape)
    msg += _field(2, 2, put_varint(len(sp)) + sp)
    if offset:
        msg += _field(4, 0, put_varint(offset))
    msg += _field(5, 0, put_varint(size))
    msg += _field(6, 5, struct.pack("<I", crc))
    return msg


# ----------------------------------------------------------------------------- table (.index) reader
def _read_block(data, offset, size, verify=True):
    contents, trailer = data[offset:offset + size], data[offset + size:offset + size + 5]
    if len(contents) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise ValueError("compressed table blocks (type %d) are not supported: tensor bundles are written uncompressed"
                         % trailer[0])
    if verify and struct.unpack("<I", trailer[1:])[0] != mask_crc(crc32c(contents + trailer[:1])):
        raise ValueError("table block checksum mismatch")
    return contents


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(data, verify=True):
    """All (key, value) pairs of a table file, in key order."""
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, p = get_varint(footer, 0)
    _, p = get_varint(footer, p)                     # metaindex handle (unused)
    ioff, p = get_varint(footer, p)
    isize, p = get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        off, q = get_varint(handle, 0)
        size, _ = get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, off, size, verify)))
    return out


def read_bundle(prefix, verify=True):
    """{variable name: ndarray} of the checkpoint ``prefix`` (the path handed to Saver.save / restore)."""
    with open(prefix + ".index", "rb") as f:
        entries = read_table(f.read(), verify)
    if not entries or entries[0][0] != b"":
        raise ValueError("checkpoint index has no bundle header")
    header = parse_proto(entries[0][1])
    num_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    shards = {}
    tensors = {}
    for key, value in entries[1:]:
        e = parse_proto(value)
        if 7 in e:
            raise ValueError("sliced (partitioned) variable %r is not supported" % key.decode())
        dtype = e.get(1, [0])[0]
        if dtype not in _NP_OF_DT:
            raise ValueError("variable %r has unsupported dtype enum %d" % (key.decode(), dtype))
        shape = []
        if 2 in e:
            for dim in parse_proto(e[2][0]).get(2, []):
                shape.append(parse_proto(dim).get(1, [0])[0])
        shard, offset, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            with open("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), "rb") as f:
                shards[shard] = f.read()
        raw = shards[shard][offset:offset + size]
        if len(raw) != size:
            raise ValueError("variable %r: data shard is truncated" % key.decode())
        if verify and 6 in e and e[6][0] != mask_crc(crc32c(raw)):
            raise ValueError("variable %r: checksum mismatch" % key.decode())
        tensors[key.decode()] = np.frombuffer(raw, dtype=_NP_OF_DT[dtype]).reshape(shape).copy()
    return tensors


# ----------------------------------------------------------------------------- writer (Saver.save equivalent)
class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count % BLOCK_RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _emit_block(out, contents):
    off = len(out)
    out += contents + b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
    return put_varint(off) + put_varint(len(contents))


def write_bundle(prefix, tensors):
    """Writes ``prefix.index`` and ``prefix.data-00000-of-00001`` for {name: ndarray} (one shard, sorted keys)."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    data = bytearray()
    items = [(b"", _field(1, 0, put_varint(1)) + _field(3, 2, put_varint(2) + _field(1, 0, put_varint(1))))]
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = np.asarray(tensors[name])
        a = np.ascontiguousarray(a).reshape(a.shape)          # (ascontiguousarray promotes 0-d to 1-d)
        if a.dtype not in _DT_OF_NP:
            raise TypeError("variable %r: dtype %s cannot be written" % (name, a.dtype))
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((name.encode(), _entry_proto(_DT_OF_NP[a.dtype], a.shape, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    out, index, block = bytearray(), _BlockBuilder(), _BlockBuilder()
    for i, (key, value) in enumerate(items):
        block.add(key, value)
        if len(block.buf) >= BLOCK_SIZE or i == len(items) - 1:
            index.add(key, _emit_block(out, block.finish()))           # separator = last key of the block
            block = _BlockBuilder()
    meta = _emit_block(out, _BlockBuilder().finish())
    idx = _emit_block(out, index.finish())
    footer = meta + idx
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
