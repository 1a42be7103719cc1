"""TensorFlow checkpoint (tensor bundle, "V2" format) reader / writer without TensorFlow.

The reference saves and restores its models with ``tf.train.Saver`` (``examples/timit/training/train_ctc.py:148,264``,
``examples/timit/evaluation/eval_ctc.py:70-88``).  This module reads and writes that on-disk format so that variables
trained there can be loaded here and vice versa (variable names are the TF names this package already uses):

    <prefix>.index                   an LevelDB-style sorted string table: key "" -> BundleHeaderProto, every other key =
                                     a tensor name -> BundleEntryProto (dtype, shape, shard, offset, size, masked CRC-32C)
    <prefix>.data-00000-of-00001     the raw little-endian tensor bytes, in key order
    checkpoint                       CheckpointState text proto naming the latest prefix

Format restated from TensorFlow's public sources (tensorflow/core/util/tensor_bundle/tensor_bundle.cc,
tensorflow/core/lib/io/{table_builder,block_builder,format}.cc, tensor_bundle.proto, tensor_shape.proto, types.proto);
TensorFlow is not installable in this environment, so interoperability is checked here only by construction
(round trips, CRC-32C / snappy known answers, a hand-assembled table): **unpinned against TensorFlow itself**.
Blocks are written uncompressed (readers accept that); snappy-compressed blocks are read.
"""
import ctypes as C
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8
# types.proto
DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 6: np.dtype("i1"),
      9: np.dtype("<i8"), 10: np.dtype("bool")}
DT_OF = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("uint8"): 4,
         np.dtype("int8"): 6, np.dtype("int64"): 9, np.dtype("bool"): 10}


# ---------------------------------------------------------------- checksums
def crc32c(data, crc=0):
    """CRC-32C through the library's host helper (b2_crc32c); bytes-like or a C-contiguous numpy array."""
    from ... import _lib
    lib = _lib.load()
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(lib.b2_crc32c(C.c_uint32(crc), C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))
    b = bytes(data)
    return int(lib.b2_crc32c(C.c_uint32(crc), C.c_char_p(b), C.c_size_t(len(b))))


def mask_crc(crc):
    """leveldb / TF crc32c::Mask: rotate right by 15 bits and add a constant"""
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ---------------------------------------------------------------- varints / protobuf wire format
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7f) << shift
        if b < 0x80:
            return v, pos
        shift += 7


def _pb_fields(buf):
    """-> list of (field number, wire type, value) of one protobuf message"""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _encode_shape(shape):
    """TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }"""
    return b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in shape)


def _decode_shape(buf):
    dims = []
    for f, _, v in _pb_fields(buf):
        if f == 2:
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _encode_entry(dtype, shape, offset, size, crc):
    """BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32, masked)"""
    return (_pb_varint(1, dtype) + _pb_bytes(2, _encode_shape(shape)) + _pb_varint(4, offset) + _pb_varint(5, size) +
            _put_varint((6 << 3) | 5) + struct.pack("<I", crc))


def _decode_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": False}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] = True
    return e


def _encode_header(num_shards=1):
    """BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0, omitted), version = 3 {producer = 1}"""
    return _pb_varint(1, num_shards) + _pb_bytes(3, _pb_varint(1, 1))


# ---------------------------------------------------------------- snappy (read side only)
def snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:                                   # copy, 1-byte offset
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:                                 # copy, 2-byte offset
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:                                           # copy, 4-byte offset
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                             # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch: %d != %d" % (len(out), n))
    return bytes(out)


# ---------------------------------------------------------------- sorted string table
def _block_entries(block):
    """-> [(key, value)] of one table block (prefix-compressed keys, restart array at the end)"""
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    out, pos, key = [], 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _read_block(buf, offset, size, verify=True):
    raw, ctype = buf[offset:offset + size], buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(bytes(buf[offset:offset + size + 1])):
            raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return bytes(raw)
    if ctype == 1:
        return snappy_uncompress(bytes(raw))
    raise ValueError("unknown block compression type %d" % ctype)


def read_table(path, verify=True):
    """-> [(key bytes, value bytes)] in key order"""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s: not a table file (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    _, pos = _get_varint(footer, 0)                    # metaindex handle
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p = _get_varint(handle, 0)
        size, p = _get_varint(handle, p)
        out += _block_entries(_read_block(buf, off, size, verify))
    return out


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            n = min(len(key), len(self.last))
            while shared < n and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))


def write_table(path, items, block_size=4096):
    """items: [(key bytes, value bytes)] sorted by key; uncompressed blocks"""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                   # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index, bb, last_key = _BlockBuilder(1), _BlockBuilder(), None
    for key, value in items:
        assert last_key is None or key > last_key, "keys must be strictly increasing"
        bb.add(key, value)
        last_key = key
        if len(bb.buf) >= block_size:
            index.add(last_key, emit(bb.finish()))
            bb = _BlockBuilder()
    if bb.count:
        index.add(last_key, emit(bb.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(out)


# ---------------------------------------------------------------- bundle
def save_tf_checkpoint(prefix, arrays, write_state=True):
    """arrays: {tensor name: numpy array} -> <prefix>.index, <prefix>.data-00000-of-00001 (+ the 'checkpoint' file)."""
    items, offset = [(b"", _encode_header(1))], 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for name in sorted(arrays, key=lambda s: s.encode()):
            a = np.asarray(arrays[name])
            if a.dtype not in DT_OF:
                raise TypeError("%s: dtype %s not supported by the bundle writer" % (name, a.dtype))
            shape = a.shape                              # (ascontiguousarray turns a scalar into shape (1,))
            a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False))
            data.write(a.tobytes())
            items.append((name.encode(), _encode_entry(DT_OF[np.dtype(a.dtype.name)], shape, offset, a.nbytes,
                                                       mask_crc(crc32c(a)))))
            offset += a.nbytes
    write_table(prefix + ".index", items)
    if write_state:
        base = os.path.basename(prefix)
        with open(os.path.join(os.path.dirname(prefix) or ".", "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def load_tf_checkpoint(prefix, verify=True):
    """-> {tensor name: numpy array} of a TF checkpoint bundle (single- or multi-shard, full tensors)."""
    entries = read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise ValueError("%s.index: no bundle header" % prefix)
    num_shards = 1
    for f, _, v in _pb_fields(entries[0][1]):
        if f == 1:
            num_shards = v
        if f == 2 and v != 0:
            raise ValueError("big-endian bundles are not supported")
    shards = {}
    out = {}
    for key, value in entries[1:]:
        e = _decode_entry(value)
        if e["slices"]:
            raise ValueError("%s: partitioned (sliced) variables are not supported" % key.decode())
        if e["dtype"] not in DT:
            continue                                    # strings / resources: not model parameters
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(np.asarray(raw)):
            raise ValueError("%s: tensor checksum mismatch" % key.decode())
        out[key.decode()] = np.frombuffer(bytes(raw), dtype=DT[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def is_tf_checkpoint(prefix):
    return os.path.isfile(prefix + ".index")
