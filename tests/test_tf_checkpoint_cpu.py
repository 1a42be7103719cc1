"""utils/io/tf_checkpoint.py: TensorFlow tensor-bundle (checkpoint V2) reader / writer without TensorFlow.  Checked by
construction only (TensorFlow is not installable here): CRC-32C and its leveldb mask against published known answers,
snappy against a hand-assembled stream, the table format through a multi-block round trip and a hand-assembled
two-entry table, the bundle through a round trip of a model-sized variable set, corruption detection."""
import os
import struct

import numpy as np
import pytest

from tensorflow_end2end_speech_recognition_b200.utils.io import tf_checkpoint as tfc


@pytest.fixture(scope="module", autouse=True)
def _library():
    """CRC-32C comes from the C-ABI library (host code, loads without a GPU): build it if this is a fresh checkout"""
    import __graft_entry__
    __graft_entry__.build()


def test_crc32c_known_answers():
    assert tfc.crc32c(b"123456789") == 0xE3069283                  # the CRC catalogue's check value for CRC-32C
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA                  # RFC 3720 B.4: 32 bytes of zeros
    assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43                  # RFC 3720 B.4: 32 bytes of ones
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E              # RFC 3720 B.4: incrementing bytes
    a = np.arange(1000, dtype=np.float32)
    assert tfc.crc32c(a) == tfc.crc32c(a.tobytes())
    assert tfc.crc32c(b"6789", tfc.crc32c(b"12345")) == 0xE3069283   # continuation
    m = tfc.mask_crc(0xE3069283)
    assert m != 0xE3069283 and tfc.unmask_crc(m) == 0xE3069283


def test_snappy_uncompress_hand_assembled_stream():
    # "abcdabcdabcdabcd": literal "abcd" (tag (4-1)<<2), then a 2-byte-offset copy of length 12 from offset 4 (overlapping)
    stream = bytes([16]) + bytes([(3 << 2) | 0]) + b"abcd" + bytes([((12 - 1) << 2) | 2, 4, 0])
    assert tfc.snappy_uncompress(stream) == b"abcd" * 4
    # 1-byte-offset copy: length 4..11, offset up to 2047
    stream = bytes([9]) + bytes([(4 << 2) | 0]) + b"hello" + bytes([((4 - 4) << 2) | 1, 5])
    assert tfc.snappy_uncompress(stream) == b"hellohell"
    with pytest.raises(ValueError):
        tfc.snappy_uncompress(bytes([5]) + bytes([(0 << 2) | 1, 9]))


def test_table_round_trip_many_blocks(tmp_path):
    rng = np.random.RandomState(0)
    items = [(("layer%03d/kernel_%d" % (i // 7, i)).encode(), rng.bytes(int(rng.randint(1, 200)))) for i in range(900)]
    items = sorted(dict(items).items())
    p = str(tmp_path / "t.index")
    tfc.write_table(p, items, block_size=512)
    assert tfc.read_table(p) == items
    raw = bytearray(open(p, "rb").read())
    assert struct.unpack_from("<Q", raw, len(raw) - 8)[0] == tfc.TABLE_MAGIC
    raw[10] ^= 0x40                                     # flip one bit inside the first data block
    open(p, "wb").write(raw)
    with pytest.raises(ValueError):
        tfc.read_table(p)


def test_table_reader_on_hand_assembled_file(tmp_path):
    """one data block with two prefix-compressed entries, written byte by byte from the format description"""
    def block(body, restarts):
        b = body + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
        return b, b + b"\x00" + struct.pack("<I", tfc.mask_crc(tfc.crc32c(b + b"\x00")))
    body = bytes([0, 3, 2]) + b"abc" + b"v1" + bytes([2, 2, 1]) + b"xy" + b"w"          # "abc" -> "v1", "abxy" -> "w"
    data, data_raw = block(body, [0])
    meta, meta_raw = block(b"", [0])
    ibody = bytes([0, 4, 2]) + b"abxy" + bytes([0, len(data)])                          # handle: offset 0, size len(data)
    index, index_raw = block(ibody, [0])
    off_meta, off_index = len(data_raw), len(data_raw) + len(meta_raw)
    footer = bytes([off_meta, len(meta)]) + bytes([off_index, len(index)])
    blob = data_raw + meta_raw + index_raw + footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tfc.TABLE_MAGIC)
    p = str(tmp_path / "hand.index")
    open(p, "wb").write(blob)
    assert tfc.read_table(p) == [(b"abc", b"v1"), (b"abxy", b"w")]


def test_bundle_round_trip_and_corruption(tmp_path):
    rng = np.random.RandomState(1)
    arrays = {"blstm_hidden1/fw/lstm_cell/kernel": rng.randn(592, 2048).astype(np.float32),
              "blstm_hidden1/fw/lstm_cell/bias": np.zeros(2048, np.float32),
              "blstm_hidden1/fw/lstm_cell/w_i_diag": rng.randn(512).astype(np.float32),
              "output/weights": rng.randn(1024, 29).astype(np.float32),
              "global_step": np.asarray(1234, np.int64),
              "scalar_f64": np.asarray(0.5, np.float64),
              "empty": np.zeros((0, 3), np.float32)}
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.save_tf_checkpoint(prefix, arrays)
    assert tfc.is_tf_checkpoint(prefix) and os.path.isfile(prefix + ".data-00000-of-00001")
    assert 'model_checkpoint_path: "model.ckpt-1234"' in open(str(tmp_path / "checkpoint")).read()
    got = tfc.load_tf_checkpoint(prefix)
    assert sorted(got) == sorted(arrays)
    for k, v in arrays.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    # entries are in key order, the data file is the concatenation in that order
    keys = [k for k, _ in tfc.read_table(prefix + ".index")]
    assert keys[0] == b"" and keys[1:] == sorted(k.encode() for k in arrays)
    size = sum(v.nbytes for v in arrays.values())
    assert os.path.getsize(prefix + ".data-00000-of-00001") == size
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(ValueError):
        tfc.load_tf_checkpoint(prefix)
    assert tfc.load_tf_checkpoint(prefix, verify=False)["global_step"] == 1234
