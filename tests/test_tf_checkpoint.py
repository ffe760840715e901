"""TF-free reader / writer of TF1 Saver checkpoints (tf_checkpoint.py; ac_irl.py:108-111,948): known answers of the
building blocks, a hand-assembled minimal bundle (the byte layout restated from the TensorFlow sources), round trips
with many variables (multi-block index, prefix compression, restart points) and corruption detection.  No GPU."""
import os
import struct

import numpy as np
import pytest

from discrete_mean_field_game_b200 import tf_checkpoint as C


def test_crc32c_and_varint_known_answers():
    assert C.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    assert C.crc32c(b"\x00" * 32) == 0x8A9136AA                        # RFC 3720 B.4
    assert C.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert C.mask_crc(0) == C.MASK_DELTA
    for n in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 - 1):
        b = C.put_varint(n)
        assert C.get_varint(b, 0) == (n, len(b))
    assert C.put_varint(300) == b"\xac\x02"                             # the protobuf documentation's example


def test_hand_assembled_bundle_is_read(tmp_path):
    """One float32 variable 'w' = [1.5, -2] laid out byte by byte as the format description says."""
    raw = struct.pack("<2f", 1.5, -2.0)
    shape = b"\x12\x02\x08\x02"                                       # TensorShapeProto{dim{size: 2}}
    entry = b"\x08\x01" + b"\x12" + bytes([len(shape)]) + shape + b"\x28\x08" + b"\x35" + struct.pack(
        "<I", C.mask_crc(C.crc32c(raw)))                                 # dtype=DT_FLOAT, shape, size=8, crc32c
    header = b"\x08\x01\x1a\x02\x08\x01"                              # num_shards=1, version{producer=1}
    block = (b"\x00\x00" + bytes([len(header)]) + header +              # key "" (shared 0, non-shared 0)
             b"\x00\x01" + bytes([len(entry)]) + b"w" + entry +          # key "w"
             struct.pack("<I", 0) + struct.pack("<I", 1))               # one restart at offset 0
    out = bytearray()
    def emit(contents):
        off = len(out)
        out.extend(contents + b"\x00" + struct.pack("<I", C.mask_crc(C.crc32c(contents + b"\x00"))))
        return C.put_varint(off) + C.put_varint(len(contents))
    h_data = emit(block)
    h_meta = emit(struct.pack("<I", 0) + struct.pack("<I", 1))
    index_block = b"\x00\x01" + bytes([len(h_data)]) + b"w" + h_data + struct.pack("<I", 0) + struct.pack("<I", 1)
    h_index = emit(index_block)
    footer = h_meta + h_index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57))
    prefix = str(tmp_path / "model.ckpt")
    open(prefix + ".index", "wb").write(bytes(out))
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    z = C.read_bundle(prefix)
    assert list(z) == ["w"] and z["w"].dtype == np.float32 and z["w"].tolist() == [1.5, -2.0]
    # and the writer produces a file the same reader accepts with the same content
    C.write_bundle(str(tmp_path / "again.ckpt"), {"w": np.float32([1.5, -2.0])})
    np.testing.assert_array_equal(C.read_bundle(str(tmp_path / "again.ckpt"))["w"], z["w"])


def test_round_trip_many_variables_multi_block(tmp_path):
    rng = np.random.RandomState(0)
    tensors = {"reward/conv1/weights": np.float32(rng.randn(5, 5, 1, 1)), "reward/conv1/biases": np.float32(rng.randn(1)),
               "reward/fc3/weights": np.float32(rng.randn(450, 8)), "beta1_power": np.float32(0.81),
               "global_step": np.int64(7), "counts": np.arange(6, dtype=np.int32).reshape(2, 3),
               "d": rng.randn(3, 2)}
    for i in range(300):                                               # > 4 KB of index entries -> several data blocks
        tensors["scope_%03d/layer/weights/Adam_1" % i] = np.float32(rng.randn(i % 5 + 1))
    prefix = str(tmp_path / "sub" / "model_none_8_4.ckpt")
    C.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + ".index") > 2 * C.BLOCK_SIZE
    z = C.read_bundle(prefix)
    assert sorted(z) == sorted(tensors)
    for k, v in tensors.items():
        assert z[k].dtype == np.asarray(v).dtype and z[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(z[k], v)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m.ckpt")
    C.write_bundle(prefix, {"a": np.float32([1, 2, 3])})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        C.read_bundle(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        C.read_bundle(prefix)
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ValueError, match="magic"):
        C.read_bundle(prefix)
