"""TFRecord reader of the reference's dataset files (F:26-65, R:138-249) without TensorFlow: known-answer CRC32C values,
a hand-assembled tf.train.Example in TensorFlow's own byte layout, round trips, corruption detection."""
import struct

import numpy as np
import pytest

from arl_conditional_normalizing_flows_b200 import tfrecords as T


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors for CRC-32C
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b"123456789") == 0xE3069283
    # TFRecord masking of the length header of an empty record (8 zero bytes)
    c = T.crc32c(bytes(8))
    assert T.masked_crc(bytes(8)) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_hand_assembled_example_in_tensorflows_layout():
    """the bytes tf.train.Example(...).SerializeToString() produces for a 1x2x1 image (field order and packed int64 as
    protobuf emits them), assembled by hand from the wire-format rules"""
    img = np.array([[[1.5], [-2.0]]], np.float32)
    label = np.array([0.0, 1.0], np.float32)

    def ld(num, payload):
        return bytes([(num << 3) | 2, len(payload)]) + payload
    def entry(key, feature):
        return ld(1, ld(1, key.encode()) + ld(2, feature))
    bytes_feat = lambda b: ld(1, ld(1, b))
    int_feat = lambda v: ld(3, ld(1, bytes([v])))           # Int64List packed: tag 0x0a, len, varints
    feats = entry("depth", int_feat(1)) + entry("height", int_feat(1)) + entry("img", bytes_feat(img.tobytes())) + \
        entry("label", bytes_feat(label.tobytes())) + entry("width", int_feat(2))
    example = ld(1, feats)
    d = T.decode_example(example)
    assert d["height"] == [1] and d["width"] == [2] and d["depth"] == [1]
    im, lb = T.parse_example(example)
    np.testing.assert_array_equal(im, img)
    np.testing.assert_array_equal(lb, label)
    assert T.encode_example(img, label) == example           # the writer emits exactly these bytes
    # unpacked int64 (one varint field per value) is accepted too
    unpacked = ld(1, entry("height", ld(3, bytes([0x08, 3]))))
    assert T.decode_example(unpacked)["height"] == [3]


def test_round_trip_and_corruption(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 28, 28, 1)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 5)]
    p = tmp_path / "x_train_mnist_c3.tfrecords"
    T.write_examples(str(p), x, y)
    xs, ys = T.load_dataset(str(p))
    np.testing.assert_array_equal(xs, x)
    np.testing.assert_array_equal(ys, y)
    assert T.load_dataset(str(p), limit=2)[0].shape == (2, 28, 28, 1)
    raw = bytearray(p.read_bytes())
    raw[40] ^= 0x01
    (tmp_path / "bad.tfrecords").write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="CRC"):
        T.load_dataset(str(tmp_path / "bad.tfrecords"))
    (tmp_path / "short.tfrecords").write_bytes(bytes(raw[:100]))
    with pytest.raises(ValueError):
        T.load_dataset(str(tmp_path / "short.tfrecords"), verify=False)
    rec = next(T.read_records(str(p)))
    length = struct.unpack("<Q", p.read_bytes()[:8])[0]
    assert len(rec) == length
    with pytest.raises(ValueError, match="missing"):
        T.parse_example(b"")
