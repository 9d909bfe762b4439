"""TFRecord files of the reference's datasets without TensorFlow.

The reference writes one `tf.train.Example` per image with `create_tfrecords.py` (R:138-249): features
`img` (bytes: float32 image, shape (height, width, depth)), `height`, `width`, `depth` (int64), `label` (bytes: float32
one-hot vector), and reads them back with `_parse_example` (conv_cINN_base_functions.py F:26-65) mapped over a
`tf.data.TFRecordDataset` (C:236-237, C:283-284, C:417-420).  This module restates both sides from the published
formats so that real MNIST / fashion-MNIST record files can feed the CUDA path:

  * TFRecord framing: uint64 length | uint32 masked CRC32C(length) | payload | uint32 masked CRC32C(payload), little endian,
    mask(c) = ((c >> 15 | c << 17) + 0xa282ead8) mod 2^32;
  * protobuf wire format of Example { Features features = 1 }, Features { map<string, Feature> feature = 1 },
    Feature { oneof { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3 } }.

`parse_example` returns (img [H,W,D] float32, label [n_classes] float32) like the reference; `load_dataset` stacks a file
into arrays (and, with `device=`, into torch tensors on the GPU: the "dataset" the data kernels and cFlow.fit consume).
`write_examples` is the inverse of the reader (the layout of R:195-249), used by the tests and for making fixtures.
"""
import struct

import numpy as np

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78                      # CRC-32C (Castagnoli), reflected
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = [int(v) for v in t]
    return _CRC_TABLE


def crc32c(data):
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- protobuf wire format (varint = 0, 64-bit = 1, length-delimited = 2, 32-bit = 5) ------------------------------
def _varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _feature(buf):
    for num, wt, val in _fields(buf):
        if num == 1:                                    # BytesList { repeated bytes value = 1 }
            return [bytes(v) for n, w, v in _fields(val) if n == 1]
        if num == 2:                                    # FloatList { repeated float value = 1 [packed] }
            out = []
            for n, w, v in _fields(val):
                if n == 1:
                    out += list(struct.unpack(f"<{len(v) // 4}f", v)) if w == 2 else list(struct.unpack("<f", v))
            return out
        if num == 3:                                    # Int64List { repeated int64 value = 1 [packed] }
            out = []
            for n, w, v in _fields(val):
                if n != 1:
                    continue
                if w == 2:
                    p = 0
                    while p < len(v):
                        x, p = _varint(v, p)
                        out.append(x - (1 << 64) if x >> 63 else x)
                else:
                    out.append(v - (1 << 64) if v >> 63 else v)
            return out
    return []


def decode_example(payload):
    """serialized tf.train.Example -> {feature name: list of bytes / floats / ints}"""
    feats = {}
    for num, wt, val in _fields(memoryview(payload)):
        if num != 1:
            continue
        for n2, w2, entry in _fields(val):              # map<string, Feature> entries
            if n2 != 1:
                continue
            key, feat = None, []
            for n3, w3, v3 in _fields(entry):
                if n3 == 1:
                    key = bytes(v3).decode("utf-8")
                elif n3 == 2:
                    feat = _feature(v3)
            feats[key] = feat
    return feats


def parse_example(serialized_example):
    """F:26-65: (img float32 [height, width, depth], label float32 [n_classes])"""
    f = decode_example(serialized_example)
    for k in ("img", "height", "width", "depth", "label"):
        if k not in f or len(f[k]) != 1:
            raise ValueError(f"parse_example: feature {k!r} missing or not a scalar (tf.io.FixedLenFeature([], ...), F:33-42)")
    h, w, d = int(f["height"][0]), int(f["width"][0]), int(f["depth"][0])
    img = np.frombuffer(f["img"][0], dtype="<f4")
    if img.size != h * w * d:
        raise ValueError(f"parse_example: img holds {img.size} floats, expected {h}x{w}x{d}")
    return img.reshape(h, w, d).copy(), np.frombuffer(f["label"][0], dtype="<f4").copy()


def read_records(path, verify=True):
    """yields the payload of every record of a TFRecord file (tf.data.TFRecordDataset, C:237)"""
    with open(path, "rb") as fh:
        while True:
            head = fh.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f"{path}: truncated record header")
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and masked_crc(head[:8]) != lcrc:
                raise ValueError(f"{path}: corrupt record length (CRC mismatch)")
            data = fh.read(length)
            tail = fh.read(4)
            if len(data) < length or len(tail) < 4:
                raise ValueError(f"{path}: truncated record")
            if verify and masked_crc(data) != struct.unpack("<I", tail)[0]:
                raise ValueError(f"{path}: corrupt record payload (CRC mismatch)")
            yield data


def load_dataset(path, device=None, limit=None, verify=True):
    """(images [N,H,W,D] float32, labels [N,n_classes] float32) of a record file; torch tensors on `device` if given"""
    imgs, labels = [], []
    for i, rec in enumerate(read_records(path, verify=verify)):
        if limit is not None and i >= limit:
            break
        im, lb = parse_example(rec)
        imgs.append(im)
        labels.append(lb)
    if not imgs:
        raise ValueError(f"{path}: no records")
    x, y = np.stack(imgs), np.stack(labels)
    if device is None:
        return x, y
    import torch
    return torch.from_numpy(x).to(device), torch.from_numpy(y).to(device)


# ---- writer (R:138-249) -----------------------------------------------------------------------------------
def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(img, label):
    """the tf.train.Example create_tfrecords.py builds (R:195-227): img / label as raw float32 bytes, three int64 sizes"""
    img = np.ascontiguousarray(img, dtype="<f4")
    label = np.ascontiguousarray(label, dtype="<f4")
    assert img.ndim == 3
    feats = {"img": _ld(1, _ld(1, img.tobytes())), "label": _ld(1, _ld(1, label.tobytes()))}
    for k, v in zip(("height", "width", "depth"), img.shape):
        feats[k] = _ld(3, _ld(1, _enc_varint(int(v))))
    body = b"".join(_ld(1, _ld(1, k.encode()) + _ld(2, v)) for k, v in sorted(feats.items()))
    return _ld(1, body)


def write_examples(path, images, labels):
    """tf.io.TFRecordWriter over encode_example (R:229-249)"""
    with open(path, "wb") as fh:
        for im, lb in zip(images, labels):
            data = encode_example(im, lb)
            head = struct.pack("<Q", len(data))
            fh.write(head + struct.pack("<I", masked_crc(head)) + data + struct.pack("<I", masked_crc(data)))
