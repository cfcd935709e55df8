"""TensorFlow tensor-bundle reader - CPU tests.  The on-disk format is restated from its published
definition (no TensorFlow here): round trips through the restated writer, multi-block tables, checksum
detection, and the EMA-shadow selection of generate.py:88-90."""
import os
import struct

import numpy as np
import pytest

import vqvae_wavenet_b200.tf_checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(b"123456789") == 0xE3069283


def _tensors(rng, n):
    out = {}
    for i in range(n):
        shape = tuple(int(x) for x in rng.integers(1, 6, size=int(rng.integers(0, 4))))
        out["decoder/cycle_%d/layer_%d/gated/kernel" % (1 + i // 10, 1 + i % 10)] = rng.standard_normal(shape).astype(np.float32)
    out["global_step"] = np.array(110640, dtype=np.int64)
    out["embedding/embedding"] = rng.standard_normal((512, 64)).astype(np.float32)
    return out


@pytest.mark.parametrize("block_entries", [1, 3, 64])
def test_round_trip(tmp_path, block_entries):
    rng = np.random.default_rng(5)
    tensors = _tensors(rng, 40)
    prefix = str(tmp_path / "weights-110640")
    ck.write_bundle(prefix, tensors, block_entries=block_entries)
    assert ck.is_bundle(prefix)
    rd = ck.BundleReader(prefix, verify="full")
    assert rd.keys() == sorted(tensors)
    for name, want in tensors.items():
        got = rd.get_tensor(name)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)
        assert rd.shape(name) == want.shape


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(6)
    prefix = str(tmp_path / "w-1")
    ck.write_bundle(prefix, {"a": rng.standard_normal(100).astype(np.float32), "b": np.arange(7, dtype=np.int32)})
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    rd = ck.BundleReader(prefix)
    with pytest.raises(ValueError):
        rd.get_tensor("a")
    assert np.array_equal(rd.get_tensor("b"), np.arange(7, dtype=np.int32))
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        ck.BundleReader(prefix)
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ValueError):
        ck.BundleReader(prefix)


def test_generator_weights_prefers_ema_shadows(tmp_path):
    """generate.py:88-90 restores through ema.variables_to_restore(): the shadow wins over the raw variable; the
    shadow may sit under the optimiser's name scope"""
    one = np.ones((2, 3), dtype=np.float32)
    tensors = {
        "decoder/skip/kernel": 1 * one, "decoder/skip/kernel/ExponentialMovingAverage": 2 * one,
        "decoder/skip/bias": 3 * one, "optimiser/decoder/skip/bias/ExponentialMovingAverage": 4 * one,
        "speaker_embedding": 5 * one,
        "decoder/skip/kernel/Adam": 9 * one, "beta1_power": np.array(0.5, dtype=np.float32),
    }
    prefix = str(tmp_path / "weights-7")
    ck.write_bundle(prefix, tensors)
    got = ck.generator_weights(prefix, ["decoder/skip/kernel", "decoder/skip/bias", "speaker_embedding", "missing/var"])
    assert sorted(got) == ["decoder/skip/bias", "decoder/skip/kernel", "speaker_embedding"]
    assert got["decoder/skip/kernel"][0, 0] == 2 and got["decoder/skip/bias"][0, 0] == 4 and got["speaker_embedding"][0, 0] == 5


def test_hand_assembled_index_block(tmp_path):
    """one table assembled byte by byte here (not by write_bundle): a single data block with prefix compression, so
    the reader is checked against the format text rather than against its own writer"""
    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)
    header = bytes([0x08, 0x01])                                   # BundleHeaderProto.num_shards = 1
    data = np.arange(6, dtype=np.float32).tobytes()
    shape = bytes([0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03])    # dim{size:2} dim{size:3}
    entry = bytes([0x08, 0x01, 0x12, len(shape)]) + shape + bytes([0x28, len(data)]) + bytes([0x35]) + struct.pack("<I", ck.masked_crc(data))
    entry2 = bytes([0x08, 0x01, 0x12, 0x00, 0x20, len(data), 0x28, 0x04])       # scalar float32 at offset 24
    block = bytearray()
    block += varint(0) + varint(0) + varint(len(header)) + header                          # key ""
    block += varint(0) + varint(7) + varint(len(entry)) + b"abc/def" + entry               # key "abc/def"
    block += varint(4) + varint(3) + varint(len(entry2)) + b"xyz" + entry2                 # key "abc/xyz" (shares "abc/")
    block += struct.pack("<I", 0) + struct.pack("<I", 1)                                   # one restart point
    block = bytes(block)
    out = bytearray(block) + b"\x00" + struct.pack("<I", ck.masked_crc(block + b"\x00"))
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)
    meta_off = len(out)
    out += meta + b"\x00" + struct.pack("<I", ck.masked_crc(meta + b"\x00"))
    handle = varint(0) + varint(len(block))
    index = varint(0) + varint(4) + varint(len(handle)) + b"abc0" + handle + struct.pack("<I", 0) + struct.pack("<I", 1)
    index_off = len(out)
    out += index + b"\x00" + struct.pack("<I", ck.masked_crc(index + b"\x00"))
    footer = varint(meta_off) + varint(len(meta)) + varint(index_off) + varint(len(index))
    out += footer + bytes(40 - len(footer)) + struct.pack("<Q", ck.TABLE_MAGIC)
    prefix = str(tmp_path / "hand-1")
    open(prefix + ".index", "wb").write(bytes(out))
    open(prefix + ".data-00000-of-00001", "wb").write(data + np.float32(2.5).tobytes())
    rd = ck.BundleReader(prefix, verify="full")
    assert rd.keys() == ["abc/def", "abc/xyz"]
    assert np.array_equal(rd.get_tensor("abc/def"), np.arange(6, dtype=np.float32).reshape(2, 3))
    assert rd.get_tensor("abc/xyz") == np.float32(2.5) and rd.shape("abc/xyz") == ()
