"""Wire formats either side of the path (generate.py:36-44,115-117): 16 kHz mono float audio in,
IEEE-float32 WAV out.  Pure NumPy (no TensorFlow / ffmpeg)."""
import struct

import numpy as np


def read_wav(path, sample_rate=16000):
    """decode_audio(content, 'wav', 16000, 1) (generate.py:36-37): mono float32 in [-1,1] at 16 kHz.
    PCM 8/16/24/32-bit and IEEE float; channels are averaged; other rates are linearly resampled."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("%s is not a RIFF/WAVE file" % path)
    pos, fmt, raw, fmt_body = 12, None, None, b""
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            fmt_body = body
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError("missing fmt/data chunk in %s" % path)
    tag, channels, rate, _, _, bits = fmt
    if tag == 0xFFFE:
        # WAVE_FORMAT_EXTENSIBLE: cbSize (2) + valid bits (2) + channel mask (4), then the sub-format GUID whose first two
        # bytes are the plain format tag (1 = PCM, 3 = IEEE float): bytes 24:26 of the fmt chunk body
        if len(fmt_body) < 26:
            raise ValueError("truncated WAVE_FORMAT_EXTENSIBLE fmt chunk in %s" % path)
        tag = struct.unpack("<H", fmt_body[24:26])[0]
    if tag == 3 and bits == 32:
        x = np.frombuffer(raw, dtype="<f4").astype(np.float32)
    elif tag == 3 and bits == 64:
        x = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    elif tag == 1 and bits == 16:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif tag == 1 and bits == 32:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (np.where(v >= 1 << 23, v - (1 << 24), v)).astype(np.float32) / 8388608.0
    else:
        raise NotImplementedError("WAV format tag %d with %d bits" % (tag, bits))
    x = x[:x.size // channels * channels].reshape(-1, channels).mean(axis=1)
    if rate != sample_rate:
        n = int(round(x.size * sample_rate / float(rate)))
        x = np.interp(np.arange(n) * (rate / float(sample_rate)), np.arange(x.size), x)
    return x.astype(np.float32)


def prepare_audio(wav, batch_size, multiple=512):
    """generate.py:39-40: trim to a multiple of 512 (largest dilation), [1,T,1], tile over the batch."""
    t = wav.shape[0] // multiple * multiple
    return np.tile(wav[:t].reshape(1, t, 1), (batch_size, 1, 1)).astype(np.float32)


def write_wav_float32(path, rate, samples):
    """scipy.io.wavfile.write(path, rate, float32 array) (generate.py:117): IEEE-float WAV, mono."""
    x = np.ascontiguousarray(samples, dtype="<f4")
    payload = x.tobytes()
    fmt = struct.pack("<HHIIHH", 3, 1, rate, rate * 4, 4, 32)
    fact = struct.pack("<I", x.size)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + \
        b"data" + struct.pack("<I", len(payload)) + payload
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
