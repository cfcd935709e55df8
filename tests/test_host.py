"""Host-side logic (no GPU): config files, speaker tables, WAV I/O, CLI argument surface."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parameter_files_keep_reference_keys():
    m = json.load(open(os.path.join(ROOT, "model_parameters.json")))
    w = json.load(open(os.path.join(ROOT, m["wavenet_parameters"])))
    assert m["encoder"] == "64" and m["use_vq"] is True and m["k"] == 512 and m["latent_dim"] == 64
    assert m["speaker_embedding"] == 64 and m["beta"] == 0.25
    assert w["dilation_rates"] == [2 ** i for i in range(10)] * 3
    assert (w["kernel_size"], w["dilation_filters"], w["skip_filters"], w["residual_filters"]) == (3, 256, 512, 256)
    assert w["preprocess"] == {"kernel_size": 32, "filters": 256} and w["quantization_channels"] == 256
    import vqvae_wavenet_b200 as pkg
    cfg = pkg.EngineConfig.from_files(os.path.join(ROOT, "model_parameters.json"))
    assert cfg.receptive_field == 6170 and cfg.cond_channels == 128
    c = cfg.to_c()
    assert c.num_layers == 30 and c.dilations[9] == 512 and c.k == 512


def test_wavenet_mirror_asserts_like_reference():
    import vqvae_wavenet_b200 as pkg
    wn = pkg.Wavenet(os.path.join(ROOT, "wavenet_parameters.json"))
    assert wn.receptive_field == 6170
    bad = dict(wn.args, dilation_rates=[1, 2, 4])
    with pytest.raises(AssertionError):
        pkg.Wavenet(bad)                                   # wavenet.py:13


def test_speaker_tables_and_onehot(tmp_path, monkeypatch):
    from conftest import write_speaker_table
    from vqvae_wavenet_b200 import utils
    assert utils.dataset_for_speakers(["p225"]) == ("vctk", 109)
    assert utils.dataset_for_speakers(["S0002"]) == ("aishell", 340)
    assert utils.dataset_for_speakers(["1034"]) == ("librispeech", 251)
    monkeypatch.delenv("VQWN_SPEAKER_TABLES", raising=False)
    with pytest.raises(FileNotFoundError):
        utils.find_speaker_table("vctk", roots=(str(tmp_path),))
    names = ["p301", "p295"] + ["p%d" % (400 + i) for i in range(107)]
    write_speaker_table(tmp_path, "vctk", names)
    path = utils.find_speaker_table("vctk", roots=(str(tmp_path),))
    monkeypatch.setenv("VQWN_SPEAKER_TABLES", str(tmp_path))
    assert utils.find_speaker_table("vctk", roots=()) == path
    table = utils.get_speaker_to_int(path)
    assert table["p301"] == 0 and len(table) == 109
    one = utils.speaker_onehot(["p301", "None", "p295"], table, 109)
    assert one.shape == (3, 1, 109) and one[0, 0, 0] == 1 and one[1].sum() == 0 and one[2, 0, 1] == 1
    assert list(np.argmax(one, -1).reshape(-1)) == [0, 0, 1]       # 'None' -> row 0
    with pytest.raises(NotImplementedError):
        utils.decode(np.zeros((1, 256), np.float32), mode="beam")
    with pytest.raises(RuntimeError):
        utils.decode(np.zeros((1, 256), np.float32), mode="greedy")   # needs an engine: no CPU path


def test_wav_roundtrip_and_prepare(tmp_path):
    from vqvae_wavenet_b200 import wavio, mu_law_ops
    x = mu_law_ops.decode_lut()[np.random.default_rng(0).integers(0, 256, 2000)]
    p = str(tmp_path / "a.wav")
    wavio.write_wav_float32(p, 16000, x)
    from scipy.io import wavfile
    sr, y = wavfile.read(p)                                  # the reader the reference's users have
    assert sr == 16000 and y.dtype == np.float32 and np.array_equal(y, x)
    assert np.array_equal(wavio.read_wav(p), x)
    wavfile.write(str(tmp_path / "b.wav"), 8000, (x[:800] * 32767).astype(np.int16))
    z = wavio.read_wav(str(tmp_path / "b.wav"))
    assert z.shape[0] == 1600 and abs(float(z[0]) - float(x[0])) < 1e-3
    w = wavio.prepare_audio(x, 3)
    assert w.shape == (3, 1536, 1) and np.array_equal(w[2, :, 0], x[:1536])


def test_cli_surface(tmp_path, monkeypatch):
    """flags and error paths of the reference CLI that do not need a device"""
    import generate
    from vqvae_wavenet_b200 import wavio
    from conftest import write_speaker_table
    write_speaker_table(tmp_path)
    monkeypatch.setenv("VQWN_SPEAKER_TABLES", str(tmp_path))
    wav = str(tmp_path / "in.wav")
    wavio.write_wav_float32(wav, 16000, np.zeros(1024, np.float32))
    restore = str(tmp_path / "weights-42")
    with pytest.raises(FileNotFoundError, match="npz"):      # neither <restore>.index (TF checkpoint) nor <restore>.npz
        generate.main(["-restore", restore, "-audio", wav, "-speakers", "p225", "None", "-mode", "greedy"])
    np.save(str(tmp_path / "z.npy"), np.zeros((16, 64), np.float32))
    with pytest.raises(FileNotFoundError, match="npz"):
        generate.main(["-restore", restore, "-audio", wav, "-speakers", "p225", "-z_e", str(tmp_path / "z.npy")])
    with pytest.raises(ValueError):
        generate.main(["-restore", str(tmp_path / "weights-x"), "-audio", wav, "-speakers", "p225"])   # gs = int(...)


def test_synthetic_magenta_encoder_weights_equal_oracle():
    """the package's seeded Encoder_Magenta weights are the oracle's (the package itself never imports oracle/)"""
    import vqvae_wavenet_b200 as pkg
    from vqvae_wavenet_b200 import synthetic
    from oracle import oracle as O
    cfg = pkg.EngineConfig(model=dict(encoder="Magenta"))
    assert cfg.to_c().encoder == 1
    a = synthetic.make_encoder_magenta_weights(cfg)
    b = O.make_encoder_magenta_weights(O.Config())
    assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
    with pytest.raises(RuntimeError):
        pkg.Encoder_Magenta(64).build(np.zeros((1, 64, 1), np.float32))
    with pytest.raises(RuntimeError):
        pkg.Encoder_2019(64).build(np.zeros((1, 320, 1), np.float32))     # like the others: no engine, no CPU fallback
    assert pkg.EngineConfig(model=dict(encoder="2019")).to_c().encoder == 2019


def test_wav_extensible_subformat(tmp_path):
    """WAVE_FORMAT_EXTENSIBLE: the plain format tag is read from the sub-format GUID (bytes 24:26 of the fmt body), so
    32-bit PCM is not mistaken for float because a 03 00 happens to occur in the header"""
    import struct
    from vqvae_wavenet_b200 import wavio
    guid_tail = bytes.fromhex("000000001000800000aa00389b71")

    def write(path, subtag, payload, bits, mask=3):           # mask = 3 puts the bytes 03 00 into the header
        fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 16000, 16000 * bits // 8, bits // 8, bits) + struct.pack("<HHI", 22, bits, mask) \
            + struct.pack("<H", subtag) + guid_tail
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
        with open(path, "wb") as f:
            f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    ints = (np.array([0, 1 << 30, -(1 << 30), (1 << 31) - 1], dtype="<i4"))
    write(tmp_path / "pcm32.wav", 1, ints.tobytes(), 32)
    x = wavio.read_wav(str(tmp_path / "pcm32.wav"))
    assert np.allclose(x, ints.astype(np.float64) / 2147483648.0)
    fl = np.array([0.0, 0.5, -0.25, 1.0], dtype="<f4")
    write(tmp_path / "f32.wav", 3, fl.tobytes(), 32)
    assert np.array_equal(wavio.read_wav(str(tmp_path / "f32.wav")), fl)


def test_crc32c_vectorised_equals_bytewise():
    from vqvae_wavenet_b200 import tf_checkpoint as T
    assert T.crc32c(b"123456789") == 0xE3069283                      # the published check value of CRC-32C
    rng = np.random.default_rng(0)
    for n in (0, 1, 16383, 16384, 16385, 100003, 300000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(d) == (T._crc_bytes(0xFFFFFFFF, d) ^ 0xFFFFFFFF), n


def test_synthetic_encoder2019_weights_equal_oracle():
    import vqvae_wavenet_b200 as pkg
    from vqvae_wavenet_b200 import synthetic
    from oracle import oracle as O
    a = synthetic.make_encoder2019_weights(pkg.EngineConfig(model=dict(encoder="2019")))
    b = O.make_encoder2019_weights(O.Config())
    assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
