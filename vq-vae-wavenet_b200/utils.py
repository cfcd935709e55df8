"""Host helpers with the reference's names (utils.py:13-46,93-100; generate.py:46-61)."""
import os

import numpy as np

from . import mu_law_ops


def decode(predictions, mode="sample", quantization_channels=256, engine=None, uniforms=None):
    """utils.decode (utils.py:30-46) executed on the device through vqwn_decode.
    Returns decoded audio [B] float32.  `uniforms` replaces np.random.rand(B) (utils.py:22)."""
    if mode not in ("sample", "greedy"):
        raise NotImplementedError("decode mode %s not implemented" % mode)
    if engine is None:
        raise RuntimeError("decode() needs the Engine that owns the device tables (no CPU fallback)")
    _, audio = engine.decode(predictions, mode=mode, uniforms=uniforms)
    return audio


def get_speaker_to_int(speaker_path):
    """'name, int' lines (utils.py:93-100)."""
    table = {}
    with open(speaker_path) as f:
        for line in f:
            line = line.strip()
            if line:
                name, number = line.split(", ")
                table[name] = int(number)
    return table


def dataset_for_speakers(speakers):
    """generate.py:46-57: dataset picked from the first character of the first speaker."""
    c = speakers[0][0]
    if c == "p":
        return "vctk", 109
    if c.lower() == "s":
        return "aishell", 340
    return "librispeech", 251


def find_speaker_table(dataset, roots=(".",)):
    """generate.py:46-57 reads data/<ds>_speakers.txt while the reference ships data/<ds>_info/<ds>_speakers.txt
    (SURVEY Q14): try both, under $VQWN_SPEAKER_TABLES first, then under the given roots (the working directory of a
    reference checkout).  The tables ("<speaker>, <index>" per line, in the order the checkpoint was trained with) are
    the reference's data and are not redistributed with this package."""
    env = os.environ.get("VQWN_SPEAKER_TABLES")
    for r in ((env,) if env else ()) + tuple(roots):
        for rel in ("data/%s_speakers.txt" % dataset, "data/%s_info/%s_speakers.txt" % (dataset, dataset)):
            p = os.path.join(r, rel)
            if os.path.exists(p):
                return p
    raise FileNotFoundError("speaker table for %s not found under %s: run from the reference checkout (it holds data/%s_info/"
                            "%s_speakers.txt) or point VQWN_SPEAKER_TABLES at a directory that contains data/"
                            % (dataset, list(roots), dataset, dataset))


def speaker_onehot(speakers, speaker_to_int, num_speakers):
    """generate.py:58-61: [B,1,N] one-hot; 'None' (any case) leaves the row all-zero."""
    one = np.zeros((len(speakers), 1, num_speakers), dtype=np.float32)
    for i, s in enumerate(speakers):
        if s.lower() != "none":
            one[i, 0, speaker_to_int[s]] = 1
    return one


mu_law_decode_np = mu_law_ops.mu_law_decode_np
