import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def write_speaker_table(root, dataset="vctk", names=None):
    """a synthetic speaker table in the reference's format ("<speaker>, <index>" per line, generate.py:46-57); the
    reference's own tables are its data and are not shipped with this repo"""
    import os
    if names is None:
        names = ["p%d" % (224 + i) for i in range(109)]      # p225 -> 1 (index 0 is what "None" maps to, model.py:22)
    d = os.path.join(str(root), "data", "%s_info" % dataset)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "%s_speakers.txt" % dataset)
    with open(path, "w") as f:
        for i, n in enumerate(names):
            f.write("%s, %d\n" % (n, i))
    return path
