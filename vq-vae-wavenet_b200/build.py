"""In-tree build of libvqwn.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(HERE, "csrc", "api.cu")]
OUTPUT = os.path.join(HERE, "libvqwn.so")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]


def _newest_source_mtime():
    m = 0.0
    for root in (os.path.join(HERE, "csrc"), os.path.join(HERE, "..", "include")):
        for dp, _, files in os.walk(root):
            for f in files:
                m = max(m, os.path.getmtime(os.path.join(dp, f)))
    return m


def build_library(force=False, verbose=False):
    if not force and os.path.exists(OUTPUT) and os.path.getmtime(OUTPUT) >= _newest_source_mtime():
        return OUTPUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUTPUT] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return OUTPUT
