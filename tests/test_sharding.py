"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes, a stand-in engine (the sharding
code only needs an object with .generate; the real engine needs a GPU and is covered by
tests/test_gpu_parity.py::test_shard_equals_unsharded)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _PerStreamEngine:
    """deterministic per-stream function of (condition, uniforms): what independence means"""

    offset = 0

    def set_stream_offset(self, offset):
        self.offset = int(offset)

    def generate(self, cond, length, mode="greedy", uniforms=None, seed=0):
        B = cond.shape[0]
        base = (np.abs(cond).sum(axis=(1, 2))[:, None] * 1000).astype(np.int64)
        t = np.arange(length, dtype=np.int64)[None, :]
        idx = (base + t * 7) % 256
        if uniforms is not None:
            idx = (idx + (uniforms.T * 256).astype(np.int64)) % 256
        return (idx / 255.0).astype(np.float32), idx.astype(np.int32)


def _worker(rank, world, port, B, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vqvae_wavenet_b200 import sharding
    rng = np.random.default_rng(0)
    cond = rng.standard_normal((B, 3, 128)).astype(np.float32)
    u = rng.random((T, B))
    a, i = sharding.generate_sharded(_PerStreamEngine(), cond, T, mode="sample", uniforms=u)
    ra, ri = _PerStreamEngine().generate(cond, T, mode="sample", uniforms=u)
    ok = np.array_equal(a, ra) and np.array_equal(i, ri)
    # time-like reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == float(world)
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.destroy_process_group()


def test_stream_slices_cover_exactly():
    from vqvae_wavenet_b200.sharding import stream_slice
    for total in (1, 5, 64, 512, 513):
        for world in (1, 2, 4, 8):
            sl = [stream_slice(total, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == total
            assert all(sl[r][1] == sl[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in sl]
            assert max(sizes) - min(sizes) <= 1
    assert stream_slice(512, 3, 8) == (192, 256)      # BASELINE config 4: 64 streams per GPU


def test_sharded_equals_unsharded_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 5, 16, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "rank0").read() == "ok" and open(tmp_path / "rank1").read() == "ok"
