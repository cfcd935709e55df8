"""Multi-GPU partitioning of the path: streams are independent, so a job is cut into
contiguous stream slices, one process + one device handle per GPU, replicated weights, no
data-path collective; the host gathers the per-rank outputs."""
import numpy as np


def stream_slice(total, rank, world):
    """contiguous slice [lo, hi) of `total` streams owned by `rank` (sizes differ by at most 1)"""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(cond, uniforms, rank, world):
    """this rank's condition [B_r,F,C] and uniforms [T,B_r] (uniforms may be None)"""
    lo, hi = stream_slice(cond.shape[0], rank, world)
    u = None if uniforms is None else np.ascontiguousarray(uniforms[:, lo:hi])
    return np.ascontiguousarray(cond[lo:hi]), u, (lo, hi)


def generate_sharded(engine, cond, length, mode="greedy", uniforms=None, seed=0, group=None):
    """Runs this rank's slice on `engine` and gathers [B,T] audio / indices on every rank through
    torch.distributed (host tensors; gloo or nccl-with-CPU fallback is the caller's choice).
    Without an initialised process group it is the single-GPU call."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return engine.generate(cond, length, mode=mode, uniforms=uniforms, seed=seed)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    c, u, (lo, hi) = shard_inputs(cond, uniforms, rank, world)
    if hi > lo:
        engine.set_stream_offset(lo)      # seeded draws (uniforms=None) are keyed on the global stream index
        audio, idx = engine.generate(c, length, mode=mode, uniforms=u, seed=seed)
        engine.set_stream_offset(0)
    else:
        audio = np.zeros((0, length), np.float32)
        idx = np.zeros((0, length), np.int32)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, audio, idx), group=group)
    B = cond.shape[0]
    out_a = np.empty((B, length), np.float32)
    out_i = np.empty((B, length), np.int32)
    for lo_r, hi_r, a, i in parts:
        out_a[lo_r:hi_r] = a
        out_i[lo_r:hi_r] = i
    return out_a, out_i
