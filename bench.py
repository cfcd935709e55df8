#!/usr/bin/env python
"""Headline benchmark: generated audio samples/s of the WaveNet fast-generation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seconds S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one whole pass of the hot path over one batch of synthetic input: init_ops + T
autoregressive steps for B streams per GPU (BASELINE.json config 3: greedy, B=64, 4 s @ 16 kHz,
4 speaker conditions; N GPUs = config 4's utterance sharding, 64 streams per GPU, weak scaling,
no data-path collective).  `value` is device-timed with the condition tensor already resident in
HBM; `e2e` goes through the public API with pinned HOST buffers (H2D of the condition, D2H of audio
and indices inside the timed region).  `--impl reference` times the CPU restatement of the
reference's TensorFlow path (oracle/oracle.py; TensorFlow 1.x cannot be installed here) on the
host cores, on a bounded window of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 36386816          # minimal algorithmic work per generated sample per stream
QUEUE_BYTES_PER_SAMPLE = 92160 + 512 + 4
NCU_DRAM_BYTES_PER_STEP = 85.3e6    # dram read + write per time step, wavenet_fp32_cluster at 64 streams (profiles/)
METRIC = "generated audio samples/sec"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(B, T, rank):
    """synthetic inputs of SURVEY 8d for this rank's shard (distinct streams per rank)"""
    import vqvae_wavenet_b200 as pkg
    from vqvae_wavenet_b200 import synthetic
    cfg = pkg.EngineConfig()
    w = synthetic.make_weights(cfg, seed=1234, peaked=True)
    F = T // 64
    z_e = synthetic.synthetic_z_e(cfg, B, F, seed=1235 + 1000 * rank)
    spk = (np.arange(B, dtype=np.int32) + rank * B) % 4
    return cfg, w, z_e, spk, F


def vq_bench(eng, vq_n):
    """BASELINE config 2: VQ lookups/s at 64 x 104 vectors (launch-sized) and at vq_n (roofline-sized),
    both kernels; algorithmic bytes per vector = 256 in + 256 out + 8 index (+131 KB codebook once)."""
    out = {}
    rng = np.random.default_rng(99)
    zbig = (0.13 * rng.standard_normal((vq_n, 64))).astype(np.float32)
    eng.vq_upload(zbig)
    hbm = load_peaks()["hbm"]
    for kern in ("tensor", "direct"):
        eng.set_vq_kernel(kern)
        for n in (6656, vq_n):
            for _ in range(3):
                eng.vq_resident(n)
            ts = []
            for _ in range(7):
                eng.vq_resident(n)
                ts.append(eng.last_kernel_ms)
            ms = float(np.median(ts))
            out["%s_n%d" % (kern, n)] = {"lookups_per_s": n / (ms * 1e-3), "ms": ms, "kernel": eng.last_kernel_name,
                                         "hbm_gbs": (n * 520 + 131072) / (ms * 1e-3) / 1e9,
                                         "hbm_frac": (n * 520 + 131072) / (ms * 1e-3) / 1e9 / hbm}
    eng.set_vq_kernel("auto")
    return out


def cpu_baseline_window(B, steps, warm):
    """the oracle port (FastWavenet: per-step matmuls + FIFO deques + NumPy decode) on host cores"""
    import torch
    from oracle import oracle as O
    cfg = O.Config()
    w = O.make_weights(cfg, seed=1234, peaked=True)
    z_e = O.synthetic_z_e(cfg, w, B, 2, seed=1235, kind="scaled")
    _, cond = O.encode_condition(z_e, np.arange(B) % 4, w)
    net = O.FastWavenet(cfg, w, B)
    audio = np.zeros((B, 1), dtype=np.float32)
    t0 = None
    for i in range(warm + steps):
        if i == warm:
            t0 = time.perf_counter()
        probs, _ = net.step(audio, cond[:, 0])
        audio = O.decode(probs, "greedy")[:, None]
    dt = time.perf_counter() - t0
    return B * steps / dt, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = oracle port (TensorFlow 1.x
    is not installable here), all host threads, bounded window per step."""
    if rank != 0:
        return
    B, win = args.batch, args.ref_window
    vals = []
    cores = None
    for i in range(args.warmup + args.steps):
        v, dt, cores = cpu_baseline_window(B, win, 4)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "greedy fast generation, B=%d streams, %d-step window of the 4 s job per bench step "
                               "(CPU restatement of the TF-1.x path: oracle/oracle.py FastWavenet)" % (B, win)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d time steps x %d streams per bench step" % (win, B)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="streams per GPU")
    ap.add_argument("--seconds", type=float, default=4.0, help="audio seconds per stream (16 kHz)")
    ap.add_argument("--mode", default="greedy", choices=["greedy", "sample"])
    ap.add_argument("--precision", default=os.environ.get("VQWN_PRECISION", "fp32"))
    ap.add_argument("--no-bf16", action="store_true", help="skip the secondary bf16 tensor-core measurement")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-stream latency measurement")
    ap.add_argument("--ref-window", type=int, default=192)
    ap.add_argument("--cpu-window", type=int, default=384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--vq-n", type=int, default=1 << 20)
    ap.add_argument("--vq-only", action="store_true", help="only the VQ lookups/s micro-benchmark (config 2)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import vqvae_wavenet_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    T = int(round(args.seconds * 16000)) // 512 * 512          # generate.py:39 trim
    cfg, w, z_e, spk, F = make_workload(B, T, rank)

    eng = pkg.Engine(cfg, device=local_rank, max_batch=B)
    eng.set_weights(w)
    eng.set_precision(args.precision)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    if args.vq_only:
        print(json.dumps({"metric": "VQ lookups/sec", "unit": "lookups/s", "vq": vq_bench(eng, args.vq_n)}), flush=True)
        eng.close()
        return

    # condition tensor = what generate.py:92 evaluates (VQ + gather + speaker concat on the device)
    _, cond = eng.encode_condition(z_e, spk)
    u = None
    if args.mode == "sample":
        u = np.random.default_rng(1236 + rank).random((T, B))
        eng.upload_uniforms(u)
    eng.upload_condition(cond)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(args.warmup):
        eng.generate_resident(B, F, T, args.mode, seed=1)
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        eng.generate_resident(B, F, T, args.mode, seed=1)
        kernel_ms.append(eng.last_kernel_ms)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    elapsed_ms = e0.elapsed_time(e1)
    t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * B * T / (ms_per_step * 1e-3)
    kernel_name = eng.last_kernel_name
    k_ms = float(np.mean(kernel_ms))

    # ---------------------------------------------------------------- end-to-end through the public API (host buffers)
    cond_pin = torch.from_numpy(cond).pin_memory()
    audio_pin = torch.empty((B, T), dtype=torch.float32).pin_memory()
    idx_pin = torch.empty((B, T), dtype=torch.int32).pin_memory()
    u_pin = torch.from_numpy(u).pin_memory().numpy() if u is not None else None
    eng.generate(cond_pin.numpy(), T, mode=args.mode, uniforms=u_pin, seed=1,
                 out_audio=audio_pin.numpy(), out_idx=idx_pin.numpy())       # warm
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 2))
    f0.record(stream)
    for _ in range(e2e_steps):
        eng.generate(cond_pin.numpy(), T, mode=args.mode, uniforms=u_pin, seed=1,
                     out_audio=audio_pin.numpy(), out_idx=idx_pin.numpy())
    f1.record(stream)
    barrier()
    t = torch.tensor([f0.elapsed_time(f1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_value = world * B * T / (e2e_ms * 1e-3)
    h2d = cond.nbytes + (u.nbytes if u is not None else 0)
    d2h = audio_pin.numel() * 4 + idx_pin.numel() * 4
    # sanity: e2e result equals the resident result
    a_res, i_res = eng.download_output(B, T)
    same = bool(np.array_equal(i_res, idx_pin.numpy()))

    # ---------------------------------------------------------------- bf16 tensor-core kernel, same workload (secondary)
    # value / e2e above stay on the float32 path (1e-3 parity with the reference); VQWN_PREC_BF16 (tcgen05, logits
    # within 2e-2) is measured beside it: one short warm launch, one timed launch of the full job.
    bf16 = None
    if args.precision == "fp32" and not args.no_bf16:
        try:
            eng.set_precision("bf16")
            eng.generate_resident(B, F, F * 8 if F * 8 < T else T, args.mode, seed=1)     # warm launch: 8 steps per frame
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            eng.generate_resident(B, F, T, args.mode, seed=1)
            g1.record(stream)
            barrier()
            tb = torch.tensor([g0.elapsed_time(g1)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            bms = float(tb.item())
            btf = B * FLOP_PER_SAMPLE / (bms * 1e-3 / T) / 1e12
            bf16 = {"value": world * B * T / (bms * 1e-3), "unit": "samples/s", "ms_per_step": bms,
                    "us_per_time_step": bms * 1e3 / T, "kernel": eng.last_kernel_name,
                    "achieved_tflops_per_gpu": btf, "tensor_frac": btf / load_peaks()["tensor_sustained"],
                    "tolerance": "teacher-forced logits within 2e-2 of max|logit| (tests/test_gpu_parity.py::test_bf16_*)"}
        except NotImplementedError:
            bf16 = None
        eng.set_precision("fp32")

    # ---------------------------------------------------------------- single-stream per-step latency (BASELINE config 1)
    # one stream, 1 s of audio (16 384 samples), greedy: the reference's own CPU-runnable case, latency-bound
    latency = None
    if rank == 0 and B > 1 and args.mode == "greedy" and not args.no_latency:
        T1 = 16384
        z1 = z_e[:1, :min(F, T1 // 64)]
        if z1.shape[1] == T1 // 64:
            _, c1 = eng.encode_condition(z1, spk[:1])
            eng.upload_condition(c1)
            eng.generate_resident(1, T1 // 64, 2048, args.mode, seed=1)            # warm
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            h0.record(stream)
            eng.generate_resident(1, T1 // 64, T1, args.mode, seed=1)
            h1.record(stream)
            torch.cuda.synchronize()
            lms = h0.elapsed_time(h1)
            latency = {"streams": 1, "time_steps": T1, "us_per_time_step": lms * 1e3 / T1, "samples_per_s": T1 / (lms * 1e-3),
                       "kernel": eng.last_kernel_name}
            eng.upload_condition(cond)

    # ---------------------------------------------------------------- VQ lookups/s (secondary metric)
    vq = vq_bench(eng, args.vq_n) if rank == 0 else {}

    # ---------------------------------------------------------------- roofline + CPU baseline + report
    if rank == 0:
        peaks = load_peaks()
        step_s = k_ms * 1e-3 / T                                 # one autoregressive step of B streams
        achieved_tf = B * FLOP_PER_SAMPLE / step_s / 1e12
        achieved_gbs = B * QUEUE_BYTES_PER_SAMPLE / step_s / 1e9
        roof_step = max(B * FLOP_PER_SAMPLE / (peaks["tensor_sustained"] * 1e12),
                        B * QUEUE_BYTES_PER_SAMPLE / (peaks["hbm"] * 1e9))
        # DRAM bytes per launch from the committed ncu --set full capture of the same kernel and stream count
        # (profiles/r1_cluster_full_summary.txt: 43.68 GB for a T = 512 launch = 85.3 MB per time step, almost all of it
        # the 78.6 MB of float32 weights, which do not fit the L2 next to the queue traffic and stream from HBM once
        # per step); other kernels / stream counts: not captured -> null
        traffic = NCU_DRAM_BYTES_PER_STEP * T if (kernel_name == "wavenet_fp32_cluster" and B == 64) else None
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": achieved_tf / peaks["tensor_sustained"], "traffic": traffic,
                    "fp32_ffma_note": "float32 CUDA-core kernel: ncu fma pipe 26.8 % of peak-active; a 3-register FFMA "
                                      "issues every 2nd cycle per SM sub-partition, so 50 % is this pipe's ceiling",
                    "kernel": kernel_name, "kernel_ms": k_ms, "us_per_time_step": step_s * 1e6,
                    "roofline_us_per_time_step": roof_step * 1e6, "hbm_achieved_gbs": achieved_gbs,
                    "hbm_frac": achieved_gbs / peaks["hbm"], "peak_source": peaks["source"] + ", sustained bf16"}
        cpu = None
        if not args.no_cpu_baseline:
            v, dt, cores = cpu_baseline_window(B, args.cpu_window, 8)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "%d time steps x %d streams of the same workload (%.1f s of CPU work)" % (args.cpu_window, B, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16(f32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "%s fast generation, %d streams/GPU x %.3f s @16 kHz (T=%d), 4 speaker conditions, "
                                   "default 30-layer WaveNet + K=512 VQ condition" % (args.mode, B, T / 16000.0, T),
                       "streams_per_gpu": B, "time_steps": T, "mode": args.mode, "precision": args.precision,
                       "l2": "dilation-queue state %.0f MB per GPU > 126 MB L2 (no flush needed)" % (B * 6.285),
                       "sharding": "contiguous stream slices per GPU, no collective"},
            "realtime_factor": value / 16000.0,
            "us_per_time_step": ms_per_step * 1e3 / T,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "matches_resident": same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "bf16_tensor_core": bf16,
            "single_stream_latency": latency,
            "vq": vq,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
