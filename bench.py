#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): batched C2C FFT throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

One "step" is one compute call of the hot path over the whole synthetic batch (steps alternate compute_forward and
compute_backward with backward_scale = 1/N so that the in-place data stays bounded and the run ends with a
full-size round-trip identity check).  Default workload = BASELINE.json configs[1] (C2): 1-D C2C fp32 N=4096,
batch 65536, in place, interleaved -- 2 GiB per GPU, far larger than the 126 MB L2, so no flush is needed.

Printed JSON line (rank 0): metric GFLOP/s = 5*N*log2(N)*batch/t (the reference's own model,
/root/reference/test/bench/utils/ops_estimate.hpp:34-50), device-timed `value`, `e2e` through the C ABI with
pinned HOST buffers (pfft_compute_host: H2D + compute + D2H per step), `roofline` for the dominant kernel against
MEASURED_PEAKS.json, `cpu_baseline` = the C oracle port timed on a bounded sample on the host cores.
Multi-GPU (torchrun): the configured batch is sharded over the ranks (65536 / N transforms per GPU for C2: strong
scaling, BASELINE.json config 2 / SURVEY 8d row "C2 per GPU at 8 GPUs"), no data-path collective; the same run
also reports the weak-scaling figure (`weak_scaling`: the full batch on every GPU) and the slab-decomposed C5
(`slab_c5`: 512^3 over the N GPUs, peer-store and NCCL exchange, rel-L2 against numpy.fft.fftn).
`--impl reference` times the reference's CPU algorithm (oracle port; the SYCL reference cannot be built here).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: lengths, batch, scalar, placement, storage
    "C1": dict(lengths=[64], batch=1024, scalar="float", inplace=False, split=False,
               desc="1D C2C fp32 N=64 batch=1024 interleaved out-of-place"),
    "C2": dict(lengths=[4096], batch=65536, scalar="float", inplace=True, split=False,
               desc="1D C2C fp32 N=4096 batch=65536 in-place interleaved"),
    "C3": dict(lengths=[1000], batch=100000, scalar="float", inplace=False, split=True,
               desc="1D C2C fp32 N=1000 batch=100000 split, fwd stride 2/dist 2048/off 7, bwd dist 1024/off 3, bwd scale 1e-3",
               forward_strides=[2], forward_distance=2048, forward_offset=7, backward_strides=[1],
               backward_distance=1024, backward_offset=3, backward_scale=1e-3),
    "C3B": dict(lengths=[1000], batch=100000, scalar="float", inplace=False, split=True,
                desc="1D C2C fp32 N=1000 batch=100000 split, batch-interleaved both domains (stride 100000, distance 1), bwd scale 1e-3",
                forward_strides=[100000], forward_distance=1, backward_strides=[100000], backward_distance=1,
                backward_scale=1e-3),
    "C4": dict(lengths=[1 << 24], batch=8, scalar="double", inplace=False, split=False,
               desc="1D C2C fp64 N=2^24 batch=8 out-of-place (global level)"),
    "C5": dict(lengths=[512, 512, 512], batch=1, scalar="float", inplace=False, split=False,
               desc="3D C2C fp32 512^3 default strides out-of-place"),
    "L1D": dict(lengths=[65536], batch=2048, scalar="float", inplace=False, split=False,
                desc="1D C2C fp32 N=65536 batch=2048 out-of-place (reference bench_float large_1d)"),
    "S16": dict(lengths=[16], batch=8 * 1024 * 1024, scalar="float", inplace=False, split=False,
                desc="1D C2C fp32 N=16 batch=8Mi out-of-place (reference bench_float small_1d)"),
    "M512": dict(lengths=[512], batch=256 * 1024, scalar="float", inplace=False, split=False,
                 desc="1D C2C fp32 N=512 batch=256Ki out-of-place (packed rows, N = 8^3 tile kernel)"),
    "M1024": dict(lengths=[1024], batch=128 * 1024, scalar="float", inplace=False, split=False,
                  desc="1D C2C fp32 N=1024 batch=128Ki out-of-place (packed rows, three-radix kernel 16x8x8)"),
    "M2048": dict(lengths=[2048], batch=64 * 1024, scalar="float", inplace=False, split=False,
                  desc="1D C2C fp32 N=2048 batch=64Ki out-of-place (packed rows, three-radix kernel 16x16x8)"),
    "M8192": dict(lengths=[8192], batch=16 * 1024, scalar="float", inplace=False, split=False,
                  desc="1D C2C fp32 N=8192 batch=16Ki out-of-place (packed rows, 16x16x32)"),
    "D512": dict(lengths=[512], batch=128 * 1024, scalar="double", inplace=False, split=False,
                 desc="1D C2C fp64 N=512 batch=128Ki out-of-place (packed rows, N = 8^3 tile kernel)"),
    "D1024": dict(lengths=[1024], batch=64 * 1024, scalar="double", inplace=False, split=False,
                  desc="1D C2C fp64 N=1024 batch=64Ki out-of-place (packed rows, 16x8x8)"),
    "D2048": dict(lengths=[2048], batch=32 * 1024, scalar="double", inplace=False, split=False,
                  desc="1D C2C fp64 N=2048 batch=32Ki out-of-place (packed rows, 16x16x8)"),
    "D4096": dict(lengths=[4096], batch=16 * 1024, scalar="double", inplace=False, split=False,
                  desc="1D C2C fp64 N=4096 batch=16Ki out-of-place (packed rows, N = 16^3)"),
    "M256": dict(lengths=[256], batch=512 * 1024, scalar="float", inplace=False, split=False,
                 desc="1D C2C fp32 N=256 batch=512Ki out-of-place (reference bench_float medium_small_1d)"),
    # REAL domain: the reference's registered real benchmark set (test/bench/utils/reference_dft_set.hpp), dense half
    # spectrum (backward_distance = N/2 + 1).  forward = real-to-complex, backward = complex-to-real.
    "R32": dict(lengths=[32], batch=8 * 1024 * 1024, scalar="float", inplace=False, split=False, real=True,
                desc="1D R2C/C2R fp32 N=32 batch=8Mi out-of-place, dense half spectrum (reference real small_1d)"),
    "R512": dict(lengths=[512], batch=512 * 1024, scalar="float", inplace=False, split=False, real=True,
                 desc="1D R2C/C2R fp32 N=512 batch=512Ki out-of-place, dense half spectrum (reference real medium_small_1d)"),
    "R8192": dict(lengths=[8192], batch=32 * 1024, scalar="float", inplace=False, split=False, real=True,
                  desc="1D R2C/C2R fp32 N=8192 batch=32Ki out-of-place, dense half spectrum (reference real medium_large_1d)"),
    "R131072": dict(lengths=[131072], batch=2048, scalar="float", inplace=False, split=False, real=True,
                    desc="1D R2C/C2R fp32 N=131072 batch=2Ki out-of-place, dense half spectrum (reference real large_1d)"),
}


def flops_of(cfg) -> float:
    n = 1
    for l in cfg["lengths"]:
        n *= l
    # (REAL: half the complex count, the convention of the reference's ops estimate for real transforms)
    return (2.5 if cfg.get("real") else 5.0) * n * math.log2(n) * cfg["batch"]


def bytes_of(cfg) -> float:
    n = 1
    for l in cfg["lengths"]:
        n *= l
    sc = 8 if cfg["scalar"] == "double" else 4
    if cfg.get("real"):  # one read of the real row + one write of the n/2 + 1 complex outputs (or the reverse)
        return (n * sc + (n // 2 + 1) * 2 * sc) * cfg["batch"]
    return 2.0 * n * cfg["batch"] * 2 * sc


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons while the timed region runs (B200_PROFILING.md recipe).  NVML through
    pynvml (~1 ms per sample, so that even a 10 ms timed region is covered); `nvidia-smi` as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (sm_mhz, max_mhz, set(reasons))
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        reasons = set()
        for name, bit in (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)):
            if mask & bit:
                reasons.add(name)
        self.samples.append((mhz, self.max_mhz, reasons))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            reasons = {nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")}
            self.samples.append((float(f[1]), float(f[2]), reasons))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.002 if self.nvml is not None else 0.1)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = max((s[1] for s in self.samples), default=0.0)
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def load_oracle():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    for name in ("pfft_oracle_fft_f32", "pfft_oracle_fft_f64"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                       ctypes.c_double, ctypes.c_int]
    return lib


def cpu_port_time(cfg, sample_batch: int, repeats: int = 1):
    """Time the C oracle (the reference's algorithm restated for the CPU) on `sample_batch` packed transforms of the
    workload's flattened length, all host cores.  Returns (seconds per call, cores)."""
    import numpy as np

    lib = load_oracle()
    n = 1
    for l in cfg["lengths"]:
        n *= l  # N-D is timed as the flattened 1-D length (same FLOP model as the reference's benches)
    dbl = cfg["scalar"] == "double"
    rng = np.random.Generator(np.random.SFC64(1))
    x = rng.uniform(-1, 1, (sample_batch, n, 2)).astype(np.float64 if dbl else np.float32)
    out = np.empty_like(x)
    fn = lib.pfft_oracle_fft_f64 if dbl else lib.pfft_oracle_fft_f32
    cores = len(os.sched_getaffinity(0))
    fn(x.ctypes.data, out.ctypes.data, n, min(sample_batch, cores), 0, 1.0, cores)  # warm tables
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc = fn(x.ctypes.data, out.ctypes.data, n, sample_batch, 0, 1.0, cores)
        best = min(best, time.perf_counter() - t0)
        assert rc == 0
    return best, cores


def cpu_sample_batch(cfg) -> int:
    """bounded sample: about 0.5 GFLOP-equivalents of the port's speed (~10-20 s on 8 cores)"""
    per = flops_of(cfg) / cfg["batch"]
    target_flops = 1.5e10
    sample = int(max(1, min(cfg["batch"], target_flops // per)))
    return cfg["batch"] if sample >= 0.9 * cfg["batch"] else sample  # nearly everything: take the whole workload


def run_reference_arm(args, cfg, name):
    # NCCL_DEBUG=VERSION makes NCCL print a banner on stdout, in front of the one JSON line this program owes
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = cpu_sample_batch(cfg)
    per = flops_of(cfg) / cfg["batch"]
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_time(cfg, min(sample, 64))
    times = []
    cores = 1
    for _ in range(args.steps):
        t, cores = cpu_port_time(cfg, sample)
        times.append(t)
    t = sum(times) / len(times)
    val = per * sample / t / 1e9
    line = {
        "impl": "reference", "metric": "batched_c2c_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if cfg["scalar"] == "double" else "f32",
        "data": "synthetic",
        "config": {"workload": f"{name}: {cfg['desc']}", "total_batch": cfg["batch"],
                   "sample": f"{sample} of {cfg['batch']} transforms per step"},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} transforms of the workload per step, C oracle port of the reference "
                                   f"algorithm (SYCL reference not buildable here)"},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_slab(args, cfg, name, rank, local_rank, world, dev):
    """--config C5 --gpus N>1: ONE 512^3 transform slab-decomposed over the N GPUs (strong scaling).  Local passes =
    the C-ABI plans; exchange = FFT stores into peer memory (symmetric memory over NVLink) or NCCL all-to-all
    (--exchange nccl).  The spectrum is left y-slab distributed (portfft_b200/distributed.py)."""
    import torch
    import torch.distributed as dist

    import portfft_b200 as pf
    from portfft_b200.distributed import slab_fft3d

    n0, n1, n2 = cfg["lengths"]
    plan = slab_fft3d(cfg["lengths"], cfg["scalar"], exchange=args.exchange, device=dev)
    g = plan.geom
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.rand(g.xl, n1, n2, 2, generator=gen, device=dev, dtype=torch.float32) * 2 - 1)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        out = plan.forward(x)
    barrier()
    e_in = (x.abs().double() ** 2).sum()
    e_out = (out.abs().double() ** 2).sum()
    t = torch.stack([e_in, e_out])
    dist.all_reduce(t)
    parseval = float(abs(t[1] / (t[0] * n0 * n1 * n2) - 1.0))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pf.total_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        plan.forward(x)
    ev1.record()
    barrier()
    launches = pf.total_launches() - l0
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    tm = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    # e2e: pinned host slab in, y-slab of the spectrum back to the host, every step
    h_in = torch.empty(x.shape, dtype=x.dtype).pin_memory()
    h_in.copy_(x)
    h_out = torch.empty((n0, g.yb, n2), dtype=x.dtype).pin_memory()
    xd = torch.empty_like(x)
    e_steps = max(1, min(args.steps, 4))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        xd.copy_(h_in, non_blocking=True)
        h_out.copy_(plan.forward(xd), non_blocking=True)
        torch.cuda.synchronize(dev)
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    te = float(te.item())
    if rank == 0:
        flops = flops_of(cfg)
        peak, peak_src = measured_peak()
        nv_bytes = (world - 1) * g.block_elems * 8
        line = {
            "metric": "batched_c2c_gflops", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{name}: {cfg['desc']}", "decomposition": f"x-slabs of {g.xl} planes per GPU, one "
                       f"exchange step ({args.exchange}), spectrum left y-slab distributed",
                       "l2": "per-GPU slab %.0f MiB" % (g.slab_elems * 8 / 2**20)},
            "hbm_gbs": bytes_of(cfg) / (ms * 1e-3) / 1e9,
            "roofline": {"bound": "nvlink", "achieved": nv_bytes / (ms * 1e-3) / 1e9, "peak": 770.0, "unit": "GB/s",
                         "frac": nv_bytes / (ms * 1e-3) / 1e9 / 770.0, "traffic": None,
                         "peak_source": "measured peer copy per direction per GPU (B200_PROFILING.md)",
                         "note": "achieved = bytes each GPU sends over NVLink / whole step time (local passes included)"},
            "cpu_baseline": None,
            "e2e": {"value": flops / te / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": g.slab_elems * 8,
                    "d2h_bytes_per_step": g.slab_elems * 8, "steps": e_steps, "ms_per_step": te * 1e3,
                    "api": "slab_fft3d.forward with pinned host slabs"},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "parseval_rel": parseval,
        }
        print(json.dumps(line), flush=True)
    plan.destroy()


def scipy_cpu_time(cfg, sample_batch: int):
    """Second CPU baseline (SURVEY 8d Plan B ii): scipy.fft (pocketfft, the library behind the numpy oracle the
    reference's tests use) on the same kind of sample, all host cores.  Returns (seconds, cores) or None."""
    try:
        import numpy as np
        import scipy.fft as sfft
    except Exception:
        return None
    dbl = cfg["scalar"] == "double"
    rng = np.random.Generator(np.random.SFC64(1))
    shape = [sample_batch] + list(cfg["lengths"])
    if cfg.get("real"):
        x = rng.uniform(-1, 1, shape).astype(np.float64 if dbl else np.float32)
        fn = sfft.rfftn
    else:
        x = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128 if dbl else np.complex64)
        fn = sfft.fftn
    cores = len(os.sched_getaffinity(0))
    axes = tuple(range(1, len(shape)))
    fn(x[:max(1, min(sample_batch, cores))], axes=axes, workers=cores)  # plan cache warm-up
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        fn(x, axes=axes, workers=cores)
        best = min(best, time.perf_counter() - t0)
    return best, cores


def copy_ceiling(torch, dist, dev, world, h_in, h_out, reps=3):
    """What the host link allows for one e2e step, all ranks at once: the step's H2D bytes and D2H bytes as two plain
    pinned cudaMemcpyAsync streams running concurrently (PCIe is full duplex), no kernels.  Seconds, max over ranks."""
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_in = [torch.empty_like(t, device=dev) for t in h_in]
    d_out = d_in if h_out is h_in else [torch.empty_like(t, device=dev) for t in h_out]
    h_back = [torch.empty_like(t).pin_memory() for t in h_out]  # D2H target distinct from the H2D source
    best = float("inf")
    for _ in range(reps + 1):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_up):
            for a, b in zip(d_in, h_in):
                a.copy_(b, non_blocking=True)
        with torch.cuda.stream(s_dn):
            for a, b in zip(h_back, d_out):
                a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(dev)
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


def run_batched(args, cfg, name, batch, rank, local_rank, world, dev, steps, warmup, with_e2e, use_graph):
    """One measurement of the batched path: every rank commits the descriptor for `batch` transforms (its shard; no
    data-path collective), runs `warmup` + `steps` alternating forward / backward computes, device-timed with CUDA
    events on the launching stream, max over ranks.  Returns a dict of the raw figures."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import portfft_b200 as pf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n_flat = 1
    for l in cfg["lengths"]:
        n_flat *= l
    real = bool(cfg.get("real"))
    d = pf.descriptor(cfg["lengths"], cfg["scalar"], pf.domain.REAL if real else pf.domain.COMPLEX)
    d.number_of_transforms = batch
    if real:
        d.backward_distance = n_flat // 2 + 1
    d.placement = pf.placement.IN_PLACE if cfg["inplace"] else pf.placement.OUT_OF_PLACE
    d.complex_storage = pf.complex_storage.SPLIT_COMPLEX if cfg["split"] else pf.complex_storage.INTERLEAVED_COMPLEX
    for k in ("forward_strides", "forward_distance", "forward_offset", "backward_strides", "backward_distance",
              "backward_offset", "backward_scale"):
        if k in cfg:
            v = cfg[k]
            if k.endswith("_strides") and batch != cfg["batch"]:
                v = [batch if x == cfg["batch"] else x for x in v]  # batch-interleaved shard: stride = local batch
            setattr(d, k, v)
    if "backward_scale" not in cfg:
        d.backward_scale = 1.0 / n_flat
    stream = torch.cuda.current_stream(dev)
    plan = d.commit(stream, local_rank)
    fdt = torch.float64 if cfg["scalar"] == "double" else torch.float32
    n_fwd, n_bwd = d.get_input_count(pf.direction.FORWARD), d.get_input_count(pf.direction.BACKWARD)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    planes = 2 if cfg["split"] else 1

    def alloc(count):
        if cfg["split"]:
            return [torch.rand(count, dtype=fdt, device=dev, generator=g) * 2 - 1 for _ in range(2)]
        return [torch.view_as_complex((torch.rand(count, 2, dtype=fdt, device=dev, generator=g) * 2 - 1))]

    fwd_buf = [torch.rand(n_fwd, dtype=fdt, device=dev, generator=g) * 2 - 1] if real else alloc(n_fwd)
    bwd_buf = fwd_buf if cfg["inplace"] else alloc(n_bwd)
    orig = [t.clone() for t in fwd_buf] if n_fwd * (16 if fdt == torch.float64 else 8) <= (8 << 30) else None

    def step(i, q=stream):
        if i % 2 == 0:
            if cfg["inplace"]:
                plan.compute_forward(*fwd_buf, queue=q)
            else:
                plan.compute_forward(*fwd_buf, *bwd_buf, queue=q)
        else:
            if cfg["inplace"]:
                plan.compute_backward(*fwd_buf, queue=q)
            else:
                plan.compute_backward(*bwd_buf, *fwd_buf, queue=q)

    if (warmup + steps) % 2 == 1:
        warmup += 1  # even number of steps: the data ends where it started (round-trip check below)
    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pf.total_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if use_graph:
        # launch-bound workloads: the K steps are captured once into a CUDA graph (stream capture of the same C-ABI
        # calls) and the timed region replays it; the work on the device is identical
        side = torch.cuda.Stream(dev)
        side.wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            cap = torch.cuda.current_stream(dev)
            for i in range(steps):
                step(warmup + i, cap)
        stream.wait_stream(side)
        with torch.cuda.stream(stream):
            graph.replay()  # warm replay (an even number of steps: the data is back where it started)
        barrier()
    barrier()
    ev0.record(stream)
    if graph is not None:
        with torch.cuda.stream(stream):
            graph.replay()
    else:
        for i in range(steps):
            step(warmup + i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pf.total_launches() - launches0
    if graph is not None:  # the replayed graph holds the launches counted at capture time
        launches = steps * plan.num_launches(pf.direction.FORWARD)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    res = {"ms_per_step": float(tmax.item()) / steps, "steps": steps, "warmup": warmup, "launches": int(launches),
           "clocks": sampler.summary() if rank == 0 else None, "batch": batch,
           "n_pass": plan.num_launches(pf.direction.FORWARD), "l2_chunk": plan.l2_chunk(), "roundtrip": None,
           "e2e": None}

    # ---- round-trip identity at full size (size-independent parity property) ------------------------------------
    if orig is not None:
        num = sum(float(torch.linalg.vector_norm((a - b).reshape(-1)).item()) ** 2 for a, b in zip(fwd_buf, orig))
        den = sum(float(torch.linalg.vector_norm(b.reshape(-1)).item()) ** 2 for b in orig)
        res["roundtrip"] = math.sqrt(num / den)
        del orig

    # ---- e2e through the C ABI with pinned host buffers ----------------------------------------------------------
    if with_e2e:
        esz = (2 if not cfg["split"] else 1)
        h_in = [torch.empty(n_fwd * (1 if real else esz), dtype=fdt).pin_memory() for _ in range(planes)]
        h_out = h_in if cfg["inplace"] else [torch.empty(n_bwd * esz, dtype=fdt).pin_memory() for _ in range(planes)]
        for t in h_in:
            t.uniform_(-1, 1)
        for t in h_out:
            if t is not h_in[0]:
                t.zero_()
        e_steps = max(1, min(steps, 4))

        def e2e_step():
            a = [t.data_ptr() for t in h_in] + [None] * (2 - planes)
            b = [t.data_ptr() for t in h_out] + [None] * (2 - planes)
            plan.compute_host(pf.direction.FORWARD, a[0], a[1], b[0], b[1])

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        barrier()
        te = (time.perf_counter() - t0) / e_steps
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        bpe = (16 if cfg["scalar"] == "double" else 8)
        del fwd_buf, bwd_buf
        torch.cuda.empty_cache()
        tc = copy_ceiling(torch, dist, dev, world, h_in, h_out)
        res["e2e"] = {"seconds": te, "h2d_bytes_per_step": n_fwd * (bpe // 2 if real else bpe), "d2h_bytes_per_step": n_bwd * bpe,
                      "steps": e_steps, "copy_only_seconds": tc}
        del h_in, h_out
    plan.destroy()
    torch.cuda.empty_cache()
    return res


def run_slab_check(args, rank, local_rank, world, dev):
    """C5 (3-D fp32 512^3) slab-decomposed over the N ranks, both exchanges, inside the default multi-GPU run: timing
    (CUDA events, max over ranks), bytes over NVLink against the measured 770 GB/s, and a real parity check --
    rank 0 draws the input (SFC64(0), uniform(-1,1), as the reference's generator), scatters the x-slabs, gathers the
    y-slabs of the spectrum and compares them with numpy.fft.fftn in complex128 (128^3 and 512^3)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from portfft_b200.distributed import slab_fft3d

    out = {}
    for exchange in ("peer", "nccl"):
        entry = {}
        try:
            for n in (128, 512):
                plan = slab_fft3d([n, n, n], "float", exchange=exchange, device=dev)
                g = plan.geom
                x = torch.empty(g.xl, n, n, dtype=torch.complex64, device=dev)
                ref = None
                if rank == 0:
                    rng = np.random.Generator(np.random.SFC64(0))
                    re = rng.uniform(-1, 1, (n, n, n)).astype(np.float32)
                    full = (re + 1j * rng.uniform(-1, 1, (n, n, n)).astype(np.float32)).astype(np.complex64)
                    del re
                    parts = [torch.view_as_real(torch.from_numpy(full[r * g.xl:(r + 1) * g.xl]).to(dev))
                             for r in range(world)]  # (NCCL moves real views: it has no complex type)
                    dist.scatter(torch.view_as_real(x), parts, src=0)
                    del parts
                    ref = np.fft.fftn(full.astype(np.complex128))
                    del full
                else:
                    dist.scatter(torch.view_as_real(x), None, src=0)
                y = torch.view_as_real(plan.forward(x).contiguous())
                torch.cuda.synchronize(dev)
                got = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
                dist.gather(y, got, dst=0)
                if rank == 0:
                    got = [torch.view_as_complex(t) for t in got]
                if rank == 0:
                    num = den = 0.0
                    for r in range(world):
                        want = ref[:, r * g.yb:(r + 1) * g.yb, :]
                        diff = got[r].cpu().numpy().astype(np.complex128) - want
                        num += float(np.vdot(diff, diff).real)
                        den += float(np.vdot(want, want).real)
                    entry[f"rel_l2_{n}"] = math.sqrt(num / den)
                    entry[f"rel_l2_bound_{n}"] = 1e-5 * math.log2(n ** 3)
                    del ref, got
                if n == 512:
                    for _ in range(3):
                        plan.forward(x)
                    torch.cuda.synchronize(dev)
                    dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 20
                    e0.record()
                    for _ in range(reps):
                        plan.forward(x)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    tm = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    ms = float(tm.item())
                    nv = (world - 1) * g.block_elems * 8
                    entry.update({"ms": ms, "gflops": 5.0 * n ** 3 * math.log2(n ** 3) / (ms * 1e-3) / 1e9,
                                  "nvlink_bytes_sent_per_gpu": nv, "nvlink_gbs": nv / (ms * 1e-3) / 1e9,
                                  "nvlink_frac_of_770": nv / (ms * 1e-3) / 1e9 / 770.0})
                plan.destroy()
                del x, y
                torch.cuda.empty_cache()
        except Exception as exc:  # reported, never silently dropped
            entry["error"] = repr(exc)[:300]
        out[exchange] = entry
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the timed steps from one CUDA graph (launch-bound configs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="slab exchange (C5 at --gpus > 1)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="--gpus > 1: strong = the configured batch sharded over the GPUs (BASELINE.json config 2), "
                         "weak = the configured batch on every GPU")
    ap.add_argument("--no-extras", action="store_true", help="--gpus > 1: skip the weak-scaling and C5 slab extras")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg, args.config)
        return
    args.warmup = max(args.warmup, 3)
    if args.graph and args.steps % 2:
        args.steps += 1  # a replay must leave the in-place data where it started

    import torch
    import torch.distributed as dist

    import portfft_b200 as pf  # noqa: F401  (fails loudly when the CUDA library is missing)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.config == "C5" and world > 1:
        run_slab(args, cfg, args.config, rank, local_rank, world, dev)
        dist.destroy_process_group()
        return

    # ---- the measurement: the configured batch, sharded over the ranks (strong) or replicated (weak) -------------
    strong = args.scaling == "strong" or world == 1
    if strong and cfg["batch"] % world:
        raise SystemExit(f"batch {cfg['batch']} does not divide over {world} GPUs")
    batch = cfg["batch"] // world if strong else cfg["batch"]
    total_batch = batch * world
    r = run_batched(args, cfg, args.config, batch, rank, local_rank, world, dev, args.steps, args.warmup,
                    not args.no_e2e, args.graph)
    ms_per_step = r["ms_per_step"]
    per_gpu_flops = flops_of(cfg) * batch / cfg["batch"]
    per_gpu_bytes = bytes_of(cfg) * batch / cfg["batch"]
    value = per_gpu_flops * world / (ms_per_step * 1e-3) / 1e9
    hbm_gbs = per_gpu_bytes * world / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (one launch per step for single-pass plans) ---------------------------
    peak, peak_src = measured_peak()
    n_pass, l2_chunk = r["n_pass"], r["l2_chunk"]
    # Multi-pass plans that run L2 resident touch HBM once per element and direction whatever their pass count: their
    # roofline is that of the whole step.
    hbm_passes = 1 if l2_chunk else n_pass
    per_launch_ms = ms_per_step / hbm_passes
    achieved = per_gpu_bytes / (per_launch_ms * 1e-3) / 1e9  # algorithmic bytes of ONE pass over the data / launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config)
        except Exception:
            traffic = None
    if traffic is not None and batch != cfg["batch"]:
        traffic = traffic * batch / cfg["batch"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of "
                                  "this kernel on this workload (profiles/traffic.json), scaled to the per-GPU batch; "
                                  "not re-measured in this run",
                "peak_source": peak_src, "kernel_launches_per_step": r["launches"] // r["steps"],
                "passes_per_transform": n_pass, "l2_chunk_transforms": l2_chunk,
                "note": ("achieved = 1 read + 1 write of every element per step / step time (CUDA events): the plan runs "
                         "L2 resident, its intermediate passes do not reach HBM" if l2_chunk else
                         "achieved = 1 read + 1 write of every element per launch / mean launch time (CUDA events)")}

    e2e = None
    if r["e2e"] is not None:
        e = r["e2e"]
        moved = e["h2d_bytes_per_step"] + e["d2h_bytes_per_step"]
        e2e = {"value": per_gpu_flops * world / e["seconds"] / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": e["h2d_bytes_per_step"], "d2h_bytes_per_step": e["d2h_bytes_per_step"],
               "steps": e["steps"], "ms_per_step": e["seconds"] * 1e3,
               "api": "pfft_compute_host (C ABI, pinned host buffers)",
               "roofline": {"bound": "host_link", "unit": "GB/s per GPU (H2D + D2H bytes of one step / time)",
                            "achieved": moved / e["seconds"] / 1e9, "peak": moved / e["copy_only_seconds"] / 1e9,
                            "frac": e["copy_only_seconds"] / e["seconds"],
                            "peak_source": "measured in this run: the same bytes as two concurrent pinned "
                                           "cudaMemcpyAsync streams (H2D || D2H), all ranks at once, no kernels",
                            "copy_only_ms": e["copy_only_seconds"] * 1e3}}

    # ---- extras of the multi-GPU run: weak-scaling figure, C5 slab ------------------------------------------------
    weak = slab = None
    if world > 1 and strong and not args.no_extras:
        w = run_batched(args, cfg, args.config, cfg["batch"], rank, local_rank, world, dev, max(4, min(args.steps, 20)),
                        args.warmup, False, False)
        weak = {"value": flops_of(cfg) * world / (w["ms_per_step"] * 1e-3) / 1e9, "unit": "GFLOP/s",
                "ms_per_step": w["ms_per_step"], "per_gpu_batch": cfg["batch"], "steps": w["steps"],
                "roundtrip_rel_l2": w["roundtrip"]}
        if args.config == "C2":
            slab = run_slab_check(args, rank, local_rank, world, dev)

    # ---- CPU baselines (rank 0, N=1 only) ---------------------------------------------------------------------
    cpu = cpu2 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and cfg.get("real"):
        # (the C port restates the reference's complex algorithm: the reference has no real transform to port)
        sample = cpu_sample_batch(cfg)
        sc = scipy_cpu_time(cfg, sample)
        if sc is not None:
            per = flops_of(cfg) / cfg["batch"]
            cpu = {"value": per * sample / sc[0] / 1e9, "unit": "GFLOP/s", "cores": sc[1], "kind": "pocketfft",
                   "sample": f"{sample} of {cfg['batch']} transforms, scipy.fft.rfft(workers={sc[1]}), {sc[0]:.2f} s; the "
                             f"reference implements no REAL transform, so there is no port of it"}
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = cpu_sample_batch(cfg)
        t, cores = cpu_port_time(cfg, sample)
        per = flops_of(cfg) / cfg["batch"]
        cpu = {"value": per * sample / t / 1e9, "unit": "GFLOP/s", "cores": cores,
               "kind": "port", "sample": f"{sample} of {cfg['batch']} transforms, C oracle port of the reference "
                                          f"algorithm, {t:.2f} s"}
        sample2 = sample
        sc = scipy_cpu_time(cfg, sample2)
        if sc is not None:
            cpu2 = {"value": per * sample2 / sc[0] / 1e9, "unit": "GFLOP/s", "cores": sc[1], "kind": "pocketfft",
                    "sample": f"{sample2} of {cfg['batch']} transforms, scipy.fft.fftn(workers={sc[1]}) -- the library "
                              f"behind the numpy oracle of the reference's tests, {sc[0]:.2f} s"}

    if rank == 0:
        line = {
            "metric": "batched_c2c_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64" if cfg["scalar"] == "double" else "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['desc']}", "total_batch": total_batch, "per_gpu_batch": batch,
                       "direction": "alternating compute_forward / compute_backward (backward_scale 1/N)",
                       "l2": "per-GPU working set %.0f MiB > 126 MB L2, rewritten by every step (no flush needed)"
                             % (per_gpu_bytes / 2 / 2**20)
                       if per_gpu_bytes / 2 > 1.5 * 126e6 else "working set fits L2: launch-latency bound",
                       "sharding": "batch-sharded: %d transforms per GPU, one plan per GPU, no collective" % batch},
            "hbm_gbs": hbm_gbs, "hbm_frac_of_measured": hbm_gbs / world / peak,
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_pocketfft": cpu2, "e2e": e2e,
            "gpu_launches": r["launches"], "clocks": r["clocks"], "roundtrip_rel_l2": r["roundtrip"],
            "cuda_graph": bool(args.graph),
        }
        if weak is not None:
            line["weak_scaling"] = weak
        if slab is not None:
            line["slab_c5"] = slab
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
