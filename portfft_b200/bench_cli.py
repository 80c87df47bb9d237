"""The reference's benchmark command line re-expressed over the C ABI.

* `parse_manual_args("d=cpx,n=64,b=1024")`: the `key=value` syntax of `bench_manual_float` / `bench_manual_double`
  (/root/reference/test/bench/portfft/register_manual_bench.hpp:36-39 keys, :77-112 tokenizer, :114-152 value parsers,
  :160-211 descriptor filling): same long / short keys, `x`-separated vectors, same rejections.
* `benchmark_names(...)`: the GBench names `average_host_time/d=cpx,prec=single,n=[64],batch=1024/<suffix>` and
  `device_time/...` of `register_host_device_benchmark` (launch_bench.hpp:260-290).
* `CANNED_FLOAT`: the four configurations of `bench_float` (bench_float.cpp:49-52).
* `run_host_device_benchmark(...)` (needs a GPU): the two timing methods of the reference --
  average_host_time = 10 chained asynchronous computes over up to 10 distinct inputs, host clock, inputs rewritten
  before every iteration (launch_bench.hpp:49-145, bench_utils.hpp:39-64); device_time = one compute per iteration
  timed on the device (launch_bench.hpp:171-234, CUDA events instead of SYCL event profiling) -- with the same
  `flops` = 5*N*log2(N)*batch/t and `throughput` = batch*(N*sizeof(in)+N*sizeof(out))/t counters
  (utils/ops_estimate.hpp:34-50).
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Tuple

ARG_KEYS = [("domain", "d"), ("lengths", "n"), ("batch", "b"), ("fwd_strides", "fs"), ("bwd_strides", "bs"),
            ("fwd_dist", "fd"), ("bwd_dist", "bd"), ("scale", "sx"), ("storage", "s"), ("placement", "p")]
RUNS_TO_AVERAGE = 10  # bench_utils.hpp:39

CANNED_FLOAT = [("small_1d", [16], 8 * 1024 * 1024), ("medium_small_1d", [256], 512 * 1024),
                ("medium_large_1d", [4096], 32 * 1024), ("large_1d", [65536], 2048)]
# register_complex_float_benchmark_set / register_real_float_benchmark_set (utils/reference_dft_set.hpp:87-112): the
# sets the reference registers for the closed-source comparison benches.  "large_1d_prime" (65537) cannot run on the
# reference (a prime factor beyond one work-item, committed_descriptor_impl.hpp:241); it runs here (Bluestein).
CANNED_COMPLEX_SET = CANNED_FLOAT + [("large_1d_prime", [65537], 2048)]
CANNED_REAL_SET = [("small_1d", [32], 8 * 1024 * 1024), ("medium_small_1d", [512], 512 * 1024),
                   ("medium_large_1d", [8192], 32 * 1024), ("large_1d", [128 * 1024], 2048)]


class bench_error(RuntimeError):
    pass


class invalid_value(bench_error):
    def __init__(self, key, value):
        super().__init__(f"Invalid '{key}' value: '{value}'")


def get_arg_map(arg: str) -> Dict[str, str]:
    """register_manual_bench.hpp:77-112"""
    out: Dict[str, str] = {}
    if arg == "":
        return out
    valid = {k for pair in ARG_KEYS for k in pair}
    for token in arg.split(","):
        if token == "":
            break
        if "=" not in token:
            raise bench_error(f"Invalid token '{token}'")
        key, value = token.split("=", 1)
        if key in out:
            raise bench_error(f"Key can only be specified once: '{key}'")
        if key not in valid:
            raise bench_error(f"Invalid key: '{key}'")
        if value == "":
            raise invalid_value(key, value)
        out[key] = value
    return out


def _get(arg_map, long_name):
    short = dict(ARG_KEYS)[long_name]
    return arg_map.get(long_name, arg_map.get(short, ""))


def _unsigned(key, value) -> int:
    try:
        v = int(value)
        if v <= 0:
            raise ValueError
        return v
    except ValueError:
        raise bench_error(f"Invalid '{key}' value: '{value}' must be a positive integer")


def _vec(key, value) -> List[int]:
    return [_unsigned(key, t) for t in value.split("x")] if value else []


def parse_manual_args(desc_str: str, scalar: str = "float"):
    """-> portfft_b200.descriptor filled as `register_manual_benchmark` + `fill_descriptor` do."""
    import portfft_b200 as pf

    m = get_arg_map(desc_str)
    dom = _get(m, "domain")
    if dom in ("complex", "cpx"):
        domain = pf.domain.COMPLEX
    elif dom in ("real", "re"):
        domain = pf.domain.REAL
    elif dom == "":
        raise bench_error("'domain' must be specified")
    else:
        raise invalid_value("domain", dom)
    lengths = _vec("lengths", _get(m, "lengths"))
    if not lengths:
        raise bench_error("'lengths' must be specified")
    d = pf.descriptor(lengths, scalar, domain)
    if _get(m, "batch"):
        d.number_of_transforms = _unsigned("batch", _get(m, "batch"))
    if _get(m, "fwd_strides"):
        d.forward_strides = _vec("fwd_strides", _get(m, "fwd_strides"))
    if _get(m, "bwd_strides"):
        d.backward_strides = _vec("bwd_strides", _get(m, "bwd_strides"))
    if _get(m, "fwd_dist"):
        d.forward_distance = _unsigned("fwd_dist", _get(m, "fwd_dist"))
    if _get(m, "bwd_dist"):
        d.backward_distance = _unsigned("bwd_dist", _get(m, "bwd_dist"))
    if _get(m, "scale"):
        d.forward_scale = d.backward_scale = float(_get(m, "scale"))
    st = _get(m, "storage")
    if st in ("complex", "cpx", "interleaved", "int"):
        d.complex_storage = pf.complex_storage.INTERLEAVED_COMPLEX
    elif st in ("real_real", "rr", "split", "sp"):
        d.complex_storage = pf.complex_storage.SPLIT_COMPLEX
    elif st:
        raise invalid_value("storage", st)
    pl = _get(m, "placement")
    if pl in ("in_place", "ip"):
        d.placement = pf.placement.IN_PLACE
    elif pl in ("out_of_place", "oop"):
        d.placement = pf.placement.OUT_OF_PLACE
    elif pl:
        raise invalid_value("placement", pl)
    return d


def benchmark_names(desc, suffix: str) -> Tuple[str, str]:
    """launch_bench.hpp:267-289"""
    core = "d=%s,prec=%s,n=[%s],batch=%d" % ("re" if int(desc.domain) == 0 else "cpx",
                                              "single" if desc.scalar == "float" else "double",
                                              ", ".join(str(x) for x in desc.lengths), desc.number_of_transforms)
    return f"average_host_time/{core}/{suffix}", f"device_time/{core}/{suffix}"


def ops_estimate(n: int, batch: int) -> float:
    """cooley_tukey_ops_estimate, utils/ops_estimate.hpp:34-36"""
    return 5.0 * batch * n * math.log2(n)


def mem_transactions(n: int, batch: int, scalar: str, real: bool = False) -> float:
    """global_mem_transactions<forward_t, complex_type>(batch, N, N), utils/ops_estimate.hpp:47-50 (one read of the
    forward type + one write of complex<scalar>; launch_bench.hpp:138-141)"""
    c = 16 if scalar == "double" else 8
    return float(batch) * n * ((c // 2 if real else c) + c)


def run_host_device_benchmark(desc, suffix: str, iterations: int = 10, device: int = 0) -> List[dict]:
    """Both reference timing methods for one descriptor (interleaved storage, forward direction, as the reference's
    benches).  Returns two result dicts (GBench-like: name, real_time ms, flops, throughput)."""
    import torch

    import portfft_b200 as pf

    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream(dev)
    n = desc.get_flattened_length()
    batch = desc.number_of_transforms
    cdt = torch.complex128 if desc.scalar == "double" else torch.complex64
    rdt = torch.float64 if desc.scalar == "double" else torch.float32
    real = int(desc.domain) == int(pf.domain.REAL)
    n_in, n_out = desc.get_input_count(pf.direction.FORWARD), desc.get_output_count(pf.direction.FORWARD)
    in_place = desc.placement == pf.placement.IN_PLACE
    if real and in_place:
        raise bench_error("REAL-domain benchmarks run out of place")
    esz = 16 if desc.scalar == "double" else 8
    isz = esz // 2 if real else esz
    total = torch.cuda.get_device_properties(dev).total_memory
    num_inputs = RUNS_TO_AVERAGE if n_in * isz * RUNS_TO_AVERAGE + (0 if in_place else n_out * esz) <= 0.9 * total else 1
    inputs = [torch.zeros(n_in, dtype=rdt if real else cdt, device=dev) for _ in range(num_inputs)]
    out = None if in_place else torch.zeros(n_out, dtype=cdt, device=dev)
    if real:
        host = (torch.rand(n_in, dtype=rdt) * 2 - 1).pin_memory()
    else:
        host = torch.view_as_complex(torch.rand(n_in, 2, dtype=rdt) * 2 - 1).pin_memory()
    plan = desc.commit(stream, device)

    def compute(buf):
        if in_place:
            plan.compute_forward(buf, queue=stream)
        else:
            plan.compute_forward(buf, out, queue=stream)

    compute(inputs[0])
    torch.cuda.synchronize(dev)
    ops, byts = ops_estimate(n, batch), mem_transactions(n, batch, desc.scalar, real)
    host_name, dev_name = benchmark_names(desc, suffix)
    # ---- average_host_time ----------------------------------------------------------------------------------------
    t_host = []
    for _ in range(iterations):
        for buf in inputs:
            buf.copy_(host, non_blocking=True)  # rewrite the inputs: defeats the cache, keeps in-place data bounded
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for r in range(RUNS_TO_AVERAGE):
            compute(inputs[r % num_inputs])
        torch.cuda.synchronize(dev)
        t_host.append((time.perf_counter() - t0) / RUNS_TO_AVERAGE)
    # ---- device_time ------------------------------------------------------------------------------------------------
    t_dev = []
    for _ in range(iterations):
        inputs[0].copy_(host, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        compute(inputs[0])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t_dev.append(e0.elapsed_time(e1) * 1e-3)
    plan.destroy()

    def result(name, ts):
        mean = sum(ts) / len(ts)
        return {"name": name, "iterations": len(ts), "real_time_ms": mean * 1e3,
                "flops": sum(ops / t for t in ts) / len(ts), "throughput": sum(byts / t for t in ts) / len(ts)}

    return [result(host_name, t_host), result(dev_name, t_dev)]
