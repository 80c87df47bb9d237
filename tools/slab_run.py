#!/usr/bin/env python
"""Multi-GPU check + timing of the slab-decomposed 3-D transform (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_run.py --lengths 512 512 512 --exchange peer --steps 20

Every rank builds the same seeded input, transforms its x-slab, compares its y-slab of the spectrum with
numpy.fft.fftn (sizes up to 128^3; larger sizes are checked through Parseval + a forward DC/Nyquist sample) and
times `steps` calls with CUDA events (max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lengths", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"])
    ap.add_argument("--scalar", default="float")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from portfft_b200.distributed import slab_fft3d

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n0, n1, n2 = args.lengths
    plan = slab_fft3d(args.lengths, args.scalar, exchange=args.exchange, device=dev,
                      backward_scale=1.0 / (n0 * n1 * n2))
    g = plan.geom
    cdt = torch.complex128 if args.scalar == "double" else torch.complex64
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    x = torch.view_as_complex(torch.rand(g.xl, n1, n2, 2, generator=gen, device=dev,
                                         dtype=torch.float64 if args.scalar == "double" else torch.float32) * 2 - 1)
    out = plan.forward(x)
    torch.cuda.synchronize()
    # parity: gather the input on rank 0 for small sizes, compare against numpy fftn
    check = {}
    if n0 * n1 * n2 <= 128 ** 3:
        xs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(xs, x)
        full = torch.cat(xs).cpu().numpy()
        ref = np.fft.fftn(full.astype(np.complex128))[:, rank * g.yb:(rank + 1) * g.yb, :]
        got = out.cpu().numpy()
        check["rel_l2"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    # Parseval (size independent): sum |X|^2 == N * sum |x|^2 over all ranks
    e_in = (x.abs().double() ** 2).sum()
    e_out = (out.abs().double() ** 2).sum()
    t = torch.stack([e_in, e_out])
    dist.all_reduce(t)
    check["parseval_rel"] = float(abs(t[1] / (t[0] * n0 * n1 * n2) - 1.0))
    # full-size round trip across the GPUs: backward(forward(x)) == x
    back = plan.backward(plan.forward(x))
    torch.cuda.synchronize()
    num = ((back - x).abs().double() ** 2).sum()
    den = (x.abs().double() ** 2).sum()
    t2 = torch.stack([num, den])
    dist.all_reduce(t2)
    check["roundtrip_rel_l2"] = float((t2[0] / t2[1]).sqrt())
    for _ in range(args.warmup):
        plan.forward(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.forward(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = n0 * n1 * n2
        esize = 16 if args.scalar == "double" else 8
        print(json.dumps({"slab_fft3d": args.lengths, "world": world, "exchange": args.exchange,
                          "ms_per_transform": float(ms), "gflops": 5 * n * np.log2(n) / float(ms) / 1e6,
                          "nvlink_bytes_sent_per_gpu": (world - 1) * g.block_elems * esize, **check}), flush=True)
    plan.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
