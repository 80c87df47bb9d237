"""Multi-GPU layer of the C ABI (include/pfft.h "Multi-GPU", csrc/multi.cu).

CPU part: the host-only pieces (pfft_partition against the Python restatement, argument checks that must fire before
any CUDA call).  GPU part: the same entry points with several ranks on ONE device ("virtual ranks": `devices` may
repeat) -- the slab transform runs exactly the code path of a multi-GPU run (peer-store exchange into the other ranks'
windows, device-side flag barrier between streams), checked against numpy.fft.fftn, the oracle of the reference's own
tests (test/common/reference_data_wrangler.hpp:117-145)."""
import ctypes

import numpy as np
import pytest

import portfft_b200 as pf
from portfft_b200 import _lib
from portfft_b200.distributed import _tensor_view, partition, slab_plan


def test_partition_matches_python():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(300):
        n, w = int(rng.integers(0, 100000)), int(rng.integers(1, 17))
        covered = 0
        for r in range(w):
            first, count = ctypes.c_size_t(), ctypes.c_size_t()
            assert lib.pfft_partition(n, w, r, ctypes.byref(first), ctypes.byref(count)) == 0
            assert (first.value, count.value) == partition(n, w, r)
            assert first.value == covered
            covered += count.value
        assert covered == n
    first, count = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.pfft_partition(10, 0, 0, ctypes.byref(first), ctypes.byref(count)) == 1
    assert lib.pfft_partition(10, 4, 4, ctypes.byref(first), ctypes.byref(count)) == 1


@pytest.mark.parametrize("lengths,world,rank,exc", [
    ((512, 512, 512), 3, 0, pf.invalid_configuration),     # lengths not divisible by the world size
    ((512, 512), 2, 0, pf.unsupported_configuration),       # not 3-D
    ((8, 8, 8), 2, 2, pf.invalid_configuration),            # rank out of range
    ((8, 8, 8), 17, 0, pf.invalid_configuration),           # more ranks than peer-table entries
    ((8, 0, 8), 2, 0, pf.invalid_configuration),            # invalid descriptor (zero length)
])
def test_slab_commit_rejects_before_touching_the_gpu(lengths, world, rank, exc):
    d = pf.descriptor(lengths)
    with pytest.raises(exc):
        slab_plan.commit(d, world, rank, "cuda:0")


def test_slab_commit_rejects_layouts():
    d = pf.descriptor((8, 8, 8))
    d.complex_storage = pf.complex_storage.SPLIT_COMPLEX
    with pytest.raises(pf.unsupported_configuration):
        slab_plan.commit(d, 2, 0, "cuda:0")
    d = pf.descriptor((8, 8, 8))
    d.number_of_transforms = 2
    d.forward_distance = d.backward_distance = 512
    with pytest.raises(pf.unsupported_configuration):
        slab_plan.commit(d, 2, 0, "cuda:0")


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
def _random_complex(shape, scalar, seed=0):
    rng = np.random.Generator(np.random.SFC64(seed))
    x = rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)
    return x.astype(np.complex128 if scalar == "double" else np.complex64)


@pytest.mark.gpu
@pytest.mark.parametrize("scalar", ["float", "double"])
@pytest.mark.parametrize("lengths,world", [((4, 4, 8), 1), ((8, 12, 5), 4), ((16, 8, 64), 8), ((64, 64, 512), 8),
                                            ((32, 16, 1000), 2), ((8, 8, 4096), 2), ((128, 128, 128), 4)])
def test_slab_c_layer_virtual_ranks(lengths, world, scalar):
    import torch

    dev = torch.device("cuda", 0)
    n0, n1, n2 = lengths
    total = n0 * n1 * n2
    x = _random_complex(lengths, scalar)
    d = pf.descriptor(lengths, scalar)
    d.backward_scale = 1.0 / total
    plans = slab_plan.commit_local(d, [0] * world)
    xl, yb = n0 // world, n1 // world
    slabs = [torch.from_numpy(np.ascontiguousarray(x[r * xl:(r + 1) * xl])).to(dev) for r in range(world)]
    torch.cuda.synchronize(dev)
    ref = np.fft.fftn(x.astype(np.complex128))
    bound = (1e-13 if scalar == "double" else 1e-5) * np.log2(total)
    for rep in range(3):  # repeated calls: the flag epochs and the window reuse barrier
        outs = [p.forward(s) for p, s in zip(plans, slabs)]
        for p in plans:
            p.sync()
        for r in range(world):
            want = ref[:, r * yb:(r + 1) * yb, :]
            rel = np.linalg.norm(outs[r].cpu().numpy() - want) / np.linalg.norm(want)
            assert rel <= bound, (rep, r, rel, bound)
    # backward from the y-slabs left in the windows: identity (backward_scale = 1 / N)
    backs = [p.backward(o) for p, o in zip(plans, outs)]
    for p in plans:
        p.sync()
    for r in range(world):
        want = x[r * xl:(r + 1) * xl]
        rel = np.linalg.norm(backs[r].cpu().numpy() - want) / np.linalg.norm(want)
        assert rel <= 2 * bound, (r, rel, bound)
    # backward from caller-owned y-slabs (copied into the window first)
    ys = [torch.from_numpy(np.ascontiguousarray(ref[:, r * yb:(r + 1) * yb, :]).astype(x.dtype)).to(dev)
          for r in range(world)]
    torch.cuda.synchronize(dev)
    backs = [p.backward(y) for p, y in zip(plans, ys)]
    for p in plans:
        p.sync()
    for r in range(world):
        want = x[r * xl:(r + 1) * xl]
        rel = np.linalg.norm(backs[r].cpu().numpy() - want) / np.linalg.norm(want)
        assert rel <= 2 * bound, (r, rel, bound)
    for p in plans:
        p.destroy()


@pytest.mark.gpu
def test_slab_c_layer_alltoall_callback_world1():
    """The caller-collective path (pfft_slab_set_alltoall) with one rank: the 'collective' is a device copy."""
    import torch

    dev = torch.device("cuda", 0)
    lengths = (16, 8, 32)
    x = _random_complex(lengths, "float")
    d = pf.descriptor(lengths)
    d.backward_scale = 1.0 / x.size
    plan = slab_plan.commit(d, 1, 0, dev)
    calls = []

    def a2a(send, recv, block_bytes, stream):
        calls.append(block_bytes)
        src = _tensor_view(send, (block_bytes,), torch.uint8, dev)
        dst = _tensor_view(recv, (block_bytes,), torch.uint8, dev)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
            dst.copy_(src)

    plan.set_alltoall(a2a)
    xs = torch.from_numpy(x).to(dev)
    torch.cuda.synchronize(dev)
    out = plan.forward(xs)
    plan.sync()
    want = np.fft.fftn(x.astype(np.complex128))
    assert np.linalg.norm(out.cpu().numpy() - want) / np.linalg.norm(want) < 1e-5 * np.log2(x.size)
    back = plan.backward(out)
    plan.sync()
    assert np.linalg.norm(back.cpu().numpy() - x) / np.linalg.norm(x) < 2e-5 * np.log2(x.size)
    assert calls == [x.size * 8, x.size * 8]
    plan.destroy()


@pytest.mark.gpu
def test_slab_barrier_timeout_is_reported():
    """A peer that never arrives: the barrier gives up after PFFT_SLAB_TIMEOUT_MS and pfft_slab_sync reports it
    instead of hanging the GPU."""
    import os

    import torch

    os.environ["PFFT_SLAB_TIMEOUT_MS"] = "200"
    try:
        d = pf.descriptor((8, 8, 16))
        plans = slab_plan.commit_local(d, [0, 0])
    finally:
        del os.environ["PFFT_SLAB_TIMEOUT_MS"]
    x = torch.zeros(4, 8, 16, dtype=torch.complex64, device="cuda")
    torch.cuda.synchronize()
    plans[0].forward(x)  # rank 1 never calls
    with pytest.raises(pf.cuda_error):
        plans[0].sync()
    for p in plans:
        p.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3, 8])
@pytest.mark.parametrize("kind", ["packed", "batch_interleaved", "strided"])
def test_commit_multi_batch_sharded(world, kind):
    """pfft_commit_multi / pfft_multi_compute with all shards on device 0, against the un-sharded oracle result."""
    import torch

    import portfft_oracle as oracle
    from portfft_b200.distributed import shard_descriptor

    lib = _lib.load()
    n, batch = 256, 37
    d = pf.descriptor([n])
    d.number_of_transforms = batch
    if kind == "batch_interleaved":
        d.forward_strides = d.backward_strides = [batch]
        d.forward_distance = d.backward_distance = 1
    elif kind == "strided":
        d.forward_strides, d.forward_distance, d.forward_offset = [2], 2 * n + 3, 5
        d.backward_strides, d.backward_distance, d.backward_offset = [1], n + 1, 2
    od = oracle.OracleDescriptor(lengths=[n], number_of_transforms=batch, forward_strides=list(d.forward_strides),
                                 backward_strides=list(d.backward_strides), forward_distance=d.forward_distance,
                                 backward_distance=d.backward_distance, forward_offset=d.forward_offset,
                                 backward_offset=d.backward_offset)
    host_in, host_ref = oracle.expected_io(od, oracle.FORWARD)
    c, keep = d._c_desc()
    devs = (ctypes.c_int * world)(*([0] * world))
    multi = ctypes.c_void_p()
    assert lib.pfft_commit_multi(ctypes.byref(c), world, devs, None, ctypes.byref(multi)) == 0, lib.pfft_last_error()
    assert lib.pfft_multi_size(multi) == world
    dev = torch.device("cuda", 0)
    j = np.arange(n)
    ins, outs, shards = [], [], []
    for r in range(world):
        info = _lib.pfft_shard_info()
        assert lib.pfft_multi_shard(multi, r, ctypes.byref(info), None) == 0
        sh = shard_descriptor(d, world, r)
        assert (info.first, info.count, info.forward_start, info.backward_start) == \
            (sh.first, sh.count, sh.forward_start, sh.backward_start)
        loc = sh.desc
        b = np.arange(sh.first, sh.first + sh.count)
        lb = np.arange(sh.count)
        local_in = np.zeros(loc.get_input_count(pf.direction.FORWARD), dtype=host_in.dtype)
        if sh.count:
            gin = d.forward_offset + b[:, None] * d.forward_distance + j[None, :] * d.forward_strides[0]
            lin = loc.forward_offset + lb[:, None] * loc.forward_distance + j[None, :] * loc.forward_strides[0]
            local_in[lin] = host_in[gin]
        ins.append(torch.from_numpy(local_in).to(dev))
        outs.append(torch.zeros(loc.get_output_count(pf.direction.FORWARD), dtype=ins[-1].dtype, device=dev))
        shards.append(sh)
    torch.cuda.synchronize(dev)
    tab_in = (ctypes.c_void_p * world)(*[t.data_ptr() for t in ins])
    tab_out = (ctypes.c_void_p * world)(*[t.data_ptr() for t in outs])
    assert lib.pfft_multi_compute(multi, 0, tab_in, None, tab_out, None) == 0, lib.pfft_last_error()
    assert lib.pfft_multi_sync(multi) == 0
    out = np.full_like(host_ref, complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE))
    for r, sh in enumerate(shards):
        if sh.count == 0:
            continue
        loc = sh.desc
        b = np.arange(sh.first, sh.first + sh.count)
        lb = np.arange(sh.count)
        lout = loc.backward_offset + lb[:, None] * loc.backward_distance + j[None, :] * loc.backward_strides[0]
        gout = d.backward_offset + b[:, None] * d.backward_distance + j[None, :] * d.backward_strides[0]
        out[gout] = outs[r].cpu().numpy()[lout]
    addressed = np.zeros(host_ref.shape, bool)
    allb = np.arange(batch)
    addressed[d.backward_offset + allb[:, None] * d.backward_distance + j[None, :] * d.backward_strides[0]] = True
    rel = np.linalg.norm(out[addressed] - host_ref[addressed]) / np.linalg.norm(host_ref[addressed])
    assert rel < 1e-5 * np.log2(n), rel
    assert lib.pfft_multi_destroy(multi) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("kind", ["packed", "strided"])
def test_multi_compute_host(world, kind):
    """pfft_multi_compute_host: un-sharded HOST buffers in and out, every GPU (here: every shard on device 0, one host
    thread each) running the chunk-pipelined host path on its batch range; untouched padding included."""
    import portfft_oracle as oracle

    lib = _lib.load()
    n, batch = 512, 101
    d = pf.descriptor([n])
    d.number_of_transforms = batch
    if kind == "strided":
        d.forward_strides, d.forward_distance, d.forward_offset = [1], n + 7, 3
        d.backward_strides, d.backward_distance, d.backward_offset = [1], n + 2, 1
    od = oracle.OracleDescriptor(lengths=[n], number_of_transforms=batch, forward_strides=list(d.forward_strides),
                                 backward_strides=list(d.backward_strides), forward_distance=d.forward_distance,
                                 backward_distance=d.backward_distance, forward_offset=d.forward_offset,
                                 backward_offset=d.backward_offset)
    host_in, host_ref = oracle.expected_io(od, oracle.FORWARD)
    c, keep = d._c_desc()
    devs = (ctypes.c_int * world)(*([0] * world))
    multi = ctypes.c_void_p()
    assert lib.pfft_commit_multi(ctypes.byref(c), world, devs, None, ctypes.byref(multi)) == 0, lib.pfft_last_error()
    out = np.full_like(host_ref, complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE))
    st = lib.pfft_multi_compute_host(multi, 0, host_in.ctypes.data, None, out.ctypes.data, None)
    assert st == 0, lib.pfft_last_error()
    j = np.arange(n)
    addressed = np.zeros(host_ref.shape, bool)
    addressed[d.backward_offset + np.arange(batch)[:, None] * d.backward_distance + j[None, :]] = True
    rel = np.linalg.norm(out[addressed] - host_ref[addressed]) / np.linalg.norm(host_ref[addressed])
    assert rel < 1e-5 * np.log2(n), rel
    assert np.all(out[~addressed] == complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE))
    # batch-interleaved shards are not contiguous in an un-sharded host buffer: refused, not mangled
    db = pf.descriptor([64])
    db.number_of_transforms = 10
    db.forward_strides = db.backward_strides = [10]
    db.forward_distance = db.backward_distance = 1
    cb, keepb = db._c_desc()
    mb = ctypes.c_void_p()
    assert lib.pfft_commit_multi(ctypes.byref(cb), 2, devs, None, ctypes.byref(mb)) == 0
    buf = np.zeros(640, np.complex64)
    assert lib.pfft_multi_compute_host(mb, 0, buf.ctypes.data, None, buf.ctypes.data, None) == 2
    lib.pfft_multi_destroy(mb)
    lib.pfft_multi_destroy(multi)
