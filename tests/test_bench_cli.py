"""The reference's benchmark CLI syntax (register_manual_bench.hpp) and names (launch_bench.hpp:267-289)."""
import pytest

import portfft_b200 as pf
from portfft_b200 import bench_cli


def test_manual_args_fill_the_descriptor_like_the_reference():
    d = bench_cli.parse_manual_args("d=cpx,n=64,b=1024")  # BASELINE config C1 verbatim
    assert d.lengths == [64] and d.number_of_transforms == 1024 and d.domain == pf.domain.COMPLEX
    assert d.placement == pf.placement.OUT_OF_PLACE and d.complex_storage == pf.complex_storage.INTERLEAVED_COMPLEX
    d = bench_cli.parse_manual_args("domain=complex,lengths=16x512,batch=3,fs=1024x2,bs=512x1,fd=9000,bd=8192,"
                                    "sx=0.5,storage=sp,placement=oop", "double")
    assert d.lengths == [16, 512] and d.forward_strides == [1024, 2] and d.backward_strides == [512, 1]
    assert d.forward_distance == 9000 and d.backward_distance == 8192
    assert d.forward_scale == 0.5 and d.backward_scale == 0.5 and d.scalar == "double"
    assert d.complex_storage == pf.complex_storage.SPLIT_COMPLEX
    assert bench_cli.parse_manual_args("d=cpx,n=8,p=ip").placement == pf.placement.IN_PLACE


@pytest.mark.parametrize("arg,msg", [
    ("n=64", "'domain' must be specified"), ("d=cpx", "'lengths' must be specified"),
    ("d=foo,n=4", "Invalid 'domain' value: 'foo'"), ("d=cpx,n=4,n=5", "Key can only be specified once: 'n'"),
    ("d=cpx,n=4,zz=1", "Invalid key: 'zz'"), ("d=cpx,n=0", "must be a positive integer"),
    ("d=cpx,n=4,b=", "Invalid 'b' value: ''"), ("d=cpx,n=4,s=weird", "Invalid 'storage' value: 'weird'"),
    ("d=cpx,n=4,p=weird", "Invalid 'placement' value: 'weird'"), ("d=cpx,n4", "Invalid token 'n4'"),
])
def test_manual_args_rejections(arg, msg):
    with pytest.raises(bench_cli.bench_error, match=msg.replace("'", ".")):
        bench_cli.parse_manual_args(arg)


def test_benchmark_names_and_counters():
    d = bench_cli.parse_manual_args("d=cpx,n=16x512,b=3")
    host, dev = bench_cli.benchmark_names(d, "f:d=cpx,n=16x512,b=3")
    assert host == "average_host_time/d=cpx,prec=single,n=[16, 512],batch=3/f:d=cpx,n=16x512,b=3"
    assert dev.startswith("device_time/d=cpx,prec=single,n=[16, 512],batch=3/")
    assert bench_cli.ops_estimate(4096, 65536) == 5.0 * 65536 * 4096 * 12
    assert bench_cli.mem_transactions(4096, 65536, "float") == 4294967296.0
    assert [c[1:] for c in bench_cli.CANNED_FLOAT] == [([16], 8388608), ([256], 524288), ([4096], 32768), ([65536], 2048)]


@pytest.mark.gpu
def test_manual_benchmark_runs():
    d = bench_cli.parse_manual_args("d=cpx,n=64,b=1024")
    res = bench_cli.run_host_device_benchmark(d, "f:d=cpx,n=64,b=1024", iterations=3)
    assert len(res) == 2 and all(r["flops"] > 0 and r["throughput"] > 0 for r in res)
