"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/pfft.h declares, and its
host-side logic (validation, descriptor arithmetic, layout classification, planner) matches the oracle restatement
of the reference.  No compute call is made (no GPU needed)."""
import ctypes
import os
import random
import re
import time

import pytest

import portfft_b200 as pf
import portfft_oracle as o
from portfft_b200 import _lib
from test_oracle import INVALID

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pfft.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pfft_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = pf.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pfft.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.pfft_version().decode().startswith("pfft_b200")


def _desc(lengths, fs, bs, fd, bd, batch, pl, scalar="float"):
    d = pf.descriptor(lengths, scalar)
    d.number_of_transforms = batch
    d.placement = pf.placement(pl)
    d.forward_strides, d.backward_strides = list(fs), list(bs)
    d.forward_distance, d.backward_distance = fd, bd
    return d


@pytest.mark.parametrize("case", INVALID, ids=[str(i) for i in range(len(INVALID))])
def test_invalid_configurations_throw_invalid_configuration(case):
    # EXPECT_THROW(desc.commit(queue), portfft::invalid_configuration) -- instantiate_fft_tests.hpp:406-411.
    # validation runs on the host before any CUDA call, so commit() raises it without a GPU as well.
    d = _desc(*case)
    with pytest.raises(pf.invalid_configuration):
        d.validate()
    with pytest.raises(pf.invalid_configuration):
        d.commit()


def test_overlap_check_is_sublinear():
    d = _desc([8], [3333333], [3333333], 1, 1, 3333334, 1)
    t0 = time.perf_counter()
    with pytest.raises(pf.invalid_configuration):
        d.validate()
    assert time.perf_counter() - t0 < 0.5


def test_mismatching_strides_length_and_real_domain():
    d = pf.descriptor([8, 4])
    d.forward_strides = [4]
    with pytest.raises(pf.invalid_configuration):
        d.validate()
    # REAL domain: the reference throws unsupported_configuration for every REAL descriptor
    # (descriptor_validation.hpp:268-270); here it is implemented (half spectrum along the last dimension)
    d = pf.descriptor([8], dom=pf.domain.REAL)
    d.validate()
    assert d.get_input_count(pf.direction.FORWARD) == 8 and d.get_input_count(pf.direction.BACKWARD) == 5
    d = pf.descriptor([4, 8], dom=pf.domain.REAL)
    d.validate()
    assert d.get_input_count(pf.direction.FORWARD) == 32 and d.get_input_count(pf.direction.BACKWARD) == 3 * 8 + 5
    d = pf.descriptor([])
    with pytest.raises(pf.invalid_configuration):
        d.validate()


def test_buffer_count_known_answers():
    # test/unit_test/descriptor.cpp:84-108
    d = pf.descriptor([2, 3])
    assert d.get_flattened_length() == 6
    d.number_of_transforms = 2
    d.forward_strides, d.forward_distance, d.forward_offset = [8, 3], 15, 3
    d.backward_strides, d.backward_distance, d.backward_offset = [2, 4], 1, 5
    assert d.get_input_count(pf.direction.FORWARD) == 33 == d.get_output_count(pf.direction.BACKWARD)
    assert d.get_input_count(pf.direction.BACKWARD) == 17 == d.get_output_count(pf.direction.FORWARD)
    d = pf.descriptor([512, 512, 512])
    assert d.forward_strides == [262144, 512, 1] and d.forward_distance == 512 ** 3


def test_validation_and_layout_agree_with_oracle_on_random_descriptors():
    rng = random.Random(7)
    lay = {o.PACKED: pf.layout.PACKED, o.UNPACKED: pf.layout.UNPACKED, o.BATCH_INTERLEAVED: pf.layout.BATCH_INTERLEAVED}
    n_invalid = 0
    for _ in range(3000):
        rank = rng.choice([1, 1, 1, 2, 3])
        lengths = [rng.choice([1, 2, 3, 4, 5, 8, 16]) for _ in range(rank)]
        batch = rng.choice([1, 2, 3, 7, 33])
        pl = rng.choice([0, 1])

        def dom():
            if rng.random() < 0.3:
                return o.get_default_strides(lengths), int(__import__("numpy").prod(lengths))
            return [rng.choice([0, 1, 2, 3, 4, 7, 8, 16, 33, 64]) for _ in range(rank)], rng.choice(
                [0, 1, 2, 3, 5, 8, 16, 40, 64, 300])

        fs, fd = dom()
        bs, bd = (fs, fd) if (pl == 0 and rng.random() < 0.7) else dom()
        od = o.OracleDescriptor(lengths, number_of_transforms=batch, placement=pl, forward_strides=fs,
                                backward_strides=bs, forward_distance=fd, backward_distance=bd)
        d = _desc(lengths, fs, bs, fd, bd, batch, pl)
        try:
            o.validate_descriptor(od)
            ok = True
        except o.InvalidConfiguration:
            ok = False
        if ok:
            d.validate()
            for dr in (0, 1):
                assert d.get_layout(pf.direction(dr)) == lay[o.get_layout(od, dr)]
                assert d.get_input_count(pf.direction(dr)) == od.get_input_count(dr)
        else:
            n_invalid += 1
            with pytest.raises(pf.invalid_configuration):
                d.validate()
    assert 300 < n_invalid < 2900


def _levels(desc_text):
    return desc_text.splitlines()[0].split(";")[0].replace("levels:", "").split()


def test_planner_levels_for_baseline_configs():
    """Dry-run planner (host only).  Level thresholds are this library's own (north_star keeps the hierarchy)."""
    d = pf.descriptor([4096])
    d.number_of_transforms = 65536
    d.placement = pf.placement.IN_PLACE
    txt = d.describe_plan()
    assert _levels(txt) == ["WORKGROUP"] and txt.count("pass ") == 1  # one launch, one HBM pass
    d = pf.descriptor([1 << 24], "double")
    d.number_of_transforms = 8
    txt = d.describe_plan()
    assert _levels(txt) == ["GLOBAL"] and "scratch_elems=134217728" in txt
    d = pf.descriptor([512, 512, 512])
    txt = d.describe_plan()
    assert len(_levels(txt)) == 3 and txt.count("pass ") == 3  # reference: 1 + 512 + 1 launches
    # a prime factor > 31: the reference throws unsupported_configuration (committed_descriptor_impl.hpp:241); here
    # it becomes a Bluestein plan: two transforms of the padded power-of-two convolution length
    d = pf.descriptor([2 * 7 * 37])
    txt = d.describe_plan()
    assert txt.count("pass ") == 2 and txt.count(" n=2048 ") == 2 and "smod=conv" in txt and "lmod=chirp" in txt
    d = pf.descriptor([65537])
    txt = d.describe_plan()
    # multi-pass Bluestein: chirp pass, then two transforms whose last passes also do the multiply by the transformed
    # chirp and the final chirp / truncation (5 launches; 7 with separate element-wise passes)
    assert _levels(txt) == ["GLOBAL"] and txt.count("kernel=ew") == 1 and txt.count("pass ") == 5
    assert "scratch2_elems=262144" in txt
    d.complex_storage = pf.complex_storage.SPLIT_COMPLEX       # split user data: the final pass stays element-wise
    assert d.describe_plan().count("kernel=ew") >= 2
    d = pf.descriptor([(1 << 24) + 1])
    with pytest.raises(pf.unsupported_configuration):
        d.describe_plan()


def test_planner_pass_factors_multiply_to_length():
    for n in [8192 * 2, 32768, 65536, 131072, 9800, 15360, 68640, 1 << 20, 1 << 24]:
        d = pf.descriptor([n])
        txt = d.describe_plan()
        ns = [int(m) for m in re.findall(r" n=(\d+) ", txt)]
        prod = 1
        for x in ns:
            prod *= x
        assert prod == n, (n, ns)
        for line in txt.splitlines()[1:]:
            rad = re.search(r"radices=([0-9x]+)", line).group(1).split("x")
            pn = int(re.search(r" n=(\d+) ", line).group(1))
            pr = 1
            for r in rad:
                pr *= int(r)
            assert pr == pn


def _kernels(d, direction=None):
    txt = d.describe_plan(direction if direction is not None else pf.direction.FORWARD)
    return re.findall(r"pass kernel=(\w+)", txt)


def test_planner_kernel_choices():
    """Which kernel the planner picks for the BASELINE configs and the packed sizes (host only).  Guards the
    selection logic: the measured numbers in DESIGN.md section 5 belong to exactly these choices."""
    def packed(n, batch, scalar="float"):
        d = pf.descriptor([n], scalar)
        d.number_of_transforms = batch
        return d

    assert _kernels(packed(4096, 65536)) == ["wg_cube"]                      # C2
    for n in (512, 1024, 2048, 8192):
        assert _kernels(packed(n, 4096)) == ["wg_cube"], n                   # cube / rows3 family
    for n in (64, 128, 256):
        assert _kernels(packed(n, 4096)) == ["wg_col"], n                    # TMA tile kernel, rows in and out
    assert _kernels(packed(16, 1 << 20)) == ["wi"] and _kernels(packed(96, 4096)) == ["sg"]
    for n in (512, 1024, 2048, 4096):
        assert _kernels(packed(n, 4096, "double")) == ["wg_cube"], n         # fp64: same kernels, half the tile
    assert _kernels(packed(8192, 64, "double")) != ["wg_cube"]               # two passes in fp64
    assert _kernels(pf.descriptor([512, 512, 512])) == ["wg_cube", "wg_col", "wg_col"]   # C5: z, y, x
    assert _kernels(packed(1 << 24, 8, "double")) == ["wg_col"] * 3          # C4: 256^3
    c3 = pf.descriptor([1000])                                               # C3: split, stride 2, offsets
    c3.number_of_transforms = 100000
    c3.complex_storage = pf.complex_storage.SPLIT_COMPLEX
    c3.forward_strides, c3.forward_distance, c3.forward_offset = [2], 2048, 7
    c3.backward_strides, c3.backward_distance, c3.backward_offset = [1], 1024, 3
    assert _kernels(c3) == ["wg_r3"] and _kernels(c3, pf.direction.BACKWARD) == ["wg_r3"]
    c3b = pf.descriptor([1000])                                              # C3b: batch-interleaved both domains
    c3b.number_of_transforms = 100000
    c3b.complex_storage = pf.complex_storage.SPLIT_COMPLEX
    c3b.forward_strides = c3b.backward_strides = [100000]
    c3b.forward_distance = c3b.backward_distance = 1
    assert _kernels(c3b) == ["wg_colg"]
    r = pf.descriptor([8192], "float", pf.domain.REAL)                       # real: half-length cube, post / pre fused
    r.number_of_transforms = 1024
    assert _kernels(r) == ["wg_cube"] and _kernels(r, pf.direction.BACKWARD) == ["wg_cube"]
    r.complex_storage = pf.complex_storage.SPLIT_COMPLEX                     # split half spectrum: separate passes
    assert _kernels(r) == ["wg_cube", "r2c_post"]
    assert _kernels(r, pf.direction.BACKWARD) == ["c2r_pre", "wg_cube"]
    for n, batch, kern in ((128, 4096, "wg_cube"), (256, 4096, "wg_cube"), (512, 4096, "wg_cube"), (1024, 4096, "wg_cube"),
                           (16384, 64, "wg_cube"), (32, 1 << 20, "wi")):
        r = pf.descriptor([n], "float", pf.domain.REAL)                      # one fused pass in both directions
        r.number_of_transforms = batch
        r.backward_distance = n // 2 + 1
        assert _kernels(r) == [kern] and _kernels(r, pf.direction.BACKWARD) == [kern], n
    r = pf.descriptor([4096], "float", pf.domain.REAL)                       # in-place layout (rows of n + 2 reals): fused
    r.number_of_transforms = 100
    r.placement = pf.placement.IN_PLACE
    r.forward_distance, r.backward_distance = 4098, 2049
    assert _kernels(r) == ["wg_cube"] and _kernels(r, pf.direction.BACKWARD) == ["wg_cube"]
    r = pf.descriptor([131072], "float", pf.domain.REAL)                     # GLOBAL level: two passes + post pass
    r.number_of_transforms = 8
    assert _kernels(r) == ["wg_col", "wg_col", "r2c_post"]
    r = pf.descriptor([1000], "float", pf.domain.REAL)                       # half length 500: no tile kernel, unfused
    r.number_of_transforms = 64
    assert _kernels(r)[-1] == "r2c_post"
    r = pf.descriptor([8192], "float", pf.domain.REAL)                       # strided real rows: pack pass first
    r.number_of_transforms = 4
    r.forward_strides, r.forward_distance = [3], 3 * 8192
    assert _kernels(r) == ["real_pack", "wg_cube", "r2c_post"]
