"""CPU oracle for the portFFT descriptor -> commit -> compute path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module.  The product (`portfft_b200/`, `include/`) never does.

It restates, in numpy, what the reference's own test-suite uses to decide whether a transform is right
(all citations relative to /root/reference):

* `gen_data`            <- test/common/reference_data_wrangler.hpp:117-145 (the embedded `python3 -c` script)
* `reshape_to_desc`     <- test/common/reference_data_wrangler.hpp:52-90
* `expected_io`         <- test/common/reference_data_wrangler.hpp:106-257 (`gen_fourier_data`)
* `verify_dft`          <- test/common/reference_data_wrangler.hpp:270-371
* `get_buffer_count`    <- src/portfft/descriptor.hpp:262-270
* `get_default_strides` <- src/portfft/utils.hpp:190-201
* `get_layout`          <- src/portfft/utils.hpp:210-246
* `validate_descriptor` <- src/portfft/descriptor_validation.hpp:38-281

Parity pinning: the reference stores no golden vectors; its expected outputs ARE numpy (`np.fft.fftn` on the
SFC64(0) stream).  This module is therefore pinned by (a) the literal known answers the reference tests hold
(buffer counts 33 / 17, flattened length 6, the invalid-configuration list, padding value -5, first element of the
SFC64(0) stream) -- see tests/test_oracle.py -- and (b) the reference's own code compiled from /root/reference
through `oracle/ref_shim` into `oracle/_ref/` (tests/test_ref_shim.py): its device arithmetic at the WORKITEM,
SUBGROUP and WORKGROUP levels (wi_dft, sg_dft, wg_dft on emulated sub-groups / work-groups), its planner predicates,
and its whole validate_descriptor / get_layout host code (4000 random descriptors, identical verdicts).  The REAL
domain follows the generator's `is_complex = False` branch (rfftn), which the reference's tests never reach because
its validation rejects REAL descriptors.

numpy >= 2 computes `np.fft.fftn(complex64)` in single precision; the reference's script was written for numpy 1.x
("outData is always double precision at this point", reference_data_wrangler.hpp:139), so the transform is
evaluated on a complex128 upcast and cast back, which reproduces what the reference authors compared against.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

FORWARD = 0
BACKWARD = 1

INTERLEAVED_COMPLEX = 0
SPLIT_COMPLEX = 1

IN_PLACE = 0
OUT_OF_PLACE = 1

PACKED = "PACKED"
UNPACKED = "UNPACKED"
BATCH_INTERLEAVED = "BATCH_INTERLEAVED"

PADDING_VALUE = -5.0  # test/unit_test/fft_test_utils.hpp:452


class InvalidConfiguration(Exception):
    """mirrors portfft::invalid_configuration (src/portfft/common/exceptions.hpp:55-58)"""


class UnsupportedConfiguration(Exception):
    """mirrors portfft::unsupported_configuration (src/portfft/common/exceptions.hpp:63-66)"""


def inv(direction: int) -> int:
    """src/portfft/enums.hpp:36-38"""
    return BACKWARD if direction == FORWARD else FORWARD


def get_default_strides(lengths: Sequence[int]) -> List[int]:
    """src/portfft/utils.hpp:190-201: row-major, last stride 1."""
    strides = [0] * len(lengths)
    total = 1
    for i in range(len(lengths) - 1, -1, -1):
        strides[i] = total
        total *= lengths[i]
    return strides


@dataclass
class OracleDescriptor:
    """Plain mirror of the public fields of portfft::descriptor (src/portfft/descriptor.hpp:59-144)."""

    lengths: List[int]
    forward_scale: float = 1.0
    backward_scale: float = 1.0
    number_of_transforms: int = 1
    complex_storage: int = INTERLEAVED_COMPLEX
    placement: int = OUT_OF_PLACE
    forward_strides: List[int] = field(default_factory=list)
    backward_strides: List[int] = field(default_factory=list)
    forward_distance: Optional[int] = None
    backward_distance: Optional[int] = None
    forward_offset: int = 0
    backward_offset: int = 0
    is_double: bool = False
    is_real: bool = False  # Domain == domain::REAL (descriptor.hpp:43-57)

    def __post_init__(self):
        self.lengths = [int(x) for x in self.lengths]
        if not self.forward_strides:
            self.forward_strides = get_default_strides(self.lengths)
        if not self.backward_strides:
            self.backward_strides = get_default_strides(self.lengths)
        # ctor default: distance = flattened length (descriptor.hpp:141-143)
        if self.forward_distance is None:
            self.forward_distance = self.get_flattened_length()
        if self.backward_distance is None:
            self.backward_distance = self.get_flattened_length()

    # descriptor.hpp:161-163
    def get_flattened_length(self) -> int:
        return int(np.prod(self.lengths, dtype=np.int64)) if self.lengths else 1

    def get_strides(self, d):
        return self.forward_strides if d == FORWARD else self.backward_strides

    def get_distance(self, d):
        return self.forward_distance if d == FORWARD else self.backward_distance

    def get_offset(self, d):
        return self.forward_offset if d == FORWARD else self.backward_offset

    def get_scale(self, d):
        return self.forward_scale if d == FORWARD else self.backward_scale

    def domain_lengths(self, d) -> List[int]:
        """Lengths of the data in one domain.  REAL descriptors hold lengths[-1] // 2 + 1 complex elements along the
        last dimension of the backward domain -- the shape numpy.fft.rfftn returns, which is what the reference's
        generator reads back (test/common/reference_data_wrangler.hpp:136-137,196)."""
        l = list(self.lengths)
        if self.is_real and d == BACKWARD:
            l[-1] = l[-1] // 2 + 1
        return l

    # descriptor.hpp:172-183
    def get_input_count(self, d) -> int:
        return get_buffer_count(self.domain_lengths(d), self.number_of_transforms, self.get_strides(d),
                                self.get_distance(d), self.get_offset(d))

    def get_output_count(self, d) -> int:
        return self.get_input_count(inv(d))


def get_buffer_count(lengths, number_of_transforms, strides, distance, offset) -> int:
    """src/portfft/descriptor.hpp:262-270."""
    last = (number_of_transforms - 1) * distance
    for n, s in zip(lengths, strides):
        last += (n - 1) * s
    return offset + last + 1


def get_layout(desc: OracleDescriptor, d: int) -> str:
    """src/portfft/utils.hpp:210-246."""
    if desc.get_strides(d) == get_default_strides(desc.lengths) and desc.get_distance(d) == desc.get_flattened_length():
        return PACKED
    if len(desc.lengths) == 1 and desc.get_distance(d) == 1 and desc.get_strides(d)[-1] == desc.number_of_transforms:
        return BATCH_INTERLEAVED
    return UNPACKED


# ---------------------------------------------------------------------------------------------------------------------
# validation (descriptor_validation.hpp)
# ---------------------------------------------------------------------------------------------------------------------

def _validate_lengths(lengths):
    """descriptor_validation.hpp:38-47"""
    if len(lengths) == 0:
        raise InvalidConfiguration("Invalid lengths, must have at least 1 dimension")
    for i, n in enumerate(lengths):
        if n == 0:
            raise InvalidConfiguration(f"Invalid lengths[{i}]=0, must be positive")


def _basic(lengths, batch, strides, distance, name):
    """descriptor_validation.hpp:92-111"""
    if len(strides) != len(lengths):
        raise InvalidConfiguration(f"Mismatching {name} strides length")
    for i, s in enumerate(strides):
        if s == 0:
            raise InvalidConfiguration(f"Invalid {name} stride[{i}]=0, must be positive")
    if batch > 1 and distance == 0:
        raise InvalidConfiguration(f"Invalid {name} distance 0, must be positive for batched FFTs")


def _multidim(lengths, batch, strides, distance, name):
    """descriptor_validation.hpp:123-151"""
    gs = list(strides)
    gn = list(lengths)
    if batch > 1:
        gs.append(distance)
        gn.append(batch)
    # std::sort is not stable; ties between equal strides resolve to an overlap either way
    idx = sorted(range(len(gn)), key=lambda a: gs[a])
    for i in range(1, len(idx)):
        if not (gs[idx[i - 1]] * gn[idx[i - 1]] <= gs[idx[i]]):
            raise InvalidConfiguration(f"Domain {name}: multi-dimension strides are not large enough to avoid overlap")


def _onedim(lengths, batch, strides, distance, name):
    """descriptor_validation.hpp:162-204"""
    fft_size = lengths[0]
    stride = strides[0]
    first_batch_limit = stride * fft_size
    first_length_limit = distance * batch
    if (stride <= distance and first_batch_limit <= distance) or (distance <= stride and first_length_limit <= stride):
        return
    b = 1
    while b < batch:
        first = b * distance
        column = first % stride
        if column == 0:
            if first >= first_batch_limit:
                return
            raise InvalidConfiguration(f"Domain {name}: batch {b} collides with first batch at index {first}")
        skip = (stride - column) // distance
        if (stride - column) % distance != 0:
            skip += 1
        b += skip


def _strides_distance_check(lengths, batch, strides, distance, name):
    """descriptor_validation.hpp:215-224"""
    _basic(lengths, batch, strides, distance, name)
    if len(lengths) > 1:
        _multidim(lengths, batch, strides, distance, name)
    else:
        _onedim(lengths, batch, strides, distance, name)


def validate_descriptor(desc: OracleDescriptor, reference_layout_limits: bool = False) -> None:
    """descriptor_validation.hpp:264-281.

    `reference_layout_limits=True` additionally applies `validate_layout` (:57-81), which throws
    `unsupported_configuration` for N-D non-PACKED layouts and for UNPACKED layouts beyond sub-group sizes.  The
    B200 library lifts those *unsupported* restrictions but must reject every *invalid* configuration.
    """
    if desc.number_of_transforms == 0:
        raise InvalidConfiguration("Invalid number of transform 0, must be positive")
    _validate_lengths(desc.lengths)
    if desc.placement == IN_PLACE:
        if list(desc.forward_strides) != list(desc.backward_strides):
            raise InvalidConfiguration("Invalid forward and backward strides must match for in-place configurations")
        if desc.forward_distance != desc.backward_distance:
            raise InvalidConfiguration("Invalid forward and backward distances must match for in-place configurations")
        _strides_distance_check(desc.lengths, desc.number_of_transforms, desc.forward_strides, desc.forward_distance,
                                "forward")
    else:
        _strides_distance_check(desc.lengths, desc.number_of_transforms, desc.forward_strides, desc.forward_distance,
                                "forward")
        _strides_distance_check(desc.lengths, desc.number_of_transforms, desc.backward_strides, desc.backward_distance,
                                "backward")
    if reference_layout_limits:
        fl, bl = get_layout(desc, FORWARD), get_layout(desc, BACKWARD)
        if len(desc.lengths) > 1 and not (fl == PACKED and bl == PACKED):
            raise UnsupportedConfiguration("Multi-dimensional transforms are only supported with default data layout")
        if (fl == UNPACKED or bl == UNPACKED) and not ref_fits_in_sg(desc.lengths[-1], 32, desc.is_double):
            raise UnsupportedConfiguration("Arbitrary strides only supported for sizes that fit a subgroup")


# ---------------------------------------------------------------------------------------------------------------------
# planner predicates of the reference (used to label cases / for the reference_layout_limits switch)
# ---------------------------------------------------------------------------------------------------------------------

def ref_factorize(n: int) -> int:
    """src/portfft/common/workitem.hpp:135-144: largest divisor <= sqrt(N), 1 when prime."""
    res = 1
    i = 2
    while i * i <= n:
        if n % i == 0:
            res = i
        i += 1
    return res


def ref_wi_temps(n: int, level: int = 0) -> int:
    """src/portfft/common/workitem.hpp:154-169 (MaxRecursionLevel = int_log2(56) - 1 = 4)."""
    f0 = ref_factorize(n)
    f1 = n // f0
    if f0 < 2 or f1 < 2:
        return n
    a = b = 2
    if level < 4:
        a = ref_wi_temps(f0, level + 1)
        b = ref_wi_temps(f1, level + 1)
    return max(a, b) + n


def ref_fits_in_wi(n: int, is_double: bool, registers_per_wi: int = 128) -> bool:
    """src/portfft/common/workitem.hpp:179-185."""
    return (n + ref_wi_temps(n)) * 2 * (8 if is_double else 4) <= registers_per_wi * 4


def ref_factorize_sg(n: int, sg: int) -> int:
    """src/portfft/common/subgroup.hpp:226-238: largest divisor of N that is <= sg."""
    for i in range(sg, 1, -1):
        if n % i == 0:
            return i
    return 1


def ref_fits_in_sg(n: int, sg: int, is_double: bool) -> bool:
    """src/portfft/common/subgroup.hpp:248-253."""
    return ref_fits_in_wi(n // ref_factorize_sg(n, sg), is_double)


# ---------------------------------------------------------------------------------------------------------------------
# data generation / expected results
# ---------------------------------------------------------------------------------------------------------------------

def gen_data(batch: int, dims: Sequence[int], is_double: bool, seed: int = 0,
             is_real: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Restatement of the numpy script at test/common/reference_data_wrangler.hpp:117-145.

    Returns (inData, outData) with shape [batch] + dims; `outData = fftn(inData, axes=1..)`, unscaled.  REAL domain
    (`is_complex` False in the script): inData is real and outData = rfftn(inData), last dimension dims[-1] // 2 + 1.
    """
    scalar_type = np.float64 if is_double else np.float32
    complex_type = np.complex128 if is_double else np.complex64
    shape = [int(batch)] + [int(d) for d in dims]
    rng = np.random.Generator(np.random.SFC64(seed))
    in_data = rng.uniform(-1, 1, shape).astype(scalar_type)
    axes = tuple(range(1, len(dims) + 1))
    if is_real:
        out_data = np.fft.rfftn(in_data.astype(np.float64), axes=axes).astype(complex_type)
        return in_data, out_data
    in_data = in_data + 1j * rng.uniform(-1, 1, shape).astype(scalar_type)
    in_data = in_data.astype(complex_type)
    out_data = np.fft.fftn(in_data.astype(np.complex128), axes=axes).astype(complex_type)
    return in_data, out_data


def reshape_to_desc(packed: np.ndarray, desc: OracleDescriptor, direction: int,
                    padding_value: float = PADDING_VALUE) -> np.ndarray:
    """test/common/reference_data_wrangler.hpp:52-90, generalised to N-D strides (the reference only scatters 1-D
    because its N-D tests are PACKED-only): element (b; i_1..i_d) -> offset + b*distance + sum(i_k*stride_k)
    (src/portfft/descriptor.hpp:91-92).  Unaddressed elements hold `padding_value` (complex: (p, p))."""
    count = desc.get_input_count(direction)
    pad = padding_value + 1j * padding_value if np.iscomplexobj(packed) else padding_value
    out = np.full(count, pad, dtype=packed.dtype)
    idx = element_indices(desc, direction)
    out[idx.reshape(-1)] = packed.reshape(-1)
    return out


def element_indices(desc: OracleDescriptor, direction: int) -> np.ndarray:
    """Flat index of every addressed element, shape [batch] + lengths (src/portfft/descriptor.hpp:91-92)."""
    idx = desc.get_offset(direction) + np.arange(desc.number_of_transforms, dtype=np.int64) * desc.get_distance(direction)
    idx = idx.reshape([desc.number_of_transforms] + [1] * len(desc.lengths))
    for k, (n, s) in enumerate(zip(desc.domain_lengths(direction), desc.get_strides(direction))):
        shape = [1] * (len(desc.lengths) + 1)
        shape[k + 1] = n
        idx = idx + (np.arange(n, dtype=np.int64) * s).reshape(shape)
    return idx


def expected_io(desc: OracleDescriptor, direction: int, seed: int = 0,
                padding_value: float = PADDING_VALUE) -> Tuple[np.ndarray, np.ndarray]:
    """`gen_fourier_data` (reference_data_wrangler.hpp:106-257): returns (input buffer, expected output buffer) in the
    descriptor's layouts, complex dtype (callers split into re / im planes for SPLIT_COMPLEX storage).

    FORWARD: input = numpy input in the forward layout, expected = fftn * forward_scale in the backward layout.
    BACKWARD: input = fftn in the backward layout, expected = numpy input * backward_scale * N in the forward layout
    (:202-210)."""
    scalar = np.float64 if desc.is_double else np.float32
    fwd, bwd = gen_data(desc.number_of_transforms, desc.lengths, desc.is_double, seed, desc.is_real)
    if direction == FORWARD:
        bwd = bwd * scalar(desc.forward_scale)
    else:
        fwd = fwd * (scalar(desc.backward_scale) * scalar(desc.get_flattened_length()))
    fwd_buf = reshape_to_desc(fwd, desc, FORWARD, padding_value)
    bwd_buf = reshape_to_desc(bwd, desc, BACKWARD, padding_value)
    return (fwd_buf, bwd_buf) if direction == FORWARD else (bwd_buf, fwd_buf)


def rel_l2_bound(flat_len: int, is_double: bool) -> float:
    """north_star accuracy bar: 1e-5*log2(N) fp32, 1e-13*log2(N) fp64 (N = flattened length; >= 1 bit)."""
    return (1e-13 if is_double else 1e-5) * max(1.0, math.log2(max(2, flat_len)))


def reference_elem_tolerance(flat_len: int, is_double: bool) -> float:
    """test/unit_test/fft_test_utils.hpp:461-464: 2 * eps * N * log2(N)."""
    eps = np.finfo(np.float64 if is_double else np.float32).eps
    return 2.0 * float(eps) * flat_len * math.log2(max(2, flat_len))


def max_rel_l2(desc: OracleDescriptor, direction: int, ref: np.ndarray, actual: np.ndarray) -> float:
    """Max over batches of ||actual - ref||_2 / ||ref||_2 on the addressed elements
    (reference_data_wrangler.hpp:325-353, evaluated in float64)."""
    idx = element_indices(desc, inv(direction)).reshape(desc.number_of_transforms, -1)
    r = ref[idx].astype(np.complex128)
    a = actual[idx].astype(np.complex128)
    err = np.sqrt(np.sum(np.abs(a - r) ** 2, axis=1))
    nrm = np.sqrt(np.sum(np.abs(r) ** 2, axis=1))
    nrm = np.where(nrm == 0, 1.0, nrm)
    return float(np.max(err / nrm))


def verify_dft(desc: OracleDescriptor, direction: int, ref: np.ndarray, actual: np.ndarray,
               rel_l2_tol: Optional[float] = None) -> float:
    """`verify_dft` (reference_data_wrangler.hpp:270-371) plus the north_star relative-L2 assertion.

    (a) the `offset` prefix must be bit-identical (:300-317);
    (b) every element after the offset -- including padding between strided elements, which still holds -5 in `ref`
        -- must satisfy abs_diff <= tol or rel_diff <= tol with tol = 2*eps*N*log2(N) (:355-370);  for unaddressed
        elements we are stricter than the reference and require exact equality;
    (c) max relative L2 over batches must be within `rel_l2_tol` (the reference only logs it, :353).
    Returns the max relative L2 error."""
    assert ref.shape == actual.shape, (ref.shape, actual.shape)
    out_dir = inv(direction)
    off = desc.get_offset(out_dir)
    if not np.array_equal(ref[:off], actual[:off]):
        bad = int(np.nonzero(ref[:off] != actual[:off])[0][0])
        raise AssertionError(f"Incorrectly written value in padding at global idx {bad}")
    addressed = np.zeros(ref.shape[0], dtype=bool)
    addressed[element_indices(desc, out_dir).reshape(-1)] = True
    if not np.array_equal(ref[~addressed], actual[~addressed]):
        bad = int(np.nonzero((ref != actual) & ~addressed)[0][0])
        raise AssertionError(f"unaddressed element {bad} was modified: {actual[bad]} (expected {ref[bad]})")
    flat = desc.get_flattened_length()
    tol = reference_elem_tolerance(flat, desc.is_double)
    a = actual[addressed].astype(np.complex128)
    r = ref[addressed].astype(np.complex128)
    abs_diff = np.abs(a - r)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel_diff = abs_diff / np.abs(a)
    bad = (abs_diff > tol) & ~(rel_diff <= tol)
    if np.any(bad):
        i = int(np.nonzero(bad)[0][0])
        raise AssertionError(f"value at addressed element #{i} does not match: ref {r[i]} vs {a[i]}, tol {tol}")
    l2 = max_rel_l2(desc, direction, ref, actual)
    bound = rel_l2_bound(flat, desc.is_double) if rel_l2_tol is None else rel_l2_tol
    if not l2 <= bound:
        raise AssertionError(f"max relative L2 error {l2:.3e} exceeds bound {bound:.3e}")
    return l2
