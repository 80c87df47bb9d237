"""portfft_b200: B200-native (sm_100a) batched C2C FFT behind portFFT's descriptor -> commit -> compute API.

The product is the C-ABI shared library `lib/libpfft_b200.so` (include/pfft.h); this package is its Python mirror
of the reference's public interface, used by the tests and the benchmark.  Importing it never falls back to a CPU
implementation: the library must be built (`make`).
"""
from .api import (base_error, committed_descriptor, complex_storage, cuda_error, descriptor, direction, domain,
                  get_default_strides, internal_error, inv, invalid_configuration, layout, level,
                  out_of_local_memory_error, placement, total_launches, unsupported_configuration)
from ._lib import LIB_PATH, load

__all__ = [
    "descriptor", "committed_descriptor", "domain", "complex_storage", "placement", "direction", "inv", "level",
    "layout", "base_error", "internal_error", "invalid_configuration", "unsupported_configuration",
    "out_of_local_memory_error", "cuda_error", "get_default_strides", "total_launches", "load", "LIB_PATH",
]
