"""The header-only C++ mirror of the reference API (include/portfft/portfft.hpp) compiled against the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headers_declare_reference_api_names():
    """Every public name of the reference API exists in the drop-in headers."""
    text = "".join(open(os.path.join(ROOT, "include", "portfft", f)).read()
                   for f in os.listdir(os.path.join(ROOT, "include", "portfft")))
    for name in ["enum class domain", "enum class complex_storage", "enum class placement", "enum class direction",
                 "struct descriptor", "class committed_descriptor", "compute_forward", "compute_backward",
                 "get_flattened_length", "get_input_count", "get_output_count", "get_strides", "get_distance",
                 "get_offset", "get_scale", "forward_scale", "backward_scale", "number_of_transforms",
                 "forward_strides", "backward_strides", "forward_distance", "backward_distance", "forward_offset",
                 "backward_offset", "struct invalid_configuration", "struct unsupported_configuration",
                 "struct out_of_local_memory_error", "struct internal_error", "class base_error", "get_real",
                 "get_domain", "scalar_type", "complex_type", "inline direction inv"]:
        assert name in text, name


@pytest.mark.gpu
def test_cpp_api_smoke():
    exe = os.path.join(ROOT, "build", "api_smoke")
    if not os.path.exists(exe):
        subprocess.run(["make", "build/api_smoke"], cwd=ROOT, check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "OK" in r.stdout
