"""CPU: the C++ host layer (include/uzkge_host.hpp) compiles warning-free against the C ABI, its host-side logic (trimming, negation,
domain-size rules, DegreeError / ParameterError before any device call) holds, and without a GPU every device-backed call fails
loudly with the UzkgeError variant INTEGRATION.md maps it to -- there is no CPU path behind it."""
import pytest

from host_cpp import run


def test_cpp_host_layer_builds_and_refuses_to_run_without_a_device():
    from uzkge_b200 import ffi

    if ffi.lib().uzkge_cuda_device_count() > 0:
        pytest.skip("a CUDA device is visible: the parity run is tests/test_gpu_zz_host_cpp.py")
    out = run("nodevice")
    assert out.returncode == 0 and "PASS nodevice" in out.stdout, out.stdout + out.stderr
