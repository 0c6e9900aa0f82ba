"""CPU: the C-ABI shared library loads, exports every symbol include/uzkge_cuda.h declares, and -- without a GPU --
refuses to compute (no CPU fallback).  Host-only entry points are checked against the golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

from uzkge_b200 import ffi
from uzkge_b200.errors import BackendUnavailable


@pytest.fixture(scope="module")
def lib():
    from uzkge_b200 import build

    build.build()
    return ffi.lib()


def test_header_declares_and_library_exports_every_symbol(lib):
    syms = ffi.header_symbols()
    assert len(syms) >= 20
    assert set(syms) == set(ffi._SIGNATURES), "ffi.py must bind exactly what the header declares"
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/uzkge_cuda.h but not exported"


def test_version_and_error_codes(lib):
    assert "sm_100a" in ffi.version()
    txt = open(ffi.HEADER_PATH).read()
    for name, val in (("UZKGE_OK", 0), ("UZKGE_ERR_NO_DEVICE", 1), ("UZKGE_ERR_SIZE", 2), ("UZKGE_ERR_CUDA", 3),
                      ("UZKGE_ERR_OOM", 4), ("UZKGE_ERR_HANDLE", 5), ("UZKGE_ERR_ARG", 6)):
        assert f"#define {name} {val}" in txt


def test_root_of_unity_host_entry_point_matches_golden(lib, bn, domain_kat):
    for e in domain_kat.values():
        w = ffi.fr_root_of_unity(e["cs_size"])
        assert bn.array_to_ints(w.reshape(1, 4), bn.FR)[0] == int(e["root"], 16)
    out = np.zeros(4, dtype=np.uint64)
    assert lib.uzkge_cuda_fr_root_of_unity(5, ffi.ptr(out)) == ffi.ERR_SIZE
    assert "fr_root_of_unity" in ffi.last_error()
    assert lib.uzkge_cuda_fr_root_of_unity(8, None) == ffi.ERR_ARG


@pytest.mark.skipif(ffi.lib().uzkge_cuda_device_count() > 0, reason="only meaningful without a GPU")
def test_no_gpu_means_no_result(lib):
    """No device: every compute entry point fails with UZKGE_ERR_NO_DEVICE; nothing is computed on the CPU."""
    buf = np.zeros((8, 4), dtype=np.uint64)
    assert lib.uzkge_cuda_init(-1) == ffi.ERR_NO_DEVICE
    assert lib.uzkge_cuda_ntt_fr(ffi.ptr(buf), 8, 8, 0, None) == ffi.ERR_NO_DEVICE
    assert not buf.any()
    h = C.c_uint64(0)
    pts = np.zeros((4, 8), dtype=np.uint64)
    assert lib.uzkge_cuda_srs_upload(ffi.ptr(pts), 4, 0, C.byref(h)) == ffi.ERR_NO_DEVICE
    # the device group: nothing to form a group from, and the group calls refuse as well
    assert lib.uzkge_cuda_init_devices(0) == ffi.ERR_NO_DEVICE
    assert lib.uzkge_cuda_group_size() == 0
    assert lib.uzkge_cuda_srs_upload_multi(ffi.ptr(pts), 4, 0, 0, C.byref(h)) == ffi.ERR_NO_DEVICE
    assert lib.uzkge_cuda_ntt_fr_multi(ffi.ptr(buf), 8, 8, 0, None) == ffi.ERR_NO_DEVICE
    assert not buf.any()
    assert lib.uzkge_cuda_plonk_params_upload_multi(None, C.byref(h)) in (ffi.ERR_ARG, ffi.ERR_NO_DEVICE)
    with pytest.raises(BackendUnavailable):
        ffi.ntt_fr(buf, 8)
    with pytest.raises(BackendUnavailable):
        from uzkge_b200 import KZGCommitmentSchemeBN254

        KZGCommitmentSchemeBN254(pts)


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under uzkge_b200/ may reference it."""
    root = os.path.join(os.path.dirname(ffi.HERE), "uzkge_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
