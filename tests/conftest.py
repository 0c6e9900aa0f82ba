import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        from uzkge_b200 import ffi

        return ffi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oc():
    from oracle import cpu

    cpu.lib()
    return cpu


@pytest.fixture(scope="session")
def bn():
    from oracle import bn254

    return bn254


@pytest.fixture(scope="session")
def domain_kat():
    return json.load(open(os.path.join(GOLDEN, "domain_kat.json")))


def _load_points_mont(name, oc):
    """Golden SRS fixtures hold canonical affine coordinates; the ABI wants Montgomery limbs."""
    raw = np.load(os.path.join(GOLDEN, name))
    flat = oc.fq_to_mont(raw.reshape(-1, 4))
    return flat.reshape(-1, 8)


@pytest.fixture(scope="session")
def lagrange_srs_4096(oc):
    return _load_points_mont("lagrange_srs_4096.npy", oc)


@pytest.fixture(scope="session")
def lagrange_srs_16384(oc):
    return _load_points_mont("lagrange_srs_16384.npy", oc)


@pytest.fixture(scope="session")
def srs_padding_head(oc):
    return _load_points_mont("srs_padding_head.npy", oc)


@pytest.fixture(scope="session")
def lagrange_srs_8192(oc):
    return _load_points_mont("lagrange_srs_8192.npy", oc)


@pytest.fixture(scope="session")
def srs_padding_tail(oc):
    """tau^n, tau^(n+1), tau^(n+2) for n = 4096, 8192, 16384 (rows 0-2, 3-5, 6-8): srs-padding.bin[2051..2060)."""
    return _load_points_mont("srs_padding_tail.npy", oc)


@pytest.fixture(scope="session")
def gpu():
    """The CUDA backend through its C ABI; fails loudly (no skip, no fallback) if it cannot start on a GPU box."""
    from uzkge_b200 import ffi

    ffi.init(0)
    if os.environ.get("UZKGE_MSM_AFFINE"):      # run the same suite over the batched-affine accumulation (2 = forced at every size)
        ffi.configure("msm_affine", int(os.environ["UZKGE_MSM_AFFINE"]))
    return ffi
