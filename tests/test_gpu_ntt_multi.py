"""GPU: ONE transform over the device group through the host-pointer C ABI (uzkge_cuda_ntt_fr_multi): four-step decomposition, both
exchanges done by the kernels' own loads / stores over peer memory, natural output folded into the last pass.  Same contract as
uzkge_cuda_ntt_fr, so every case is held to the oracle's single transform (FpPolynomial::{fft,ifft,coset_fft,coset_ifft}_with_domain,
field_polynomial.rs:583-607).  Groups are virtual on a one-GPU box (G members on device 0) and real under gpurun --gpus N."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _group(gpu, members: int) -> int:
    gpu.configure("virtual_devices", members)
    return gpu.init_devices(0)


def _multi(gpu, data, n, inverse=False, coset=None):
    buf = np.zeros((n, 4), dtype=np.uint64)
    buf[: data.shape[0]] = data
    gpu.ntt_fr_multi_inplace(buf, data.shape[0], n, inverse, coset)
    return buf


@pytest.mark.parametrize("members", [0, 2, 4, 8])
def test_group_transform_matches_the_oracle(gpu, oc, members):
    try:
        G = _group(gpu, members)
        assert G == (members or gpu.device_count())
        k = oc.random_fr(1, 31)[0]
        k_inv = oc.fr_inv(k)
        for log_n in (6, 9, 13, 16):
            n = 1 << log_n
            x = oc.random_fr(n, 900 + log_n)
            ev = oc.ntt_fr(x, n)
            assert np.array_equal(_multi(gpu, x, n), ev), (G, log_n)
            assert np.array_equal(_multi(gpu, ev, n, inverse=True), x), (G, log_n, "inverse")
            cev = oc.ntt_fr(x, n, coset=k)
            assert np.array_equal(_multi(gpu, x, n, coset=k), cev), (G, log_n, "coset")
            assert np.array_equal(_multi(gpu, cev, n, inverse=True, coset=k_inv), x), (G, log_n, "coset inverse")
            # ragged input: a polynomial of fewer coefficients than the domain (zero padding on the device), incl. one that ends inside a slice
            for m in (1, n // 3 + 1, n - 1):
                assert np.array_equal(_multi(gpu, x[:m], n, coset=k), oc.ntt_fr(x[:m], n, coset=k)), (G, log_n, m)
        # where the four-step does not apply the call is the single-device transform: 3 * 2^k domains, sizes below G^2
        for n in (3 << 7, 2, 16):
            x = oc.random_fr(n, 77 + n)
            assert np.array_equal(_multi(gpu, x, n), oc.ntt_fr(x, n)), (G, n)
        with pytest.raises(Exception):
            _multi(gpu, oc.random_fr(9, 1), 8)          # input longer than the domain
        with pytest.raises(Exception):
            _multi(gpu, oc.random_fr(4, 1), 5)          # not a domain size
    finally:
        _group(gpu, 0)


def test_group_transform_at_the_benchmark_size(gpu, oc):
    """2^22 (BASELINE's NTT size) over a virtual group of 8 against the single-device transform, and a forward / inverse round trip."""
    try:
        assert _group(gpu, 8) == 8
        n = 1 << 22
        x = oc.random_fr(n, 4242)
        want = gpu.ntt_fr(x, n)
        got = _multi(gpu, x, n)
        assert np.array_equal(got, want)
        assert np.array_equal(_multi(gpu, got, n, inverse=True), x)
    finally:
        _group(gpu, 0)
