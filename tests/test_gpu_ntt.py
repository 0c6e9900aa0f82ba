"""GPU: Fr NTT / iNTT / coset transforms through the C ABI (uzkge_cuda_ntt_fr) and the FpPolynomial mirror,
bit-exact against the oracle; reference tests mirrored: test_fft / check_fft
(/root/reference/uzkge/src/poly_commit/field_polynomial.rs:632-719)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K1 = 0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A  # golden coset shift k[1]

RADIX2 = [1 << k for k in range(0, 17)]
MIXED = [3 << k for k in range(0, 16)]


def k1(bn):
    return bn.ints_to_array([K1], bn.FR)[0]


@pytest.mark.parametrize("n", RADIX2 + MIXED)
def test_ntt_all_variants_match_oracle(gpu, oc, bn, n):
    x = oc.random_fr(n, 1000 + n % 101)
    k = k1(bn)
    kinv = bn.ints_to_array([bn.inv_mod(K1, bn.FR)], bn.FR)[0]
    for len_in in sorted({n, n // 2 + 1, 1}):
        xi = x[:len_in]
        assert np.array_equal(gpu.ntt_fr(xi, n), oc.ntt_fr(xi, n))
        assert np.array_equal(gpu.ntt_fr(xi, n, inverse=True), oc.ntt_fr(xi, n, inverse=True))
        assert np.array_equal(gpu.ntt_fr(xi, n, coset_shift=k), oc.ntt_fr(xi, n, coset=k))
        assert np.array_equal(gpu.ntt_fr(xi, n, inverse=True, coset_shift=kinv), oc.ntt_fr(xi, n, inverse=True, coset=kinv))


@pytest.mark.parametrize("n", [1 << 18, 1 << 20, 3 << 17, 98304 * 4])
def test_ntt_medium_sizes_match_oracle(gpu, oc, bn, n):
    x = oc.random_fr(n, 5)
    k = k1(bn)
    assert np.array_equal(gpu.ntt_fr(x, n), oc.ntt_fr(x, n))
    assert np.array_equal(gpu.ntt_fr(x[: n // 6 + 3], n, coset_shift=k), oc.ntt_fr(x[: n // 6 + 3], n, coset=k))
    assert np.array_equal(gpu.ntt_fr(x, n, inverse=True), oc.ntt_fr(x, n, inverse=True))


@pytest.mark.parametrize("n", [1, 2, 3])
def test_check_fft(gpu, oc, n):
    """fft[i] == poly.eval(root^i) with root = domain.group_gen, sizes 1, 2, 3 (field_polynomial.rs:632-646)."""
    from uzkge_b200 import FpPolynomial

    coefs = oc.random_fr(n, 60 + n)
    poly = FpPolynomial.from_coefs(coefs)
    dom = FpPolynomial.quotient_evaluation_domain(n)
    ev = poly.fft_with_domain(dom)
    assert np.array_equal(dom.group_gen, oc.fr_root_of_unity(n))
    for i in range(n):
        assert np.array_equal(ev[i], oc.fr_eval(poly.coefs, oc.fr_pow(dom.group_gen, i)))


@pytest.mark.parametrize("n", [16, 32, 3, 48])
def test_fft_ifft_round_trip_reference_sizes(gpu, oc, n):
    """test_fft's round trips: 16, 32 through the radix-2 domain; 3, 48 through the mixed-radix domain."""
    from uzkge_b200 import FpPolynomial

    poly = FpPolynomial.from_coefs(oc.random_fr(n, 70 + n))
    dom = FpPolynomial.evaluation_domain(n) if n & (n - 1) == 0 else FpPolynomial.quotient_evaluation_domain(n)
    ev = poly.fft_with_domain(dom)
    assert ev.shape[0] == n
    assert FpPolynomial.ifft_with_domain(dom, ev) == poly
    assert np.array_equal(poly.fft(n), ev)


def test_polynomial_layer_semantics(gpu, oc, bn):
    from uzkge_b200 import FpPolynomial

    n = 64
    c = oc.random_fr(40, 9)
    c[30:] = 0  # trailing zeros are trimmed by from_coefs
    poly = FpPolynomial.from_coefs(c)
    assert poly.degree() == 29
    dom = FpPolynomial.quotient_evaluation_domain(6 * n)
    k = k1(bn)
    kinv = bn.ints_to_array([bn.inv_mod(K1, bn.FR)], bn.FR)[0]
    ev = poly.coset_fft_with_domain(dom, k)
    assert np.array_equal(ev, oc.ntt_fr(poly.coefs, 6 * n, coset=k))
    back = FpPolynomial.coset_ifft_with_domain(dom, ev, kinv)
    assert back == poly and back.degree() == 29
    zero = FpPolynomial.from_coefs(np.zeros((5, 4), dtype=np.uint64))
    assert zero.degree() == 0 and zero.is_zero()
    assert not dom.fft(zero.coefs).any()
    with pytest.raises(AssertionError):
        poly.fft_with_domain(FpPolynomial.evaluation_domain(16))  # domain.size() > degree is asserted


def test_special_vectors(gpu, oc, bn):
    n = 4096
    one = bn.ints_to_array([1], bn.FR)
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[0] = one[0]
    assert np.array_equal(gpu.ntt_fr(delta, n), np.repeat(one, n, axis=0))  # fft(delta) = all ones
    ones = np.repeat(one, n, axis=0)
    nn = bn.ints_to_array([n], bn.FR)
    want = np.zeros((n, 4), dtype=np.uint64)
    want[0] = nn[0]
    assert np.array_equal(gpu.ntt_fr(ones, n), want)
    assert np.array_equal(gpu.ntt_fr(np.zeros((0, 4), dtype=np.uint64), n), np.zeros((n, 4), dtype=np.uint64))


def test_bad_sizes_are_errors(gpu):
    from uzkge_b200.errors import FFTError, ParameterError

    x = np.zeros((8, 4), dtype=np.uint64)
    for n in (0, 5, 7, 9, 18, 3 << 29):
        with pytest.raises((FFTError, ParameterError)):
            gpu.ntt_fr(x[: min(8, n)], n)
    with pytest.raises(ParameterError):
        gpu.ntt_fr(x, 4)


@pytest.mark.parametrize("n", [1 << 22, 1 << 24, 3 << 21, 3 << 23])
def test_full_size_properties(gpu, oc, bn, n):
    """BASELINE sizes, through size-independent properties: round trip, linearity, Horner spot checks."""
    x = oc.random_fr(n, 11)
    ev = gpu.ntt_fr(x, n)
    assert np.array_equal(gpu.ntt_fr(ev, n, inverse=True), x)
    w = oc.fr_root_of_unity(n)
    rng = np.random.default_rng(5)
    for i in [0, 1, n - 1] + list(rng.integers(0, n, size=3)):
        assert np.array_equal(ev[int(i)], oc.fr_eval(x, oc.fr_pow(w, int(i))))
    # linearity: fft(x + y) = fft(x) + fft(y) at sampled positions; x + y built with the oracle's field add via
    # fr_mul-free trick: y = x * c  =>  fft(y) = c * fft(x)
    c = oc.random_fr(1, 12)
    y = oc.fr_mul(x, np.repeat(c, n, axis=0))
    evy = gpu.ntt_fr(y, n)
    idx = rng.integers(0, n, size=4096)
    assert np.array_equal(evy[idx], oc.fr_mul(ev[idx], np.repeat(c, idx.size, axis=0)))
    # coset round trip with the golden shift
    k = k1(bn)
    kinv = bn.ints_to_array([bn.inv_mod(K1, bn.FR)], bn.FR)[0]
    cev = gpu.ntt_fr(x[: n // 2], n, coset_shift=k)
    back = gpu.ntt_fr(cev, n, inverse=True, coset_shift=kinv)
    assert np.array_equal(back[: n // 2], x[: n // 2]) and not back[n // 2 :].any()


@pytest.mark.parametrize("n", [64, 1 << 14, 3 << 12])
def test_batch_host_transforms_match_single_calls(gpu, oc, n):
    """uzkge_cuda_ntt_fr_batch: k host vectors through the three-buffer copy / transform / copy pipeline (more vectors than buffers,
    ragged inputs, forward, inverse and coset) against the oracle."""
    k = 7
    shift = oc.random_fr(1, 5)[0]
    lens = [n, n // 2 + 1, 1, n, 3, n - 1, n]
    for inverse, cs in ((False, None), (True, None), (False, shift), (True, shift)):
        src = [oc.random_fr(n, 300 + j) for j in range(k)]
        bufs = []
        for j in range(k):
            b = np.zeros((n, 4), dtype=np.uint64)
            b[: lens[j]] = src[j][: lens[j]]
            bufs.append(b)
        want = [oc.ntt_fr(src[j][: lens[j]], n, inverse=inverse, coset=cs) for j in range(k)]
        gpu.ntt_fr_batch_inplace(bufs, lens, n, inverse=inverse, coset_shift=cs)
        for j in range(k):
            assert np.array_equal(bufs[j], want[j]), (inverse, cs is not None, j)


@pytest.mark.parametrize("n,k", [(1 << 10, 3), (1 << 14, 8), (3 << 13, 5), (3 << 14, 16), (1 << 17, 2)])
def test_batched_device_transforms_match_oracle(gpu, oc, n, k):
    """uzkge_cuda_ntt_fr_batch_device: k vectors of one domain in one launch per pass (the prover's rounds), ragged inputs, every
    variant, in place and out of place -- each vector equals the oracle's single transform."""
    import torch

    shift = oc.random_fr(1, 909)[0]
    xs = [oc.random_fr(n - 37 * j if j % 2 else n, 3000 + j + n % 11) for j in range(k)]
    d_in = [torch.zeros(4 * n, dtype=torch.int64, device="cuda") for _ in range(k)]
    d_out = [torch.zeros(4 * n, dtype=torch.int64, device="cuda") for _ in range(k)]
    scratch = torch.empty(4 * n * k, dtype=torch.int64, device="cuda")
    for inverse in (False, True):
        for cs in (None, shift):
            for j in range(k):
                d_in[j].zero_()
                d_in[j][: 4 * xs[j].shape[0]].copy_(torch.from_numpy(xs[j].view(np.int64).reshape(-1)))
            gpu.ntt_fr_batch_device([t.data_ptr() for t in d_in], [t.data_ptr() for t in d_out], scratch.data_ptr(),
                                    [x.shape[0] for x in xs], n, inverse, cs)
            torch.cuda.synchronize()
            for j in range(k):
                want = oc.ntt_fr(xs[j], n, inverse, cs)
                assert np.array_equal(d_out[j].cpu().numpy().view(np.uint64).reshape(n, 4), want), (n, k, j, inverse, cs is not None)
    # in place
    gpu.ntt_fr_batch_device([t.data_ptr() for t in d_in], [t.data_ptr() for t in d_in], scratch.data_ptr(), [n] * k, n)
    torch.cuda.synchronize()
    for j in range(k):
        full = np.zeros((n, 4), dtype=np.uint64)
        full[: xs[j].shape[0]] = xs[j]
        assert np.array_equal(d_in[j].cpu().numpy().view(np.uint64).reshape(n, 4), oc.ntt_fr(full, n))


def test_radix4_stage_walk_matches_the_oracle_at_every_size(gpu, oc):
    """The radix-4 walk of the tile passes (two butterfly stages per shared-memory round trip; default from 2^19 points on) forced onto
    every size (uzkge_cuda_configure("ntt_radix4", 1)): 2^1..2^18 and 3 * 2^1..3 * 2^16, forward / inverse / coset / coset inverse,
    against the oracle's transforms -- odd and even stage counts, 1-, 2- and 3-pass plans."""
    try:
        gpu.configure("ntt_radix4", 1)
        k = oc.random_fr(1, 31)[0]
        for n in [1 << l for l in range(1, 19)] + [3 << l for l in range(1, 17)]:
            x = oc.random_fr(n, 100 + n % 97)
            for inv, cs in ((False, None), (True, None), (False, k), (True, k)):
                assert np.array_equal(gpu.ntt_fr(x, n, inv, cs), oc.ntt_fr(x, n, inverse=inv, coset=cs)), (n, inv, cs is not None)
    finally:
        gpu.configure("ntt_radix4", 19)
