"""CPU: pin the oracle (oracle/bn254.py big-int restatement and oracle/oracle.c) to the reference's own fixtures
(tests/golden, generated from /root/reference by tests/golden/make_golden.py) and to the executable NTT contract
`check_fft` (/root/reference/uzkge/src/poly_commit/field_polynomial.rs:632-646, 648-719)."""
import numpy as np
import pytest


def hx(s):
    return int(s, 16)


def test_domain_generators_match_solidity_verifier_keys(bn, oc, domain_kat):
    # VerifierKey_{20,52}.sol: `root` is domain.group_gen written by gen_params/solidity.rs:91-95
    for cards, e in domain_kat.items():
        n = e["cs_size"]
        w = bn.root_of_unity(n)
        assert w == hx(e["root"]), cards
        got = bn.array_to_ints(oc.fr_root_of_unity(n).reshape(1, 4), bn.FR)[0]
        assert got == w
        assert pow(w, n, bn.FR) == 1 and pow(w, n // 2, bn.FR) != 1


def test_public_input_locations_are_domain_elements(bn, domain_kat):
    # VerifierKeyExtra1: w^idx ; VerifierKeyExtra2: w^idx / n
    for cards, e in domain_kat.items():
        n = e["cs_size"]
        n_inv = bn.inv_mod(n, bn.FR)
        w = bn.root_of_unity(n)
        powers = {}
        acc = 1
        for i in range(n):
            powers[acc] = i
            acc = acc * w % bn.FR
        loc = [hx(v) for v in e["PI_POLY_INDICES_LOC"]]
        lag = [hx(v) for v in e["PI_POLY_LAGRANGE_LOC"]]
        assert len(loc) == len(lag) > 0
        for a, b in zip(loc, lag):
            assert a in powers
            assert b == a * n_inv % bn.FR


def test_mixed_domain_generator_relations(bn, oc):
    # w_{6n}^6 = w_n ; values cross-checked by the survey's probe
    assert bn.root_of_unity(98304) == hx("0x1b45e5e4772c0cf2342cd7a41541cf97777bdab72ccced8e415b2b12528f0eb1")
    for n in (4096, 8192, 16384):
        assert pow(bn.root_of_unity(6 * n), 6, bn.FR) == bn.root_of_unity(n)
        got = bn.array_to_ints(oc.fr_root_of_unity(6 * n).reshape(1, 4), bn.FR)[0]
        assert got == bn.root_of_unity(6 * n)


def test_srs_fixtures_are_on_curve(oc, lagrange_srs_4096, srs_padding_head, bn):
    for p in list(lagrange_srs_4096[:50]) + list(srs_padding_head):
        assert oc.g1_on_curve(p)
    g = bn.array_to_affine(srs_padding_head[:1])[0]
    assert g == bn.G1_GEN  # srs-padding.bin point 0 is the generator (1, 2)


@pytest.mark.parametrize("j", [0, 1, 2, 5, 63])
def test_msm_known_answer_lagrange_srs_4096(oc, bn, lagrange_srs_4096, srs_padding_head, j):
    """MSM(lagrange-srs-n, [w^(i*j)]_i) == srs-padding[j] = tau^j * G   (sum_i w^(ij) L_i(tau) = tau^j)."""
    n = 4096
    w = bn.root_of_unity(n)
    wj = pow(w, j, bn.FR)
    sc, acc = [], 1
    for _ in range(n):
        sc.append(acc)
        acc = acc * wj % bn.FR
    scalars = bn.ints_to_array(sc, bn.FR)
    got = oc.g1_to_affine(oc.msm_g1(lagrange_srs_4096, scalars))
    assert np.array_equal(got, srs_padding_head[j])


def test_msm_known_answer_lagrange_srs_16384(oc, bn, lagrange_srs_16384, srs_padding_head):
    n, j = 16384, 3
    w = bn.root_of_unity(n)
    wj = pow(w, j, bn.FR)
    sc, acc = [], 1
    for _ in range(n):
        sc.append(acc)
        acc = acc * wj % bn.FR
    got = oc.g1_to_affine(oc.msm_g1(lagrange_srs_16384, bn.ints_to_array(sc, bn.FR)))
    assert np.array_equal(got, srs_padding_head[j])


def test_c_msm_matches_bigint_naive(oc, bn):
    pts = oc.g1_random_points(40, 11)
    sc = oc.random_fr(40, 12)
    P = bn.array_to_affine(pts)
    s = bn.array_to_ints(sc, bn.FR)
    want = bn.msm_naive(P, s)
    got = bn.jac_array_to_affine(oc.msm_g1(pts, sc))
    assert got == want
    assert bn.msm_pippenger(P, s) == want


@pytest.mark.parametrize("n", [1, 2, 3])
def test_check_fft_contract_small(oc, bn, n):
    """check_fft: fft(poly)[i] == poly.eval(root^i), root = domain.group_gen (field_polynomial.rs:632-646)."""
    coefs = oc.random_fr(n, 20 + n)
    ev = oc.ntt_fr(coefs, n)
    w = oc.fr_root_of_unity(n)
    for i in range(n):
        assert np.array_equal(ev[i], oc.fr_eval(coefs, oc.fr_pow(w, i)))


@pytest.mark.parametrize("n", [16, 32, 3, 48, 96, 4096, 49152])
def test_fft_ifft_round_trip(oc, n):
    # test_fft: sizes 16, 32 (radix-2) and 3, 48 (mixed radix), field_polynomial.rs:648-719
    x = oc.random_fr(n, 30 + n % 13)
    assert np.array_equal(oc.ntt_fr(oc.ntt_fr(x, n), n, inverse=True), x)


@pytest.mark.parametrize("n", [8, 12, 24, 64, 192])
def test_c_ntt_matches_bigint_dft(oc, bn, n):
    x = oc.random_fr(n - 1, 40 + n)  # shorter than the domain: zero padding
    xi = bn.array_to_ints(x, bn.FR)
    want = bn.dft_naive(xi, n)
    assert bn.array_to_ints(oc.ntt_fr(x, n), bn.FR) == want
    k = 0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A  # golden k[1]
    kk = bn.ints_to_array([k], bn.FR)[0]
    assert bn.array_to_ints(oc.ntt_fr(x, n, coset=kk), bn.FR) == bn.coset_fft(xi, n, k)
    kinv = bn.inv_mod(k, bn.FR)
    ev = bn.coset_fft(xi, n, k)
    back = oc.ntt_fr(bn.ints_to_array(ev, bn.FR), n, inverse=True, coset=bn.ints_to_array([kinv], bn.FR)[0])
    assert bn.trim(bn.array_to_ints(back, bn.FR)) == bn.trim(xi)


def test_coset_shift_is_golden_k1(domain_kat, bn):
    # k[1] of choose_ks (indexer.rs:211-235) as written to both verifier keys: a quadratic non-residue
    k1 = hx(domain_kat["52"]["k"][1])
    assert k1 == hx(domain_kat["20"]["k"][1])
    assert pow(k1, (bn.FR - 1) // 2, bn.FR) == bn.FR - 1
