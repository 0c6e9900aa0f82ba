"""GPU: the quotient polynomial's pointwise map (uzkge_cuda_plonk_quotient_fr_device) against the big-integer restatement of
t_poly's loop body (oracle/plonk.py, /root/reference/uzkge/src/plonk/helpers.rs:284-669), on random coset evaluations, and the
assembled pipeline  coset FFT -> map -> coset iFFT  on a satisfied toy circuit (divisibility by the vanishing polynomial)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K = [1,
     0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A,
     0x2042A587A90C187B0A087C03E29C968B950B1DB26D5C82D666905A6895790C0A,
     0x2DB4944E13E6E33CF0EF0734796FF332D73B5FA160DCA733BF529E9B758E4960,
     0x1D9E3A4AAF01052D9925138DC6D7D05AA614E311040142458B045D0053D22F46]  # golden k[0..5] (tests/golden/domain_kat.json)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda()


@pytest.mark.parametrize("n,factor", [(8, 16), (64, 6), (1024, 6)])
def test_quotient_map_matches_restatement(gpu, oc, bn, n, factor):
    from oracle import plonk

    m = n * factor
    rnd = lambda seed: oc.random_fr(m, seed)
    names = [("w", 5), ("q", 9), ("s", 5), ("q_prk", 4)]
    arrays, seed = {}, 100
    for name, cnt in names:
        arrays[name] = [rnd(seed + i) for i in range(cnt)]
        seed += cnt
    for name in ("pi", "z", "coset_quotient", "l1", "qb"):
        arrays[name] = rnd(seed)
        seed += 1
    sc = oc.random_fr(5, 999)
    alpha, beta, gamma, g = sc[0], sc[1], sc[2], sc[3]
    g_inv = oc.fr_inv(g)
    zh = oc.random_fr(factor, 77)
    k = bn.ints_to_array(K, bn.FR)
    keep = {nm: ([dev(a) for a in v] if isinstance(v, list) else dev(v)) for nm, v in arrays.items()}
    out = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    ptr = lambda t: t.data_ptr()
    gpu.plonk_quotient_fr_device(
        [ptr(t) for t in keep["w"]], [ptr(t) for t in keep["q"]], ptr(keep["pi"]), ptr(keep["z"]), [ptr(t) for t in keep["s"]],
        ptr(keep["coset_quotient"]), ptr(keep["l1"]), ptr(keep["qb"]), [ptr(t) for t in keep["q_prk"]], k, alpha, beta, gamma, g, g_inv,
        zh, m, factor, out.data_ptr())
    torch.cuda.synchronize()
    got = bn.array_to_ints(out.cpu().numpy().view(np.uint64).reshape(m, 4), bn.FR)
    I = lambda a: bn.array_to_ints(a, bn.FR)
    want = plonk.quotient_coset_evals(
        [I(a) for a in arrays["w"]], [I(a) for a in arrays["q"]], I(arrays["pi"]), I(arrays["z"]), [I(a) for a in arrays["s"]],
        I(arrays["coset_quotient"]), I(arrays["l1"]), I(arrays["qb"]), [I(a) for a in arrays["q_prk"]], K,
        I(alpha.reshape(1, 4))[0], I(beta.reshape(1, 4))[0], I(gamma.reshape(1, 4))[0], I(g.reshape(1, 4))[0], I(g_inv.reshape(1, 4))[0],
        I(zh), factor)
    assert got == want


def test_quotient_pipeline_divisibility(gpu, oc, bn):
    """A satisfied circuit of addition / multiplication gates (identity permutation, z = 1): the numerator vanishes on the
    n-th roots of unity, so  coset_ifft(map(coset_fft(...)))  is a polynomial of degree < 3n (two multiplied degree-n wire
    polynomials and a selector) -- the check the verifier relies on (t(X) Z_H(X) = numerator)."""
    from oracle import plonk
    from uzkge_b200 import FpPolynomial

    n, factor = 64, 6
    m = n * factor
    F = bn.FR
    rng = np.random.default_rng(3)
    dom_n = FpPolynomial.evaluation_domain(n)
    dom_m = FpPolynomial.quotient_evaluation_domain(m)
    k1 = bn.ints_to_array([K[1]], F)[0]
    k1_inv = bn.ints_to_array([bn.inv_mod(K[1], F)], F)[0]
    # gates: even rows  w0 + w1 - w4 = 0 (q0 = q1 = 1, q8 = 1); odd rows  w0 * w1 - w4 = 0 (q4 = 1, q8 = 1)
    w = [[int(x) for x in rng.integers(1, 1 << 60, size=n)] for _ in range(5)]
    q = [[0] * n for _ in range(9)]
    for i in range(n):
        if i % 2 == 0:
            q[0][i] = q[1][i] = 1
            w[4][i] = (w[0][i] + w[1][i]) % F
        else:
            q[4][i] = 1
            w[4][i] = w[0][i] * w[1][i] % F
        q[8][i] = 1
    omega_n = bn.root_of_unity(n)
    group = [pow(omega_n, i, F) for i in range(n)]

    def coset_evals(values):   # evaluations on H -> polynomial -> evaluations on the coset k1 * <w_m>
        poly = FpPolynomial.ifft_with_domain(dom_n, bn.ints_to_array(values, F))
        return poly.coset_fft_with_domain(dom_m, k1)

    arrays = {
        "w": [coset_evals(w[j]) for j in range(5)],
        "q": [coset_evals(q[j]) for j in range(9)],
        "s": [coset_evals([K[j] * group[i] % F for i in range(n)]) for j in range(5)],   # identity permutation: s_j = k_j X
        "q_prk": [np.zeros((m, 4), dtype=np.uint64) for _ in range(4)],
        "pi": np.zeros((m, 4), dtype=np.uint64),
        "z": coset_evals([1] * n),
        "l1": coset_evals([1] + [0] * (n - 1)),
        "qb": np.zeros((m, 4), dtype=np.uint64),
    }
    omega_m = bn.root_of_unity(m)
    arrays["coset_quotient"] = bn.ints_to_array([K[1] * pow(omega_m, i, F) % F for i in range(m)], F)
    zh = bn.ints_to_array(plonk.z_h_inv_coset_evals(K[1], omega_m, n, factor), F)
    sc = oc.random_fr(4, 5)
    g = sc[3]
    keep = {nm: ([dev(a) for a in v] if isinstance(v, list) else dev(v)) for nm, v in arrays.items()}
    out = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    ptr = lambda t: t.data_ptr()
    gpu.plonk_quotient_fr_device(
        [ptr(t) for t in keep["w"]], [ptr(t) for t in keep["q"]], ptr(keep["pi"]), ptr(keep["z"]), [ptr(t) for t in keep["s"]],
        ptr(keep["coset_quotient"]), ptr(keep["l1"]), ptr(keep["qb"]), [ptr(t) for t in keep["q_prk"]], bn.ints_to_array(K, F),
        sc[0], sc[1], sc[2], g, oc.fr_inv(g), zh, m, factor, out.data_ptr())
    torch.cuda.synchronize()
    t_evals = out.cpu().numpy().view(np.uint64).reshape(m, 4)
    t_poly = FpPolynomial.coset_ifft_with_domain(dom_m, t_evals, k1_inv)
    assert t_poly.degree() < 3 * n, t_poly.degree()
    # and it is not trivially zero: break one gate and the quotient stops being a low-degree polynomial
    w[4][5] = (w[4][5] + 1) % F
    bad = dev(coset_evals(w[4]))
    gpu.plonk_quotient_fr_device(
        [ptr(t) for t in keep["w"][:4]] + [bad.data_ptr()], [ptr(t) for t in keep["q"]], ptr(keep["pi"]), ptr(keep["z"]),
        [ptr(t) for t in keep["s"]], ptr(keep["coset_quotient"]), ptr(keep["l1"]), ptr(keep["qb"]), [ptr(t) for t in keep["q_prk"]],
        bn.ints_to_array(K, F), sc[0], sc[1], sc[2], g, oc.fr_inv(g), zh, m, factor, out.data_ptr())
    torch.cuda.synchronize()
    t_bad = FpPolynomial.coset_ifft_with_domain(dom_m, out.cpu().numpy().view(np.uint64).reshape(m, 4), k1_inv)
    assert t_bad.degree() >= 5 * n


@pytest.mark.parametrize("n,factor", [(8, 16), (64, 6), (2048, 6)])
def test_quotient_map_coset_by_coset(gpu, oc, bn, n, factor):
    """uzkge_cuda_plonk_quotient_range_fr_device on (start, step, count) = (j, factor, n) -- the coset g_j <w_n> of the quotient domain,
    the unit by which a device group splits the round: the `factor` coset launches together write exactly what the whole-domain launch
    writes, and each touches only its own points."""
    m = n * factor
    arrays, seed = {}, 300
    for name, cnt in (("w", 5), ("q", 9), ("s", 5), ("q_prk", 4)):
        arrays[name] = [dev(oc.random_fr(m, seed + i)) for i in range(cnt)]
        seed += cnt
    for name in ("pi", "z", "coset_quotient", "l1", "qb"):
        arrays[name] = dev(oc.random_fr(m, seed))
        seed += 1
    sc = oc.random_fr(5, 998)
    zh = oc.random_fr(factor, 78)
    k = bn.ints_to_array(K, bn.FR)
    ptr = lambda t: t.data_ptr()

    def run(out, point_range=None):
        gpu.plonk_quotient_fr_device(
            [ptr(t) for t in arrays["w"]], [ptr(t) for t in arrays["q"]], ptr(arrays["pi"]), ptr(arrays["z"]), [ptr(t) for t in arrays["s"]],
            ptr(arrays["coset_quotient"]), ptr(arrays["l1"]), ptr(arrays["qb"]), [ptr(t) for t in arrays["q_prk"]], k, sc[0], sc[1], sc[2],
            sc[3], oc.fr_inv(sc[3]), zh, m, factor, out.data_ptr(), point_range=point_range)

    whole = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    run(whole)
    parts = torch.full((4 * m,), -1, dtype=torch.int64, device="cuda")
    for j in range(factor):
        before = parts.clone()
        run(parts, (j, factor, n))
        changed = (parts.view(m, 4) != before.view(m, 4)).any(dim=1).nonzero().flatten()
        assert bool(((changed % factor) == j).all())                 # only coset j's points were written
    assert torch.equal(parts, whole)
    with pytest.raises(Exception):
        run(parts, (1, factor, n + 1))                               # runs past the domain


@pytest.mark.parametrize("log_n,factor", [(3, 16), (6, 6), (12, 6)])
def test_coefficients_from_per_coset_inverse_transforms(gpu, oc, bn, log_n, factor):
    """uzkge_cuda_plonk_coset_combine_fr_device: a polynomial t of factor * n coefficients, evaluated on the coset k1 <w_m> (the oracle's
    coset FFT), is recovered from its values coset by coset: strided copy of coset j, size-n coset iFFT with shift g_j^-1
    (g_j = k1 w_m^j), and the factor-point inverse DFT over the cosets -- what coset_ifft_with_domain over the whole domain returns."""
    n = 1 << log_n
    m = n * factor
    t = oc.random_fr(m, 555 + log_n)
    k1 = bn.ints_to_array([K[1]], bn.FR)[0]
    evals = dev(oc.ntt_fr(t, m, coset=k1))
    w_m = bn.root_of_unity(m)
    u = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    tmp, scr = torch.empty(4 * n, dtype=torch.int64, device="cuda"), torch.empty(4 * n, dtype=torch.int64, device="cuda")
    for j in range(factor):
        g_inv = bn.ints_to_array([pow(K[1] * pow(w_m, j, bn.FR) % bn.FR, -1, bn.FR)], bn.FR)[0]
        gpu.fr_strided_copy_device(evals.data_ptr(), j, factor, tmp.data_ptr(), 0, 1, n)
        gpu.ntt_fr_device(tmp.data_ptr(), u.data_ptr() + 32 * j * n, scr.data_ptr(), n, n, True, g_inv)
    out = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    gpu.plonk_coset_combine_fr_device(u.data_ptr(), n, factor, k1, out.data_ptr())
    assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(m, 4), t)
