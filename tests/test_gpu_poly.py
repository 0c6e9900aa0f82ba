"""GPU: the polynomial glue kernels (uzkge_cuda_poly_eval_fr, _poly_div_linear_fr, _grand_product_fr) through the C ABI,
bit-exact against the oracle; the reference's doc-tests for FpPolynomial::eval / div_rem
(/root/reference/uzkge/src/poly_commit/field_polynomial.rs:62-85, 198-209, 500-550) restated on the host mirror."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fr(bn, vals):
    return bn.ints_to_array([v % bn.FR for v in vals], bn.FR)


def ints(bn, a):
    return bn.array_to_ints(np.asarray(a).reshape(-1, 4), bn.FR)


def test_eval_doc_test(gpu, bn):
    # from_coefs doc-test: poly = 1 + X^2: eval(0) = 1, eval(1) = 2, eval(2) = 5; trailing zeros do not matter
    from uzkge_b200 import FpPolynomial

    poly = FpPolynomial.from_coefs(fr(bn, [1, 0, 1]))
    assert poly.degree() == 2
    for x, want in ((0, 1), (1, 2), (2, 5)):
        assert ints(bn, poly.eval(fr(bn, [x])[0])) == [want]
    poly2 = FpPolynomial.from_coefs(fr(bn, [1, 0, 1, 0, 0, 0]))
    assert poly2.degree() == 2 and poly == poly2


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 255, 1023, 1024, 1025, 4097, 100000, (1 << 20) + 7])
def test_eval_matches_oracle(gpu, oc, n):
    c = oc.random_fr(n, 500 + n % 97)
    for seed in (1, 2):
        x = oc.random_fr(1, seed)[0]
        assert np.array_equal(gpu.poly_eval_fr(c, x), oc.fr_eval(c, x))
    zero = np.zeros(4, dtype=np.uint64)
    assert np.array_equal(gpu.poly_eval_fr(c, zero), c[0])


def test_div_rem_doc_test(gpu, bn):
    # div_rem doc-test: (1 + X + X^2) / (1 + X) = X remainder 1   (divisor X - z with z = -1)
    from uzkge_b200 import FpPolynomial

    poly = FpPolynomial.from_coefs(fr(bn, [1, 1, 1]))
    q, r = poly.div_rem_linear(fr(bn, [-1])[0])
    assert q == FpPolynomial.from_coefs(fr(bn, [0, 1]))
    assert r == FpPolynomial.from_coefs(fr(bn, [1]))
    # degree-0 dividend: quotient zero, remainder the polynomial itself (l > k branch)
    q0, r0 = FpPolynomial.from_coefs(fr(bn, [7])).div_rem_linear(fr(bn, [3])[0])
    assert q0.is_zero() and ints(bn, r0.coefs) == [7]


@pytest.mark.parametrize("n", [2, 3, 5, 64, 1024, 1025, 2049, 5000])
def test_div_linear_matches_bigint_ruffini(gpu, oc, bn, n):
    c = oc.random_fr(n, 600 + n)
    z = oc.random_fr(1, 7)[0]
    q, rem = gpu.poly_div_linear_fr(c, z)
    ci, zi = ints(bn, c), ints(bn, z)[0]
    want = [0] * (n - 1)
    acc = 0
    for k in range(n - 1, 0, -1):  # q_{k-1} = c_k + z q_k
        acc = (ci[k] + zi * acc) % bn.FR
        want[k - 1] = acc
    assert ints(bn, q) == want
    assert ints(bn, rem) == [(ci[0] + zi * acc) % bn.FR]


@pytest.mark.parametrize("n", [1 << 16, (1 << 22) + 3])
def test_div_linear_full_size_identity(gpu, oc, n):
    """p(X) = q(X) (X - z) + r checked at random points with the oracle's Horner evaluation; r = p(z)."""
    c = oc.random_fr(n, 11)
    z = oc.random_fr(1, 12)[0]
    q, rem = gpu.poly_div_linear_fr(c, z)
    assert np.array_equal(rem, oc.fr_eval(c, z))
    for seed in (21, 22):
        x = oc.random_fr(1, seed)[0]
        lhs = oc.fr_eval(c, x)
        qx = oc.fr_eval(q, x)
        # (x - z) as a field element: x + (r - z)
        import numpy as _np
        from oracle import bn254 as b

        xm = b.ints_to_array([(b.array_to_ints(x.reshape(1, 4), b.FR)[0] - b.array_to_ints(z.reshape(1, 4), b.FR)[0]) % b.FR], b.FR)
        prod = oc.fr_mul(qx.reshape(1, 4), xm)[0]
        rhs = b.ints_to_array([(b.array_to_ints(prod.reshape(1, 4), b.FR)[0] + b.array_to_ints(rem.reshape(1, 4), b.FR)[0]) % b.FR], b.FR)[0]
        assert _np.array_equal(lhs, rhs)


@pytest.mark.parametrize("n", [1, 2, 7, 1023, 1024, 1025, 16383, 100000])
def test_grand_product_matches_reference_loop(gpu, oc, bn, n):
    """z_poly's loop (plonk/helpers.rs:204-217): batch_inversion(denominators); prev *= num * den^-1."""
    num = oc.random_fr(n, 700 + n % 89)
    den = oc.random_fr(n, 800 + n % 83)
    got = gpu.grand_product_fr(num, den)
    assert got.shape == (n + 1, 4)
    ni, di = ints(bn, num), ints(bn, den)
    if n <= 20000:
        want, prev = [1], 1
        for a, b in zip(ni, di):
            prev = prev * a % bn.FR * bn.inv_mod(b, bn.FR) % bn.FR
            want.append(prev)
        assert ints(bn, got) == want
    else:
        g = ints(bn, got)
        assert g[0] == 1
        for i in list(range(0, n, n // 50)) + [n - 1]:
            assert g[i + 1] * di[i] % bn.FR == g[i] * ni[i] % bn.FR


def test_grand_product_edge_cases(gpu, oc, bn):
    from uzkge_b200.errors import UzkgeError

    assert ints(bn, gpu.grand_product_fr(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))) == [1]
    num = oc.random_fr(10, 1)
    den = oc.random_fr(10, 2)
    den[4] = 0
    with pytest.raises(UzkgeError):
        gpu.grand_product_fr(num, den)
    # num == den: z stays 1 (the permutation argument's invariant z(1) = 1 ... prod = 1, helpers.rs:1441-1473)
    same = gpu.grand_product_fr(num, num)
    assert ints(bn, same) == [1] * 11


def test_eval_batch_matches_oracle(gpu, oc):
    """uzkge_cuda_poly_eval_batch_fr_device: ragged polynomials (one coefficient ... several hundred tiles), two points, k = 1 and
    k = 32, against the oracle's Horner evaluation."""
    import torch

    lens = [1, 2, 5, 1023, 1024, 1025, 4099, 16387, 300007, (1 << 18) + 3]
    lens = (lens * 4)[:32]
    polys = [oc.random_fr(n, 700 + i) for i, n in enumerate(lens)]
    pts = oc.random_fr(2, 9)
    d = [torch.from_numpy(p.view(np.int64).reshape(-1)).cuda() for p in polys]
    for k in (1, 7, 32):
        idx = [(j * 5 + 1) % 2 for j in range(k)]
        vals = torch.zeros(4 * k, dtype=torch.int64, device="cuda")
        gpu.poly_eval_batch_fr_device([t.data_ptr() for t in d[:k]], lens[:k], idx, pts, vals.data_ptr())
        torch.cuda.synchronize()
        got = vals.cpu().numpy().view(np.uint64).reshape(k, 4)
        for j in range(k):
            assert np.array_equal(got[j], oc.fr_eval(polys[j], pts[idx[j]])), (k, j, lens[j])
    # a single point
    vals = torch.zeros(8, dtype=torch.int64, device="cuda")
    gpu.poly_eval_batch_fr_device([d[3].data_ptr(), d[8].data_ptr()], [lens[3], lens[8]], [0, 0], pts[:1], vals.data_ptr())
    got = vals.cpu().numpy().view(np.uint64).reshape(2, 4)
    assert np.array_equal(got[0], oc.fr_eval(polys[3], pts[0])) and np.array_equal(got[1], oc.fr_eval(polys[8], pts[0]))


# ---------------------------------------------------------------- the prover's elementwise glue (csrc/plonk_glue.cu), kernel by kernel
def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda()


def _host(t, n):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)[:n]


def test_fr_lincomb_matches_bigint(gpu, oc, bn):
    """out = sum_j c_j p_j with ragged lengths (r_poly_or_comm, batch_prove), including aliasing of the output with an input."""
    import torch

    lens = [1, 5, 1000, 1027, 4096, 3]
    polys = [oc.random_fr(n, 800 + i) for i, n in enumerate(lens)]
    coefs = oc.random_fr(len(lens), 9)
    out_len = 4200
    d = [_dev(p) for p in polys]
    out = torch.zeros(4 * out_len, dtype=torch.int64, device="cuda")
    gpu.fr_lincomb_device([t.data_ptr() for t in d], lens, coefs, out.data_ptr(), out_len)
    pi, ci = [ints(bn, p) for p in polys], ints(bn, coefs)
    want = [sum(ci[j] * pi[j][i] for j in range(len(lens)) if i < lens[j]) % bn.FR for i in range(out_len)]
    assert ints(bn, _host(out, out_len)) == want
    # in place: p_4 <- 2 p_4 + p_2
    two_one = fr(bn, [2, 1])
    gpu.fr_lincomb_device([d[4].data_ptr(), d[2].data_ptr()], [4096, 1000], two_one, d[4].data_ptr(), 4096)
    want = [(2 * pi[4][i] + (pi[2][i] if i < 1000 else 0)) % bn.FR for i in range(4096)]
    assert ints(bn, _host(d[4], 4096)) == want


def test_fr_sparse_add_powers_gather_mul_trim(gpu, oc, bn):
    import torch

    n = 5000
    a = oc.random_fr(n, 31)
    ai = ints(bn, a)
    # add_coef_assign on a few coefficients, a repeated index included (applied in order)
    t = _dev(a)
    gpu.fr_add_sparse_device(t.data_ptr(), [0, 4999, 7, 7], fr(bn, [5, -1, 11, 13]))
    got = ints(bn, _host(t, n))
    want = list(ai)
    want[0] = (want[0] + 5) % bn.FR
    want[4999] = (want[4999] - 1) % bn.FR
    want[7] = (want[7] + 24) % bn.FR
    assert got == want
    # powers: scale * base^i (domain.elements(), coset_quotient)
    base, scale = oc.random_fr(2, 5)
    out = torch.zeros(4 * n, dtype=torch.int64, device="cuda")
    gpu.fr_powers_device(base, n, out.data_ptr(), scale=scale)
    b, s = ints(bn, base)[0], ints(bn, scale)[0]
    assert ints(bn, _host(out, n)) == [s * pow(b, i, bn.FR) % bn.FR for i in range(n)]
    gpu.fr_powers_device(base, 7, out.data_ptr())
    assert ints(bn, _host(out, 7)) == [pow(b, i, bn.FR) for i in range(7)]
    # gather (extend_witness) and pointwise product
    idx = np.random.default_rng(1).integers(0, n, 3 * n).astype(np.uint32)
    d_idx = torch.from_numpy(idx.view(np.int32)).cuda()
    g = torch.zeros(4 * 3 * n, dtype=torch.int64, device="cuda")
    da = _dev(a)                 # keep the tensors alive: a temporary's memory is recycled as soon as its pointer is taken
    gpu.fr_gather_device(da.data_ptr(), d_idx.data_ptr(), 3 * n, g.data_ptr())
    assert np.array_equal(_host(g, 3 * n), a[idx])
    bvec = oc.random_fr(n, 32)
    db = _dev(bvec)
    prod = torch.zeros(4 * n, dtype=torch.int64, device="cuda")
    gpu.fr_mul_device(da.data_ptr(), db.data_ptr(), n, prod.data_ptr())
    assert np.array_equal(_host(prod, n), oc.fr_mul(a, bvec))
    # trailing-zero trim
    z = a.copy()
    z[3000:] = 0
    dz = _dev(z)
    assert gpu.fr_trimmed_len_device(dz.data_ptr(), n) == 3000
    z[:] = 0
    dz = _dev(z)
    assert gpu.fr_trimmed_len_device(dz.data_ptr(), n) == 0
    z[n - 1, 2] = 1
    dz = _dev(z)
    assert gpu.fr_trimmed_len_device(dz.data_ptr(), n) == n


@pytest.mark.parametrize("n", [2, 8, 1000, 4097])
def test_z_evals_match_restated_z_poly(gpu, oc, bn, n):
    """uzkge_cuda_plonk_z_evals_fr_device against z_poly's loop (plonk/helpers.rs:160-220) in big integers: random wires, a random
    encoded permutation, device-resident grand product."""
    import torch

    F = bn.FR
    w = [oc.random_fr(n, 40 + j) for j in range(5)]
    sig = [oc.random_fr(n, 50 + j) for j in range(5)]
    group = oc.random_fr(n, 60)
    k = oc.random_fr(5, 61)
    beta, gamma = oc.random_fr(2, 62)
    dw, ds, dg = [_dev(x) for x in w], [_dev(x) for x in sig], _dev(group)
    z = torch.zeros(4 * n, dtype=torch.int64, device="cuda")
    tmp = torch.zeros(4 * 4 * n, dtype=torch.int64, device="cuda")
    gpu.plonk_z_evals_fr_device([t.data_ptr() for t in dw], [t.data_ptr() for t in ds], dg.data_ptr(), k, beta, gamma, n, z.data_ptr(), tmp.data_ptr())
    wi, si, gi, ki = [ints(bn, x) for x in w], [ints(bn, x) for x in sig], ints(bn, group), ints(bn, k)
    b, g_ = ints(bn, beta)[0], ints(bn, gamma)[0]
    want, prev = [1], 1
    for i in range(n - 1):
        num = den = 1
        for j in range(5):
            num = num * (wi[j][i] + g_ + b * ki[j] * gi[i]) % F
            den = den * (wi[j][i] + g_ + b * si[j][i]) % F
        prev = prev * num % F * pow(den, -1, F) % F
        want.append(prev)
    assert ints(bn, _host(z, n)) == want
