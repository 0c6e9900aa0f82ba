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
