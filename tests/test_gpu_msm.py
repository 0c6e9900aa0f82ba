"""GPU: BN254 G1 MSM through the C ABI (uzkge_cuda_srs_upload / uzkge_cuda_msm_g1*) and the KZG commit mirror,
bit-exact (affine coordinates) against the oracle and the reference's fixtures.  Reference tests mirrored:
test_commit, test_homomorphic_poly_com_elem (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:483-548)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same_point(oc, a_jac, b_jac):
    return np.array_equal(oc.g1_to_affine(a_jac), oc.g1_to_affine(b_jac))


def fr(bn, vals):
    return bn.ints_to_array([v % bn.FR for v in vals], bn.FR)


def witness_like(oc, bn, n, seed):
    """50 % zero, 30 % in {0, 1}, 10 % < 2^16, 10 % uniform."""
    rng = np.random.default_rng(seed)
    s = oc.random_fr(n, seed)
    u = rng.random(n)
    small = fr(bn, [int(v) for v in rng.integers(0, 1 << 16, size=n)])
    bit = fr(bn, [int(v) for v in rng.integers(0, 2, size=n)])
    s[u < 0.5] = 0
    m = (u >= 0.5) & (u < 0.8)
    s[m] = bit[m]
    m = (u >= 0.8) & (u < 0.9)
    s[m] = small[m]
    return s


@pytest.mark.parametrize("n,c", [(1, 0), (2, 0), (3, 2), (31, 3), (32, 0), (100, 5), (1000, 0), (1000, 11), (5000, 14), (16384, 0)])
def test_msm_matches_oracle(gpu, oc, n, c):
    pts = oc.g1_random_points(n, 100 + n)
    sc = oc.random_fr(n, 200 + n)
    h = gpu.srs_upload(pts, c)
    try:
        for lanes in (0, 1, 4, 32):
            gpu.configure("msm_lanes", lanes)
            assert same_point(oc, gpu.msm_g1(h, sc), oc.msm_g1(pts, sc)), (n, c, lanes)
        gpu.configure("msm_lanes", 0)
        # prefix and offset ranges (commit uses a prefix of the SRS; blinds use an offset)
        m = max(1, n // 3)
        assert same_point(oc, gpu.msm_g1(h, sc[:m]), oc.msm_g1(pts[:m], sc[:m]))
        assert same_point(oc, gpu.msm_g1(h, sc[:m], base_offset=n - m), oc.msm_g1(pts[n - m :], sc[:m]))
    finally:
        gpu.configure("msm_lanes", 0)
        gpu.srs_free(h)


@pytest.mark.parametrize("n", [1 << 16, 1 << 18])
def test_msm_medium_matches_oracle(gpu, oc, bn, n):
    pts = oc.g1_random_points(n, 7)
    h = gpu.srs_upload(pts)
    try:
        for sc in (oc.random_fr(n, 8), witness_like(oc, bn, n, 9)):
            assert same_point(oc, gpu.msm_g1(h, sc), oc.msm_g1(pts, sc))
    finally:
        gpu.srs_free(h)


def test_msm_edge_cases(gpu, oc, bn):
    n = 2048
    pts = oc.g1_random_points(n, 31)
    pts[5] = 0                # identity bases are skipped (padded SRS, gen_params/mod.rs:160-161)
    pts[100:110] = 0
    pts[201] = pts[200]       # duplicated point
    neg = pts[300].copy()
    negy = bn.array_to_ints(neg[4:].reshape(1, 4), bn.FQ)[0]
    neg[4:] = bn.ints_to_array([bn.FQ - negy], bn.FQ)[0]
    pts[301] = neg            # P and -P
    h = gpu.srs_upload(pts)
    try:
        cases = {
            "zeros": np.zeros((n, 4), dtype=np.uint64),
            "ones": np.repeat(fr(bn, [1]), n, axis=0),
            "minus_one": np.repeat(fr(bn, [-1]), n, axis=0),
            "half": np.repeat(fr(bn, [(bn.FR - 1) // 2]), n, axis=0),
            "pow2": fr(bn, [1 << (i % 254) for i in range(n)]),
            "window_edges": fr(bn, [(1 << 11) * (i % 7) + (1 << 10) - 1 + (i % 3) for i in range(n)]),
            "all_digits_max": np.repeat(fr(bn, [int("1" * 253, 2)]), n, axis=0),
        }
        single = np.zeros((n, 4), dtype=np.uint64)
        single[777] = oc.random_fr(1, 4)[0]
        cases["single"] = single
        cancel = np.zeros((n, 4), dtype=np.uint64)
        cancel[300] = cancel[301] = oc.random_fr(1, 5)[0]
        cases["cancel"] = cancel
        for name, sc in cases.items():
            got = gpu.msm_g1(h, sc)
            assert same_point(oc, got, oc.msm_g1(pts, sc)), name
        ident = gpu.msm_g1(h, cases["cancel"])
        assert not ident[8:].any()                      # Z == 0
        assert not gpu.g1_to_affine(ident).any()        # affine identity encodes as zeros
        assert not gpu.msm_g1(h, np.zeros((0, 4), dtype=np.uint64))[8:].any()   # n == 0
    finally:
        gpu.srs_free(h)


def test_msm_errors(gpu, oc):
    from uzkge_b200.errors import CommitmentError, ParameterError, UzkgeError

    pts = oc.g1_random_points(8, 1)
    h = gpu.srs_upload(pts)
    sc = oc.random_fr(9, 2)
    with pytest.raises(UzkgeError):
        gpu.msm_g1(h, sc)                    # more scalars than points
    with pytest.raises(UzkgeError):
        gpu.msm_g1(h, sc[:4], base_offset=6)
    gpu.srs_free(h)
    with pytest.raises(UzkgeError):
        gpu.msm_g1(h, sc[:4])                # stale handle
    with pytest.raises(UzkgeError):
        gpu.srs_free(h)
    with pytest.raises((CommitmentError, ParameterError)):
        gpu.srs_upload(np.zeros((0, 8), dtype=np.uint64))


@pytest.mark.parametrize("n,fixture", [(4096, "lagrange_srs_4096"), (16384, "lagrange_srs_16384")])
def test_lagrange_srs_known_answers(gpu, oc, bn, srs_padding_head, request, n, fixture):
    """The reference's bundled parameters: MSM(lagrange-srs-n, [w_n^(i*j)]_i) == srs-padding[j] = tau^j G."""
    srs = request.getfixturevalue(fixture)
    h = gpu.srs_upload(srs)
    try:
        w = bn.root_of_unity(n)
        vecs = []
        js = [0, 1, 2, 5, 40, 63]
        for j in js:
            wj, acc, sc = pow(w, j, bn.FR), 1, []
            for _ in range(n):
                sc.append(acc)
                acc = acc * wj % bn.FR
            vecs.append(bn.ints_to_array(sc, bn.FR))
        outs = gpu.msm_g1_batch(h, vecs)
        for j, o in zip(js, outs):
            assert np.array_equal(gpu.g1_to_affine(o), srs_padding_head[j]), j
            assert np.array_equal(oc.g1_to_affine(o), srs_padding_head[j]), j
        assert same_point(oc, gpu.msm_g1(h, vecs[3]), outs[3])
    finally:
        gpu.srs_free(h)


def test_batched_msm_matches_one_by_one(gpu, oc, bn):
    """uzkge_cuda_msm_g1_batch: independent MSMs of one prover round in one pass (ragged lengths, an empty and an
    all-zero vector, skewed scalars, more vectors than batch slots)."""
    n = 3000
    pts = oc.g1_random_points(n, 77)
    h = gpu.srs_upload(pts)
    try:
        slots = gpu.srs_info(h)["batch_slots"]
        assert slots >= 1
        vecs = [oc.random_fr(n, 300 + j)[: max(1, (n * (j + 1)) // 21)] for j in range(19)]
        vecs[3] = np.zeros((0, 4), dtype=np.uint64)
        vecs[5] = np.zeros((100, 4), dtype=np.uint64)
        vecs[7] = witness_like(oc, bn, n, 5)
        outs = gpu.msm_g1_batch(h, vecs)
        for j, v in enumerate(vecs):
            want = oc.msm_g1(pts[: v.shape[0]], v) if v.shape[0] else np.zeros(12, dtype=np.uint64)
            assert same_point(oc, outs[j], want), j
    finally:
        gpu.srs_free(h)


def test_pipelined_batch_matches_one_by_one(gpu, oc, bn):
    """More MSMs than one pass carries: groups alternate between the SRS's two workspaces on the engine's internal streams
    (MsmEngine::run_pipelined).  Host-pointer and device-pointer entry points, ragged lengths, empty groups, repeated calls."""
    import torch

    n = 2500
    pts = oc.g1_random_points(n, 78)
    h = gpu.srs_upload(pts)
    try:
        slots = gpu.srs_info(h)["batch_slots"]
        k = 3 * slots + 5                       # four groups: both workspaces are reused
        vecs = [oc.random_fr(n, 900 + j)[: 1 + (j * 131) % n] for j in range(k)]
        vecs[1] = np.zeros((0, 4), dtype=np.uint64)
        vecs[slots + 2] = witness_like(oc, bn, n, 6)
        for j in range(2 * slots, 3 * slots):   # one whole group of all-zero scalars
            vecs[j] = np.zeros((7, 4), dtype=np.uint64)
        want = [oc.msm_g1(pts[: v.shape[0]], v) if v.shape[0] else np.zeros(12, dtype=np.uint64) for v in vecs]
        for _ in range(2):
            outs = gpu.msm_g1_batch(h, vecs)
            for j in range(k):
                assert same_point(oc, outs[j], want[j]), j
        d_vecs = [torch.from_numpy(np.ascontiguousarray(v).view(np.int64).reshape(-1)).cuda() if v.shape[0] else torch.zeros(4, dtype=torch.int64).cuda()
                  for v in vecs]
        d_out = torch.zeros(12 * k, dtype=torch.int64, device="cuda")
        for _ in range(2):
            d_out.zero_()
            gpu.msm_g1_batch_device(h, [t.data_ptr() for t in d_vecs], [v.shape[0] for v in vecs], d_out.data_ptr())
            torch.cuda.synchronize()
            got = d_out.cpu().numpy().view(np.uint64).reshape(k, 12)
            for j in range(k):
                assert same_point(oc, got[j], want[j]), j
    finally:
        gpu.srs_free(h)


def test_batched_affine_accumulation_matches_oracle(gpu, oc, bn):
    """The pairwise affine tree with shared inversions (csrc/msm_affine.cu; off by default, `msm_affine` = 2 forces it at every
    size): uniform and skewed scalars, duplicated points (P + P inside a bucket), P and -P (cancellation to the identity),
    identity bases, and a size where the tree runs several rounds."""
    try:
        gpu.configure("msm_affine", 2)
        for n, seed in ((64, 1), (1000, 2), (1 << 15, 3)):
            pts = oc.g1_random_points(n, 50 + seed)
            pts[n // 2] = pts[0]                      # duplicate base: equal digits meet as P + P
            pts[n // 3] = 0                           # identity base
            neg = pts[1].copy()
            neg[4:] = bn.ints_to_array([(bn.FQ - v) % bn.FQ for v in bn.array_to_ints(pts[1][4:].reshape(1, 4), bn.FQ)], bn.FQ)[0]
            pts[n // 4] = neg                         # -P_1
            sc = oc.random_fr(n, 60 + seed)
            sc[n // 2] = sc[0]
            sc[n // 4] = sc[1]                        # s * P_1 + s * (-P_1) = identity
            h = gpu.srs_upload(pts)
            try:
                for vec in (sc, witness_like(oc, bn, n, 70 + seed)):
                    assert same_point(oc, gpu.msm_g1(h, vec), oc.msm_g1(pts, vec)), (n, seed)
            finally:
                gpu.srs_free(h)
    finally:
        gpu.configure("msm_affine", 0)


def test_group_helpers(gpu, oc):
    pts = oc.g1_random_points(4, 3)
    a = oc.g1_mul(pts[0], oc.random_fr(1, 1)[0])
    b = oc.g1_mul(pts[1], oc.random_fr(1, 2)[0])
    assert same_point(oc, gpu.g1_add(a, b), oc.g1_add_jac(a, b))
    assert same_point(oc, gpu.g1_add(a, a), oc.g1_add_jac(a, a))          # doubling path
    ident = np.zeros(12, dtype=np.uint64)
    assert same_point(oc, gpu.g1_add(a, ident), a) and same_point(oc, gpu.g1_add(ident, b), b)
    assert np.array_equal(gpu.g1_to_affine(a), oc.g1_to_affine(a))


# ---------------------------------------------------------------- the KZG mirror (reference tests restated)
def small_srs(oc, bn, n, seed=0x5eed):
    """A synthetic SRS with a known trapdoor: tau^i * G as affine points (KZGCommitmentScheme::new,
    kzg_poly_commitment.rs:183-204)."""
    tau = bn.array_to_ints(oc.random_fr(1, seed), bn.FR)[0]
    pts, acc = [], 1
    for _ in range(n):
        pts.append(bn.g1_mul(bn.G1_GEN, acc))
        acc = acc * tau % bn.FR
    return bn.affine_to_array(pts), tau


def test_commit_like_reference_test_commit(gpu, oc, bn):
    from uzkge_b200 import FpPolynomial, KZGCommitmentSchemeBN254

    srs, tau = small_srs(oc, bn, 11)
    pcs = KZGCommitmentSchemeBN254(srs)
    assert pcs.max_degree() == 10
    poly = FpPolynomial.from_coefs(fr(bn, [2, 3, 6]))
    commitment = pcs.commit(poly)
    expected = None   # "doing the multiexp by hand"
    P = bn.array_to_affine(srs)
    for i, coef in enumerate([2, 3, 6]):
        expected = bn.g1_add(expected, bn.g1_mul(P[i], coef))
    assert bn.array_to_affine(commitment.to_affine().reshape(1, 8))[0] == expected
    assert expected == bn.g1_mul(bn.G1_GEN, (2 + 3 * tau + 6 * tau * tau) % bn.FR)
    x, y = expected
    assert commitment.to_transcript_bytes() == x.to_bytes(32, "big") + y.to_bytes(32, "big")
    pcs.close()


def test_homomorphism_degree_error_and_blinds(gpu, oc, bn):
    from uzkge_b200 import DegreeError, FpPolynomial, KZGCommitmentSchemeBN254

    srs, tau = small_srs(oc, bn, 21)
    pcs = KZGCommitmentSchemeBN254(srs)
    p1 = FpPolynomial.from_coefs(fr(bn, [2, 3, 6]))
    p2 = FpPolynomial.from_coefs(fr(bn, [1, 8, 4]))
    c1, c2 = pcs.commit(p1), pcs.commit(p2)
    psum = FpPolynomial.from_coefs(fr(bn, [3, 11, 10]))
    assert pcs.commit(psum) == c1.add(c2)
    # zero polynomial: coefs = [0], degree 0 -> identity, transcript bytes are 64 zeros
    z = pcs.commit(FpPolynomial.zero())
    assert z.is_identity() and z.to_transcript_bytes() == bytes(64)
    with pytest.raises(DegreeError):
        pcs.commit(FpPolynomial.from_coefs(oc.random_fr(22, 1)))
    # batch == one by one
    polys = [p1, p2, psum, FpPolynomial.from_coefs(oc.random_fr(21, 2))]
    for a, b in zip(pcs.commit_batch(polys), [pcs.commit(p) for p in polys]):
        assert a == b
    # apply_blind_factors: C + sum b_i (SRS[i] - SRS[zd + i])  (kzg_poly_commitment.rs:299-313)
    blinds = [5, 7]
    zd = 16
    want = bn.g1_mul(bn.G1_GEN, (2 + 3 * tau + 6 * tau * tau) % bn.FR)
    for i, b in enumerate(blinds):
        want = bn.g1_add(want, bn.g1_mul(bn.G1_GEN, b * (pow(tau, i, bn.FR) - pow(tau, zd + i, bn.FR)) % bn.FR))
    got = pcs.apply_blind_factors(c1, fr(bn, blinds), zd)
    assert bn.array_to_affine(got.to_affine().reshape(1, 8))[0] == want
    pcs.close()


def test_prove_like_reference_test_eval(gpu, oc, bn):
    """test_eval (kzg_poly_commitment.rs:550-586): prove an evaluation; with the trapdoor known the proof must equal
    ((P(tau) - P(x)) / (tau - x)) * G, which is what the pairing check e(proof, [tau - x]_2) = e(C - P(x) G, H) verifies."""
    from uzkge_b200 import DegreeError, FpPolynomial, KZGCommitmentSchemeBN254

    srs, tau = small_srs(oc, bn, 11)
    pcs = KZGCommitmentSchemeBN254(srs)
    coefs = [1, 2, 3, 4, 5, 6, 7]
    poly = FpPolynomial.from_coefs(fr(bn, coefs))
    x = 123456789
    xm = fr(bn, [x])[0]
    px = sum(c * pow(x, i, bn.FR) for i, c in enumerate(coefs)) % bn.FR
    assert bn.array_to_ints(pcs.eval(poly, xm).reshape(1, 4), bn.FR) == [px]
    proof = pcs.prove(poly, xm, 10)
    ptau = sum(c * pow(tau, i, bn.FR) for i, c in enumerate(coefs)) % bn.FR
    qtau = (ptau - px) * bn.inv_mod((tau - x) % bn.FR, bn.FR) % bn.FR
    assert bn.array_to_affine(proof.to_affine().reshape(1, 8))[0] == bn.g1_mul(bn.G1_GEN, qtau)
    with pytest.raises(DegreeError):
        pcs.prove(poly, xm, 5)   # "wrong degree" branch of the reference test
    pcs.close()


def test_full_size_trapdoor_property(gpu, oc, bn):
    """BASELINE size 2^20 through a size-independent property: for bases P_i = P0 + i*Q,
    sum s_i P_i = (sum s_i) P0 + (sum i s_i) Q  (two scalar multiplications on the CPU)."""
    n = 1 << 20
    two = oc.g1_random_points(2, 99)
    pts = oc.g1_progression(two[0], two[1], n)
    h = gpu.srs_upload(pts)
    try:
        for sc in (oc.random_fr(n, 17), witness_like(oc, bn, n, 18)):
            s0, s1 = oc.fr_weighted_sums(sc)
            want = oc.g1_add_jac(oc.g1_mul(two[0], s0), oc.g1_mul(two[1], s1))
            assert same_point(oc, gpu.msm_g1(h, sc), want)
    finally:
        gpu.srs_free(h)


def witness_like_fast(oc, n, seed):
    """witness_like without per-element Python integers (for 2^22 and above): the small values are written as canonical limbs and
    converted to Montgomery form by the oracle in one call."""
    rng = np.random.default_rng(seed)
    s = oc.random_fr(n, seed)
    u = rng.random(n)
    canon = np.zeros((n, 4), dtype=np.uint64)
    canon[:, 0] = np.where(u < 0.8, rng.integers(0, 2, size=n), rng.integers(0, 1 << 16, size=n)).astype(np.uint64)
    small = oc.fr_to_mont(canon)
    s[u < 0.5] = 0
    m = (u >= 0.5) & (u < 0.9)
    s[m] = small[m]
    return s


@pytest.mark.parametrize("log_n", [22, 24])
def test_top_of_the_sweep_trapdoor_property(gpu, oc, log_n):
    """BASELINE configs[1] above 2^20 (2^22: 13 windows of c = 20; 2^24: the widest window, the largest `f n + i | sign` entry codes,
    13 GiB of tables): the same arithmetic-progression property, uniform and witness-like scalars, plus a prefix and an offset range."""
    n = 1 << log_n
    two = oc.g1_random_points(2, 1000 + log_n)
    pts = oc.g1_progression(two[0], two[1], n)
    h = gpu.srs_upload(pts)

    def want(sc, first):
        s0, s1 = oc.fr_weighted_sums(sc)      # sum s_i (P0 + (first + i) Q) = (sum s_i) (P0 + first Q) + (sum i s_i) Q
        return oc.g1_add_jac(oc.g1_mul(pts[first], s0), oc.g1_mul(two[1], s1))

    try:
        for sc in (oc.random_fr(n, 17 + log_n), witness_like_fast(oc, n, 18 + log_n)):
            assert same_point(oc, gpu.msm_g1(h, sc), want(sc, 0))
        m = n // 2 + 12345
        assert same_point(oc, gpu.msm_g1(h, sc[:m]), want(sc[:m], 0))
        assert same_point(oc, gpu.msm_g1(h, sc[:m], base_offset=n - m), want(sc[:m], n - m))
    finally:
        gpu.srs_free(h)


def test_largest_size_against_the_trapdoor(gpu, oc):
    """2^24 over a powers-of-tau SRS generated on the device: MSM(srs, f) = f(tau) G (SURVEY 8c golden 6) -- O(n) field work on the
    CPU and one scalar multiplication; bases of full entropy in every coordinate (the progression above has structured bases)."""
    n = 1 << 24
    tau = oc.random_fr(1, 4242)[0]
    srs = gpu.srs_generate(tau, n)
    h = gpu.srs_upload(srs)
    try:
        f = oc.random_fr(n, 4243)
        want = oc.g1_mul(srs[0], oc.fr_eval(f, tau))
        assert same_point(oc, gpu.msm_g1(h, f), want)
    finally:
        gpu.srs_free(h)


def test_partial_sums_are_combined_on_the_device(gpu, oc):
    """uzkge_cuda_g1_sum_device: out[j] = sum_r parts[r * stride + j] -- the N - 1 projective additions after the all-gather of a
    point-split MSM batch -- against the oracle's Jacobian addition, with identity and repeated points among the parts."""
    import torch

    pts = oc.g1_random_points(12, 9100)
    one = oc.fq_to_mont(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    jac = np.zeros((4, 3, 12), dtype=np.uint64)                   # 4 "ranks" x 3 results
    for r in range(4):
        for j in range(3):
            jac[r, j, :8] = pts[3 * r + j]
            jac[r, j, 8:] = one
    jac[1, 0] = 0                                                 # the identity (Z = 0)
    jac[2, 1] = jac[0, 1]                                         # P + P inside the sum
    d = torch.from_numpy(jac.view(np.int64).reshape(-1)).cuda()
    out = torch.zeros(3 * 12, dtype=torch.int64, device="cuda")
    gpu.g1_sum_device(d.data_ptr(), 4, out.data_ptr(), 0, k=3)
    got = out.cpu().numpy().view(np.uint64).reshape(3, 12)
    for j in range(3):
        acc = jac[0, j]
        for r in range(1, 4):
            acc = oc.g1_add_jac(acc, jac[r, j])
        assert np.array_equal(oc.g1_to_affine(got[j]), oc.g1_to_affine(acc)), j
    gpu.g1_sum_device(d.data_ptr(), 1, out.data_ptr(), 0)          # a single part: a copy
    assert np.array_equal(oc.g1_to_affine(out.cpu().numpy().view(np.uint64)[:12]), oc.g1_to_affine(jac[0, 0]))
