"""GPU: the device-resident TurboPlonK prover (uzkge_b200/plonk.py over the C ABI) against the big-integer restatement of the
reference's prover (oracle/plonk_prover.py: prover.rs:88-394) -- same circuit, same SRS trapdoor, same ChaCha seed, same transcript
label: every commitment and every evaluation of the proof must be identical, and the restated verifier must accept it."""
import numpy as np
import pytest

from plonk_circuits import FR, build_circuit

pytestmark = pytest.mark.gpu

TAU = 0x1234567890ABCDEF1234567890ABCDEF


def _aff(bn, cm):
    """KZGCommitment -> canonical affine tuple or None."""
    a = cm.to_affine()
    if not a.any():
        return None
    x, y = bn.array_to_ints(a.reshape(2, 4), bn.FQ)
    return (x, y)


def _proof_as_oracle_dict(bn, proof):
    return {
        "cm_w_vec": [_aff(bn, c) for c in proof.cm_w_vec], "cm_t_vec": [_aff(bn, c) for c in proof.cm_t_vec], "cm_z": _aff(bn, proof.cm_z),
        "prk_3_poly_eval_zeta": proof.prk_3_poly_eval_zeta, "prk_4_poly_eval_zeta": proof.prk_4_poly_eval_zeta,
        "w_polys_eval_zeta": list(proof.w_polys_eval_zeta), "w_polys_eval_zeta_omega": list(proof.w_polys_eval_zeta_omega),
        "z_eval_zeta_omega": proof.z_eval_zeta_omega, "s_polys_eval_zeta": list(proof.s_polys_eval_zeta),
        "opening_witness_zeta": _aff(bn, proof.opening_witness_zeta), "opening_witness_zeta_omega": _aff(bn, proof.opening_witness_zeta_omega),
    }


@pytest.mark.parametrize("n_gates,n_public,n_boolean", [(2, 1, 1), (25, 1, 1), (100, 3, 2), (200, 0, 0)])
def test_prover_matches_restatement(gpu, bn, n_gates, n_public, n_boolean):
    from oracle import plonk_prover as pp
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    seed = 11 + n_gates
    cs = build_circuit(plonk.TurboCS(), n_gates, seed, n_public, n_boolean)
    ocs = build_circuit(pp.TurboCS(), n_gates, seed, n_public, n_boolean)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    opcs = pp.Kzg(cs.size + 2, TAU)
    params = plonk.indexer(cs, pcs)
    oparams = pp.indexer(ocs, opcs)
    vp, ovp = params.verifier_params, oparams["vp"]
    assert vp.k == ovp["k"]
    assert [_aff(bn, c) for c in vp.cm_q_vec] == ovp["cm_q_vec"]
    assert [_aff(bn, c) for c in vp.cm_s_vec] == ovp["cm_s_vec"]
    assert _aff(bn, vp.cm_qb) == ovp["cm_qb"]
    assert [_aff(bn, c) for c in vp.cm_prk_vec] == ovp["cm_prk_vec"]
    assert vp.lagrange_constants == ovp["lagrange_constants"]
    # preprocessed coset evaluations and coefficient forms
    for got, want in zip(params.s_coset_evals + params.q_coset_evals + [params.l1_coset_evals, params.coset_quotient],
                         list(oparams["s_coset"]) + list(oparams["q_coset"]) + [oparams["l1_coset"], oparams["coset_quotient"]]):
        assert bn.array_to_ints(got.numpy(), bn.FR) == list(want)

    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, cs.get_witness_array())
    want = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"test"), opcs, ocs, oparams, ocs.witness)
    want.pop("_u")
    got = _proof_as_oracle_dict(bn, proof)
    for key in want:
        assert got[key] == want[key], key
    pi = [ocs.witness[i] for i in ocs.public_vars_witness_indices]
    assert pp.verifier(pp.Transcript(b"test"), opcs, ovp, pi, got)
    # a different RNG seed changes the blinds, hence the proof, and it still verifies
    proof2 = plonk.prover(ChaChaRng.from_seed(bytes([1] * 32)), Transcript(b"test"), pcs, cs, params, cs.get_witness_array())
    got2 = _proof_as_oracle_dict(bn, proof2)
    assert got2["cm_w_vec"] != got["cm_w_vec"]
    assert pp.verifier(pp.Transcript(b"test"), opcs, ovp, pi, got2)
    pcs.close()


def test_prover_rejects_an_unsatisfied_witness(gpu, bn):
    """A wrong witness makes the numerator non-divisible by Z_H: the quotient no longer fits the SRS (DegreeError in the reference's
    commit, kzg_poly_commitment.rs:283-285) -- no proof comes out."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.errors import UzkgeError
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = build_circuit(plonk.TurboCS(), 30, 5)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs)
    w = cs.get_witness_array().copy()
    w[7] = plonk.mont(plonk.unmont(w[7]) + 1)
    with pytest.raises(UzkgeError):
        plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, w)
    pcs.close()


@pytest.mark.parametrize("log_size", [10, 14, 20, 22])
def test_synthetic_circuit_proof_verifies(gpu, bn, oc, log_size):
    """BASELINE configs[4] up to its full size (2^22 gates): a synthetic circuit of add / mul gates built in bulk, proved on the GPU, checked by the
    restated verifier under the SRS trapdoor (O(1) group operations, independent of n); the witness satisfies every gate."""
    from oracle import plonk_prover as pp
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = plonk.TurboCS.synthetic(log_size, seed=log_size)
    n = cs.size
    assert n == 1 << log_size
    # gate equations on the host for a sample of rows
    wit = cs.get_witness_array()
    rows = np.random.default_rng(0).integers(0, n, 200)
    for r in rows:
        w = [plonk.unmont(wit[cs.wiring[j][r]]) for j in range(5)]
        q = [plonk.unmont(cs.selectors[j][r]) for j in range(9)]
        assert (q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3] + q[4] * w[0] * w[1] + q[5] * w[2] * w[3] + q[6]
                + q[7] * w[0] * w[1] * w[2] * w[3] * w[4] - q[8] * w[4]) % FR == 0
    pcs = KZGCommitmentSchemeBN254.new(n + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs)
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, plonk.mont(TAU)) if log_size == 14 else None   # one size through prover_with_lagrange
    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"synthetic"), pcs, cs, params, wit, lagrange_pcs=lagrange)
    vp = params.verifier_params

    class Trapdoor:
        tau = TAU

    ovp = {"cm_q_vec": [_aff(bn, c) for c in vp.cm_q_vec], "cm_s_vec": [_aff(bn, c) for c in vp.cm_s_vec], "cm_qb": _aff(bn, vp.cm_qb),
           "cm_prk_vec": [_aff(bn, c) for c in vp.cm_prk_vec], "anemoi_generator": 0, "anemoi_generator_inv": 0, "k": vp.k, "cs_size": n,
           "public_vars_constraint_indices": [], "lagrange_constants": []}
    got = _proof_as_oracle_dict(bn, proof)
    assert pp.verifier(pp.Transcript(b"synthetic"), Trapdoor, ovp, [], got)
    got["w_polys_eval_zeta"][2] = (got["w_polys_eval_zeta"][2] + 1) % FR
    assert not pp.verifier(pp.Transcript(b"synthetic"), Trapdoor, ovp, [], got)
    pcs.close()


def test_lagrange_srs_known_answers(gpu, oc, bn):
    """uzkge_cuda_srs_generate_lagrange: sum_i w^(i j) L_i(tau) G = tau^j G -- the circuit-free known answer the reference's bundled
    lagrange-srs files satisfy against srs-padding.bin (tests/test_oracle_golden.py), here for a synthetic trapdoor."""
    from uzkge_b200 import plonk

    n = 512
    tau_m = plonk.mont(TAU)
    lag = gpu.srs_generate_lagrange(tau_m, n)
    mono = gpu.srs_generate(tau_m, 8)
    assert all(oc.g1_on_curve(p) for p in lag[:16])
    w = bn.root_of_unity(n)
    h = gpu.srs_upload(lag)
    try:
        for j in (0, 1, 2, 5):
            sc = bn.ints_to_array([pow(w, i * j, bn.FR) for i in range(n)], bn.FR)
            got = gpu.g1_to_affine(gpu.msm_g1(h, sc))
            assert np.array_equal(got, mono[j]), j
    finally:
        gpu.srs_free(h)


@pytest.mark.parametrize("n_gates", [25, 200])
def test_prover_with_lagrange_gives_the_same_proof(gpu, bn, n_gates):
    """prover_with_lagrange (prover.rs:88-146): committing the evaluation vectors against the Lagrange SRS and folding the blind
    factors in (apply_blind_factors, kzg_poly_commitment.rs:299-313) yields the same commitments as the coefficient path."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = build_circuit(plonk.TurboCS(), n_gates, 3, 2, 1)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    lagrange_pcs = KZGCommitmentSchemeBN254.new_lagrange(cs.size, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs)
    wit = cs.get_witness_array()
    a = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, wit)
    b = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, wit, lagrange_pcs=lagrange_pcs)
    assert _proof_as_oracle_dict(bn, a) == _proof_as_oracle_dict(bn, b)
    # a Lagrange SRS of another size is ignored, as in the reference (prover.rs:119-124)
    other = KZGCommitmentSchemeBN254.new_lagrange(2 * cs.size, plonk.mont(TAU))
    c = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, wit, lagrange_pcs=other)
    assert _proof_as_oracle_dict(bn, a) == _proof_as_oracle_dict(bn, c)
    for p in (pcs, lagrange_pcs, other):
        p.close()


@pytest.mark.parametrize("n", [1, 2, 8, 256, 4096])
def test_lagrange_srs_from_monomial_without_trapdoor(gpu, oc, bn, n):
    """uzkge_cuda_srs_lagrange_from_monomial (the inverse transform over G1 points) must give exactly the points
    L_i(tau) * G that the trapdoor route gives; for n = 8 also against the definition (1 / n) sum_j w^(-i j) P_j in big integers,
    with an identity point among the inputs."""
    from uzkge_b200 import plonk

    tau_m = plonk.mont(TAU)
    mono = gpu.srs_generate(tau_m, n)
    got = gpu.srs_lagrange_from_monomial(mono, n)
    assert np.array_equal(got, gpu.srs_generate_lagrange(tau_m, n))
    if n == 8:
        pts = mono.copy()
        pts[5] = 0                                                   # identity (the padded SRS holds such entries)
        got = gpu.srs_lagrange_from_monomial(pts, n)
        P = bn.array_to_affine(pts)
        w_inv, n_inv = pow(bn.root_of_unity(n), -1, bn.FR), pow(n, -1, bn.FR)
        for i in range(n):
            want = bn.msm_naive(P, [n_inv * pow(w_inv, i * j, bn.FR) % bn.FR for j in range(n)])
            assert bn.array_to_affine(got[i].reshape(1, 8))[0] == want, i


@pytest.mark.parametrize("n_gates,n_public", [(2, 1), (60, 2)])
def test_shuffle_feature_prover_matches_restatement_and_golden_pinned_verifier(gpu, bn, n_gates, n_public):
    """The `shuffle` feature set (what zshuffle is compiled with): witness-selector polynomials, quotient terms 12-18, q_ecc / w_sel
    openings, linearisation parts 6-9.  The GPU proof must equal the restatement's byte for byte (PlonkProof::to_bytes_be) and be
    accepted by oracle/plonk_verifier_shuffle.py -- the verifier restatement that accepts the reference's own golden proofs
    (tests/test_oracle_golden_proof.py) -- under the synthetic SRS's trapdoor."""
    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = build_circuit(plonk.TurboCS(), n_gates, 21, n_public, 1)
    ocs = build_circuit(pp.TurboCS(), n_gates, 21, n_public, 1)
    pcs, opcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU)), pp.Kzg(cs.size + 2, TAU)
    params, oparams = plonk.indexer(cs, pcs, shuffle=True), pp.indexer(ocs, opcs, shuffle=True)
    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"Plonk shuffle Proof"), pcs, cs, params, cs.get_witness_array())
    want = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"Plonk shuffle Proof"), opcs, ocs, oparams, ocs.witness)
    raw = proof.to_bytes_be()
    assert len(raw) == 1632 and raw == pp.proof_to_bytes_be(want)
    pi = [ocs.witness[i] for i in ocs.public_vars_witness_indices]
    assert vs.verifier(pp.Transcript(b"Plonk shuffle Proof"), oparams["vp"], pi, vs.parse_proof(raw), trapdoor=TAU)
    bad = bytearray(raw)
    bad[900] ^= 1                          # inside prk_3 / a wire evaluation
    try:
        rejected = not vs.verifier(pp.Transcript(b"Plonk shuffle Proof"), oparams["vp"], pi, vs.parse_proof(bytes(bad)), trapdoor=TAU)
    except AssertionError:                 # not a field element any more
        rejected = True
    assert rejected
    pcs.close()


def test_shuffle_quotient_terms_match_restatement(gpu, oc, bn):
    """uzkge_cuda_plonk_quotient_shuffle_fr_device on random coset evaluations (every selector non-zero) against t_poly's loop body with
    terms 1-18 in big integers (oracle/plonk.py)."""
    import torch

    from oracle import plonk

    n, factor = 32, 6
    m = n * factor
    K5 = [1, 0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A, 3, 5, 7]
    rnd = lambda seed: oc.random_fr(m, seed)
    groups = {"w": 5, "q": 9, "s": 5, "q_prk": 4, "w_sel": 3, "pk": 12, "gen": 12}
    arrays, seed = {}, 500
    for name, cnt in groups.items():
        arrays[name] = [rnd(seed + i) for i in range(cnt)]
        seed += cnt
    for name in ("pi", "z", "coset_quotient", "l1", "qb", "q_ecc"):
        arrays[name] = rnd(seed)
        seed += 1
    sc = oc.random_fr(6, 77)
    alpha, beta, gamma, g, ed_a = sc[0], sc[1], sc[2], sc[3], sc[4]
    g_inv = oc.fr_inv(g)
    zh = oc.random_fr(factor, 78)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda()
    keep = {nm: ([dev(a) for a in v] if isinstance(v, list) else dev(v)) for nm, v in arrays.items()}
    out = torch.empty(4 * m, dtype=torch.int64, device="cuda")
    P_ = lambda t: t.data_ptr()
    gpu.plonk_quotient_fr_device(
        [P_(t) for t in keep["w"]], [P_(t) for t in keep["q"]], P_(keep["pi"]), P_(keep["z"]), [P_(t) for t in keep["s"]],
        P_(keep["coset_quotient"]), P_(keep["l1"]), P_(keep["qb"]), [P_(t) for t in keep["q_prk"]], bn.ints_to_array(K5, bn.FR),
        alpha, beta, gamma, g, g_inv, zh, m, factor, out.data_ptr(),
        shuffle={"w_sel": [P_(t) for t in keep["w_sel"]], "q_ecc": P_(keep["q_ecc"]), "pk": [P_(t) for t in keep["pk"]],
                 "gen": [P_(t) for t in keep["gen"]], "edwards_a": ed_a})
    torch.cuda.synchronize()
    got = bn.array_to_ints(out.cpu().numpy().view(np.uint64).reshape(m, 4), bn.FR)
    I = lambda a: bn.array_to_ints(a, bn.FR)
    S = lambda a: I(a.reshape(1, 4))[0]
    want = plonk.quotient_coset_evals(
        [I(a) for a in arrays["w"]], [I(a) for a in arrays["q"]], I(arrays["pi"]), I(arrays["z"]), [I(a) for a in arrays["s"]],
        I(arrays["coset_quotient"]), I(arrays["l1"]), I(arrays["qb"]), [I(a) for a in arrays["q_prk"]], K5, S(alpha), S(beta), S(gamma),
        S(g), S(g_inv), I(zh), factor,
        shuffle={"w_sel": [I(a) for a in arrays["w_sel"]], "q_ecc": I(arrays["q_ecc"]), "pk": [I(a) for a in arrays["pk"]],
                 "gen": [I(a) for a in arrays["gen"]], "edwards_a": S(ed_a)})
    assert got == want


@pytest.mark.parametrize("n_gates,n_public", [(25, 2), (300, 0)])
def test_quotient_round_by_cosets_gives_the_same_proof(gpu, bn, n_gates, n_public):
    """The quotient round evaluated coset by coset (six size-n coset transforms per polynomial on the folded coefficients, the
    pointwise map per coset with factor 1, interleave, one 6n inverse transform) -- the decomposition dist.SplitCommitter deals to the
    GPUs of a box -- must give the proof of the natural 6n path."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = build_circuit(plonk.TurboCS(), n_gates, 9, n_public, 1)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs)
    wit = cs.get_witness_array()
    a = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, wit)
    b = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, wit, quotient_by_cosets=True)
    assert a.to_bytes_be() == b.to_bytes_be()
    pcs.close()
