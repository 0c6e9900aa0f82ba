"""CPU: the restated verifier against the reference's OWN golden proofs.

/root/reference/contracts/solidity/test/plonk_{20,52}.js hold a shuffle proof, its public inputs and public-key commitments that the
reference's verifier accepts (fixtures: tests/golden/plonk_{20,52}_golden.json, extracted by tests/golden/make_golden_proof.py).
oracle/plonk_verifier_shuffle.py restates verify_shuffle + verifier (plonk/verifier.rs:17-164, `shuffle` feature) down to the
pairing (oracle/pairing.py).  Accepting the golden proofs -- and rejecting every modification -- pins what the prover restatement and
the GPU prover share with it: the Keccak transcript and the order of everything appended to it, the challenges, the linearisation
(r_eval_zeta / r_commitment), eval_pi_poly, and the batch-opening algebra of KZG."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(cards, domain_kat):
    fx = json.load(open(os.path.join(GOLDEN, f"plonk_{cards}_golden.json")))
    e = domain_kat[str(cards)]
    return fx, [int(x, 16) for x in e["PI_POLY_INDICES_LOC"]], [int(x, 16) for x in e["PI_POLY_LAGRANGE_LOC"]]


def test_pairing_bilinear_and_reference_srs(bn):
    """e(a P, Q) = e(P, Q)^a, and the reference's own parameters: e(tau G, H) = e(G, tau H) with tau G from srs-padding.bin and
    tau H from the generated Solidity verifier."""
    from oracle import pairing as pr

    e1 = pr.pairing(pr.G2_GEN, (1, 2))
    assert e1 != pr.Fp12.one() and e1 ** bn.FR == pr.Fp12.one()
    assert pr.pairing(pr.G2_GEN, bn.g1_mul((1, 2), 7)) == e1 ** 7
    fx = json.load(open(os.path.join(GOLDEN, "plonk_20_golden.json")))
    x1, x0, y1, y0 = (int(v, 16) for v in fx["g2_tau_h_eip197"])
    tau_h = ((x0, x1), (y0, y1))
    h1, h0, k1, k0 = (int(v, 16) for v in fx["g2_h_eip197"])
    assert ((h0, h1), (k0, k1)) == pr.G2_GEN and pr.g2_is_on_twist(tau_h)
    head = np.load(os.path.join(GOLDEN, "srs_padding_head.npy"))
    tau_g = (bn.limbs_to_int(head[1][:4]), bn.limbs_to_int(head[1][4:]))     # canonical (non-Montgomery) limbs in the fixture
    assert bn.g1_is_on_curve(tau_g)
    assert pr.multi_pairing_is_one([(tau_g, pr.G2_GEN), (bn.g1_neg((1, 2)), tau_h)])


@pytest.mark.parametrize("cards", [20, 52])
def test_reference_golden_proof_is_accepted(domain_kat, cards):
    from oracle import plonk_verifier_shuffle as vs

    fx, pts, lag = _load(cards, domain_kat)
    assert vs.verify_shuffle_proof(fx, pts, lag)


def test_modified_golden_proofs_are_rejected(domain_kat):
    from oracle import plonk_verifier_shuffle as vs
    from oracle.bn254 import FR

    fx, pts, lag = _load(20, domain_kat)

    def bump(field, idx=None):
        def f(p, pi):
            if idx is None:
                p[field] = (p[field] + 1) % FR
            else:
                p[field][idx] = (p[field][idx] + 1) % FR
        return f

    def bump_pi(p, pi):
        pi[17] = (pi[17] + 1) % FR

    def swap_commitments(p, pi):
        p["cm_w_vec"][0], p["cm_w_vec"][1] = p["cm_w_vec"][1], p["cm_w_vec"][0]

    for tamper in (bump("z_eval_zeta_omega"), bump("w_polys_eval_zeta", 3), bump("w_sel_polys_eval_zeta", 2), bump("q_ecc_poly_eval_zeta"),
                   bump_pi, swap_commitments):
        assert not vs.verify_shuffle_proof(fx, pts, lag, tamper=tamper)
    wrong_cards = dict(fx, n_cards=21)       # the transcript prefix binds the number of cards (build_cs.rs:108-109)
    assert not vs.verify_shuffle_proof(wrong_cards, pts, lag)


@pytest.mark.parametrize("n_gates", [3, 25])
def test_restated_shuffle_prover_is_accepted_by_the_golden_pinned_verifier(n_gates):
    """The prover restatement for the `shuffle` feature set (witness selectors, quotient terms 12-18, linearisation parts 6-9) emits
    1632-byte proofs that the verifier above -- the one that accepts the reference's golden proofs -- accepts under the synthetic SRS's
    trapdoor, and rejects after a change.  The GPU prover is compared byte for byte with this restatement (tests/test_gpu_plonk.py)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from plonk_circuits import FR, build_circuit

    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs

    tau = 0x1234567890ABCDEF1234567890ABCDEF
    cs = build_circuit(pp.TurboCS(), n_gates, 7, n_public=2)
    pcs = pp.Kzg(cs.size + 2, tau)
    params = pp.indexer(cs, pcs, shuffle=True)
    proof = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"Plonk shuffle Proof"), pcs, cs, params, cs.witness)
    raw = pp.proof_to_bytes_be(proof)
    assert len(raw) == 1632
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    parsed = vs.parse_proof(raw)
    assert vs.verifier(pp.Transcript(b"Plonk shuffle Proof"), params["vp"], pi, parsed, trapdoor=tau)
    bad = dict(parsed, w_sel_polys_eval_zeta=[(parsed["w_sel_polys_eval_zeta"][0] + 1) % FR] + parsed["w_sel_polys_eval_zeta"][1:])
    assert not vs.verifier(pp.Transcript(b"Plonk shuffle Proof"), params["vp"], pi, bad, trapdoor=tau)
    assert not vs.verifier(pp.Transcript(b"Plonk shuffle Proof"), params["vp"], [(pi[0] + 1) % FR] + pi[1:], parsed, trapdoor=tau)
