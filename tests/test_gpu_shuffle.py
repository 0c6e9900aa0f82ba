"""GPU: the zshuffle circuit (remark + permutation gadgets) through the device-resident indexer / prover.

* small decks: every commitment and evaluation of the GPU proof equals the big-integer restatement's (oracle/plonk_prover.py), which
  the golden-pinned verifier accepts (tests/test_shuffle_host.py);
* the Lagrange paths (partial and all-Lagrange, incl. an SRS with the production files' holes) give the same bytes;
* the reference's PRODUCTION parameters (bundled Lagrange SRS, srs-padding.bin, VerifierKey_{20,52}.sol): the GPU indexer reproduces
  the deployed verifier key, and a GPU proof verifies under that key with the pairing check against the deployed G2 elements.
"""
import json
import os

import numpy as np
import pytest

from plonk_circuits import build_shuffle_circuit, shuffle_inputs

pytestmark = pytest.mark.gpu

TAU = 0x1234567890ABCDEF1234567890ABCDEF
LABEL = b"Plonk shuffle Proof"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _aff(bn, cm):
    a = cm.to_affine()
    if not a.any():
        return None
    x, y = bn.array_to_ints(a.reshape(2, 4), bn.FQ)
    return (x, y)


def _transcript(cards):
    from uzkge_b200.transcript import Transcript

    tr = Transcript(LABEL)
    tr.append_u64(cards)
    return tr


@pytest.mark.parametrize("n_cards", [1, 2])
def test_remark_circuit_prover_matches_restatement(gpu, bn, n_cards):
    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng

    inp = shuffle_inputs(n_cards, 5)
    cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
    ocs, _ = build_shuffle_circuit(pp.TurboCS(), inp)
    n = cs.size
    pcs, opcs = KZGCommitmentSchemeBN254.new(n + 2, plonk.mont(TAU)), pp.Kzg(n + 2, TAU)
    params, oparams = plonk.indexer(cs, pcs, shuffle=True), pp.indexer(ocs, opcs, shuffle=True)
    vp, ovp = params.verifier_params, oparams["vp"]
    assert [_aff(bn, c) for c in vp.cm_q_vec] == ovp["cm_q_vec"] and [_aff(bn, c) for c in vp.cm_s_vec] == ovp["cm_s_vec"]
    assert _aff(bn, vp.cm_q_ecc) == ovp["cm_q_ecc"] and _aff(bn, vp.cm_qb) == ovp["cm_qb"]
    assert [_aff(bn, c) for c in vp.cm_shuffle_generator_vec] == ovp["cm_shuffle_generator_vec"]
    assert [_aff(bn, c) for c in vp.cm_shuffle_public_key_vec] == ovp["cm_shuffle_generator_vec"]
    cms = plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"])
    assert [_aff(bn, c) for c in cms] == pp.refresh_public_key(oparams, ocs, opcs, inp["pk"])

    otr = pp.Transcript(LABEL)
    otr.u64(n_cards)
    want = pp.proof_to_bytes_be(pp.prover(pp.ChaCha(bytes(32)), otr, opcs, ocs, oparams, ocs.witness))
    wit = cs.get_witness_array()
    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_cards), pcs, cs, params, wit)
    raw = proof.to_bytes_be()
    assert len(raw) == 1632 and raw == want
    pi = [ocs.witness[i] for i in ocs.public_vars_witness_indices]
    otr = pp.Transcript(LABEL)
    otr.u64(n_cards)
    assert vs.verifier(otr, ovp, pi, vs.parse_proof(raw), trapdoor=TAU)

    # the Lagrange routes: wires and z only; everything; everything because the monomial SRS has the production files' holes
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, plonk.mont(TAU))
    for kw in ({}, {"lagrange_all": True}):
        got = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_cards), pcs, cs, params, wit, lagrange_pcs=lagrange, **kw)
        assert got.to_bytes_be() == want, kw
    holes = pcs.public_parameter_group_1.copy()
    holes[3:n] = 0
    sparse = KZGCommitmentSchemeBN254(holes)
    params.workspace.pop("lagrange_scheme", None)
    params.workspace.pop("srs_truncated", None)
    got = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_cards), sparse, cs, params, wit, lagrange_pcs=lagrange)
    assert got.to_bytes_be() == want
    # the indexer and the key refresh over the Lagrange SRS give the same commitments
    params2 = plonk.indexer(cs, sparse, shuffle=True, lagrange_pcs=lagrange)
    vp2 = params2.verifier_params
    for a, b in zip(vp.cm_q_vec + vp.cm_s_vec + [vp.cm_qb, vp.cm_q_ecc] + vp.cm_shuffle_generator_vec,
                    vp2.cm_q_vec + vp2.cm_s_vec + [vp2.cm_qb, vp2.cm_q_ecc] + vp2.cm_shuffle_generator_vec):
        assert _aff(bn, a) == _aff(bn, b)
    cms2 = plonk.refresh_prover_params_public_key(cs, params2, sparse, inp["pk"], lagrange_pcs=lagrange)
    assert [_aff(bn, c) for c in cms2] == [_aff(bn, c) for c in cms]
    for p in (pcs, lagrange, sparse):
        p.close()


@pytest.mark.parametrize("cards,n,tail", [(20, 4096, 0), (52, 16384, 6)])
def test_zshuffle_with_the_production_parameters(gpu, bn, domain_kat, srs_padding_head, srs_padding_tail, lagrange_srs_4096,
                                                 lagrange_srs_16384, cards, n, tail):
    """load_srs_params + load_lagrange_params (gen_params/mod.rs:140-171) -> indexer_with_lagrange -> refresh_prover_params_public_key
    -> prove_shuffle -> verify_shuffle, with the bundled parameter files and the deployed verifier key."""
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng

    fx = json.load(open(os.path.join(GOLDEN, f"plonk_{cards}_golden.json")))
    vk = vs.parse_vk(fx["vk_words"], fx["public_key_commitments"], [], [])
    inp = shuffle_inputs(cards, 2)
    cs, out = build_shuffle_circuit(plonk.TurboCS(), inp)
    assert cs.size == n
    srs = np.zeros((n + 3, 8), dtype=np.uint64)              # tau^i G for i < 64 (of the file's 2051) and i in [n, n + 3)
    srs[:64] = srs_padding_head
    srs[n:n + 3] = srs_padding_tail[tail:tail + 3]
    pcs = KZGCommitmentSchemeBN254(srs)
    lagrange = KZGCommitmentSchemeBN254(lagrange_srs_4096 if n == 4096 else lagrange_srs_16384)
    params = plonk.indexer(cs, pcs, shuffle=True, lagrange_pcs=lagrange)
    vp = params.verifier_params
    assert [_aff(bn, c) for c in vp.cm_q_vec] == vk["cm_q_vec"]
    assert [_aff(bn, c) for c in vp.cm_s_vec] == vk["cm_s_vec"]
    assert _aff(bn, vp.cm_qb) == vk["cm_qb"] and _aff(bn, vp.cm_q_ecc) == vk["cm_q_ecc"]
    assert [_aff(bn, c) for c in vp.cm_shuffle_generator_vec] == vk["cm_shuffle_generator_vec"]
    assert vp.k == vk["k"] and vp.edwards_a == vk["edwards_a"] and params.root == vk["root"]

    cms = plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"], lagrange_pcs=lagrange)
    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(cards), pcs, cs, params, cs.get_witness_array(), lagrange_pcs=lagrange)
    raw = proof.to_bytes_be()
    assert len(raw) == 1632
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    pk_words = [hex(c) for cm in cms for c in _aff(bn, cm)]
    e = domain_kat[str(cards)]
    pts, lag = [int(x, 16) for x in e["PI_POLY_INDICES_LOC"]], [int(x, 16) for x in e["PI_POLY_LAGRANGE_LOC"]]
    ours = dict(fx, proof="0x" + raw.hex(), public_inputs=[hex(v) for v in pi], public_key_commitments=pk_words)
    assert vs.verify_shuffle_proof(ours, pts, lag)
    other_deck = dict(ours, public_inputs=[hex(v) for v in pi[:-1]] + [hex((pi[-1] + 1) % bn.FR)])
    assert not vs.verify_shuffle_proof(other_deck, pts, lag)
    assert not vs.verify_shuffle_proof(dict(ours, public_key_commitments=fx["public_key_commitments"]), pts, lag)
    for p in (pcs, lagrange):
        p.close()
