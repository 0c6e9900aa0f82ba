"""GPU: zmatchmaking's circuit (Anemoi gates: non-zero round-key selectors, quotient terms 8-11, the prk parts of the linearisation)
through the device-resident indexer / prover.

* small circuits (3 and 5 inputs): the GPU proof equals the big-integer restatement's byte for byte in both feature sets;
* the PRODUCTION parameters (bundled Lagrange SRS of size 8192, srs-padding.bin, matchmaking/parameters/vk-specific.bin): the GPU
  indexer reproduces the bundled verifier key, and a GPU proof of the 50-input circuit verifies under that key's commitments with
  the pairing check against the deployed G2 elements.
"""
import json
import os
import random

import numpy as np
import pytest

from plonk_circuits import FR, transplant

pytestmark = pytest.mark.gpu

TAU = 0x1234567890ABCDEF1234567890ABCDEF
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _aff(bn, cm):
    a = cm.to_affine()
    if not a.any():
        return None
    x, y = bn.array_to_ints(a.reshape(2, 4), bn.FQ)
    return (x, y)


def _transcript(n_inputs):
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200.transcript import Transcript

    tr = Transcript(mm.PLONK_PROOF_TRANSCRIPT)
    tr.append_u64(n_inputs)
    return tr


@pytest.mark.parametrize("n_inputs", [3, 5])
def test_anemoi_circuit_prover_matches_restatement(gpu, bn, n_inputs):
    from oracle import plonk_prover as pp
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200.rng import ChaChaRng

    rnd = random.Random(9 + n_inputs)
    cs, _ = mm.build_cs(plonk.TurboCS(), [rnd.randrange(FR) for _ in range(n_inputs)], rnd.randrange(FR), rnd.randrange(FR))
    ocs = transplant(cs)
    n = cs.size
    pcs, opcs = KZGCommitmentSchemeBN254.new(n + 2, plonk.mont(TAU)), pp.Kzg(n + 2, TAU)
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, plonk.mont(TAU))
    wit = cs.get_witness_array()
    for shuffle in (False, True):
        params, oparams = plonk.indexer(cs, pcs, shuffle=shuffle), pp.indexer(ocs, opcs, shuffle=shuffle)
        vp, ovp = params.verifier_params, oparams["vp"]
        assert [_aff(bn, c) for c in vp.cm_prk_vec] == list(ovp["cm_prk_vec"]) and None not in ovp["cm_prk_vec"]
        assert [_aff(bn, c) for c in vp.cm_q_vec] == ovp["cm_q_vec"] and [_aff(bn, c) for c in vp.cm_s_vec] == ovp["cm_s_vec"]
        assert (vp.anemoi_generator, vp.anemoi_generator_inv) == (ovp["anemoi_generator"], ovp["anemoi_generator_inv"])
        otr = pp.Transcript(mm.PLONK_PROOF_TRANSCRIPT)
        otr.u64(n_inputs)
        want = pp.proof_to_bytes_be(pp.prover(pp.ChaCha(bytes(32)), otr, opcs, ocs, oparams, ocs.witness))
        proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_inputs), pcs, cs, params, wit)
        assert proof.prk_3_poly_eval_zeta != 0
        assert proof.to_bytes_be() == want, shuffle
        for kw in ({}, {"lagrange_all": True}):
            got = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_inputs), pcs, cs, params, wit, lagrange_pcs=lagrange, **kw)
            assert got.to_bytes_be() == want, (shuffle, kw)
        if not shuffle and n >= 16:
            got = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(n_inputs), pcs, cs, params, wit, quotient_by_cosets=True)
            assert got.to_bytes_be() == want
    for p in (pcs, lagrange):
        p.close()


def test_matchmaking_with_the_production_parameters(gpu, bn, srs_padding_head, srs_padding_tail, lagrange_srs_8192):
    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200.rng import ChaChaRng

    vk = json.load(open(os.path.join(GOLDEN, "matchmaking_vk.json")))
    pt = lambda v: None if v is None else (int(v[0], 16), int(v[1], 16))
    rnd = random.Random(4)
    cs, _ = mm.build_cs(plonk.TurboCS(), list(range(1, mm.N + 1)), rnd.randrange(FR), rnd.randrange(FR))
    n = cs.size
    assert n == 8192
    srs = np.zeros((n + 3, 8), dtype=np.uint64)
    srs[:64] = srs_padding_head
    srs[n:n + 3] = srs_padding_tail[3:6]
    pcs, lagrange = KZGCommitmentSchemeBN254(srs), KZGCommitmentSchemeBN254(lagrange_srs_8192)
    params = plonk.indexer(cs, pcs, shuffle=True, lagrange_pcs=lagrange)
    vp = params.verifier_params
    assert [_aff(bn, vp.cm_q_vec[j]) for j in (0, 1, 2, 3, 4, 5, 6, 8)] == [pt(c) for c in vk["cm_q_vec"]]
    assert [_aff(bn, c) for c in vp.cm_s_vec] == [pt(c) for c in vk["cm_s_vec"]]
    assert _aff(bn, vp.cm_qb) == pt(vk["cm_qb"]) and [_aff(bn, c) for c in vp.cm_prk_vec] == [pt(c) for c in vk["cm_prk_vec"]]
    assert _aff(bn, vp.cm_q_vec[7]) is None and _aff(bn, vp.cm_q_ecc) is None

    proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), _transcript(mm.N), pcs, cs, params, cs.get_witness_array(), lagrange_pcs=lagrange)
    raw = proof.to_bytes_be()
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    num = lambda v: int(v, 16) if isinstance(v, str) else v
    root = bn.root_of_unity(n)
    key = {   # the bundled key's commitments in the layout of the `shuffle` feature set (identity for the columns it predates)
        "cm_q_vec": [pt(c) for c in vk["cm_q_vec"][:7]] + [None, pt(vk["cm_q_vec"][7])], "cm_s_vec": [pt(c) for c in vk["cm_s_vec"]],
        "cm_qb": pt(vk["cm_qb"]), "cm_prk_vec": [pt(c) for c in vk["cm_prk_vec"]], "cm_q_ecc": None,
        "cm_shuffle_generator_vec": [None] * 12, "cm_shuffle_public_key_vec": [None] * 12,
        "anemoi_generator": num(vk["anemoi_generator"]), "anemoi_generator_inv": num(vk["anemoi_generator_inv"]),
        "k": [num(v) for v in vk["k"]], "edwards_a": 0, "root": root, "cs_size": n,
        "pi_points": [pow(root, i, FR) for i in vk["public_vars_constraint_indices"]],
        "pi_lagrange": [num(v) for v in vk["lagrange_constants"]],
    }
    fx = json.load(open(os.path.join(GOLDEN, "plonk_52_golden.json")))         # the same SRS: its G2 elements
    g2 = (vs._g2(fx["g2_tau_h_eip197"]), vs._g2(fx["g2_h_eip197"]))

    def transcript():
        tr = pp.Transcript(mm.PLONK_PROOF_TRANSCRIPT)
        tr.u64(mm.N)
        return tr

    assert vs.verifier(transcript(), key, pi, vs.parse_proof(raw), g2=g2)
    assert not vs.verifier(transcript(), key, pi[:-1] + [(pi[-1] + 1) % FR], vs.parse_proof(raw), g2=g2)
    for p in (pcs, lagrange):
        p.close()
