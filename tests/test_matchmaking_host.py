"""CPU: the Anemoi-Jive primitives, the Anemoi gates of TurboCS and zmatchmaking's circuit on the host mirror
(uzkge_b200/anemoi.py, uzkge_b200/matchmaking.py) against the reference's own data:

* anemoi/bn254/mod.rs's parameter tables and anemoi/tests.rs's known answers (tests/golden/anemoi_bn254.json): every DERIVED constant
  and the sponge / stream-cipher outputs;
* the reference's gadget tests (constraint_system/anemoi/mod.rs:536-626) restated: verify_witness accepts the gadgets' witnesses;
* zmatchmaking's bundled verifier key (matchmaking/parameters/vk-specific.bin -> tests/golden/matchmaking_vk.json): the circuit built
  HERE, committed over the bundled Lagrange SRS, reproduces its 18 commitments, public-input rows and Lagrange constants.
"""
import json
import os
import random

import numpy as np
import pytest

from plonk_circuits import FR, transplant

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLDEN, "anemoi_bn254.json")))


def test_derived_anemoi_parameters_match_the_reference(golden):
    from uzkge_b200.anemoi import AnemoiJive254 as A

    import hashlib

    def matches(table, g):
        rows = [[str(v) for v in row] for row in table]
        return len(rows) == 14 and hashlib.sha256(";".join(",".join(r) for r in rows).encode()).hexdigest() == g["sha256"] and rows[0] == g["first"]

    assert matches(A.ROUND_KEYS_X, golden["round_keys_x"]) and matches(A.ROUND_KEYS_Y, golden["round_keys_y"])
    px, py = A.preprocessed_round_keys()
    assert matches(px, golden["preprocessed_round_keys_x"]) and matches(py, golden["preprocessed_round_keys_y"])
    assert not matches(py, golden["preprocessed_round_keys_x"])
    assert A.MDS_MATRIX == [[int(v) for v in row] for row in golden["mds_matrix"]]
    assert A.GENERATOR == int(golden["generator"]) and A.GENERATOR_INV == int(golden["generator_inv"])
    assert A.ALPHA_INV == int(golden["alpha_inv"]) and A.ALPHA * A.ALPHA_INV % (FR - 1) == 1


def test_anemoi_known_answers(golden):
    """test_anemoi_variable_length_hash, test_eval_stream_cipher (anemoi/tests.rs:10-21, 210-237) and the trace variants."""
    from uzkge_b200.anemoi import AnemoiJive254 as A

    want = [int(v) for v in golden["stream_cipher_1234"]]
    assert A.eval_variable_length_hash([1, 2, 3, 4]) == int(golden["variable_length_hash_1234"]) == want[0]
    for n in (1, 2, 3, 4, 5, 6, 7):
        assert A.eval_stream_cipher([1, 2, 3, 4], n) == want[:n]
        trace = A.eval_stream_cipher_with_trace([1, 2, 3, 4], n)
        assert trace.output == want[:n]
        assert len(trace.before_permutation) == len(trace.after_permutation) == len(trace.intermediate_values_before_constant_additions)
    trace = A.eval_variable_length_hash_with_trace([1, 2, 3, 4])
    assert trace.output == want[0] and len(trace.before_permutation) == 2
    x, y = A.anemoi_permutation(*trace.before_permutation[0])
    assert (x, y) == trace.after_permutation[0]


def test_anemoi_gadgets_are_satisfied():
    """test_anemoi_variable_length_hash_constraint_system, test_anemoi_stream_cipher (constraint_system/anemoi/mod.rs:545-625)."""
    from uzkge_b200 import plonk
    from uzkge_b200.anemoi import AnemoiJive254 as A
    from uzkge_b200.errors import UzkgeError

    for values in ([1, 2, 3, 4], [7], [1, 2, 3], [1, 2, 3, 4, 5, 6, 7, 8]):
        trace = A.eval_variable_length_hash_with_trace(values)
        cs = plonk.TurboCS()
        cs.load_anemoi_parameters()
        cs.anemoi_variable_length_hash(trace, [cs.new_variable(v) for v in values], cs.new_variable(trace.output))
        cs.pad()
        cs.verify_witness(cs.witness, [])
    for output_len in range(1, 8):
        for input_len in (3, 4):
            values = list(range(1, input_len + 1))
            trace = A.eval_stream_cipher_with_trace(values, output_len)
            cs = plonk.TurboCS()
            cs.load_anemoi_parameters()
            cs.anemoi_stream_cipher(trace, [cs.new_variable(v) for v in values], [cs.new_variable(v) for v in trace.output])
            cs.pad()
            cs.verify_witness(cs.witness, [])
    bad = list(cs.witness)
    bad[cs.wiring[2, cs.anemoi_constraints_indices[0] + 3]] += 1
    with pytest.raises(UzkgeError):
        cs.verify_witness(bad, [])


def test_matchmaking_circuit_reproduces_the_bundled_verifier_key(oc, bn):
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200 import plonk
    from uzkge_b200.anemoi import AnemoiJive254 as A
    from uzkge_b200.rng import ChaChaRng, choose_ks

    vk = json.load(open(os.path.join(GOLDEN, "matchmaking_vk.json")))
    pt = lambda v: None if v is None else (int(v[0], 16), int(v[1], 16))
    num = lambda v: int(v, 16) if isinstance(v, str) else v
    rnd = random.Random(1)
    seed, number = rnd.randrange(FR), rnd.randrange(FR)
    inputs = list(range(1, mm.N + 1))                                    # test_matchmaking (matchmaking/src/test.rs:10-44)
    cs, out = mm.build_cs(plonk.TurboCS(), inputs, seed, number)
    n = cs.size
    assert n == vk["cs_size"] == vk["shrunk_cs"]["size"] == 8192 and cs.num_vars == vk["shrunk_cs"]["num_vars"]
    online = [cs.witness[i] for i in cs.public_vars_witness_indices]
    cs.verify_witness(cs.witness, online)
    assert online[:mm.N] == inputs and sorted(online[mm.N:2 * mm.N]) == inputs and online[mm.N:2 * mm.N] != inputs
    assert online[-2:] == [number, A.eval_variable_length_hash([seed])]
    srs = oc.fq_to_mont(np.load(os.path.join(GOLDEN, "lagrange_srs_8192.npy")).reshape(-1, 4)).reshape(-1, 8)

    def commit(evals_mont):
        a = oc.fq_from_mont(oc.g1_to_affine(oc.msm_g1(srs, evals_mont)).reshape(2, 4))
        x, y = (sum(int(a[c][i]) << (64 * i) for i in range(4)) for c in range(2))
        return None if x == 0 and y == 0 else (x, y)

    # the bundled key predates the q_ecc column: 8 selector commitments = today's selectors without index 7
    assert [commit(cs.selectors[j]) for j in (0, 1, 2, 3, 4, 5, 6, 8)] == [pt(c) for c in vk["cm_q_vec"]]
    assert not cs.selectors[7].any()
    k = choose_ks(ChaChaRng.from_seed(bytes(32)), 5)
    assert k == [num(v) for v in vk["k"]]
    root = bn.root_of_unity(n)
    group = [1] * n
    for i in range(1, n):
        group[i] = group[i - 1] * root % FR
    perm = cs.compute_permutation()
    sigma = [[k[int(p) // n] * group[int(p) % n] % FR for p in perm[c * n:(c + 1) * n]] for c in range(5)]
    assert [commit(bn.ints_to_array(s, bn.FR)) for s in sigma] == [pt(c) for c in vk["cm_s_vec"]]
    qb = [0] * n
    for i in cs.boolean_constraint_indices:
        qb[i] = 1
    assert commit(bn.ints_to_array(qb, bn.FR)) == pt(vk["cm_qb"])
    prk = cs.compute_anemoi_jive_selectors()
    assert [commit(prk[i]) for i in range(4)] == [pt(c) for c in vk["cm_prk_vec"]]
    assert [plonk.unmont(r) for r in prk[2]] == cs.compute_anemoi_jive_selectors_int()[2]
    assert num(vk["anemoi_generator"]) == cs.anemoi_generator and num(vk["anemoi_generator_inv"]) == cs.anemoi_generator_inv
    assert cs.public_vars_constraint_indices == vk["public_vars_constraint_indices"]
    assert [pow(n * pow(root, -ci, FR) % FR, -1, FR) for ci in cs.public_vars_constraint_indices] == [num(v) for v in vk["lagrange_constants"]]


@pytest.mark.parametrize("n_inputs", [3, 5])
def test_restated_prover_on_anemoi_circuits(n_inputs):
    """Small matchmaking circuits (real Anemoi gates: non-zero q_prk, quotient terms 8-11, the prk parts of the linearisation)
    through the restated indexer / prover in both feature sets; accepted by the restated verifiers, rejected for other outputs."""
    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200 import plonk

    tau = 0x1234567890ABCDEF1234567890ABCDEF
    rnd = random.Random(9 + n_inputs)
    cs, _ = mm.build_cs(plonk.TurboCS(), [rnd.randrange(FR) for _ in range(n_inputs)], rnd.randrange(FR), rnd.randrange(FR))
    ocs = transplant(cs)
    pcs = pp.Kzg(cs.size + 2, tau)
    pi = [ocs.witness[i] for i in ocs.public_vars_witness_indices]
    other = pi[:-1] + [(pi[-1] + 1) % FR]

    def transcript():
        tr = pp.Transcript(mm.PLONK_PROOF_TRANSCRIPT)
        tr.u64(n_inputs)                                     # build_cs.rs:81-82
        return tr

    P = pp.indexer(ocs, pcs)
    proof = pp.prover(pp.ChaCha(bytes(32)), transcript(), pcs, ocs, P, ocs.witness)
    assert proof["prk_3_poly_eval_zeta"] != 0 and len(pp.proof_to_bytes_be(proof)) == 1312
    assert pp.verifier(transcript(), pcs, P["vp"], pi, proof) and not pp.verifier(transcript(), pcs, P["vp"], other, proof)
    P = pp.indexer(ocs, pcs, shuffle=True)
    raw = pp.proof_to_bytes_be(pp.prover(pp.ChaCha(bytes(32)), transcript(), pcs, ocs, P, ocs.witness))
    assert len(raw) == 1632
    assert vs.verifier(transcript(), P["vp"], pi, vs.parse_proof(raw), trapdoor=tau)
    assert not vs.verifier(transcript(), P["vp"], other, vs.parse_proof(raw), trapdoor=tau)
