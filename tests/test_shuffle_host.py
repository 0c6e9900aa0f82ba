"""CPU: the shuffle gadgets of the host mirror (uzkge_b200/shuffle.py, the shuffle methods of plonk.TurboCS) and of the restatement
(oracle/babyjubjub.py, oracle/plonk_prover.py) against the reference's own data:

* shuffle/babyjubjub.rs's preprocessed generator tables (tests/golden/babyjubjub_generators.json) pin the curve arithmetic;
* the deployed verifier keys VerifierKey_{20,52}.sol (tests/golden/plonk_{20,52}_golden.json) hold the commitments of every
  preprocessed polynomial of the zshuffle circuits over the bundled Lagrange SRS: building the circuit here (build_cs.rs:26-56) and
  committing its selector / permutation / q_ecc / generator-selector columns must reproduce all 28 of them, which pins the gate
  layout, the variable numbering, compute_permutation, choose_ks and the remark tables -- SURVEY 8c's "golden 3".
"""
import json
import os

import numpy as np
import pytest

from plonk_circuits import FR, build_shuffle_circuit, shuffle_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def tables():
    return json.load(open(os.path.join(GOLDEN, "babyjubjub_generators.json")))


def _matches(table, golden) -> bool:
    """A derived 84 x 4 table against the fixture: SHA-256 of the decimal entries ("," within a row, ";" between rows), plus the
    first and last rows in clear."""
    import hashlib

    rows = [[str(v) for v in row] for row in table]
    digest = hashlib.sha256(";".join(",".join(r) for r in rows).encode()).hexdigest()
    return len(rows) == golden["rows"] and digest == golden["sha256"] and rows[0] == golden["first"] and rows[-1] == golden["last"]


def test_generator_tables_match_the_reference(tables):
    from oracle import babyjubjub as bj
    from uzkge_b200 import shuffle as sh

    segs = bj.segments(bj.GEN)
    assert _matches([[p[0] for p in s] for s in segs], tables["x"])
    assert _matches([[p[1] for p in s] for s in segs], tables["y"])
    assert _matches([[bj.D * p[0] * p[1] % FR for p in s] for s in segs], tables["dxy"])
    g = sh.BabyJubjubShuffle
    assert _matches(g.get_preprocessed_generators_x(), tables["x"])
    assert _matches(g.get_preprocessed_generators_y(), tables["y"])
    assert _matches(g.get_preprocessed_generators_dxy(), tables["dxy"])
    assert not _matches(g.get_preprocessed_generators_y(), tables["x"])
    assert g.NUM_ITERATIONS == bj.NUM_ITERATIONS == 84
    # the constants: curve equation at the generator, d from the table, prime-order subgroup
    x, y = bj.GEN
    assert (bj.A * x * x + y * y - 1 - bj.D * x * x * y * y) % FR == 0
    assert int(tables["dxy"]["first"][0]) * pow(x * y, -1, FR) % FR == bj.D == sh.COEFF_D and [str(x), tables["y"]["first"][0]] == [tables["x"]["first"][0], str(y)]
    e = bj._ext(bj.GEN)
    acc, k = (0, 1, 1, 0), bj.ORDER
    while k:                                  # ORDER * G without the reduction `mul` applies
        if k & 1:
            acc = bj._ext_add(acc, e)
        e = bj._ext_add(e, e)
        k >>= 1
    assert bj._affine(acc) == (0, 1)
    assert sh.SUBGROUP_ORDER == bj.ORDER and sh.GENERATOR == bj.GEN


def test_remark_is_a_rerandomisation(tables):
    """test_babyjubjub_remark (shuffle/babyjubjub.rs:3568-3598): the remarked card decrypts to the same message; the trace's last
    row is the remarked card; product and restatement agree on every intermediate point."""
    from oracle import babyjubjub as bj
    from uzkge_b200 import shuffle as sh
    from uzkge_b200.rng import ChaChaRng

    inp = shuffle_inputs(1, 3)
    card, bits = inp["cards"][0], inp["bits"][0]
    fb, iv = bj.remark_trace(card, bits, inp["pk"])
    trace = sh.BabyJubjubShuffle.eval_remark_with_trace(sh.Ciphertext(*card), bits, inp["pk"])
    assert trace.bits == fb and trace.intermediate_values == iv and trace.n_round == 84
    out = sh.BabyJubjubShuffle.eval_remark(sh.Ciphertext(*card), bits, inp["pk"])
    assert out.flatten() == iv[-1]
    assert out.verify(inp["messages"][0], inp["sk"]) and bj.decrypt_ok((out.e1, out.e2), inp["messages"][0], inp["sk"])
    assert out != sh.Ciphertext(*card)
    prng = ChaChaRng.from_seed(bytes(32))
    drawn = sh.BabyJubjubShuffle.sample_random_scalar_bits(prng)
    assert len(drawn) == 84 and all(len(b) == 3 for b in drawn)
    c = sh.Ciphertext.rand(prng)
    assert sh.ed_is_on_curve(c.e1) and sh.ed_is_on_curve(c.e2)
    p = sh.Permutation.rand(prng, 52)
    p.sanity_check()


@pytest.mark.parametrize("n_cards", [1, 2, 5])
def test_shuffle_circuit_matches_restatement_and_is_satisfied(n_cards):
    """test_remark_constraint_system / the permutation tests (constraint_system/shuffle/*.rs): same gates, wiring, witness and selector
    tables in both mirrors; verify_witness accepts the honest witness and rejects a changed one; the output deck is the permuted,
    re-randomised input deck."""
    from oracle import babyjubjub as bj
    from oracle import plonk_prover as pp
    from uzkge_b200 import plonk
    from uzkge_b200.errors import UzkgeError

    inp = shuffle_inputs(n_cards, 40 + n_cards)
    ocs, oout = build_shuffle_circuit(pp.TurboCS(), inp)
    cs, out = build_shuffle_circuit(plonk.TurboCS(), inp)
    assert [list(v) for v in out] == oout and cs.size == ocs.size and cs.num_vars == ocs.num_vars
    assert cs.witness == ocs.witness
    assert np.array_equal(cs.wiring, np.array(ocs.wiring, dtype=np.uint32))
    un = lambda a: [plonk.unmont(r) for r in a]
    assert all(un(cs.selectors[j]) == ocs.selectors[j] for j in range(9))
    assert cs.public_vars_constraint_indices == ocs.public_vars_constraint_indices
    assert cs.boolean_constraint_indices == ocs.boolean_constraint_indices
    assert cs.shuffle_remark_constraint_indices() == [r for r, _ in ocs.remark]
    assert all(un(a) == b for a, b in zip(cs.compute_witness_selectors(), ocs.compute_witness_selectors()))
    assert all(un(a) == b for a, b in zip(cs.compute_shuffle_generator_selectors(), ocs.table_selectors(ocs.gen_table)))
    assert all(un(a) == b for a, b in zip(cs.compute_shuffle_public_key_selectors(), ocs.table_selectors(ocs.pk_table)))
    assert un(cs.compute_q_ecc()) == ocs.q_ecc() and sum(ocs.q_ecc()) == 84 * n_cards
    assert cs.compute_witness_selectors_int() == ocs.compute_witness_selectors()
    assert cs.compute_shuffle_public_key_selectors_int() == ocs.table_selectors(ocs.pk_table)

    online = [cs.witness[i] for i in cs.public_vars_witness_indices]
    assert len(online) == 8 * n_cards
    cs.verify_witness(cs.witness, online)
    assert ocs.check_remark_equations(ocs.witness)
    bad = list(cs.witness)
    bad[cs.wiring[0, cs.shuffle_remark_constraint_indices()[0] + 7]] += 1       # a point inside the first remark chain
    with pytest.raises(UzkgeError):
        cs.verify_witness(bad, [bad[i] for i in cs.public_vars_witness_indices])
    assert not ocs.check_remark_equations(bad)
    with pytest.raises(UzkgeError):
        cs.verify_witness(cs.witness, [online[0] + 1] + online[1:])
    for i, cv in enumerate(oout):                 # output card i = re-randomised input card perm[i]
        w = ocs.witness
        assert bj.decrypt_ok(((w[cv[2]], w[cv[3]]), (w[cv[0]], w[cv[1]])), inp["messages"][inp["perm"][i]], inp["sk"])
    assert [w for card in inp["cards"] for w in (card[1][0], card[1][1], card[0][0], card[0][1])] == online[: 4 * n_cards]


@pytest.mark.parametrize("cards,n", [(20, 4096), (52, 16384)])
def test_zshuffle_circuit_reproduces_the_deployed_verifier_key(oc, bn, domain_kat, cards, n):
    """The 28 preprocessed commitments of VerifierKey_{20,52}.sol from the circuit built HERE and the bundled Lagrange SRS, plus the
    public-input rows (VerifierKeyExtra1: w^row) and k."""
    from oracle import plonk_verifier_shuffle as vs
    from uzkge_b200 import plonk
    from uzkge_b200.rng import ChaChaRng, choose_ks

    fx = json.load(open(os.path.join(GOLDEN, f"plonk_{cards}_golden.json")))
    vk = vs.parse_vk(fx["vk_words"], fx["public_key_commitments"], [], [])
    cs, _ = build_shuffle_circuit(plonk.TurboCS(), shuffle_inputs(cards, 1))
    assert cs.size == n == vk["cs_size"]
    raw = np.load(os.path.join(GOLDEN, f"lagrange_srs_{n}.npy"))
    srs = oc.fq_to_mont(raw.reshape(-1, 4)).reshape(-1, 8)

    def commit(evals_mont):
        a = oc.fq_from_mont(oc.g1_to_affine(oc.msm_g1(srs, evals_mont)).reshape(2, 4))
        x, y = (sum(int(a[c][i]) << (64 * i) for i in range(4)) for c in range(2))
        return None if x == 0 and y == 0 else (x, y)

    assert [commit(cs.selectors[j]) for j in range(9)] == vk["cm_q_vec"]
    k = choose_ks(ChaChaRng.from_seed(bytes(32)), 5)
    assert k == vk["k"]
    root = bn.root_of_unity(n)
    assert root == vk["root"]
    group = [1] * n
    for i in range(1, n):
        group[i] = group[i - 1] * root % FR
    perm = cs.compute_permutation()
    sigma = [[k[int(p) // n] * group[int(p) % n] % FR for p in perm[c * n:(c + 1) * n]] for c in range(5)]
    assert [commit(bn.ints_to_array(s, bn.FR)) for s in sigma] == vk["cm_s_vec"]
    qb = [0] * n
    for i in cs.boolean_constraint_indices:
        qb[i] = 1
    assert commit(bn.ints_to_array(qb, bn.FR)) == vk["cm_qb"]
    assert commit(cs.compute_q_ecc()) == vk["cm_q_ecc"]
    gen = cs.compute_shuffle_generator_selectors()
    assert [commit(gen[i]) for i in range(12)] == vk["cm_shuffle_generator_vec"]
    assert vk["cm_prk_vec"] == [None] * 4 and vk["edwards_a"] == cs.edwards_a == 1
    e = domain_kat[str(cards)]
    assert [group[i] for i in cs.public_vars_constraint_indices] == [int(x, 16) for x in e["PI_POLY_INDICES_LOC"]]
    assert [group[i] * pow(n, -1, FR) % FR for i in cs.public_vars_constraint_indices] == [int(x, 16) for x in e["PI_POLY_LAGRANGE_LOC"]]


@pytest.mark.parametrize("n_cards", [1, 2])
def test_restated_prover_on_remark_circuits(n_cards):
    """The restatement's indexer / refresh_public_key / prover on circuits WITH remark gates (non-zero q_ecc, witness selectors,
    generator and public-key selectors): accepted by the verifier that accepts the reference's golden proofs; rejected when the
    verifier is given the commitments of another key (the un-refreshed parameters), another deck, or another card count."""
    from oracle import plonk_prover as pp
    from oracle import plonk_verifier_shuffle as vs

    tau = 0x1234567890ABCDEF1234567890ABCDEF
    label = b"Plonk shuffle Proof"

    def transcript(cards=n_cards):
        tr = pp.Transcript(label)
        tr.u64(cards)                                       # build_cs.rs:71-72
        return tr

    inp = shuffle_inputs(n_cards, 5)
    cs, _ = build_shuffle_circuit(pp.TurboCS(), inp)
    pcs = pp.Kzg(cs.size + 2, tau)
    P = pp.indexer(cs, pcs, shuffle=True)
    assert P["vp"]["cm_shuffle_public_key_vec"] == P["vp"]["cm_shuffle_generator_vec"]          # indexer.rs:471-476
    stale = list(P["vp"]["cm_shuffle_public_key_vec"])
    cms = pp.refresh_public_key(P, cs, pcs, inp["pk"])
    assert cms != stale and len(cms) == 12
    proof = pp.prover(pp.ChaCha(bytes(32)), transcript(), pcs, cs, P, cs.witness)
    raw = pp.proof_to_bytes_be(proof)
    assert len(raw) == 1632
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    assert vs.verifier(transcript(), P["vp"], pi, vs.parse_proof(raw), trapdoor=tau)
    assert not vs.verifier(transcript(), dict(P["vp"], cm_shuffle_public_key_vec=stale), pi, vs.parse_proof(raw), trapdoor=tau)
    assert not vs.verifier(transcript(), P["vp"], pi[:-1] + [(pi[-1] + 1) % FR], vs.parse_proof(raw), trapdoor=tau)
    assert not vs.verifier(transcript(n_cards + 1), P["vp"], pi, vs.parse_proof(raw), trapdoor=tau)
