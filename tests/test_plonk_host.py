"""CPU: the host-side pieces of the device-resident prover (transcript, RNG, circuit bookkeeping) against known answers, and the
big-integer restatement of the reference's prover against its own restated verifier (oracle/plonk_prover.py)."""
import random

import numpy as np
import pytest

from plonk_circuits import FR, build_circuit
from uzkge_b200 import rng as prng_mod
from uzkge_b200 import transcript as tr_mod

K_GOLDEN = [1,
            0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A,
            0x2042A587A90C187B0A087C03E29C968B950B1DB26D5C82D666905A6895790C0A,
            0x2DB4944E13E6E33CF0EF0734796FF332D73B5FA160DCA733BF529E9B758E4960,
            0x1D9E3A4AAF01052D9925138DC6D7D05AA614E311040142458B045D0053D22F46]


def test_keccak256_known_answers():
    from oracle import plonk_prover as pp

    kats = {b"": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
            b"abc": "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"}
    for msg, want in kats.items():
        assert tr_mod.keccak256(msg).hex() == want        # csrc/hostutil.c
        assert tr_mod.keccak256_py(msg).hex() == want
        assert pp.keccak256(msg).hex() == want
    rnd = random.Random(1)
    for n in (1, 31, 32, 135, 136, 137, 272, 1000):     # around the 136-byte rate
        msg = bytes(rnd.randrange(256) for _ in range(n))
        assert tr_mod.keccak256(msg) == pp.keccak256(msg) == tr_mod.keccak256_py(msg)


def test_chacha_fr_rand_reproduces_the_golden_ks(domain_kat):
    """choose_ks (plonk/indexer.rs:211-235) with ChaChaRng::from_seed([0; 32]) must give the k[1..4] the reference wrote into its
    verifier keys (tests/golden/domain_kat.json): pins ChaCha20, next_u64 order and Fr::rand's Montgomery interpretation."""
    from oracle import plonk_prover as pp

    k = prng_mod.choose_ks(prng_mod.ChaChaRng.from_seed(bytes(32)), 5)
    assert k == K_GOLDEN
    assert pp.choose_ks(pp.ChaCha(bytes(32)), 5) == K_GOLDEN
    for e in domain_kat.values():
        if "k" in e:
            assert [int(x, 16) for x in e["k"]] == K_GOLDEN[: len(e["k"])]
    c, d = prng_mod.ChaChaRng.from_seed(bytes(range(32))), prng_mod.ChaChaRng.from_seed(bytes(range(32)))
    assert [c._block() for _ in range(3)] == [d._block_py() for _ in range(3)]     # csrc/hostutil.c against its Python statement
    a, b = prng_mod.ChaChaRng.from_seed(bytes(range(32))), pp.ChaCha(bytes(range(32)))
    assert [prng_mod.fr_rand(a) for _ in range(40)] == [b.fr() for _ in range(40)]


def test_transcript_matches_restatement():
    from oracle import plonk_prover as pp

    a, b = tr_mod.Transcript(b"Plonk test"), pp.Transcript(b"Plonk test")
    a.append_message(b"PLONK"); b.msg(b"PLONK")
    a.append_u64(1 << 40); b.u64(1 << 40)
    a.append_challenge(12345); b.fr(12345)
    assert a.get_challenge_field_elem() == b.challenge()
    a.append_single_byte(1); b.byte(1)
    a.append_message(bytes(range(64))); b.msg(bytes(range(64)))
    c = a.get_challenge_field_elem()
    assert c == b.challenge() and 0 <= c < FR
    assert bytes(a.state) == c.to_bytes(32, "big")


def test_turbo_cs_bookkeeping_matches_restatement():
    """Selectors, wiring, padding, public-input rows and the copy-constraint permutation of the product's TurboCS against the
    restatement (whose compute_permutation follows constraint_system/mod.rs:54-84 position by position)."""
    from oracle import plonk_prover as pp
    from uzkge_b200.plonk import TurboCS, unmont

    for n_gates, seed in ((3, 1), (20, 2), (100, 3)):
        a = build_circuit(TurboCS(), n_gates, seed, n_public=2, n_boolean=2)
        b = build_circuit(pp.TurboCS(), n_gates, seed, n_public=2, n_boolean=2)
        assert a.size == b.size and a.num_vars == b.num_vars
        assert a.witness == b.witness
        assert a.public_vars_constraint_indices == b.public_vars_constraint_indices
        assert a.public_vars_witness_indices == b.public_vars_witness_indices
        assert a.boolean_constraint_indices == b.boolean_constraint_indices
        for j in range(9):
            assert [unmont(r) for r in a.selectors[j]] == b.selectors[j]
        for j in range(5):
            assert a.wiring[j].tolist() == b.wiring[j]
        assert a.compute_permutation().tolist() == b.compute_permutation()
        # the quadratic loop of the reference, literally, on the small case
        if n_gates <= 20:
            v = [x for w in b.wiring for x in w]
            perm, marked = [0] * len(v), set()
            for i, val in enumerate(v):
                if val in marked:
                    continue
                prev = i
                for j in range(i + 1, len(v)):
                    if v[j] == val:
                        perm[prev] = j
                        prev = j
                perm[prev] = i
                marked.add(val)
            assert perm == b.compute_permutation()


@pytest.mark.parametrize("n_gates", [2, 25])
def test_restated_prover_is_accepted_by_restated_verifier(n_gates):
    """prover.rs:88-394 and verifier.rs:17-164 restated with big integers: a proof verifies, a tampered one and one over another
    public input do not.  n = 8 uses the 16 n quotient domain, n = 64 the 6 n one (turbo/mod.rs:89-95)."""
    from oracle import plonk_prover as pp

    cs = build_circuit(pp.TurboCS(), n_gates, 7)
    pcs = pp.Kzg(cs.size + 2, 0x1234567890ABCDEF1234567890ABCDEF)
    params = pp.indexer(cs, pcs)
    proof = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"test"), pcs, cs, params, cs.witness)
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    assert pp.verifier(pp.Transcript(b"test"), pcs, params["vp"], pi, proof)
    bad = dict(proof, z_eval_zeta_omega=(proof["z_eval_zeta_omega"] + 1) % FR)
    assert not pp.verifier(pp.Transcript(b"test"), pcs, params["vp"], pi, bad)
    assert not pp.verifier(pp.Transcript(b"test"), pcs, params["vp"], [(pi[0] + 1) % FR], proof)
    assert not pp.verifier(pp.Transcript(b"other"), pcs, params["vp"], pi, proof)
    # commitments through the trapdoor equal commitments through the SRS (the MSM definition)
    assert pcs.commit(params["q_polys"][0]) == pcs.commit_msm(params["q_polys"][0])


def test_turbo_cs_helper_gadgets():
    """range_check, select, is_equal / is_not_equal (turbo/mod.rs:703-836) on the host mirror: the honest witness satisfies every gate
    (verify_witness), a value outside the range or a flipped flag does not, and the restated prover / verifier accept the circuit."""
    from plonk_circuits import transplant

    from oracle import plonk_prover as pp
    from uzkge_b200 import plonk
    from uzkge_b200.errors import UzkgeError

    def build(x_value, n_bits):
        cs = plonk.TurboCS()
        x, y, z = cs.new_variable(x_value), cs.new_variable(77), cs.new_variable(77)
        bits = cs.range_check(x, n_bits)
        eq, ne = cs.is_equal_or_not_equal(y, z)
        ne2 = cs.is_not_equal(x, y)
        picked = cs.select(x, y, cs.is_equal(x, x))
        cs.prepare_pi_variable(picked)
        cs.pad()
        return cs, bits, (eq, ne, ne2, picked)

    for n_bits in (2, 3, 4, 5, 8, 13):
        value = (1 << n_bits) - 3 if n_bits > 2 else 2
        cs, bits, (eq, ne, ne2, picked) = build(value, n_bits)
        w = cs.witness
        assert sum(w[b] << i for i, b in enumerate(bits)) == value and len(bits) == n_bits
        assert (w[eq], w[ne], w[ne2], w[picked]) == (1, 0, 1, 77)
        cs.verify_witness(w, [77])
    cs, bits, _ = build(1 << 9, 8)                                   # 512 does not fit 8 bits
    with pytest.raises(UzkgeError):
        cs.verify_witness(cs.witness, [77])
    cs, bits, (eq, ne, ne2, picked) = build(200, 8)
    bad = list(cs.witness)
    bad[eq] = 0
    with pytest.raises(UzkgeError):
        cs.verify_witness(bad, [77])
    ocs = transplant(cs)
    tau = 0x1234567890ABCDEF1234567890ABCDEF
    pcs = pp.Kzg(cs.size + 2, tau)
    P = pp.indexer(ocs, pcs)
    proof = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"gadgets"), pcs, ocs, P, ocs.witness)
    assert pp.verifier(pp.Transcript(b"gadgets"), pcs, P["vp"], [77], proof)
    assert not pp.verifier(pp.Transcript(b"gadgets"), pcs, P["vp"], [78], proof)
