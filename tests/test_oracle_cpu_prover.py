"""The compiled CPU prover (oracle/cpu_prover.py + oracle.c: the proofs/s baseline of bench.py) against the big-integer restatement
of the reference's prover (oracle/plonk_prover.py): same circuit, SRS, RNG seed and transcript label -> the same proof, byte for
byte, and the restated verifier accepts it.  CPU only."""
import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import cpu_prover as cp
from oracle import plonk_prover as pp
from plonk_circuits import build_circuit


@pytest.mark.parametrize("gates,seed,n_public", [(5, 1, 1), (12, 2, 3), (40, 3, 2), (100, 4, 0)])
def test_cpu_prover_equals_the_restatement(gates, seed, n_public):
    cs = build_circuit(pp.TurboCS(), gates, seed, n_public=n_public, n_boolean=2)
    tau = 0x1234567890ABCDEF + seed
    pcs = pp.Kzg(cs.size + 2, tau)
    cpcs = cp.CpuKzg(bn.affine_to_array(pcs.srs))
    P, cP = pp.indexer(cs, pcs), cp.indexer(cs, cpcs)
    assert cP["vp"]["cm_q_vec"] == P["vp"]["cm_q_vec"] and cP["vp"]["cm_s_vec"] == P["vp"]["cm_s_vec"]
    assert cP["vp"]["lagrange_constants"] == P["vp"]["lagrange_constants"] and cP["vp"]["cm_qb"] == P["vp"]["cm_qb"]
    want = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"cpu"), pcs, cs, P, cs.witness)
    got = cp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"cpu"), cpcs, cs, cP, cs.witness)
    assert pp.proof_to_bytes_be(got) == pp.proof_to_bytes_be(want)
    pi = [cs.witness[i] for i in cs.public_vars_witness_indices]
    assert pp.verifier(pp.Transcript(b"cpu"), pcs, cP["vp"], pi, got)


def test_pieces_against_python_integers():
    rng = np.random.default_rng(5)
    xs = [int(x) for x in rng.integers(1, 1 << 62, size=9)]
    a = cp.A(xs)
    assert cp.I(a) == xs
    q, rem = cp.div_linear(a, 7)
    wq, wrem = pp.p_div_linear(xs, 7)
    assert cp.I(q) == wq and rem == wrem
    lc = cp.lincomb([(3, a), (5, a[:4])])
    assert cp.I(lc) == [(3 * x + (5 * x if i < 4 else 0)) % bn.FR for i, x in enumerate(xs)]
    assert cp.p_eval(a, 11) == pp.p_eval(xs, 11)
    assert cp.I(cp.trim(cp.A([1, 2, 0, 0]))) == [1, 2] and cp.I(cp.trim(cp.A([0, 0]))) == [0]


def test_array_circuit_gives_the_same_proof():
    """The array hand-over bench.py uses (the product's TurboCS layout) against the list-based circuit."""
    cs = build_circuit(pp.TurboCS(), 30, 9, n_public=2, n_boolean=1)
    pcs = pp.Kzg(cs.size + 2, 0xABCDEF)
    cpcs = cp.CpuKzg(bn.affine_to_array(pcs.srs))
    acs = cp.ArrayCS([cp.A(s) for s in cs.selectors], np.array(cs.wiring), cs.boolean_constraint_indices,
                     cs.public_vars_constraint_indices, cs.public_vars_witness_indices)
    want = cp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"arr"), cpcs, cs, cp.indexer(cs, cpcs), cs.witness)
    got = cp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"arr"), cpcs, acs, cp.indexer(acs, cpcs), cp.A(cs.witness))
    assert pp.proof_to_bytes_be(got) == pp.proof_to_bytes_be(want)
