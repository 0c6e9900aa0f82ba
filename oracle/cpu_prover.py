"""A COMPILED CPU TurboPlonK prover: oracle/plonk_prover.py's indexer and prover (default feature set) with every O(n) step in C
(oracle/oracle.c, OpenMP) instead of Python integers.  TEST INFRASTRUCTURE ONLY: the `cpu_baseline` of bench.py's proofs/s block and
tests/ -- never imported by the product package.

What runs where -- the same split as the reference's prover_with_lagrange (/root/reference/uzkge/src/plonk/prover.rs:88-394):
  commitments          oracle_msm_g1 (arkworks' Pippenger restated, one task per window)      kzg_poly_commitment.rs:278-293
  transforms           oracle_ntt_fr (Radix2 / MixedRadix domains, serial coset power loop)   field_polynomial.rs:583-607
  z evaluations        oracle_plonk_z_evals (batch inversion, serial running product)         helpers.rs:160-220
  quotient map         oracle_plonk_quotient (parallel over the 6n points, terms 1-11)        helpers.rs:284-669
  evaluations          oracle_fr_eval (serial Horner)                                         field_polynomial.rs eval
  r / h polynomials    oracle_fr_lincomb                                                      helpers.rs:681-999, pcs.rs:107-147
  division by X - z    oracle_fr_div_linear                                                   pcs.rs:140-147
  transcript, RNG, challenges, O(1) scalar arithmetic: Python integers, shared with oracle/plonk_prover.py

Parity: tests/test_oracle_cpu_prover.py holds its proofs to oracle/plonk_prover.py's byte for byte (same circuit, SRS, seed and
transcript label); bench.py compares the proof it times with the GPU's.  Labelled kind = "port": arkworks itself cannot be built here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import bn254 as bn
from . import cpu as oc
from . import plonk as qmap
from . import plonk_prover as pp
from .bn254 import FR, inv_mod

N_WIRES, N_SELECTORS = pp.N_WIRES, pp.N_SELECTORS
_U64P = C.POINTER(C.c_uint64)
_bound = False


def _lib():
    global _bound
    L = oc.lib()
    if not _bound:
        pp_t = C.POINTER(_U64P)
        L.oracle_plonk_quotient.argtypes = [pp_t, _U64P, _U64P, C.c_size_t, C.c_size_t, _U64P]
        L.oracle_plonk_quotient.restype = None
        L.oracle_plonk_z_evals.argtypes = [_U64P, _U64P, _U64P, _U64P, _U64P, _U64P, C.c_size_t, _U64P]
        L.oracle_plonk_z_evals.restype = C.c_int
        L.oracle_fr_lincomb.argtypes = [pp_t, _U64P, _U64P, C.c_size_t, _U64P, C.c_size_t]
        L.oracle_fr_lincomb.restype = None
        L.oracle_fr_div_linear.argtypes = [_U64P, C.c_size_t, _U64P, _U64P, _U64P]
        L.oracle_fr_div_linear.restype = None
        _bound = True
    return L


# ---------------------------------------------------------------- representation: (len, 4) uint64 Montgomery limbs
def A(xs) -> np.ndarray:
    if isinstance(xs, np.ndarray) and xs.dtype == np.uint64:
        return np.ascontiguousarray(xs).reshape(-1, 4)        # already Montgomery limbs
    return bn.ints_to_array([int(x) % FR for x in xs], FR, mont=True)


def I(a) -> list[int]:
    return bn.array_to_ints(a, FR, mont=True)


def _p(a):
    return oc._p(a)


def trim(a: np.ndarray) -> np.ndarray:
    """FpPolynomial::from_coefs: drop trailing zeros, the zero polynomial keeps one coefficient."""
    nz = np.nonzero(a.any(axis=1))[0]
    return np.ascontiguousarray(a[: (int(nz[-1]) + 1 if nz.size else 1)])


def add_coef(a: np.ndarray, v: int, i: int) -> np.ndarray:
    """FpPolynomial::add_coef_assign: grows the vector when needed."""
    if a.shape[0] <= i:
        a = np.concatenate([a, np.zeros((i + 1 - a.shape[0], 4), dtype=np.uint64)])
    a[i] = A([(I(a[i:i + 1])[0] + v) % FR])[0]
    return a


def ifft(evals: np.ndarray, n: int) -> np.ndarray:
    return trim(oc.ntt_fr(evals, n, inverse=True))


def coset_fft(coefs: np.ndarray, m: int, k: int) -> np.ndarray:
    return oc.ntt_fr(coefs, m, coset=A([k]))


def coset_ifft(evals: np.ndarray, m: int, k_inv: int) -> np.ndarray:
    return oc.ntt_fr(evals, m, inverse=True, coset=A([k_inv]))


def p_eval(coefs: np.ndarray, x: int) -> int:
    return I(oc.fr_eval(coefs, A([x])[0]).reshape(1, 4))[0]


def _ptr_table(arrs):
    tab = (_U64P * len(arrs))(*[_p(a) for a in arrs])
    return tab


def lincomb(terms) -> np.ndarray:
    """sum s_i p_i over (scalar int, coefficient array) pairs."""
    polys = [np.ascontiguousarray(p, dtype=np.uint64) for _, p in terms]
    L = max(p.shape[0] for p in polys)
    out = np.empty((L, 4), dtype=np.uint64)
    lens = np.array([p.shape[0] for p in polys], dtype=np.uint64)
    sc = A([s for s, _ in terms])
    _lib().oracle_fr_lincomb(_ptr_table(polys), _p(lens), _p(sc), len(polys), _p(out), L)
    return out


def div_linear(coefs: np.ndarray, z: int):
    n = coefs.shape[0]
    quot = np.zeros((max(n - 1, 1), 4), dtype=np.uint64)
    rem = np.zeros(4, dtype=np.uint64)
    _lib().oracle_fr_div_linear(_p(np.ascontiguousarray(coefs)), n, _p(A([z])), _p(quot), _p(rem))
    return quot[: max(n - 1, 0)], I(rem.reshape(1, 4))[0]


class ArrayCS:
    """A circuit handed over as arrays (the layout of the product's TurboCS: selectors (9, n, 4) Montgomery limbs, wiring (5, n)
    variable numbers) with the interface the indexer / prover below use.  No Anemoi rows unless `anemoi_prk` (4, n, 4) is given."""

    def __init__(self, selectors, wiring, boolean_constraint_indices=(), public_vars_constraint_indices=(),
                 public_vars_witness_indices=(), anemoi_prk=None, anemoi_generator=0, anemoi_generator_inv=0):
        self.selectors = [np.ascontiguousarray(s_, dtype=np.uint64) for s_ in selectors]
        self.wiring = np.asarray(wiring, dtype=np.int64)
        self.size = int(self.wiring.shape[1])
        self.boolean_constraint_indices = list(boolean_constraint_indices)
        self.public_vars_constraint_indices = list(public_vars_constraint_indices)
        self.public_vars_witness_indices = list(public_vars_witness_indices)
        self.anemoi_prk = anemoi_prk
        self.anemoi_generator, self.anemoi_generator_inv = anemoi_generator, anemoi_generator_inv

    def quot_eval_dom_size(self):
        return self.size * 6 if self.size > 8 else self.size * 16

    def compute_permutation(self):
        """constraint_system/mod.rs:54-84 (as in oracle.plonk_prover.TurboCS)."""
        v = self.wiring.reshape(-1).tolist()
        perm = [0] * len(v)
        last, first = {}, {}
        for i, var in enumerate(v):
            if var in last:
                perm[last[var]] = i
            else:
                first[var] = i
            last[var] = i
        for var, i in last.items():
            perm[i] = first[var]
        return perm

    def compute_anemoi_jive_selectors(self):
        if self.anemoi_prk is not None:
            return [np.ascontiguousarray(e, dtype=np.uint64) for e in self.anemoi_prk]
        return [np.zeros((self.size, 4), dtype=np.uint64) for _ in range(4)]

    def extend_witness(self, witness):
        return np.ascontiguousarray(np.asarray(witness, dtype=np.uint64).reshape(-1, 4)[self.wiring.reshape(-1)])

    @staticmethod
    def get_hiding_degree(idx):
        return 3 if idx < 3 else 2


class CpuKzg:
    """KZG over an explicit SRS (affine points, (len, 8) uint64 Montgomery): every commitment is a real Pippenger MSM."""

    def __init__(self, srs_affine: np.ndarray):
        self.srs = np.ascontiguousarray(srs_affine, dtype=np.uint64).reshape(-1, 8)

    def commit(self, coefs: np.ndarray):
        c = trim(coefs)
        assert c.shape[0] <= self.srs.shape[0], "DegreeError"
        return bn.jac_array_to_affine(oc.msm_g1(self.srs[: c.shape[0]], c))


# ---------------------------------------------------------------- indexer (indexer.rs:248-536, default feature set)
def indexer(cs, pcs: CpuKzg):
    n, m = cs.size, cs.quot_eval_dom_size()
    factor = m // n
    root, root_m = bn.root_of_unity(n), bn.root_of_unity(m)
    group = [1] * n
    for i in range(1, n):
        group[i] = group[i - 1] * root % FR
    k = pp.choose_ks(pp.ChaCha(bytes(32)), N_WIRES)
    cq = [k[1]] * m
    for i in range(1, m):
        cq[i] = cq[i - 1] * root_m % FR
    perm = [int(p) for p in cs.compute_permutation()]
    enc = [k[p // n] * group[p % n] % FR for p in perm]

    def pre(evals_ints):
        coefs = ifft(A(evals_ints), n)
        return coefs, coset_fft(coefs, m, k[1])

    P = {"n": n, "m": m, "factor": factor, "group": A(group), "coset_quotient": A(cq), "root": root,
         "permutation": np.array(perm, dtype=np.uint64)}
    P["s_polys"], P["s_coset"] = zip(*[pre(enc[i * n:(i + 1) * n]) for i in range(N_WIRES)])
    P["q_polys"], P["q_coset"] = zip(*[pre(cs.selectors[i]) for i in range(N_SELECTORS)])
    P["l1_coefs"], P["l1_coset"] = pre([n % FR] + [0] * (n - 1))
    P["z_h_inv"] = A(qmap.z_h_inv_coset_evals(k[1], root_m, n, factor))
    qb = [0] * n
    for i in cs.boolean_constraint_indices:
        qb[i] = 1
    P["qb_poly"], P["qb_coset"] = pre(qb)
    P["q_prk_polys"], P["q_prk_coset"] = zip(*[pre(e) for e in cs.compute_anemoi_jive_selectors()])
    lagrange_constants = []
    for ci in cs.public_vars_constraint_indices:
        # prod_{i != ci} (g^ci - g^i) = n * g^(-ci)  (derivative of X^n - 1 at g^ci): the constant is g^ci / n
        lagrange_constants.append(group[ci] * inv_mod(n, FR) % FR)
    P["vp"] = {
        "cm_q_vec": [pcs.commit(p) for p in P["q_polys"]], "cm_s_vec": [pcs.commit(p) for p in P["s_polys"]],
        "cm_qb": pcs.commit(P["qb_poly"]), "cm_prk_vec": [pcs.commit(p) for p in P["q_prk_polys"]],
        "anemoi_generator": cs.anemoi_generator, "anemoi_generator_inv": cs.anemoi_generator_inv, "k": k, "cs_size": n,
        "public_vars_constraint_indices": list(cs.public_vars_constraint_indices), "lagrange_constants": lagrange_constants,
    }
    return P


# ---------------------------------------------------------------- prover (prover.rs:88-394, lagrange_pcs = None)
def hide_polynomial(rng, coefs, hiding_degree, zeroing_degree):
    for i in range(hiding_degree):
        b = rng.fr()
        coefs = add_coef(coefs, b, i)
        coefs = add_coef(coefs, (FR - b) % FR, zeroing_degree + i)
    return coefs


def z_evals(P, w_ext: np.ndarray, beta: int, gamma: int) -> np.ndarray:
    n = P["n"]
    out = np.empty((n, 4), dtype=np.uint64)
    rc = _lib().oracle_plonk_z_evals(_p(w_ext), _p(P["permutation"]), _p(A(P["vp"]["k"])), _p(P["group"]), _p(A([beta])), _p(A([gamma])),
                                     n, _p(out))
    if rc:
        raise MemoryError("oracle_plonk_z_evals")
    return out


def t_poly(P, w_polys, z_poly, alpha, beta, gamma, pi_poly) -> np.ndarray:
    m, vp = P["m"], P["vp"]
    k = vp["k"]
    co = lambda c: coset_fft(c, m, k[1])
    cols = [co(p) for p in w_polys] + list(P["q_coset"]) + [co(pi_poly), co(z_poly)] + list(P["s_coset"]) + \
           [P["coset_quotient"], P["l1_coset"], P["qb_coset"]] + list(P["q_prk_coset"])
    assert len(cols) == 28 and all(c.shape == (m, 4) for c in cols)
    sc = A(list(k) + [alpha, beta, gamma, vp["anemoi_generator"], vp["anemoi_generator_inv"]])
    out = np.empty((m, 4), dtype=np.uint64)
    _lib().oracle_plonk_quotient(_ptr_table(cols), _p(sc), _p(P["z_h_inv"]), P["factor"], m, _p(out))
    return trim(coset_ifft(out, m, inv_mod(k[1], FR)))


def split_t(rng, t: np.ndarray, n_pieces: int, n: int):
    """helpers.rs:1323-1408 without the commitments."""
    pieces, prev, L = [], 0, t.shape[0]
    for i in range(n_pieces):
        start, end = i * n, (L if i == n_pieces - 1 else (i + 1) * n)
        coefs = np.array(t[start:min(L, end)]) if start < L else np.zeros((0, 4), dtype=np.uint64)
        r = rng.fr()
        if i != n_pieces - 1:
            coefs = add_coef(coefs, r, n)
            coefs = add_coef(coefs, (FR - prev) % FR, 0)
        elif coefs.shape[0] == 0:
            coefs = A([(FR - prev) % FR])
        else:
            coefs = add_coef(coefs, (FR - prev) % FR, 0)
        prev = r
        pieces.append(trim(coefs))
    return pieces


def batch_prove(tr, pcs, polys, point, max_degree):
    pp._init_batch_eval(tr, max_degree, point)
    alpha = tr.challenge()
    terms, mult, const = [], 1, 0
    for p in polys:
        const = (const + mult * p_eval(p, point)) % FR
        terms.append((mult, p))
        mult = mult * alpha % FR
    h = add_coef(lincomb(terms), (FR - const) % FR, 0)
    quo, rem = div_linear(trim(h), point)
    assert rem == 0, "PCSProveEvalError"
    return pcs.commit(quo) if quo.shape[0] else None


def prover(rng, tr, pcs: CpuKzg, cs, P, witness):
    """Same arguments and the same proof dictionary as oracle.plonk_prover.prover (default feature set)."""
    n, vp = P["n"], P["vp"]
    k, root = vp["k"], P["root"]
    if isinstance(witness, np.ndarray):
        witness = witness.reshape(-1, 4)
        online = I(witness[list(cs.public_vars_witness_indices)]) if cs.public_vars_witness_indices else []
    else:
        online = [witness[i] for i in cs.public_vars_witness_indices]
    pp.transcript_init_plonk(tr, vp, online, root)
    pi_arr = np.zeros((n, 4), dtype=np.uint64)
    if online:
        pi_arr[list(cs.public_vars_constraint_indices)] = A(online)
    pi = ifft(pi_arr, n)
    w_ext = A(cs.extend_witness(witness))
    w_polys, cm_w = [], []
    for i in range(N_WIRES):
        f = hide_polynomial(rng, ifft(w_ext[i * n:(i + 1) * n], n), cs.get_hiding_degree(i), n)
        cm = pcs.commit(f)
        tr.point(cm)
        w_polys.append(f)
        cm_w.append(cm)
    beta = tr.challenge()
    tr.byte(0x01)
    gamma = tr.challenge()
    z = hide_polynomial(rng, ifft(z_evals(P, w_ext, beta, gamma), n), 3, n)
    cm_z = pcs.commit(z)
    tr.point(cm_z)
    alpha = tr.challenge()
    t = t_poly(P, w_polys, z, alpha, beta, gamma, pi)
    t_polys = split_t(rng, t, N_WIRES, n + 2)
    cm_t = [pcs.commit(p) for p in t_polys]
    for c in cm_t:
        tr.point(c)
    zeta = tr.challenge()
    w_ev = [p_eval(p, zeta) for p in w_polys]
    s_ev = [p_eval(p, zeta) for p in P["s_polys"][:N_WIRES - 1]]
    prk3, prk4 = p_eval(P["q_prk_polys"][2], zeta), p_eval(P["q_prk_polys"][3], zeta)
    zeta_omega = root * zeta % FR
    z_ev_omega = p_eval(z, zeta_omega)
    w_ev_omega = [p_eval(p, zeta_omega) for p in w_polys[:3]]
    for v in w_ev + s_ev:
        tr.fr(v)
    tr.fr(prk3)
    tr.fr(prk4)
    tr.fr(z_ev_omega)
    for v in w_ev_omega:
        tr.fr(v)
    u = tr.challenge()
    z_h_ev, l1_ev = pp.first_lagrange(zeta, n)
    src = {"q": P["q_polys"], "z": [z], "s_last": [P["s_polys"][N_WIRES - 1]], "qb": [P["qb_poly"]], "prk": P["q_prk_polys"], "t": t_polys}
    terms = pp.r_scalars(k, w_ev, s_ev, prk3, z_ev_omega, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n + 2)
    r = lincomb([(s_, src[name][i]) for s_, (name, i) in terms])
    open_zeta = w_polys + list(P["s_polys"][:N_WIRES - 1]) + [P["q_prk_polys"][2], P["q_prk_polys"][3], r]
    proof = {
        "cm_w_vec": cm_w, "cm_t_vec": cm_t, "cm_z": cm_z, "prk_3_poly_eval_zeta": prk3, "prk_4_poly_eval_zeta": prk4,
        "w_polys_eval_zeta": w_ev, "w_polys_eval_zeta_omega": w_ev_omega, "z_eval_zeta_omega": z_ev_omega, "s_polys_eval_zeta": s_ev,
    }
    proof["opening_witness_zeta"] = batch_prove(tr, pcs, open_zeta, zeta, n + 2)
    proof["opening_witness_zeta_omega"] = batch_prove(tr, pcs, [z, w_polys[0], w_polys[1], w_polys[2]], zeta_omega, n + 2)
    proof["_u"] = u
    return proof
