"""The reference's PlonK verifier for the `shuffle` feature set, restated with Python integers, INCLUDING the pairing check.
TEST INFRASTRUCTURE ONLY.

Purpose: pin the restatement (oracle/plonk_prover.py shares the transcript, the challenge order, the linearisation terms 1-5 and
the batch-opening algebra with this file) against something the reference itself produced: the golden proofs of
/root/reference/contracts/solidity/test/plonk_{20,52}.js, which the reference's verifier accepts.  If this verifier accepts them
too -- and rejects any modification -- transcript, challenges, r_eval_zeta, r_commitment, eval_pi_poly and the KZG batch check
are the reference's.

Follows
  verifier                      /root/reference/uzkge/src/plonk/verifier.rs:17-164 (+ compute_challenges :167-222), `shuffle` lines
  r_eval_zeta                   /root/reference/uzkge/src/plonk/helpers.rs:1181-1320 (terms 7-10 are the shuffle ones)
  r_poly_or_comm (commitments)  /root/reference/uzkge/src/plonk/helpers.rs:681-999 (parts 6-9 are the shuffle ones)
  eval_pi_poly                  /root/reference/uzkge/src/plonk/helpers.rs:1129-1165
  batch / batch_verify_diff_points  poly_commit/pcs.rs:170-191, kzg_poly_commitment.rs:407-460
  verify_shuffle (transcript prefix, public inputs)  /root/reference/shuffle/src/build_cs.rs:99-129
  PlonkProof::from_bytes_be     /root/reference/uzkge/src/plonk/indexer.rs:592-760
The fixture tests/golden/plonk_{20,52}_golden.json is produced by tests/golden/make_golden_proof.py.
"""
from __future__ import annotations

from . import bn254 as bn
from . import pairing
from .bn254 import FR
from .plonk_prover import Transcript, _g1_lin, _init_batch_eval, first_lagrange, transcript_init_plonk

N_WIRES = 5


def _pt(x: int, y: int):
    return None if x == 0 and y == 0 else (x, y)


def parse_proof(raw: bytes) -> dict:
    """PlonkProof::to_bytes_be with the `shuffle` feature (indexer.rs:538-590): 64-byte affine points (x BE, y BE), 32-byte scalars."""
    pos = 0

    def point():
        nonlocal pos
        x, y = int.from_bytes(raw[pos:pos + 32], "big"), int.from_bytes(raw[pos + 32:pos + 64], "big")
        pos += 64
        P = _pt(x, y)
        assert bn.g1_is_on_curve(P)
        return P

    def scalar():
        nonlocal pos
        v = int.from_bytes(raw[pos:pos + 32], "big")
        pos += 32
        assert v < FR
        return v

    p = {}
    p["cm_w_vec"] = [point() for _ in range(5)]
    p["cm_w_sel_vec"] = [point() for _ in range(3)]
    p["cm_t_vec"] = [point() for _ in range(5)]
    p["cm_z"] = point()
    p["prk_3_poly_eval_zeta"] = scalar()
    p["prk_4_poly_eval_zeta"] = scalar()
    p["w_polys_eval_zeta"] = [scalar() for _ in range(5)]
    p["w_polys_eval_zeta_omega"] = [scalar() for _ in range(3)]
    p["z_eval_zeta_omega"] = scalar()
    p["s_polys_eval_zeta"] = [scalar() for _ in range(4)]
    p["q_ecc_poly_eval_zeta"] = scalar()
    p["w_sel_polys_eval_zeta"] = [scalar() for _ in range(3)]
    p["opening_witness_zeta"] = point()
    p["opening_witness_zeta_omega"] = point()
    assert pos == len(raw) == 1632
    return p


def parse_vk(words: dict, public_key_commitments, pi_points, pi_lagrange) -> dict:
    """The generated VerifierKey_N.sol (offsets relative to CM_Q0_X_LOC, PlonkVerifier.sol:83-193) plus the run-time inputs."""
    w = lambda off: int(words.get(hex(off), "0x0"), 16)
    pts = lambda first, count: [_pt(w(first + 0x40 * i), w(first + 0x40 * i + 0x20)) for i in range(count)]
    pk = [int(x, 16) for x in public_key_commitments]
    return {
        "cm_q_vec": pts(0x0, 9), "cm_s_vec": pts(0x240, 5), "cm_qb": pts(0x380, 1)[0], "cm_prk_vec": pts(0x3C0, 4),
        "cm_q_ecc": pts(0x4C0, 1)[0], "cm_shuffle_generator_vec": pts(0x500, 12),
        "cm_shuffle_public_key_vec": [_pt(pk[2 * i], pk[2 * i + 1]) for i in range(12)],
        "anemoi_generator": w(0xB00), "anemoi_generator_inv": w(0xB20), "k": [w(0xB40 + 0x20 * i) for i in range(5)],
        "edwards_a": w(0xBE0), "root": w(0xC00), "cs_size": w(0xC20),
        "pi_points": pi_points,       # w^idx of every public-input row (VerifierKeyExtra1)
        "pi_lagrange": pi_lagrange,   # the Lagrange constants w^idx / n (VerifierKeyExtra2)
    }


def _g2(e):
    """EIP-197 order (x1, x0, y1, y0) -> ((x0, x1), (y0, y1))."""
    x1, x0, y1, y0 = (int(v, 16) for v in e)
    return ((x0, x1), (y0, y1))


def r_eval_zeta(p, alpha, beta, gamma, pi_ev, l1_ev, g, g_inv):
    a = [pow(alpha, i, FR) for i in range(17)]
    w, wo, s, ws = p["w_polys_eval_zeta"], p["w_polys_eval_zeta_omega"], p["s_polys_eval_zeta"], p["w_sel_polys_eval_zeta"]
    prk3, prk4, q_ecc = p["prk_3_poly_eval_zeta"], p["prk_4_poly_eval_zeta"], p["q_ecc_poly_eval_zeta"]
    term1 = alpha * p["z_eval_zeta_omega"] % FR
    for i in range(N_WIRES - 1):
        term1 = term1 * (w[i] + beta * s[i] + gamma) % FR
    term1 = term1 * (w[4] + gamma) % FR
    term2 = l1_ev * a[2] % FR
    w30, w21 = w[3] + w[0], w[2] + w[1]
    w320, w221 = w30 + w[0], w21 + w[1]
    tmp = (w30 + g * w21 + prk3) % FR
    term3 = a[6] * prk3 % FR * (pow(tmp - wo[2], 5, FR) + g * tmp * tmp - (w320 + g * w221)) % FR
    term5 = a[8] * prk3 % FR * (pow(tmp - wo[2], 5, FR) + g * wo[2] * wo[2] + g_inv - wo[0]) % FR
    g2p1 = (g * g + 1) % FR
    tmp = (g * w30 + g2p1 * w21 + prk4) % FR
    term4 = a[7] * prk3 % FR * (pow(tmp - w[4], 5, FR) + g * tmp * tmp - (g * w320 + g2p1 * w221)) % FR
    term6 = a[9] * prk3 % FR * (pow(tmp - w[4], 5, FR) + g * w[4] * w[4] + g_inv - wo[1]) % FR
    sel_00 = ((1 - ws[0]) * (1 - ws[1]) + q_ecc - 1) % FR
    sel_01 = ws[0] * (1 - ws[1]) % FR
    sel_10 = (1 - ws[0]) * ws[1] % FR
    sel_11 = ws[0] * ws[1] % FR
    term7 = ws[2] * (a[10] * wo[0] + a[11] * wo[1] + a[12] * wo[2] + a[13] * w[4]) % FR * (sel_00 + sel_01 + sel_10 + sel_11) % FR
    term8 = a[14] * (q_ecc * ws[0] % FR * (1 - ws[0]) + (1 - q_ecc) * ws[0]) % FR
    term9 = a[15] * (q_ecc * ws[1] % FR * (1 - ws[1]) + (1 - q_ecc) * ws[1]) % FR
    term10 = a[16] * q_ecc % FR * (1 - ws[2]) % FR * (1 + ws[2]) % FR
    return (term1 + term2 - pi_ev + term3 + term4 + term5 + term6 - term7 - term8 - term9 - term10) % FR


def r_terms(k, ed_a, p, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n_t_polys):
    """r_poly_or_comm (helpers.rs:681-999, `shuffle` feature) as (scalar, (family, index)) terms over the families
    q[0..9), z, s_last, qb, prk[0..2), pk[0..12), gen[0..12), t[0..5): the verifier sums commitments, the prover polynomials."""
    a = [pow(alpha, i, FR) for i in range(14)]
    w, wo, s, ws = p["w_polys_eval_zeta"], p["w_polys_eval_zeta_omega"], p["s_polys_eval_zeta"], p["w_sel_polys_eval_zeta"]
    prk3, q_ecc = p["prk_3_poly_eval_zeta"], p["q_ecc_poly_eval_zeta"]
    sel_mult = [w[0], w[1], w[2], w[3], w[0] * w[1], w[2] * w[3], 1, w[0] * w[1] * w[2] * w[3] * w[4], -w[4]]
    terms = [(sel_mult[i], ("q", i)) for i in range(9)]
    z_scalar = alpha
    for i in range(N_WIRES):
        z_scalar = z_scalar * (w[i] + k[i] * beta % FR * zeta + gamma) % FR
    z_scalar += l1_ev * a[2]
    terms.append((z_scalar, ("z", 0)))
    s_last = alpha * p["z_eval_zeta_omega"] % FR * beta % FR
    for i in range(N_WIRES - 1):
        s_last = s_last * (w[i] + beta * s[i] + gamma) % FR
    terms.append((-s_last, ("s_last", 0)))
    terms.append((w[1] * (w[1] - 1) * a[3] + w[2] * (w[2] - 1) * a[4] + w[3] * (w[3] - 1) * a[5], ("qb", 0)))
    terms.append((prk3 * a[6], ("prk", 0)))
    terms.append((prk3 * a[7], ("prk", 1)))
    sel = [((1 - ws[0]) * (1 - ws[1]) + q_ecc - 1) % FR, ws[0] * (1 - ws[1]) % FR, (1 - ws[0]) * ws[1] % FR, ws[0] * ws[1] % FR]
    for c in range(4):                                          # x: 0..3, y: 4..7, dxy: 8..11
        # 6. alpha^10: dxy * w0 w1 w0' - y * wsel2 w0 - x * w1                (public key)
        terms += [(a[10] * sel[c] % FR * (w[0] * w[1] % FR * wo[0]) % FR, ("pk", 8 + c)),
                  (-a[10] * sel[c] % FR * (ws[2] * w[0]) % FR, ("pk", 4 + c)), (-a[10] * sel[c] % FR * w[1] % FR, ("pk", c))]
        # 7. alpha^11: -dxy * w0 w1 w1' + x * a w0 - y * wsel2 w1
        terms += [(-a[11] * sel[c] % FR * (w[0] * w[1] % FR * wo[1]) % FR, ("pk", 8 + c)),
                  (a[11] * sel[c] % FR * (w[0] * ed_a) % FR, ("pk", c)), (-a[11] * sel[c] % FR * (ws[2] * w[1]) % FR, ("pk", 4 + c))]
        # 8. alpha^12: dxy * w2 w3 w2' - y * wsel2 w2 - x * w3                (generator)
        terms += [(a[12] * sel[c] % FR * (w[2] * w[3] % FR * wo[2]) % FR, ("gen", 8 + c)),
                  (-a[12] * sel[c] % FR * (ws[2] * w[2]) % FR, ("gen", 4 + c)), (-a[12] * sel[c] % FR * w[3] % FR, ("gen", c))]
        # 9. alpha^13: -dxy * w2 w3 w4 + x * a w2 - y * wsel2 w3
        terms += [(-a[13] * sel[c] % FR * (w[2] * w[3] % FR * w[4]) % FR, ("gen", 8 + c)),
                  (a[13] * sel[c] % FR * (w[2] * ed_a) % FR, ("gen", c)), (-a[13] * sel[c] % FR * (ws[2] * w[3]) % FR, ("gen", 4 + c))]
    f = pow(zeta, n_t_polys, FR)
    e = z_h_ev
    for i in range(N_WIRES):
        terms.append((-e, ("t", i)))
        e = e * f % FR
    return [(sc % FR, ref) for sc, ref in terms]


def r_commitment(vp, p, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n_t_polys):
    fam = {"q": vp["cm_q_vec"], "z": [p["cm_z"]], "s_last": [vp["cm_s_vec"][4]], "qb": [vp["cm_qb"]], "prk": vp["cm_prk_vec"],
           "pk": vp["cm_shuffle_public_key_vec"], "gen": vp["cm_shuffle_generator_vec"], "t": p["cm_t_vec"]}
    return _g1_lin([(sc, fam[name][i]) for sc, (name, i) in r_terms(vp["k"], vp["edwards_a"], p, alpha, beta, gamma, zeta, l1_ev, z_h_ev,
                                                                      n_t_polys)])


def _batch(tr, cms, max_degree, point, evals):
    _init_batch_eval(tr, max_degree, point)
    alpha = tr.challenge()
    mult, ev, terms = 1, 0, []
    for e, c in zip(evals, cms):
        terms.append((mult, c))
        ev = (ev + e * mult) % FR
        mult = mult * alpha % FR
    return _g1_lin(terms), ev


def verify_shuffle_proof(fixture: dict, pi_points, pi_lagrange, tamper=None) -> bool:
    """verify_shuffle (shuffle/src/build_cs.rs:99-129) + verifier (plonk/verifier.rs:17-164) on a golden fixture."""
    p = parse_proof(bytes.fromhex(fixture["proof"][2:]))
    pi = [int(x, 16) for x in fixture["public_inputs"]]
    if tamper:
        tamper(p, pi)
    vp = parse_vk(fixture["vk_words"], fixture["public_key_commitments"], pi_points, pi_lagrange)
    tr = Transcript(b"Plonk shuffle Proof")
    tr.u64(fixture["n_cards"])
    return verifier(tr, vp, pi, p, g2=(_g2(fixture["g2_tau_h_eip197"]), _g2(fixture["g2_h_eip197"])))


def verifier(tr, vp, pi, p, g2=None, trapdoor=None) -> bool:
    """plonk/verifier.rs:17-164 with the `shuffle` feature.  The final check is the pairing (g2 = (tau H, H)) or, for a synthetic SRS
    with a known trapdoor, its G1 equivalent  sum u^i W_i * tau == right-hand side."""
    n, root = vp["cs_size"], vp["root"]
    transcript_init_plonk(tr, vp, pi, root)
    for c in p["cm_w_vec"] + p["cm_w_sel_vec"]:
        tr.point(c)
    beta = tr.challenge()
    tr.byte(0x01)
    gamma = tr.challenge()
    tr.point(p["cm_z"])
    alpha = tr.challenge()
    for c in p["cm_t_vec"]:
        tr.point(c)
    zeta = tr.challenge()
    for v in p["w_polys_eval_zeta"] + p["s_polys_eval_zeta"] + p["w_sel_polys_eval_zeta"]:
        tr.fr(v)
    tr.fr(p["prk_3_poly_eval_zeta"])
    tr.fr(p["prk_4_poly_eval_zeta"])
    tr.fr(p["z_eval_zeta_omega"])
    tr.fr(p["q_ecc_poly_eval_zeta"])
    for v in p["w_polys_eval_zeta_omega"]:
        tr.fr(v)
    u = tr.challenge()
    z_h_ev, l1_ev = first_lagrange(zeta, n)
    # eval_pi_poly (helpers.rs:1129-1165): sum_i pi_i * c_i / (zeta - w^idx_i) * Z_H(zeta)
    pi_ev = 0
    for v, c, wp in zip(pi, vp["pi_lagrange"], vp["pi_points"]):
        pi_ev = (pi_ev + c * pow((zeta - wp) % FR, -1, FR) % FR * v) % FR
    pi_ev = pi_ev * z_h_ev % FR
    r_ev = r_eval_zeta(p, alpha, beta, gamma, pi_ev, l1_ev, vp["anemoi_generator"], vp["anemoi_generator_inv"])
    cm_r = r_commitment(vp, p, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n + 2)
    commitments = (p["cm_w_vec"] + vp["cm_s_vec"][:4] + [vp["cm_prk_vec"][2], vp["cm_prk_vec"][3], vp["cm_q_ecc"]] + p["cm_w_sel_vec"] + [cm_r])
    values = (p["w_polys_eval_zeta"] + p["s_polys_eval_zeta"] + [p["prk_3_poly_eval_zeta"], p["prk_4_poly_eval_zeta"], p["q_ecc_poly_eval_zeta"]]
              + p["w_sel_polys_eval_zeta"] + [r_ev])
    zeta_omega = zeta * root % FR
    comm, val = _batch(tr, commitments, n + 2, zeta, values)
    comm_o, val_o = _batch(tr, [p["cm_z"]] + p["cm_w_vec"][:3], n + 2, zeta_omega, [p["z_eval_zeta_omega"]] + p["w_polys_eval_zeta_omega"])
    # batch_verify_diff_points (kzg_poly_commitment.rs:407-460): e(sum u^i W_i, tau H) = e(sum u^i (z_i W_i - v_i G + C_i), H)
    W, Wo = p["opening_witness_zeta"], p["opening_witness_zeta_omega"]
    left = _g1_lin([(1, W), (u, Wo)])
    right = _g1_lin([(zeta, W), (u * zeta_omega, Wo), (-(val + u * val_o), (1, 2)), (1, comm), (u, comm_o)])
    if trapdoor is not None:
        return _g1_lin([(trapdoor, left)]) == right
    return pairing.multi_pairing_is_one([(left, g2[0]), (bn.g1_neg(right), g2[1])])
