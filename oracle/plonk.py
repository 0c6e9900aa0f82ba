"""CPU restatement (Python big integers) of the pointwise quotient map of the TurboPlonK prover.  TEST INFRASTRUCTURE ONLY.

Follows t_poly's loop body, /root/reference/uzkge/src/plonk/helpers.rs:284-669, for the build without the `shuffle`
feature (terms 1-11, helpers.rs:643-652), and the gate function TurboCS::eval_gate_func,
/root/reference/uzkge/src/plonk/constraint_system/turbo/mod.rs:193-222.  All values are canonical integers mod r.
"""
from __future__ import annotations

from .bn254 import FR, inv_mod


def eval_gate_func(w, q, pub_input):
    """turbo/mod.rs:193-222: (w1, w2, w3, w4, w1*w2, w3*w4, 1, w1*w2*w3*w4*wo, -wo) . selectors + public input."""
    r = q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3]
    r += q[4] * w[0] * w[1] + q[5] * w[2] * w[3] + q[6] + pub_input
    r += q[7] * w[0] * w[1] * w[2] * w[3] * w[4]
    r -= q[8] * w[4]
    return r % FR


def z_h_inv_coset_evals(k1: int, group_gen_m: int, n: int, factor: int):
    """helpers.rs:244-253: 1 / ((k1 w_m^i)^n - 1), i < factor."""
    out, mult, step = [], pow(k1, n, FR), pow(group_gen_m, n, FR)
    for _ in range(factor):
        out.append(inv_mod((mult - 1) % FR, FR))
        mult = mult * step % FR
    return out


def shuffle_terms(wv, w0n, w1n, w2n, ws, q_ecc, pk, gen, edwards_a, alpha):
    """Terms 12-18 of t_poly's numerator (helpers.rs:416-640, `shuffle` feature) at one point.  ws: the 3 witness-selector values,
    pk / gen: the 12 public-key / generator selector values (x_00..x_11, y_00..y_11, dxy_00..dxy_11)."""
    a = [pow(alpha, i, FR) for i in range(17)]
    sel = [((1 - ws[0]) * (1 - ws[1]) + q_ecc - 1) % FR, ws[0] * (1 - ws[1]) % FR, (1 - ws[0]) * ws[1] % FR, ws[0] * ws[1] % FR]
    t12 = t13 = t14 = t15 = 0
    for c in range(4):
        t12 += sel[c] * (ws[2] * w0n - ws[2] * wv[0] * pk[4 + c] - wv[1] * pk[c] + wv[0] * wv[1] * w0n * pk[8 + c])
        t13 += sel[c] * (ws[2] * w1n + wv[0] * edwards_a * pk[c] - ws[2] * wv[1] * pk[4 + c] - wv[0] * wv[1] * w1n * pk[8 + c])
        t14 += sel[c] * (ws[2] * w2n - ws[2] * wv[2] * gen[4 + c] - wv[3] * gen[c] + wv[2] * wv[3] * w2n * gen[8 + c])
        t15 += sel[c] * (ws[2] * wv[4] + wv[2] * edwards_a * gen[c] - ws[2] * wv[3] * gen[4 + c] - wv[2] * wv[3] * wv[4] * gen[8 + c])
    t16 = q_ecc * ws[0] * (1 - ws[0]) + (1 - q_ecc) * ws[0]
    t17 = q_ecc * ws[1] * (1 - ws[1]) + (1 - q_ecc) * ws[1]
    t18 = q_ecc * (1 + ws[2]) * (1 - ws[2])
    return (a[10] * t12 + a[11] * t13 + a[12] * t14 + a[13] * t15 + a[14] * t16 + a[15] * t17 + a[16] * t18) % FR


def quotient_coset_evals(w, q, pi, z, s, coset_quotient, l1, qb, q_prk, k, alpha, beta, gamma, g, g_inv, z_h_inv, factor, shuffle=None):
    """helpers.rs:284-669.  w: 5 lists of m values, q: 9, s: 5, q_prk: 4; the rest lists of m values or scalars.
    shuffle: None (default feature set, terms 1-11) or a dict {w_sel: 3 lists, q_ecc: list, pk: 12 lists, gen: 12 lists,
    edwards_a: scalar} adding terms 12-18."""
    m = len(z)
    a = [pow(alpha, i, FR) for i in range(10)]
    g2p1 = (g * g + 1) % FR
    out = []
    for p in range(m):
        pn = (p + factor) % m
        wv = [w[j][p] for j in range(5)]
        term1 = eval_gate_func(wv, [q[j][p] for j in range(9)], pi[p])
        term2 = alpha * z[p]
        term3 = alpha * z[pn]
        for j in range(5):
            term2 = term2 * (wv[j] + gamma + beta * k[j] * coset_quotient[p]) % FR
            term3 = term3 * (wv[j] + gamma + beta * s[j][p]) % FR
        term4 = a[2] * l1[p] * (z[p] - 1)
        term5 = a[3] * qb[p] * wv[1] * (wv[1] - 1)
        term6 = a[4] * qb[p] * wv[2] * (wv[2] - 1)
        term7 = a[5] * qb[p] * wv[3] * (wv[3] - 1)
        w0n, w1n, w2n = w[0][pn], w[1][pn], w[2][pn]
        prk1, prk2, prk3, prk4 = (q_prk[j][p] for j in range(4))
        w30, w21 = wv[0] + wv[3], wv[1] + wv[2]
        w320, w221 = wv[0] + w30, wv[1] + w21
        tmp = (w30 + g * w21 + prk3) % FR
        term8 = a[6] * prk3 * (pow(tmp - w2n, 5, FR) + g * tmp * tmp - (w320 + g * w221 + prk1))
        term10 = a[8] * prk3 * (pow(tmp - w2n, 5, FR) + g * w2n * w2n + g_inv - w0n)
        tmp = (g * w30 + g2p1 * w21 + prk4) % FR
        term9 = a[7] * prk3 * (pow(tmp - wv[4], 5, FR) + g * tmp * tmp - (g * w320 + g2p1 * w221 + prk2))
        term11 = a[9] * prk3 * (pow(tmp - wv[4], 5, FR) + g * wv[4] * wv[4] + g_inv - w1n)
        num = term1 + term2 + (term4 - term3) + term5 + term6 + term7 - term8 - term9 - term10 - term11
        if shuffle is not None:
            sh = shuffle
            num += shuffle_terms(wv, w0n, w1n, w2n, [sh["w_sel"][j][p] for j in range(3)], sh["q_ecc"][p], [sh["pk"][j][p] for j in range(12)],
                                 [sh["gen"][j][p] for j in range(12)], sh["edwards_a"], alpha)
        out.append(num % FR * z_h_inv[p % factor] % FR)
    return out
