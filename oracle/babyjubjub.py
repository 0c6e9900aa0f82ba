"""CPU restatement of the reference's shuffle primitives over Baby Jubjub (ark-ed-on-bn254).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/uzkge/src/shuffle/remark.rs:39-84 (crate_generators / crate_public_keys), :149-231
(eval_remark_with_trace) and /root/reference/uzkge/src/shuffle/mod.rs:17-69 (ElGamal ciphertexts).  The curve constants are not
in the reference repository (they live in the un-vendored crate ark-ed-on-bn254 0.4): a = 1, d and the generator below are pinned
by the reference's preprocessed tables shuffle/babyjubjub.rs:24-3566 (tests/golden/babyjubjub_generators.json): the first table
entry is the generator, dxy / (x y) gives d, and the curve equation at the generator gives a.
Points are affine integer pairs in extended-coordinate arithmetic (X : Y : Z : T) internally; the identity is (0, 1).
"""
from __future__ import annotations

from .bn254 import FR as Q          # Baby Jubjub's base field is BN254's scalar field

A = 1
D = 9706598848417545097372247223557719406784115219466060233080913168975159366771
GEN = (19698561148652590122159747500897617769866003486955115824547446575314762165298,
       19298250018296453272277890825869354524455968081175474282777126169995084727839)
ORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041
NUM_ITERATIONS = 84          # shuffle/babyjubjub.rs:22
N_SELECT_BITS = 4            # shuffle/mod.rs:15


def _ext(p):
    return (p[0], p[1], 1, p[0] * p[1] % Q)


def _ext_add(p, q):
    """add-2008-hwcd for a x^2 + y^2 = 1 + d x^2 y^2 (extended coordinates, T = XY / Z)."""
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = q
    a_ = x1 * x2 % Q
    b_ = y1 * y2 % Q
    c_ = D * t1 % Q * t2 % Q
    d_ = z1 * z2 % Q
    e_ = ((x1 + y1) * (x2 + y2) - a_ - b_) % Q
    f_ = (d_ - c_) % Q
    g_ = (d_ + c_) % Q
    h_ = (b_ - A * a_) % Q
    return (e_ * f_ % Q, g_ * h_ % Q, f_ * g_ % Q, e_ * h_ % Q)


def _affine(p):
    zi = pow(p[2], -1, Q)
    return (p[0] * zi % Q, p[1] * zi % Q)


def add(p, q):
    return _affine(_ext_add(_ext(p), _ext(q)))


def neg(p):
    return (-p[0] % Q, p[1])


def mul(k, p):
    acc, base = (0, 1, 1, 0), _ext(p)
    k %= ORDER
    while k:
        if k & 1:
            acc = _ext_add(acc, base)
        base = _ext_add(base, base)
        k >>= 1
    return _affine(acc)


def on_curve(p):
    x, y = p
    return (A * x * x + y * y - 1 - D * x * x * y * y) % Q == 0


def segments(base):
    """remark.rs:39-84: round i holds (j + 1) 16^i base, j < 4."""
    out, g = [], _ext(base)
    for _ in range(NUM_ITERATIONS):
        seg, cur = [], g
        for _ in range(N_SELECT_BITS):
            seg.append(_affine(cur))
            cur = _ext_add(cur, g)
        for _ in range(N_SELECT_BITS):
            g = _ext_add(g, g)
        out.append(seg)
    return out


def remark_trace(card, bits, pk):
    """remark.rs:149-231.  card = (e1, e2); bits: NUM_ITERATIONS triples of booleans.  Returns (field_bits, intermediate_values):
    field_bits[i] = [b0, b1, +-1], intermediate_values[i] = [c2.x, c2.y, c1.x, c1.y] after round i."""
    gens, pks = segments(GEN), segments(pk)
    c1, c2 = card
    fb, iv = [], []
    for b, gseg, pseg in zip(bits, gens, pks):
        j = (1 if b[0] else 0) + (2 if b[1] else 0)
        g_, p_ = (gseg[j], pseg[j]) if b[2] else (neg(gseg[j]), neg(pseg[j]))
        c1, c2 = add(c1, g_), add(c2, p_)
        fb.append([1 if b[0] else 0, 1 if b[1] else 0, 1 if b[2] else Q - 1])
        iv.append([c2[0], c2[1], c1[0], c1[1]])
    return fb, iv


def encrypt(r, m, pk):
    """shuffle/mod.rs:43-51: (r G, M + r pk)."""
    return (mul(r, GEN), add(m, mul(r, pk)))


def decrypt_ok(card, m, sk):
    """shuffle/mod.rs:53-55."""
    return m == add(card[1], neg(mul(sk, card[0])))
