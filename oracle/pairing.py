"""BN254 optimal ate pairing in plain Python integers.  TEST INFRASTRUCTURE ONLY.

Needed for one thing: checking the reference's own golden proofs (contracts/solidity/test/plonk_20.js, plonk_52.js) with the
restated verifier (oracle/plonk_verifier_shuffle.py), i.e. KZGCommitmentSchemeBN254::batch_verify_diff_points
(/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:407-460: `Bn254::multi_pairing`).  The construction is the textbook
one (as in py_ecc's bn128): Fp12 = Fp[w] / (w^12 - 18 w^6 + 82), G2 points on the sextic twist y^2 = x^3 + 3 / (9 + i) over
Fp2 = Fp[i] / (i^2 + 1) mapped into Fp12, Miller loop over 6u + 2 with the two Frobenius correction steps, final exponentiation by
(p^12 - 1) / r.  Slow (about two seconds per pairing) and only ever run on a handful of points.
"""
from __future__ import annotations

from .bn254 import FQ as P, FR as R

ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63
FQ12_MOD = (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0)   # w^12 = 18 w^6 - 82


class Fp12:
    """Polynomials of degree < 12 in w over Fp; the subfields Fp2 and Fp sit inside (i = w^6 - 9)."""

    __slots__ = ("c",)

    def __init__(self, c):
        self.c = tuple(int(x) % P for x in c)
        assert len(self.c) == 12

    @classmethod
    def one(cls):
        return cls((1,) + (0,) * 11)

    @classmethod
    def zero(cls):
        return cls((0,) * 12)

    @classmethod
    def from_int(cls, v):
        return cls((v,) + (0,) * 11)

    def __eq__(self, o):
        return isinstance(o, Fp12) and self.c == o.c

    def __hash__(self):
        return hash(self.c)

    def __add__(self, o):
        return Fp12(a + b for a, b in zip(self.c, o.c))

    def __sub__(self, o):
        return Fp12(a - b for a, b in zip(self.c, o.c))

    def __neg__(self):
        return Fp12(-a for a in self.c)

    def scale(self, k: int):
        return Fp12(a * k for a in self.c)

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        b = [0] * 23
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        # reduce with w^12 = 18 w^6 - 82
        for k in range(22, 11, -1):
            t = b[k]
            if t:
                b[k - 6] += 18 * t
                b[k - 12] -= 82 * t
        return Fp12(b[:12])

    def is_zero(self):
        return not any(self.c)

    def inv(self):
        """Extended Euclid on polynomials over Fp."""
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = list(self.c) + [0], list(x % P for x in FQ12_MOD) + [1]

        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        def poly_rounded_div(a, b):
            dega, degb = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            for i in range(dega - degb, -1, -1):
                q = temp[degb + i] * pow(b[degb], -1, P) % P
                o[i] = (o[i] + q) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - q * b[c]) % P
            return o[: deg(o) + 1]

        while deg(low):
            r = poly_rounded_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                    new[i + j] = (new[i + j] - low[i] * r[j]) % P
            lm, low, hm, high = nm, new, lm, low
        inv0 = pow(low[0], -1, P)
        return Fp12(x * inv0 for x in lm[:12])

    def __truediv__(self, o):
        return self * o.inv()

    def __pow__(self, e: int):
        result, base = Fp12.one(), self
        while e:
            if e & 1:
                result = result * base
            base = base * base
            e >>= 1
        return result


W = Fp12((0, 1) + (0,) * 10)


def fp2_to_fp12(a0: int, a1: int) -> Fp12:
    """a0 + a1 i with i = w^6 - 9."""
    return Fp12((a0 - 9 * a1, 0, 0, 0, 0, 0, a1, 0, 0, 0, 0, 0))


def twist(Q):
    """G2 point ((x0, x1), (y0, y1)) on the twist -> point on y^2 = x^3 + 3 over Fp12."""
    if Q is None:
        return None
    (x0, x1), (y0, y1) = Q
    return (fp2_to_fp12(x0, x1) * (W * W), fp2_to_fp12(y0, y1) * (W * W * W))


def cast_g1(Pt):
    if Pt is None:
        return None
    return (Fp12.from_int(Pt[0]), Fp12.from_int(Pt[1]))


def _double(pt):
    x, y = pt
    m = (x * x).scale(3) / y.scale(2)
    nx = m * m - x.scale(2)
    return (nx, m * (x - nx) - y)


def _add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            return _double(p1)
        return None
    m = (y2 - y1) / (x2 - x1)
    nx = m * m - x1 - x2
    return (nx, m * (x1 - nx) - y1)


def _linefunc(p1, p2, t):
    """The line through p1 and p2 evaluated at t."""
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) / y1.scale(2)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def _frobenius(pt):
    return (pt[0] ** P, pt[1] ** P)


def miller_loop(Q, Pt) -> Fp12:
    """Q on the twisted curve over Fp12, Pt a G1 point cast into Fp12; without the final exponentiation."""
    if Q is None or Pt is None:
        return Fp12.one()
    Rr, f = Q, Fp12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(Rr, Rr, Pt)
        Rr = _double(Rr)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(Rr, Q, Pt)
            Rr = _add(Rr, Q)
    Q1 = _frobenius(Q)
    nQ2 = _frobenius(Q1)
    nQ2 = (nQ2[0], -nQ2[1])
    f = f * _linefunc(Rr, Q1, Pt)
    Rr = _add(Rr, Q1)
    f = f * _linefunc(Rr, nQ2, Pt)
    return f


def final_exponentiate(f: Fp12) -> Fp12:
    return f ** ((P ** 12 - 1) // R)


def pairing(Q, Pt) -> Fp12:
    """e(Pt, Q) for Pt in G1 (affine ints or None), Q in G2 (((x0, x1), (y0, y1)) or None)."""
    return final_exponentiate(miller_loop(twist(Q), cast_g1(Pt)))


def multi_pairing_is_one(pairs) -> bool:
    """prod_i e(P_i, Q_i) == 1 with ONE final exponentiation (Bn254::multi_pairing(..) == Fp12::one())."""
    f = Fp12.one()
    for Pt, Q in pairs:
        f = f * miller_loop(twist(Q), cast_g1(Pt))
    return final_exponentiate(f) == Fp12.one()


# the G2 generator of ark-bn254 / EIP-197 (x = x0 + x1 i, y = y0 + y1 i)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def g2_is_on_twist(Q) -> bool:
    """y^2 = x^3 + 3 / (9 + i) over Fp2."""
    (x0, x1), (y0, y1) = Q
    x, y = fp2_to_fp12(x0, x1), fp2_to_fp12(y0, y1)
    b2 = fp2_to_fp12(3, 0) / fp2_to_fp12(9, 1)
    return y * y - x * x * x == b2
