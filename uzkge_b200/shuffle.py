"""Host mirror of the reference's shuffle gadgets: the callers that put non-zero data into the `shuffle` feature set's selector
polynomials.  Plain Python integers; nothing here touches the GPU -- these functions build the circuit and the
witness that `plonk.indexer` / `plonk.prover` / `plonk.refresh_prover_params_public_key` then process on the device.

  Ciphertext, N_SELECT_BITS            /root/reference/uzkge/src/shuffle/mod.rs:17-69
  Remark (BabyJubjubShuffle)           /root/reference/uzkge/src/shuffle/remark.rs:10-232, shuffle/babyjubjub.rs:15-22
  RemarkTrace                          /root/reference/uzkge/src/shuffle/trace.rs:8-18
  Permutation                          /root/reference/uzkge/src/shuffle/permutation.rs:4-41
  CardVar, new_card_variable, ...      /root/reference/uzkge/src/plonk/constraint_system/shuffle/mod.rs:13-85
  eval_card_remark                     /root/reference/uzkge/src/plonk/constraint_system/shuffle/remark.rs:11-94
  shuffle_card                         /root/reference/uzkge/src/plonk/constraint_system/shuffle/permutation.rs:8-216
  load_shuffle_remark_parameters, compute_*_selectors, verify_witness
                                       /root/reference/uzkge/src/plonk/constraint_system/turbo/mod.rs:155-191, 310-364, 905-966, 1041-1396
  build_cs                             /root/reference/shuffle/src/build_cs.rs:26-56

The curve is ark-ed-on-bn254 (Baby Jubjub in the a = 1 form) over BN254's scalar field: a x^2 + y^2 = 1 + d x^2 y^2.  a, d and the
generator are pinned by the reference's preprocessed tables (tests/golden/babyjubjub_generators.json, tests/test_shuffle_host.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .rng import FR_MODULUS as FQ      # the base field of Baby Jubjub is BN254's Fr

COEFF_A = 1
COEFF_D = 9706598848417545097372247223557719406784115219466060233080913168975159366771
GENERATOR = (19698561148652590122159747500897617769866003486955115824547446575314762165298,
             19298250018296453272277890825869354524455968081175474282777126169995084727839)
SUBGROUP_ORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041   # ark_ed_on_bn254::Fr
IDENTITY = (0, 1)
N_SELECT_BITS = 4
N_WIRE_SELECTORS = 3


# ---------------------------------------------------------------------------------------------------------- twisted Edwards arithmetic
def ed_add(p, q):
    """The unified affine addition law (complete on this curve: d is a non-residue)."""
    x1, y1 = p
    x2, y2 = q
    t = COEFF_D * x1 % FQ * x2 % FQ * y1 % FQ * y2 % FQ
    x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, FQ) % FQ
    y3 = (y1 * y2 - COEFF_A * x1 * x2) * pow(1 - t, -1, FQ) % FQ
    return (x3, y3)


def _ext(p):
    return (p[0], p[1], 1, p[0] * p[1] % FQ)


def _ext_add(p, q):
    """The same law in extended coordinates (X : Y : Z : T), T = XY / Z: no inversion per addition."""
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = q
    a = x1 * x2 % FQ
    b = y1 * y2 % FQ
    c = COEFF_D * t1 % FQ * t2 % FQ
    d = z1 * z2 % FQ
    e = ((x1 + y1) * (x2 + y2) - a - b) % FQ
    f, g, h = (d - c) % FQ, (d + c) % FQ, (b - COEFF_A * a) % FQ
    return (e * f % FQ, g * h % FQ, f * g % FQ, e * h % FQ)


def _batch_affine(points):
    """Extended -> affine for a list of points with ONE modular inversion (prefix products)."""
    prefix, acc = [], 1
    for p in points:
        prefix.append(acc)
        acc = acc * p[2] % FQ
    inv = pow(acc, -1, FQ)
    out = [None] * len(points)
    for i in range(len(points) - 1, -1, -1):
        zi = inv * prefix[i] % FQ
        inv = inv * points[i][2] % FQ
        out[i] = (points[i][0] * zi % FQ, points[i][1] * zi % FQ)
    return out


def ed_neg(p):
    return ((-p[0]) % FQ, p[1])


def ed_mul(k: int, p):
    acc, base = (0, 1, 1, 0), _ext(p)
    k %= SUBGROUP_ORDER
    while k:
        if k & 1:
            acc = _ext_add(acc, base)
        base = _ext_add(base, base)
        k >>= 1
    return _batch_affine([acc])[0]


def ed_is_on_curve(p) -> bool:
    x, y = p
    return (COEFF_A * x * x + y * y - 1 - COEFF_D * x * x % FQ * y * y) % FQ == 0


def _rand_scalar(prng) -> int:
    """A scalar below the subgroup order from 4 x next_u64 (top bits cleared, rejection) -- the shape of arkworks' Fp::rand."""
    while True:
        v = 0
        for i in range(4):
            v |= prng.next_u64() << (64 * i)
        v &= (1 << 251) - 1
        if v < SUBGROUP_ORDER:
            return v


def rand_point(prng):
    """A random element of the prime-order subgroup.  (arkworks' EdwardsProjective::rand draws a coordinate and clears the
    cofactor; only the point's distribution differs, nothing downstream depends on it.)"""
    return ed_mul(_rand_scalar(prng) or 1, GENERATOR)


@dataclass(frozen=True)
class Ciphertext:
    """shuffle/mod.rs:17-69: ElGamal (e1 = r G, e2 = M + r pk) over Baby Jubjub, affine points as integer pairs."""
    e1: tuple
    e2: tuple

    @classmethod
    def encrypt(cls, prng, m, pk) -> "Ciphertext":
        r = _rand_scalar(prng)
        return cls(ed_mul(r, GENERATOR), ed_add(m, ed_mul(r, pk)))

    @classmethod
    def rand(cls, prng) -> "Ciphertext":
        return cls.encrypt(prng, rand_point(prng), rand_point(prng))

    def verify(self, m, sk: int) -> bool:
        return m == ed_add(self.e2, ed_neg(ed_mul(sk, self.e1)))

    def get_first(self):
        return self.e1

    def get_second(self):
        return self.e2

    def flatten(self) -> list[int]:
        return [self.e2[0], self.e2[1], self.e1[0], self.e1[1]]


MaskedCard = Ciphertext       # shuffle/src/lib.rs: `pub type MaskedCard = Ciphertext<EdwardsProjective>`


@dataclass
class RemarkTrace:
    """shuffle/trace.rs:8-18."""
    bits: list = field(default_factory=list)                  # per round [s1, s2, s3] with s1, s2 in {0, 1}, s3 in {1, -1}
    intermediate_values: list = field(default_factory=list)   # per round [c2.x, c2.y, c1.x, c1.y]
    output: list = field(default_factory=list)
    n_round: int = 0


class BabyJubjubShuffle:
    """shuffle/remark.rs `trait Remark` with shuffle/babyjubjub.rs's constants.  Round i adds +-(j + 1) 16^i G to the first
    component and +-(j + 1) 16^i pk to the second, j in 0..4 chosen by two bits and the sign by the third."""
    COFF_A = COEFF_A
    COFF_D = COEFF_D
    NUM_ITERATIONS = 84
    _generators = None

    @classmethod
    def sample_random_scalar_bits(cls, prng) -> list:
        """remark.rs:19-27: `rng.gen::<[bool; 3]>()` per round; rand 0.8 draws a bool as the sign bit of one next_u32."""
        return [[bool(prng.next_u32() >> 31) for _ in range(N_WIRE_SELECTORS)] for _ in range(cls.NUM_ITERATIONS)]

    @classmethod
    def _segments(cls, base) -> list:
        flat, g = [], _ext(base)
        for _ in range(cls.NUM_ITERATIONS):
            cur = g
            for _ in range(N_SELECT_BITS):
                flat.append(cur)
                cur = _ext_add(cur, g)
            for _ in range(N_SELECT_BITS):
                g = _ext_add(g, g)
        aff = _batch_affine(flat)
        return [aff[i:i + N_SELECT_BITS] for i in range(0, len(aff), N_SELECT_BITS)]

    @classmethod
    def crate_generators(cls) -> list:
        """remark.rs:39-60."""
        if cls._generators is None:
            cls._generators = cls._segments(GENERATOR)
        return cls._generators

    @classmethod
    def crate_public_keys(cls, pk) -> list:
        """remark.rs:63-84."""
        return cls._segments(pk)

    @classmethod
    def get_preprocessed_generators_x(cls) -> list:
        return [[p[0] for p in seg] for seg in cls.crate_generators()]

    @classmethod
    def get_preprocessed_generators_y(cls) -> list:
        return [[p[1] for p in seg] for seg in cls.crate_generators()]

    @classmethod
    def get_preprocessed_generators_dxy(cls) -> list:
        return [[COEFF_D * p[0] % FQ * p[1] % FQ for p in seg] for seg in cls.crate_generators()]

    @classmethod
    def eval_remark_with_trace(cls, card: Ciphertext, r_bits, pk, pks=None) -> RemarkTrace:
        """remark.rs:149-231.  pks: crate_public_keys(pk) when the caller already has it."""
        if len(r_bits) != cls.NUM_ITERATIONS:
            raise ValueError("r_bits must hold NUM_ITERATIONS entries")
        gens = cls.crate_generators()
        pks = pks if pks is not None else cls.crate_public_keys(pk)
        c1, c2 = _ext(card.get_first()), _ext(card.get_second())
        trace = RemarkTrace(n_round=cls.NUM_ITERATIONS)
        chain = []
        for bits, gen, pkseg in zip(r_bits, gens, pks):
            j = int(bool(bits[0])) + 2 * int(bool(bits[1]))
            g_, p_ = (gen[j], pkseg[j]) if bits[2] else (ed_neg(gen[j]), ed_neg(pkseg[j]))
            c1, c2 = _ext_add(c1, _ext(g_)), _ext_add(c2, _ext(p_))
            trace.bits.append([int(bool(bits[0])), int(bool(bits[1])), 1 if bits[2] else FQ - 1])
            chain += [c2, c1]
        aff = _batch_affine(chain)
        for i in range(0, len(aff), 2):
            trace.intermediate_values.append([aff[i][0], aff[i][1], aff[i + 1][0], aff[i + 1][1]])
        trace.output = list(trace.intermediate_values[-1])
        return trace

    @classmethod
    def eval_remark(cls, card: Ciphertext, r_bits, pk) -> Ciphertext:
        """remark.rs:87-146."""
        out = cls.eval_remark_with_trace(card, r_bits, pk).output
        return Ciphertext((out[2], out[3]), (out[0], out[1]))


class Permutation:
    """shuffle/permutation.rs:4-41: an n x n 0/1 matrix with one 1 per row and column."""

    def __init__(self, matrix):
        self.matrix = matrix

    @classmethod
    def from_indices(cls, indices) -> "Permutation":
        n = len(indices)
        if sorted(indices) != list(range(n)):
            raise ValueError("not a permutation")
        return cls([[1 if indices[i] == j else 0 for j in range(n)] for i in range(n)])

    @classmethod
    def rand(cls, prng, n: int) -> "Permutation":
        """permutation.rs:8-23: draw without replacement with `gen_range(0..remainder.len())`.  rand 0.8's single-sample rule for
        a 64-bit range (widening multiply, zone = (range << lzcnt) - 1) is restated from memory: it only decides WHICH
        permutation is drawn, and no reference fixture depends on it."""
        remainder, idx = list(range(n)), []
        for _ in range(n):
            rng_len = len(remainder)
            zone = ((rng_len << (64 - rng_len.bit_length())) - 1) & 0xFFFFFFFFFFFFFFFF
            while True:
                prod = prng.next_u64() * rng_len
                if (prod & 0xFFFFFFFFFFFFFFFF) <= zone:
                    r = prod >> 64
                    break
            idx.append(remainder.pop(r))
        return cls.from_indices(idx)

    def __len__(self) -> int:
        return len(self.matrix)

    def get_matrix(self):
        return self.matrix

    def sanity_check(self) -> None:
        n = len(self.matrix)
        assert all(sum(row) == 1 for row in self.matrix)
        assert all(sum(self.matrix[i][j] for i in range(n)) == 1 for j in range(n))


_SEL_CODES = {0: 0, 1: 1, FQ - 1: 2}


class CardVar(list):
    """constraint_system/shuffle/mod.rs:13-62: the 4 variable indices of a card, [second.x, second.y, first.x, first.y]."""

    def get_raw(self):
        return list(self)

    def get_first_x(self):
        return self[0]

    def get_first_y(self):
        return self[1]

    def get_second_x(self):
        return self[2]

    def get_second_y(self):
        return self[3]


class ShuffleGates:
    """The shuffle methods of TurboCS (mixed into plonk.TurboCS).  The host class supplies new_variable, witness, size, zero_var,
    one_var, _push_gate, insert_lc_gate, equal, prepare_pi_variable, attach_boolean_constraint_to_gate."""

    def _init_shuffle(self) -> None:
        self.edwards_a = 0
        self.n_iteration_shuffle_scalar_mul = 0
        self.shuffle_public_keys = None       # [round][4] -> (x, y, dxy)
        self.shuffle_generators = None
        self.shuffle_remark_constraints: list = []     # (first gate, [s1 list, s2 list, s3 list])
        self._remark_sel_codes: list = []              # the same columns as indices into (0, 1, -1), or None for other values

    # ---- turbo/mod.rs:639-662
    def linear_combine(self, wires_in, q1: int, q2: int, q3: int, q4: int) -> int:
        w = self.witness
        out = self.new_variable(w[wires_in[0]] * q1 + w[wires_in[1]] * q2 + w[wires_in[2]] * q3 + w[wires_in[3]] * q4)
        self.insert_lc_gate(wires_in, out, q1, q2, q3, q4)
        return out

    # ---- constraint_system/shuffle/mod.rs:64-85
    def new_card_variable(self, card: Ciphertext) -> CardVar:
        fx, fy = card.get_first()
        sx, sy = card.get_second()
        first_x, first_y = self.new_variable(fx), self.new_variable(fy)
        second_x, second_y = self.new_variable(sx), self.new_variable(sy)
        return CardVar([second_x, second_y, first_x, first_y])

    def prepare_pi_card_variable(self, card_var) -> None:
        for var in card_var:
            self.prepare_pi_variable(var)

    # ---- turbo/mod.rs:905-966
    def load_shuffle_remark_parameters(self, shuffle_pk, remark=BabyJubjubShuffle) -> None:
        def table(segments):
            return [[(p[0], p[1], remark.COFF_D * p[0] % FQ * p[1] % FQ) for p in seg] for seg in segments]

        self.shuffle_public_keys = table(remark.crate_public_keys(shuffle_pk))
        self.shuffle_generators = table(remark.crate_generators())
        self.edwards_a = remark.COFF_A
        self.n_iteration_shuffle_scalar_mul = remark.NUM_ITERATIONS

    def attach_shuffle_remark_constraints_to_gate(self, wiring_selectors) -> None:
        if len(wiring_selectors) != N_WIRE_SELECTORS or any(len(x) != self.n_iteration_shuffle_scalar_mul for x in wiring_selectors):
            raise ValueError("one value per iteration and wire selector expected")
        self.shuffle_remark_constraints.append((self.size, [list(x) for x in wiring_selectors]))
        codes = [[_SEL_CODES.get(v % FQ, -1) for v in x] for x in wiring_selectors]
        self._remark_sel_codes.append(None if any(c < 0 for x in codes for c in x) else np.asarray(codes, dtype=np.int32))

    def shuffle_remark_constraint_indices(self) -> list:
        return [i for i, _ in self.shuffle_remark_constraints]

    # ---- constraint_system/shuffle/remark.rs:11-94
    def eval_card_remark(self, trace: RemarkTrace, input_var) -> CardVar:
        if not (len(trace.bits) == len(trace.intermediate_values) == trace.n_round == self.n_iteration_shuffle_scalar_mul):
            raise ValueError("trace length does not match the loaded remark parameters")
        self.attach_shuffle_remark_constraints_to_gate([[b[i] for b in trace.bits] for i in range(N_WIRE_SELECTORS)])
        iv = [[self.new_variable(x) for x in values] for values in trace.intermediate_values]
        zero4 = (0, 0, 0, 0)
        self._push_gate(zero4, (0, 0), 0, 0, 0, [input_var[0], input_var[1], input_var[2], input_var[3], iv[0][3]])
        for r in range(trace.n_round - 1):
            self._push_gate(zero4, (0, 0), 0, 0, 0, [iv[r][0], iv[r][1], iv[r][2], iv[r][3], iv[r + 1][3]])
        last = iv[trace.n_round - 1]
        self._push_gate(zero4, (0, 0), 0, 0, 0, [last[0], last[1], last[2], last[3], self.zero_var()])
        return CardVar(last)

    # ---- constraint_system/shuffle/permutation.rs:8-216
    def _sum_chunks(self, vars_, boolean: bool) -> int:
        zero_var, s = self.zero_var(), self.zero_var()
        for i in range(0, len(vars_), 3):
            c = vars_[i:i + 3]
            if len(c) == 3:
                s = self.linear_combine([s, c[0], c[1], c[2]], 1, 1, 1, 1)
            elif len(c) == 2:
                s = self.linear_combine([s, c[0], c[1], zero_var], 1, 1, 1, 0)
            else:
                s = self.linear_combine([s, c[0], zero_var, zero_var], 1, 1, 0, 0)
            if boolean:
                self.attach_boolean_constraint_to_gate()
        return s

    def shuffle_card(self, card_vars, permutation: Permutation) -> list:
        n = len(permutation)
        if len(card_vars) != n:
            raise ValueError("one card per row of the permutation expected")
        zero_var, one_var = self.zero_var(), self.one_var()
        pm = [[self.new_variable(y) for y in row] for row in permutation.get_matrix()]
        for row in pm:                                                    # rows: 0/1 entries summing to 1
            self.equal(self._sum_chunks(row, True), one_var)
        for j in range(n):                                                # columns sum to 1
            self.equal(self._sum_chunks([pm[i][j] for i in range(n)], False), one_var)
        split = [[cv[i] for cv in card_vars] for i in range(len(card_vars[0]))]
        w = self.witness
        out = []
        for row in pm:
            permuted = CardVar([0, 0, 0, 0])
            for i, coord in enumerate(split):
                r_vars = []
                for c in range(0, n, 2):
                    p, v = row[c:c + 2], coord[c:c + 2]
                    if len(p) == 2:
                        r_var = self.new_variable(w[p[0]] * w[v[0]] + w[p[1]] * w[v[1]])
                        self._push_gate((0, 0, 0, 0), (1, 1), 0, 0, 1, [p[0], v[0], p[1], v[1], r_var])
                    else:
                        r_var = self.new_variable(w[p[0]] * w[v[0]])
                        self._push_gate((0, 0, 0, 0), (1, 1), 0, 0, 1, [p[0], v[0], zero_var, zero_var, r_var])
                    r_vars.append(r_var)
                permuted[i] = self._sum_chunks(r_vars, False)
            out.append(permuted)
        return out

    # ---- turbo/mod.rs:171-191, 310-364: selector tables as lists of n integers
    def _remark_rows(self):
        for first, sel in self.shuffle_remark_constraints:
            for j in range(self.n_iteration_shuffle_scalar_mul):
                yield first + j, j, sel

    def compute_witness_selectors_int(self) -> list:
        polys = [[0] * self.size for _ in range(N_WIRE_SELECTORS)]
        for row, j, sel in self._remark_rows():
            for t in range(N_WIRE_SELECTORS):
                polys[t][row] = sel[t][j]
        return polys

    def _table_selectors_int(self, table) -> list:
        polys = [[0] * self.size for _ in range(12)]
        for row, j, _ in self._remark_rows():
            for c in range(3):               # x, y, dxy
                for t in range(4):
                    polys[4 * c + t][row] = table[j][t][c]
        return polys

    def compute_shuffle_generator_selectors_int(self) -> list:
        return self._table_selectors_int(self.shuffle_generators) if self.shuffle_remark_constraints else [[0] * self.size for _ in range(12)]

    def compute_shuffle_public_key_selectors_int(self) -> list:
        return self._table_selectors_int(self.shuffle_public_keys) if self.shuffle_remark_constraints else [[0] * self.size for _ in range(12)]

    def q_ecc_int(self) -> list:
        """plonk/indexer.rs:417-424: 1 on the NUM_ITERATIONS rows of every remark gate."""
        q = [0] * self.size
        for row, _, _ in self._remark_rows():
            q[row] = 1
        return q


def build_cs(cs, prng, aggregate_public_key, input_cards, permutation: Permutation | None = None, remark=BabyJubjubShuffle):
    """shuffle/src/build_cs.rs:26-56 on an empty TurboCS `cs`: remark every card with fresh random bits, shuffle the remarked
    cards, expose input and output decks as public inputs, pad.  Returns (cs, output card variables)."""
    n = len(input_cards)
    cs.load_shuffle_remark_parameters(aggregate_public_key, remark)
    pks = remark.crate_public_keys(aggregate_public_key)
    remark_card_vars = []
    for card in input_cards:
        bits = remark.sample_random_scalar_bits(prng)
        trace = remark.eval_remark_with_trace(card, bits, aggregate_public_key, pks)
        input_var = cs.new_card_variable(card)
        cs.prepare_pi_card_variable(input_var)
        remark_card_vars.append(cs.eval_card_remark(trace, input_var))
    if permutation is None:
        permutation = Permutation.rand(prng, n)
    shuffle_card_vars = cs.shuffle_card(remark_card_vars, permutation)
    for cv in shuffle_card_vars:
        cs.prepare_pi_card_variable(cv)
    cs.pad()
    return cs, shuffle_card_vars
