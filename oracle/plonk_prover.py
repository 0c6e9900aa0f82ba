"""CPU restatement (Python big integers) of the reference's TurboPlonK indexer, prover and verifier: the default feature set, and
(indexer / prover; the verifier for it is oracle/plonk_verifier_shuffle.py) the `shuffle` feature set with the remark and permutation
gadgets of zshuffle's circuit.  TEST INFRASTRUCTURE ONLY: imported by tests/ alone, never by the product package.

Follows, function by function:
  TurboCS (subset)            /root/reference/uzkge/src/plonk/constraint_system/turbo/mod.rs:395-537, 853-891, 968-977
  compute_permutation         /root/reference/uzkge/src/plonk/constraint_system/mod.rs:54-84
  indexer                     /root/reference/uzkge/src/plonk/indexer.rs:248-536
  prover                      /root/reference/uzkge/src/plonk/prover.rs:88-394
  helpers                     /root/reference/uzkge/src/plonk/helpers.rs (pi_poly :111-131, hide_polynomial :139-154, z_poly :160-220,
                              t_poly :223-678, r_poly :681-999, split_t_and_commit :1323-1408, first_lagrange_poly :1412-1423,
                              r_eval_zeta :1181-1320, eval_pi_poly :1129-1165)
  batch_prove / batch         /root/reference/uzkge/src/poly_commit/pcs.rs:107-191
  verifier                    /root/reference/uzkge/src/plonk/verifier.rs:17-164, compute_challenges :167-222
  Transcript                  /root/reference/uzkge/src/utils/transcript.rs, plonk/transcript.rs
  ChaChaRng / Fr::rand        rand_chacha 0.3 / ark-ff 0.4 (pinned by the golden k[1..4] of the verifier keys)

The pairing check of batch_verify_diff_points (kzg_poly_commitment.rs:407-460) is replaced by its G1 equivalent under a KNOWN
trapdoor tau (the SRS here is synthetic): e(A, H) = e(B, tau H)  <=>  A = tau * B.

Everything is a canonical integer mod r; G1 points are affine (x, y) tuples or None; MSMs go through oracle.bn254.
Parity status: the Rust prover cannot run here (no toolchain), so a proof is pinned by (1) this restatement's own verifier and
oracle/plonk_verifier_shuffle.py -- the verifier that accepts the reference's golden proofs -- accepting it, (2) the golden RNG /
domain values, (3) the circuits' preprocessed commitments reproducing the reference's deployed / bundled verifier keys
(tests/test_shuffle_host.py, tests/test_matchmaking_host.py), (4) bit-equality with the GPU pipeline on the same seed.
"""
from __future__ import annotations

from . import bn254 as bn
from .bn254 import FR, inv_mod
from . import plonk as qmap

N_WIRES = 5
N_SELECTORS = 9


# ---------------------------------------------------------------- Keccak-256 / transcript (second, independent implementation)
def _keccak_f1600(lanes):
    """lanes[x][y], the specification's form (rho offsets and round constants derived, not tabulated)."""
    R = 1
    for _ in range(24):
        C = [lanes[x][0] ^ lanes[x][1] ^ lanes[x][2] ^ lanes[x][3] ^ lanes[x][4] for x in range(5)]
        D = [C[(x + 4) % 5] ^ (((C[(x + 1) % 5] << 1) | (C[(x + 1) % 5] >> 63)) & (2**64 - 1)) for x in range(5)]
        lanes = [[lanes[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        x, y, cur = 1, 0, lanes[1][0]
        for t in range(24):
            x, y = y, (2 * x + 3 * y) % 5
            s = ((t + 1) * (t + 2) // 2) % 64
            cur, lanes[x][y] = lanes[x][y], ((cur << s) | (cur >> (64 - s))) & (2**64 - 1) if s else cur
        for y in range(5):
            T = [lanes[x][y] for x in range(5)]
            for x in range(5):
                lanes[x][y] = T[x] ^ ((~T[(x + 1) % 5]) & T[(x + 2) % 5] & (2**64 - 1))
        for j in range(7):
            R = ((R << 1) ^ ((R >> 7) * 0x71)) % 256
            if R & 2:
                lanes[0][0] ^= 1 << ((1 << j) - 1)
    return lanes


def keccak256(data: bytes) -> bytes:
    rate = 136
    p = bytearray(data) + b"\x01"
    p += b"\x00" * (-len(p) % rate)
    p[-1] |= 0x80
    lanes = [[0] * 5 for _ in range(5)]
    for off in range(0, len(p), rate):
        for i in range(rate // 8):
            lanes[i % 5][i // 5] ^= int.from_bytes(p[off + 8 * i: off + 8 * i + 8], "little")
        lanes = _keccak_f1600(lanes)
    return b"".join(lanes[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


class Transcript:
    def __init__(self, msg: bytes):
        self.state = b""
        self.msg(msg)

    def msg(self, m: bytes):
        if len(m) < 32:
            m = bytes(32 - len(m)) + m
        assert len(m) % 32 == 0
        self.state += m

    def u64(self, a: int):
        self.state += bytes(24) + a.to_bytes(8, "big")

    def byte(self, b: int):
        self.state += bytes([b])

    def point(self, P):
        self.msg(bytes(64) if P is None else P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big"))

    def fr(self, x: int):
        self.msg(x.to_bytes(32, "big"))

    def challenge(self) -> int:
        c = int.from_bytes(keccak256(self.state), "big") % FR
        self.state = c.to_bytes(32, "big")
        return c


def transcript_init_plonk(tr, vp, pi_values, root):
    tr.msg(b"PLONK")
    tr.u64(vp["cs_size"])
    tr.msg(FR.to_bytes(32, "big"))
    for c in vp["cm_q_vec"]:
        tr.point(c)
    for c in vp["cm_s_vec"]:
        tr.point(c)
    tr.fr(root)
    for k in vp["k"]:
        tr.fr(k)
    for v in pi_values:
        tr.fr(v)


def _init_batch_eval(tr, max_degree, point):
    tr.msg(b"New PCS-Batch-Eval Protocol")
    tr.msg(FR.to_bytes(32, "big"))
    tr.u64(max_degree)
    tr.fr(point)


# ---------------------------------------------------------------- ChaCha20 RNG (second implementation)
class ChaCha:
    def __init__(self, seed: bytes = bytes(32)):
        self.key = [int.from_bytes(seed[i:i + 4], "little") for i in range(0, 32, 4)]
        self.ctr = 0
        self.words = []

    def _refill(self):
        st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *self.key, self.ctr & 0xFFFFFFFF, self.ctr >> 32, 0, 0]
        w = st[:]

        def qr(a, b, c, d):
            for (p, q, r, rot) in ((a, b, d, 16), (c, d, b, 12), (a, b, d, 8), (c, d, b, 7)):
                w[p] = (w[p] + w[q]) % 2**32
                v = w[r] ^ w[p]
                w[r] = ((v << rot) | (v >> (32 - rot))) % 2**32

        for _ in range(10):
            for i in range(4):
                qr(i, 4 + i, 8 + i, 12 + i)
            for i in range(4):
                qr(i, 4 + (i + 1) % 4, 8 + (i + 2) % 4, 12 + (i + 3) % 4)
        self.words = [(w[i] + st[i]) % 2**32 for i in range(16)]
        self.ctr += 1

    def u64(self) -> int:
        out = 0
        for h in range(2):
            if not self.words:
                self._refill()
            out |= self.words.pop(0) << (32 * h)
        return out

    def fr(self) -> int:
        while True:
            raw = sum(self.u64() << (64 * i) for i in range(4)) % 2**254
            if raw < FR:
                return raw * inv_mod(2**256 % FR, FR) % FR


def choose_ks(rng, n):
    k = [1]
    while len(k) < n:
        ki = rng.fr()
        if ki and ki not in k and pow(ki, (FR - 1) // 2, FR) != 1:
            k.append(ki)
    return k


# ---------------------------------------------------------------- polynomials as coefficient lists
def p_eval(c, x):
    return bn.poly_eval(c, x)


def p_add_coef(c, v, i):
    """FpPolynomial::add_coef_assign (field_polynomial.rs): grows the vector when needed."""
    while len(c) <= i:
        c.append(0)
    c[i] = (c[i] + v) % FR


def p_div_linear(c, z):
    """(quotient, remainder) of c / (X - z)."""
    q = [0] * max(len(c) - 1, 0)
    acc = 0
    for i in range(len(c) - 1, -1, -1):
        acc = (acc * z + c[i]) % FR
        if i > 0:
            q[i - 1] = acc
    return q, acc


# ---------------------------------------------------------------- TurboCS (the subset a synthetic circuit needs)
class TurboCS:
    def __init__(self):
        self.selectors = [[] for _ in range(N_SELECTORS)]
        self.wiring = [[] for _ in range(N_WIRES)]
        self.num_vars = 2
        self.size = 0
        self.witness = [0, 1]
        self.public_vars_constraint_indices = []
        self.public_vars_witness_indices = []
        self.boolean_constraint_indices = []
        self.edwards_a = 0
        self.rounds = 0                 # n_iteration_shuffle_scalar_mul
        self.pk_table = self.gen_table = None      # [round][4] -> (x, y, dxy)
        self.remark = []                # shuffle_remark_constraint_indices: (first row, [s1 column, s2 column, s3 column])
        self.anemoi_constraints_indices = []          # first rows of the 14-row Anemoi permutations
        self.anemoi_generator = self.anemoi_generator_inv = 0
        self.anemoi_prk = None          # (x[14][2], y[14][2]): the preprocessed round keys
        self.insert_constant_gate(0, 0)
        self.insert_constant_gate(1, 1)

    def _push(self, q_add, q_mul, q_c, q_ecc, q_out, wires):
        for i in range(4):
            self.selectors[i].append(q_add[i] % FR)
        self.selectors[4].append(q_mul[0] % FR)
        self.selectors[5].append(q_mul[1] % FR)
        self.selectors[6].append(q_c % FR)
        self.selectors[7].append(q_ecc % FR)
        self.selectors[8].append(q_out % FR)
        for i in range(N_WIRES):
            self.wiring[i].append(wires[i])
        self.size += 1

    def new_variable(self, value):
        self.num_vars += 1
        self.witness.append(value % FR)
        return self.num_vars - 1

    def insert_lc_gate(self, wires_in, wire_out, q1, q2, q3, q4):
        self._push((q1, q2, q3, q4), (0, 0), 0, 0, 1, list(wires_in) + [wire_out])

    def insert_add_gate(self, l, r, o):
        self.insert_lc_gate((l, r, 0, 0), o, 1, 1, 0, 0)

    def insert_sub_gate(self, l, r, o):
        self.insert_lc_gate((l, r, 0, 0), o, 1, FR - 1, 0, 0)

    def insert_mul_gate(self, l, r, o):
        self._push((0, 0, 0, 0), (1, 0), 0, 0, 1, [l, r, 0, 0, o])

    def insert_constant_gate(self, var, constant):
        self._push((0, 0, 0, 0), (0, 0), constant, 0, 1, [var] * 5)

    def insert_boolean_gate(self, var):
        self.insert_mul_gate(var, var, var)

    def prepare_pi_variable(self, var):
        self.public_vars_witness_indices.append(var)
        self.public_vars_constraint_indices.append(self.size)
        self.insert_constant_gate(var, 0)

    def attach_boolean_constraint_to_gate(self):
        self.boolean_constraint_indices.append(self.size - 1)

    def add(self, l, r):
        o = self.new_variable(self.witness[l] + self.witness[r])
        self.insert_add_gate(l, r, o)
        return o

    def mul(self, l, r):
        o = self.new_variable(self.witness[l] * self.witness[r])
        self.insert_mul_gate(l, r, o)
        return o

    # ---- the shuffle gadgets
    def equal(self, l, r):
        self.insert_sub_gate(l, r, 0)

    def linear_combine(self, wires_in, q1, q2, q3, q4):
        """turbo/mod.rs:639-662."""
        w = self.witness
        o = self.new_variable(w[wires_in[0]] * q1 + w[wires_in[1]] * q2 + w[wires_in[2]] * q3 + w[wires_in[3]] * q4)
        self.insert_lc_gate(wires_in, o, q1, q2, q3, q4)
        return o

    def load_shuffle_remark_parameters(self, pk):
        """turbo/mod.rs:926-966 with BabyJubjubShuffle's constants."""
        from . import babyjubjub as bj

        tab = lambda segs: [[(x, y, bj.D * x * y % FR) for x, y in seg] for seg in segs]
        self.pk_table, self.gen_table = tab(bj.segments(pk)), tab(bj.segments(bj.GEN))
        self.edwards_a, self.rounds = bj.A, bj.NUM_ITERATIONS

    def new_card_variable(self, card):
        """constraint_system/shuffle/mod.rs:64-78: card = (e1, e2); variables created first.x, first.y, second.x, second.y and
        returned as [second.x, second.y, first.x, first.y]."""
        (fx, fy), (sx, sy) = card
        a, b, c, d = self.new_variable(fx), self.new_variable(fy), self.new_variable(sx), self.new_variable(sy)
        return [c, d, a, b]

    def prepare_pi_card_variable(self, cv):
        for v in cv:
            self.prepare_pi_variable(v)

    def eval_card_remark(self, field_bits, intermediate_values, input_var):
        """constraint_system/shuffle/remark.rs:11-94: rounds + 1 gates with all-zero selectors; row r holds the running pair of
        points in wires 0..3 and the next round's g.y in the output wire (0 in the last row)."""
        assert len(field_bits) == len(intermediate_values) == self.rounds
        self.remark.append((self.size, [[b[i] for b in field_bits] for i in range(3)]))
        iv = [[self.new_variable(x) for x in vals] for vals in intermediate_values]
        z4 = (0, 0, 0, 0)
        self._push(z4, (0, 0), 0, 0, 0, list(input_var) + [iv[0][3]])
        for r in range(self.rounds - 1):
            self._push(z4, (0, 0), 0, 0, 0, iv[r] + [iv[r + 1][3]])
        self._push(z4, (0, 0), 0, 0, 0, iv[-1] + [0])
        return list(iv[-1])

    def _sum3(self, vs, boolean=False):
        s = 0
        for i in range(0, len(vs), 3):
            c = list(vs[i:i + 3])
            q = [1] * len(c) + [0] * (3 - len(c))
            s = self.linear_combine([s] + c + [0] * (3 - len(c)), 1, q[0], q[1], q[2])
            if boolean:
                self.attach_boolean_constraint_to_gate()
        return s

    def shuffle_card(self, card_vars, matrix):
        """constraint_system/shuffle/permutation.rs:8-216."""
        n = len(matrix)
        pm = [[self.new_variable(y) for y in row] for row in matrix]
        for row in pm:
            self.equal(self._sum3(row, True), 1)
        for j in range(n):
            self.equal(self._sum3([pm[i][j] for i in range(n)]), 1)
        w, out = self.witness, []
        for row in pm:
            permuted = []
            for i in range(4):
                coord = [cv[i] for cv in card_vars]
                rv = []
                for c in range(0, n, 2):
                    if c + 1 < n:
                        r = self.new_variable(w[row[c]] * w[coord[c]] + w[row[c + 1]] * w[coord[c + 1]])
                        self._push((0, 0, 0, 0), (1, 1), 0, 0, 1, [row[c], coord[c], row[c + 1], coord[c + 1], r])
                    else:
                        r = self.new_variable(w[row[c]] * w[coord[c]])
                        self._push((0, 0, 0, 0), (1, 1), 0, 0, 1, [row[c], coord[c], 0, 0, r])
                    rv.append(r)
                permuted.append(self._sum3(rv))
            out.append(permuted)
        return out

    def _remark_rows(self):
        for first, sel in self.remark:
            for j in range(self.rounds):
                yield first + j, j, sel

    def compute_witness_selectors(self):
        """turbo/mod.rs:171-191."""
        polys = [[0] * self.size for _ in range(3)]
        for row, j, sel in self._remark_rows():
            for t in range(3):
                polys[t][row] = sel[t][j]
        return polys

    def table_selectors(self, table):
        """turbo/mod.rs:310-364: x_00..x_11, y_00..y_11, dxy_00..dxy_11 on the remark rows."""
        polys = [[0] * self.size for _ in range(12)]
        for row, j, _ in self._remark_rows():
            for c in range(3):
                for t in range(4):
                    polys[4 * c + t][row] = table[j][t][c]
        return polys

    def q_ecc(self):
        """plonk/indexer.rs:417-424."""
        q = [0] * self.size
        for row, _, _ in self._remark_rows():
            q[row] = 1
        return q

    def compute_anemoi_jive_selectors(self):
        """turbo/mod.rs:285-304."""
        polys = [[0] * self.size for _ in range(4)]
        for first in self.anemoi_constraints_indices:
            for j in range(14):
                polys[0][first + j], polys[1][first + j] = self.anemoi_prk[0][j]
                polys[2][first + j], polys[3][first + j] = self.anemoi_prk[1][j]
        return polys

    def check_remark_equations(self, witness):
        """The four remark equations of verify_witness (turbo/mod.rs:1100-1330) on every remark row: the twisted Edwards addition
        of +-(the selected multiple of pk / G) to the running pair of points, written without divisions."""
        w_of = lambda j, row: witness[self.wiring[j][row]]
        for row, r, sel in self._remark_rows():
            a, b, c, d, o = (w_of(j, row) for j in range(5))
            an, bn_, cn = (w_of(j, row + 1) for j in range(3))
            s1, s2, s3 = sel[0][r], sel[1][r], sel[2][r]
            pick = [(1 - s1) * (1 - s2), s1 * (1 - s2), (1 - s1) * s2, s1 * s2]
            e = [0, 0, 0, 0]
            for t in range(4):
                px, py, pd = self.pk_table[r][t]
                gx, gy, gd = self.gen_table[r][t]
                e[0] += pick[t] * (s3 * an - s3 * a * py - b * px + a * b * an * pd)
                e[1] += pick[t] * (s3 * bn_ + self.edwards_a * a * px - s3 * b * py - a * b * bn_ * pd)
                e[2] += pick[t] * (s3 * cn - s3 * c * gy - d * gx + c * d * cn * gd)
                e[3] += pick[t] * (s3 * o + self.edwards_a * c * gx - s3 * d * gy - c * d * o * gd)
            if any(v % FR for v in e):
                return False
        return True

    def pad(self):
        n = 1
        while n < self.size:
            n *= 2
        diff = n - self.size
        for s in self.selectors:
            s.extend([0] * diff)
        for w in self.wiring:
            w.extend([0] * diff)
        self.size += diff

    def quot_eval_dom_size(self):
        return self.size * 6 if self.size > 8 else self.size * 16

    def compute_permutation(self):
        """constraint_system/mod.rs:54-84: one cycle per variable over its positions in increasing order."""
        v = [x for w in self.wiring for x in w]
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

    def extend_witness(self, witness):
        return [witness[i] for w in self.wiring for i in w]

    @staticmethod
    def get_hiding_degree(idx):
        return 3 if idx < 3 else 2


def eval_selector_multipliers(w):
    return [w[0], w[1], w[2], w[3], w[0] * w[1] % FR, w[2] * w[3] % FR, 1, w[0] * w[1] % FR * w[2] % FR * w[3] % FR * w[4] % FR,
            (FR - w[4]) % FR]


# ---------------------------------------------------------------- KZG over a synthetic SRS with a known trapdoor
class Kzg:
    def __init__(self, max_degree: int, tau: int):
        self.tau = tau % FR
        G = (1, 2)
        self.srs, t = [], 1
        for _ in range(max_degree + 1):
            self.srs.append(bn.g1_mul(G, t))
            t = t * self.tau % FR

    def commit(self, coefs):
        c = bn.trim(coefs)
        assert len(c) <= len(self.srs), "DegreeError"
        # with a known trapdoor, sum c_i tau^i G is one scalar multiplication; the MSM itself is covered by tests/test_gpu_msm.py
        return bn.g1_mul((1, 2), p_eval(c, self.tau))

    def commit_msm(self, coefs):
        c = bn.trim(coefs)
        return bn.msm_naive(self.srs[: len(c)], c)


def _g1_lin(terms):
    """sum s_i * P_i over affine points (None = identity)."""
    acc = None
    for s, P in terms:
        if P is None or s % FR == 0:
            continue
        Q = bn.g1_mul(P, s % FR)
        acc = Q if acc is None else bn.g1_add(acc, Q)
    return acc


# ---------------------------------------------------------------- indexer (plonk/indexer.rs:248-536)
def indexer(cs: TurboCS, pcs: Kzg, shuffle: bool = False):
    """shuffle = True adds what the `shuffle` feature adds (indexer.rs:414-476): q_ecc and the 12 + 12 shuffle selector polynomials
    (all zero for circuits without remark gates)."""
    n, m = cs.size, cs.quot_eval_dom_size()
    factor = m // n
    root = bn.root_of_unity(n)
    root_m = bn.root_of_unity(m)
    group = [pow(root, i, FR) for i in range(n)]
    k = choose_ks(ChaCha(bytes(32)), N_WIRES)
    coset_quotient = [k[1] * pow(root_m, i, FR) % FR for i in range(m)]
    perm = cs.compute_permutation()
    enc = [k[p // n] * group[p % n] % FR for p in perm]     # encode_perm_to_group (indexer.rs:195-208)

    def pre(evals):
        coefs = bn.trim(bn.ifft(evals, n))
        return coefs, bn.coset_fft(coefs, m, k[1])

    P = {"n": n, "m": m, "factor": factor, "group": group, "coset_quotient": coset_quotient, "permutation": perm, "root": root}
    P["s_polys"], P["s_coset"] = zip(*[pre(enc[i * n:(i + 1) * n]) for i in range(N_WIRES)])
    P["q_polys"], P["q_coset"] = zip(*[pre(cs.selectors[i]) for i in range(N_SELECTORS)])
    P["l1_coefs"], P["l1_coset"] = pre([n % FR] + [0] * (n - 1))
    P["z_h_inv"] = qmap.z_h_inv_coset_evals(k[1], root_m, n, factor)
    qb = [0] * n
    for i in cs.boolean_constraint_indices:
        qb[i] = 1
    P["qb_poly"], P["qb_coset"] = pre(qb)
    P["q_prk_polys"], P["q_prk_coset"] = zip(*[pre(e) for e in cs.compute_anemoi_jive_selectors()])      # indexer.rs:385-412
    lagrange_constants = []
    for ci in cs.public_vars_constraint_indices:
        inv = 1
        for i, e in enumerate(group):
            if i != ci:
                inv = inv * (group[ci] - e) % FR
        lagrange_constants.append(inv_mod(inv, FR))
    P["vp"] = {
        "cm_q_vec": [pcs.commit(p) for p in P["q_polys"]], "cm_s_vec": [pcs.commit(p) for p in P["s_polys"]],
        "cm_qb": pcs.commit(P["qb_poly"]), "cm_prk_vec": [pcs.commit(p) for p in P["q_prk_polys"]],
        "anemoi_generator": cs.anemoi_generator, "anemoi_generator_inv": cs.anemoi_generator_inv, "k": k, "cs_size": n,
        "public_vars_constraint_indices": list(cs.public_vars_constraint_indices), "lagrange_constants": lagrange_constants,
    }
    P["shuffle"] = shuffle
    if shuffle:
        # indexer.rs:414-476: q_ecc, the generator selectors, and the public-key selectors as a copy of them
        P["q_ecc_poly"], P["q_ecc_coset"] = pre(cs.q_ecc())
        gen_evals = cs.table_selectors(cs.gen_table) if cs.remark else [[0] * n for _ in range(12)]
        P["q_gen_polys"], P["q_gen_coset"] = zip(*[pre(e) for e in gen_evals])
        P["q_pk_polys"], P["q_pk_coset"] = P["q_gen_polys"], P["q_gen_coset"]
        P["vp"].update({"cm_q_ecc": pcs.commit(P["q_ecc_poly"]), "cm_shuffle_generator_vec": [pcs.commit(p) for p in P["q_gen_polys"]],
                        "cm_shuffle_public_key_vec": [pcs.commit(p) for p in P["q_pk_polys"]], "edwards_a": cs.edwards_a, "root": root,
                        "pi_points": [pow(root, ci, FR) for ci in cs.public_vars_constraint_indices], "pi_lagrange": lagrange_constants})
    return P


def refresh_public_key(P, cs, pcs, pk):
    """shuffle/src/gen_params/params.rs:57-129 (monomial branch): reload the key, rebuild the 12 public-key selector polynomials,
    their coset evaluations and commitments."""
    cs.load_shuffle_remark_parameters(pk)
    n, m, k = P["n"], P["m"], P["vp"]["k"]
    polys, cosets = [], []
    for e in cs.table_selectors(cs.pk_table):
        c = bn.trim(bn.ifft(e, n))
        polys.append(c)
        cosets.append(bn.coset_fft(c, m, k[1]))
    P["q_pk_polys"], P["q_pk_coset"] = polys, cosets
    P["vp"]["cm_shuffle_public_key_vec"] = [pcs.commit(p) for p in polys]
    return P["vp"]["cm_shuffle_public_key_vec"]


# ---------------------------------------------------------------- prover (plonk/prover.rs:88-394, lagrange_pcs = None)
def hide_polynomial(rng, coefs, hiding_degree, zeroing_degree):
    blinds = []
    for i in range(hiding_degree):
        b = rng.fr()
        blinds.append(b)
        p_add_coef(coefs, b, i)
        p_add_coef(coefs, (FR - b) % FR, zeroing_degree + i)
    return blinds


def z_evals(P, w_ext, beta, gamma):
    n, k, group, perm = P["n"], P["vp"]["k"], P["group"], P["permutation"]
    out, prev = [1], 1
    for i in range(n - 1):
        num = den = 1
        for j in range(N_WIRES):
            f = w_ext[j * n + i]
            num = num * (f + gamma + beta * k[j] * group[i]) % FR
            p = perm[j * n + i]
            den = den * (f + gamma + beta * k[p // n] * group[p % n]) % FR
        prev = prev * num % FR * inv_mod(den, FR) % FR
        out.append(prev)
    return out


def t_poly(P, w_polys, z_poly, alpha, beta, gamma, pi_poly, w_sel_polys=None):
    m, k = P["m"], P["vp"]["k"]
    co = lambda c: bn.coset_fft(c, m, k[1])
    sh = None
    if w_sel_polys is not None:
        sh = {"w_sel": [co(p) for p in w_sel_polys], "q_ecc": P["q_ecc_coset"], "pk": P["q_pk_coset"], "gen": P["q_gen_coset"],
              "edwards_a": P["vp"]["edwards_a"]}
    evals = qmap.quotient_coset_evals(
        [co(p) for p in w_polys], P["q_coset"], co(pi_poly), co(z_poly), P["s_coset"], P["coset_quotient"], P["l1_coset"], P["qb_coset"],
        P["q_prk_coset"], k, alpha, beta, gamma, P["vp"]["anemoi_generator"], P["vp"]["anemoi_generator_inv"], P["z_h_inv"], P["factor"],
        shuffle=sh)
    return bn.trim(bn.coset_ifft(evals, m, inv_mod(k[1], FR)))


def split_t(rng, t, n_pieces, n):
    """helpers.rs:1323-1408 without the commitments: returns the piece coefficient lists."""
    pieces, prev, L = [], 0, len(t)
    for i in range(n_pieces):
        start, end = i * n, (L if i == n_pieces - 1 else (i + 1) * n)
        coefs = list(t[start:min(L, end)]) if start < L else []
        r = rng.fr()
        if i != n_pieces - 1:
            coefs += [0] * (n + 1 - len(coefs))
            coefs[n] = (coefs[n] + r) % FR
            coefs[0] = (coefs[0] - prev) % FR
        elif not coefs:
            coefs = [(FR - prev) % FR]
        else:
            coefs[0] = (coefs[0] - prev) % FR
        prev = r
        pieces.append(bn.trim(coefs))
    return pieces


def first_lagrange(zeta, n):
    z_h = (pow(zeta, n, FR) - 1) % FR
    return z_h, z_h * inv_mod((zeta - 1) % FR, FR) % FR


def r_scalars(P_k, w_ev, s_ev, prk3, z_ev_omega, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n_t_polys, n_pieces=N_WIRES):
    """r_poly_or_comm (helpers.rs:681-999) as (coefficient, name) pairs over {q_i, z, s_last, qb, prk1, prk2, t_i}."""
    a = [pow(alpha, i, FR) for i in range(8)]
    wm = eval_selector_multipliers(w_ev)
    terms = [(wm[i], ("q", i)) for i in range(N_SELECTORS)]
    z_scalar = alpha
    for i in range(N_WIRES):
        z_scalar = z_scalar * (w_ev[i] + P_k[i] * beta % FR * zeta + gamma) % FR
    z_scalar = (z_scalar + l1_ev * a[2]) % FR
    terms.append((z_scalar, ("z", 0)))
    s_last = alpha * z_ev_omega % FR * beta % FR
    for i in range(N_WIRES - 1):
        s_last = s_last * (w_ev[i] + beta * s_ev[i] + gamma) % FR
    terms.append(((FR - s_last) % FR, ("s_last", 0)))
    qb = (w_ev[1] * (w_ev[1] - 1) % FR * a[3] + w_ev[2] * (w_ev[2] - 1) % FR * a[4] + w_ev[3] * (w_ev[3] - 1) % FR * a[5]) % FR
    terms.append((qb, ("qb", 0)))
    terms.append((prk3 * a[6] % FR, ("prk", 0)))
    terms.append((prk3 * a[7] % FR, ("prk", 1)))
    f = pow(zeta, n_t_polys, FR)
    e = z_h_ev
    for i in range(n_pieces):
        terms.append(((FR - e) % FR, ("t", i)))
        e = e * f % FR
    return terms


def _lin_polys(terms):
    L = max(len(p) for _, p in terms)
    out = [0] * L
    for s, p in terms:
        for i, c in enumerate(p):
            out[i] = (out[i] + s * c) % FR
    return out


def batch_prove(tr, pcs, polys, point, max_degree):
    _init_batch_eval(tr, max_degree, point)
    alpha = tr.challenge()
    h, mult = [0], 1
    for p in polys:
        q = list(p)
        q[0] = (q[0] - p_eval(p, point)) % FR
        h = _lin_polys([(1, h), (mult, q)])
        mult = mult * alpha % FR
    quo, rem = p_div_linear(bn.trim(h), point)
    assert rem == 0, "PCSProveEvalError"
    return pcs.commit(quo)


def prover(rng, tr, pcs, cs, P, witness):
    """prover.rs:88-394, lagrange_pcs = None.  With P["shuffle"] (indexer(..., shuffle=True)) the `shuffle` feature set: witness-selector
    polynomials committed after the wires, quotient terms 12-18, q_ecc / w_sel openings, linearisation parts 6-9."""
    n, vp = P["n"], P["vp"]
    k, root = vp["k"], P["root"]
    shuffle = P.get("shuffle", False)
    online = [witness[i] for i in cs.public_vars_witness_indices]
    transcript_init_plonk(tr, vp, online, root)
    pi_evals = [0] * n
    for pos, ci in enumerate(cs.public_vars_constraint_indices):
        pi_evals[ci] = online[pos]
    pi = bn.trim(bn.ifft(pi_evals, n))
    w_ext = cs.extend_witness(witness)
    w_polys, cm_w = [], []
    for i in range(N_WIRES):
        f = bn.trim(bn.ifft(w_ext[i * n:(i + 1) * n], n))
        hide_polynomial(rng, f, cs.get_hiding_degree(i), n)
        cm = pcs.commit(f)
        tr.point(cm)
        w_polys.append(f)
        cm_w.append(cm)
    w_sel_polys, cm_w_sel = [], []
    if shuffle:
        for sel_evals in cs.compute_witness_selectors():      # prover.rs:177-191
            f = bn.trim(bn.ifft(sel_evals, n))
            hide_polynomial(rng, f, 2, n)
            cm = pcs.commit(f)
            tr.point(cm)
            w_sel_polys.append(f)
            cm_w_sel.append(cm)
    beta = tr.challenge()
    tr.byte(0x01)
    gamma = tr.challenge()
    z = bn.trim(bn.ifft(z_evals(P, w_ext, beta, gamma), n))
    hide_polynomial(rng, z, 3, n)
    cm_z = pcs.commit(z)
    tr.point(cm_z)
    alpha = tr.challenge()
    t = t_poly(P, w_polys, z, alpha, beta, gamma, pi, w_sel_polys if shuffle else None)
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
    q_ecc_ev = p_eval(P["q_ecc_poly"], zeta) if shuffle else 0
    w_sel_ev = [p_eval(p, zeta) for p in w_sel_polys]
    for v in w_ev + s_ev + w_sel_ev:
        tr.fr(v)
    tr.fr(prk3)
    tr.fr(prk4)
    tr.fr(z_ev_omega)
    if shuffle:
        tr.fr(q_ecc_ev)
    for v in w_ev_omega:
        tr.fr(v)
    u = tr.challenge()
    z_h_ev, l1_ev = first_lagrange(zeta, n)
    src = {"q": P["q_polys"], "z": [z], "s_last": [P["s_polys"][N_WIRES - 1]], "qb": [P["qb_poly"]], "prk": P["q_prk_polys"], "t": t_polys}
    proof = {
        "cm_w_vec": cm_w, "cm_t_vec": cm_t, "cm_z": cm_z, "prk_3_poly_eval_zeta": prk3, "prk_4_poly_eval_zeta": prk4,
        "w_polys_eval_zeta": w_ev, "w_polys_eval_zeta_omega": w_ev_omega, "z_eval_zeta_omega": z_ev_omega, "s_polys_eval_zeta": s_ev,
    }
    if shuffle:
        from .plonk_verifier_shuffle import r_terms      # the term list the golden-pinned verifier sums over commitments

        proof.update({"cm_w_sel_vec": cm_w_sel, "q_ecc_poly_eval_zeta": q_ecc_ev, "w_sel_polys_eval_zeta": w_sel_ev})
        src.update({"pk": P["q_pk_polys"], "gen": P["q_gen_polys"]})
        terms = r_terms(k, vp["edwards_a"], proof, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n + 2)
    else:
        terms = r_scalars(k, w_ev, s_ev, prk3, z_ev_omega, alpha, beta, gamma, zeta, l1_ev, z_h_ev, n + 2)
    r = _lin_polys([(s_, src[name][i]) for s_, (name, i) in terms])
    open_zeta = w_polys + list(P["s_polys"][:N_WIRES - 1]) + [P["q_prk_polys"][2], P["q_prk_polys"][3]]
    if shuffle:
        open_zeta += [P["q_ecc_poly"]] + w_sel_polys
    open_zeta.append(r)
    proof["opening_witness_zeta"] = batch_prove(tr, pcs, open_zeta, zeta, n + 2)
    proof["opening_witness_zeta_omega"] = batch_prove(tr, pcs, [z, w_polys[0], w_polys[1], w_polys[2]], zeta_omega, n + 2)
    proof["_u"] = u
    return proof


def proof_to_bytes_be(proof) -> bytes:
    """PlonkProof::to_bytes_be (indexer.rs:538-590), either feature set."""
    pt = lambda P_: bytes(64) if P_ is None else P_[0].to_bytes(32, "big") + P_[1].to_bytes(32, "big")
    sc = lambda v: v.to_bytes(32, "big")
    out = b"".join(pt(c) for c in proof["cm_w_vec"])
    out += b"".join(pt(c) for c in proof.get("cm_w_sel_vec", []))
    out += b"".join(pt(c) for c in proof["cm_t_vec"]) + pt(proof["cm_z"])
    out += sc(proof["prk_3_poly_eval_zeta"]) + sc(proof["prk_4_poly_eval_zeta"])
    out += b"".join(sc(v) for v in proof["w_polys_eval_zeta"] + proof["w_polys_eval_zeta_omega"]) + sc(proof["z_eval_zeta_omega"])
    out += b"".join(sc(v) for v in proof["s_polys_eval_zeta"])
    if "q_ecc_poly_eval_zeta" in proof:
        out += sc(proof["q_ecc_poly_eval_zeta"]) + b"".join(sc(v) for v in proof["w_sel_polys_eval_zeta"])
    return out + pt(proof["opening_witness_zeta"]) + pt(proof["opening_witness_zeta_omega"])


# ---------------------------------------------------------------- verifier (plonk/verifier.rs:17-164) under a known trapdoor
def r_eval_zeta(proof, alpha, beta, gamma, pi_ev, l1_ev, g, g_inv):
    """helpers.rs:1181-1320 without the shuffle terms."""
    a = [pow(alpha, i, FR) for i in range(10)]
    w, wo, s = proof["w_polys_eval_zeta"], proof["w_polys_eval_zeta_omega"], proof["s_polys_eval_zeta"]
    prk3, prk4 = proof["prk_3_poly_eval_zeta"], proof["prk_4_poly_eval_zeta"]
    term1 = alpha * proof["z_eval_zeta_omega"] % FR
    for i in range(N_WIRES - 1):
        term1 = term1 * (w[i] + beta * s[i] + gamma) % FR
    term1 = term1 * (w[N_WIRES - 1] + gamma) % FR
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
    return (term1 + term2 - pi_ev + term3 + term4 + term5 + term6) % FR


def eval_pi_poly(vp, pi, z_h_ev, point, root):
    ev = 0
    for v, lc, ci in zip(pi, vp["lagrange_constants"], vp["public_vars_constraint_indices"]):
        ev = (ev + lc * inv_mod((point - pow(root, ci, FR)) % FR, FR) % FR * v) % FR
    return ev * z_h_ev % FR


def _batch(tr, cms, max_degree, point, evals):
    _init_batch_eval(tr, max_degree, point)
    alpha = tr.challenge()
    mult, ev, terms = 1, 0, []
    for e, c in zip(evals, cms):
        terms.append((mult, c))
        ev = (ev + e * mult) % FR
        mult = mult * alpha % FR
    return _g1_lin(terms), ev


def verifier(tr, pcs: Kzg, vp, pi, proof) -> bool:
    n = vp["cs_size"]
    root = bn.root_of_unity(n)
    transcript_init_plonk(tr, vp, pi, root)
    for c in proof["cm_w_vec"]:                      # compute_challenges (verifier.rs:167-222)
        tr.point(c)
    beta = tr.challenge()
    tr.byte(0x01)
    gamma = tr.challenge()
    tr.point(proof["cm_z"])
    alpha = tr.challenge()
    for c in proof["cm_t_vec"]:
        tr.point(c)
    zeta = tr.challenge()
    for v in proof["w_polys_eval_zeta"] + proof["s_polys_eval_zeta"]:
        tr.fr(v)
    tr.fr(proof["prk_3_poly_eval_zeta"])
    tr.fr(proof["prk_4_poly_eval_zeta"])
    tr.fr(proof["z_eval_zeta_omega"])
    for v in proof["w_polys_eval_zeta_omega"]:
        tr.fr(v)
    u = tr.challenge()
    z_h_ev, l1_ev = first_lagrange(zeta, n)
    pi_ev = eval_pi_poly(vp, pi, z_h_ev, zeta, root)
    r_ev = r_eval_zeta(proof, alpha, beta, gamma, pi_ev, l1_ev, vp["anemoi_generator"], vp["anemoi_generator_inv"])
    cms = {"q": vp["cm_q_vec"], "z": [proof["cm_z"]], "s_last": [vp["cm_s_vec"][N_WIRES - 1]], "qb": [vp["cm_qb"]], "prk": vp["cm_prk_vec"],
           "t": proof["cm_t_vec"]}
    cm_r = _g1_lin([(s, cms[name][i]) for s, (name, i) in r_scalars(
        vp["k"], proof["w_polys_eval_zeta"], proof["s_polys_eval_zeta"], proof["prk_3_poly_eval_zeta"], proof["z_eval_zeta_omega"],
        alpha, beta, gamma, zeta, l1_ev, z_h_ev, n + 2)])
    commitments = proof["cm_w_vec"] + vp["cm_s_vec"][:N_WIRES - 1] + [vp["cm_prk_vec"][2], vp["cm_prk_vec"][3], cm_r]
    values = proof["w_polys_eval_zeta"] + proof["s_polys_eval_zeta"] + [proof["prk_3_poly_eval_zeta"], proof["prk_4_poly_eval_zeta"], r_ev]
    zeta_omega = zeta * root % FR
    comm, val = _batch(tr, commitments, n + 2, zeta, values)
    comm_o, val_o = _batch(tr, [proof["cm_z"]] + proof["cm_w_vec"][:3], n + 2, zeta_omega,
                           [proof["z_eval_zeta_omega"]] + proof["w_polys_eval_zeta_omega"])
    # batch_verify_diff_points (kzg_poly_commitment.rs:407-460): sum_i u^i (C_i - v_i G + x_i W_i) = tau * sum_i u^i W_i
    G = (1, 2)
    W, Wo = proof["opening_witness_zeta"], proof["opening_witness_zeta_omega"]
    lhs = _g1_lin([(1, comm), ((FR - val) % FR, G), (zeta, W), (u, comm_o), ((FR - u * val_o) % FR, G), (u * zeta_omega, Wo)])
    rhs = _g1_lin([(pcs.tau, W), (pcs.tau * u, Wo)])
    return lhs == rhs
