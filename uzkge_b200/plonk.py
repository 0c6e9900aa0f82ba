"""Device-resident TurboPlonK prover: the host-side mirror of the reference's `plonk` module on top of the C ABI.

No Rust toolchain exists in this image, so what in production is the patched body of `prover_with_lagrange` is restated here in
Python with the reference's names and order of operations (the RNG draws and transcript appends are order-sensitive):

  TurboCS (gates, helper gadgets, verify_witness) /root/reference/uzkge/src/plonk/constraint_system/turbo/mod.rs:395-1396; the shuffle
                                                 and Anemoi gadgets are mixed in from shuffle.py / anemoi.py
  compute_permutation / extend_witness           /root/reference/uzkge/src/plonk/constraint_system/mod.rs:54-84, 103-111
  indexer                                        /root/reference/uzkge/src/plonk/indexer.rs:248-536   (both feature sets)
  refresh_prover_params_public_key               /root/reference/shuffle/src/gen_params/params.rs:57-129
  prover                                         /root/reference/uzkge/src/plonk/prover.rs:76-394     (lagrange_pcs = None)
  pi_poly, hide_polynomial, z_poly, t_poly,
  r_poly, split_t_and_commit, first_lagrange_poly /root/reference/uzkge/src/plonk/helpers.rs:111-131, 139-154, 160-220, 223-678, 681-999,
                                                 1323-1408, 1412-1423
  batch_prove                                    /root/reference/uzkge/src/poly_commit/pcs.rs:107-168

Every polynomial lives in HBM as 4 x u64 Montgomery limbs per coefficient (arkworks' layout) between the rounds; transforms, MSMs,
scans and the pointwise maps are the library's CUDA kernels; the host does the Fiat-Shamir transcript, the RNG and O(1) scalar
arithmetic per round (Python integers).  Only 64-byte commitments and 32-byte evaluations cross PCIe.  torch is used for device
allocations and device-to-device slice copies only.  There is no CPU path: without the CUDA library every step raises.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import ffi
from .errors import DegreeError, ParameterError, UzkgeError
from .poly_commit import FR_MODULUS, KZGCommitment, KZGCommitmentSchemeBN254
from .rng import ChaChaRng, choose_ks, fr_rand
from .anemoi import N_ANEMOI_ROUNDS, AnemoiGates
from .shuffle import ShuffleGates
from .transcript import Transcript, init_pcs_batch_eval_transcript, transcript_init_plonk

N_WIRES_PER_GATE = 5
N_SELECTORS = 9
_R = (1 << 256) % FR_MODULUS
_R_INV = pow(_R, -1, FR_MODULUS)
_M64 = 0xFFFFFFFFFFFFFFFF


# ---------------------------------------------------------------------------------------------- scalars
def mont(x: int) -> np.ndarray:
    """Canonical integer -> 4 Montgomery limbs."""
    v = x % FR_MODULUS * _R % FR_MODULUS
    return np.array([(v >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


def mont_rows(xs) -> np.ndarray:
    return np.stack([mont(x) for x in xs]) if len(xs) else np.zeros((0, 4), dtype=np.uint64)


def unmont(row) -> int:
    """4 Montgomery limbs -> canonical integer."""
    v = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
    return v * _R_INV % FR_MODULUS


_ZERO, _ONE = mont(0), mont(1)


# ---------------------------------------------------------------------------------------------- device vectors
class DevVec:
    """`cap` field elements in HBM (a torch int64 tensor of 4 * cap words); `len` = coefficients in use."""

    def __init__(self, cap: int, device, zero: bool = True, length: int | None = None):
        alloc = torch.zeros if zero else torch.empty
        self.t = alloc(4 * max(cap, 1), dtype=torch.int64, device=device)
        self.cap = cap
        self.len = cap if length is None else length

    @classmethod
    def from_numpy(cls, a: np.ndarray, device, cap: int | None = None) -> "DevVec":
        a = ffi.as_u64(a, 4)
        v = cls(max(cap or 0, a.shape[0]), device, zero=cap is not None and cap > a.shape[0], length=a.shape[0])
        v.t[: 4 * a.shape[0]].copy_(torch.from_numpy(a.view(np.int64).reshape(-1)))
        return v

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    def at(self, i: int) -> int:
        return self.t.data_ptr() + 32 * i

    def numpy(self, n: int | None = None) -> np.ndarray:
        n = self.len if n is None else n
        return self.t[: 4 * n].cpu().numpy().view(np.uint64).reshape(n, 4)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------- TurboCS
class TurboCS(ShuffleGates, AnemoiGates):
    """constraint_system/turbo/mod.rs.  Gates are appended one by one (Python lists) or in bulk (`synthetic`); `pad()` freezes
    the circuit into numpy arrays: selectors (9, n, 4) Montgomery limbs, wiring (5, n) uint32.  The shuffle gadgets
    (constraint_system/shuffle/*.rs) come from shuffle.ShuffleGates, the Anemoi ones (constraint_system/anemoi/mod.rs) from
    anemoi.AnemoiGates."""

    def __init__(self):
        self._sel = [[] for _ in range(N_SELECTORS)]     # small codes: index into _sel_values
        self._sel_values: list[int] = []
        self._sel_index: dict[int, int] = {}
        self._wir = [[] for _ in range(N_WIRES_PER_GATE)]
        self.num_vars = 2
        self.size = 0
        self.witness: list[int] | None = [0, 1]
        self.witness_array: np.ndarray | None = None
        self.public_vars_constraint_indices: list[int] = []
        self.public_vars_witness_indices: list[int] = []
        self.boolean_constraint_indices: list[int] = []
        self.selectors: np.ndarray | None = None
        self.wiring: np.ndarray | None = None
        self.verifier_only = False
        self._init_shuffle()
        self._init_anemoi()
        self.insert_constant_gate(self.zero_var(), 0)
        self.insert_constant_gate(self.one_var(), 1)

    # ---- ConstraintSystem trait
    @staticmethod
    def n_wires_per_gate() -> int:
        return N_WIRES_PER_GATE

    @staticmethod
    def num_selectors() -> int:
        return N_SELECTORS

    def quot_eval_dom_size(self) -> int:
        return self.size * 6 if self.size > 8 else self.size * 16

    @staticmethod
    def get_hiding_degree(idx: int) -> int:
        return 3 if idx < 3 else 2

    def is_verifier_only(self) -> bool:
        return self.verifier_only

    def zero_var(self) -> int:
        return 0

    def one_var(self) -> int:
        return 1

    def _code(self, v: int) -> int:
        v %= FR_MODULUS
        c = self._sel_index.get(v)
        if c is None:
            c = len(self._sel_values)
            self._sel_index[v] = c
            self._sel_values.append(v)
        return c

    def _push_gate(self, q_add, q_mul, q_c, q_ecc, q_out, wires) -> None:
        if self.selectors is not None:
            raise UzkgeError("the circuit is frozen (pad() was called)")
        if any(w >= self.num_vars for w in wires):
            raise ParameterError("wire index out of bound")
        vals = list(q_add) + list(q_mul) + [q_c, q_ecc, q_out]
        for j in range(N_SELECTORS):
            self._sel[j].append(self._code(vals[j]))
        for j in range(N_WIRES_PER_GATE):
            self._wir[j].append(wires[j])
        self.size += 1

    def new_variable(self, value: int) -> int:
        self.num_vars += 1
        self.witness.append(value % FR_MODULUS)
        return self.num_vars - 1

    def add_variables(self, values) -> None:
        for v in values:
            self.new_variable(v)

    def insert_lc_gate(self, wires_in, wire_out: int, q1: int, q2: int, q3: int, q4: int) -> None:
        self._push_gate((q1, q2, q3, q4), (0, 0), 0, 0, 1, list(wires_in) + [wire_out])

    def insert_add_gate(self, left_var: int, right_var: int, out_var: int) -> None:
        self.insert_lc_gate((left_var, right_var, 0, 0), out_var, 1, 1, 0, 0)

    def insert_sub_gate(self, left_var: int, right_var: int, out_var: int) -> None:
        self.insert_lc_gate((left_var, right_var, 0, 0), out_var, 1, FR_MODULUS - 1, 0, 0)

    def insert_mul_gate(self, left_var: int, right_var: int, out_var: int) -> None:
        self._push_gate((0, 0, 0, 0), (1, 0), 0, 0, 1, [left_var, right_var, 0, 0, out_var])

    def insert_constant_gate(self, var: int, constant: int) -> None:
        self._push_gate((0, 0, 0, 0), (0, 0), constant, 0, 1, [var] * N_WIRES_PER_GATE)

    def insert_boolean_gate(self, var: int) -> None:
        self.insert_mul_gate(var, var, var)

    def prepare_pi_variable(self, var: int) -> None:
        self.public_vars_witness_indices.append(var)
        self.public_vars_constraint_indices.append(self.size)
        self.insert_constant_gate(var, 0)

    def attach_boolean_constraint_to_gate(self) -> None:
        self.boolean_constraint_indices.append(self.size - 1)

    def add(self, left_var: int, right_var: int) -> int:
        out = self.new_variable(self.witness[left_var] + self.witness[right_var])
        self.insert_add_gate(left_var, right_var, out)
        return out

    def sub(self, left_var: int, right_var: int) -> int:
        out = self.new_variable(self.witness[left_var] - self.witness[right_var])
        self.insert_sub_gate(left_var, right_var, out)
        return out

    def mul(self, left_var: int, right_var: int) -> int:
        out = self.new_variable(self.witness[left_var] * self.witness[right_var])
        self.insert_mul_gate(left_var, right_var, out)
        return out

    def select(self, var0: int, var1: int, bit: int) -> int:
        """turbo/mod.rs:767-796: var_bit = (1 - bit) var0 + bit var1; wires (bit, var0, bit, var1), q2 = qm2 = qo = 1, qm1 = -1."""
        out = self.new_variable(self.witness[var0] if self.witness[bit] == 0 else self.witness[var1])
        self._push_gate((0, 1, 0, 0), (-1, 1), 0, 0, 1, [bit, var0, bit, var1, out])
        return out

    def range_check(self, var: int, n_bits: int) -> list:
        """turbo/mod.rs:703-763: constrain 0 <= witness[var] < 2^n_bits; returns the bit variables, little endian.  The bits are
        boolean-constrained through the qb selector of the accumulation gates (wires 2-4) and one boolean gate for the top bit."""
        if n_bits < 2:
            raise ParameterError("the number of bits is less than two")
        value = self.witness[var]
        b = [self.new_variable((value >> i) & 1) for i in range(n_bits)]
        acc = b[n_bits - 1]
        self.insert_boolean_gate(b[n_bits - 1])
        m = (n_bits - 2) // 3
        for i in range(m):
            acc = self.linear_combine([acc, b[n_bits - 1 - i * 3 - 1], b[n_bits - 1 - i * 3 - 2], b[n_bits - 1 - i * 3 - 3]], 8, 4, 2, 1)
            self.attach_boolean_constraint_to_gate()
        rest = (n_bits - 1) - 3 * m
        if rest == 1:
            self.insert_lc_gate([acc, b[0], 0, 0], var, 2, 1, 0, 0)
        elif rest == 2:
            self.insert_lc_gate([acc, b[1], b[0], 0], var, 4, 2, 1, 0)
        else:
            self.insert_lc_gate([acc, b[2], b[1], b[0]], var, 8, 4, 2, 1)
        self.attach_boolean_constraint_to_gate()
        return b

    def is_equal_or_not_equal(self, left_var: int, right_var: int) -> tuple:
        """turbo/mod.rs:814-836: two boolean variables, (1, 0) iff the values are equal, (0, 1) otherwise."""
        diff = self.sub(left_var, right_var)
        d = self.witness[diff]
        inv_diff = self.new_variable(pow(d, -1, FR_MODULUS) if d else 0)
        mul_var = self.mul(diff, inv_diff)
        diff_is_zero = self.sub(self.one_var(), mul_var)
        self.insert_mul_gate(diff, diff_is_zero, self.zero_var())
        return diff_is_zero, mul_var

    def is_equal(self, left_var: int, right_var: int) -> int:
        return self.is_equal_or_not_equal(left_var, right_var)[0]

    def is_not_equal(self, left_var: int, right_var: int) -> int:
        return self.is_equal_or_not_equal(left_var, right_var)[1]

    def equal(self, left_var: int, right_var: int) -> None:
        self.insert_sub_gate(left_var, right_var, self.zero_var())

    def pad(self) -> None:
        """turbo/mod.rs:968-977 (zero selectors, wires on variable 0), then freeze into arrays."""
        n = 1
        while n < self.size:
            n *= 2
        table = mont_rows(self._sel_values + [0])
        zero_code = len(self._sel_values)
        sel = np.full((N_SELECTORS, n), zero_code, dtype=np.int64)
        wir = np.zeros((N_WIRES_PER_GATE, n), dtype=np.uint32)
        for j in range(N_SELECTORS):
            sel[j, : self.size] = self._sel[j]
        for j in range(N_WIRES_PER_GATE):
            wir[j, : self.size] = self._wir[j]
        self.selectors = table[sel]
        self.wiring = wir
        self.size = n
        self._sel = self._wir = None
        self._sel_codes, self._sel_table = sel, self._sel_values + [0]       # kept for verify_witness

    # ---- `shuffle` feature set: selector tables as Montgomery arrays (turbo/mod.rs:171-191, 310-364; plonk/indexer.rs:417-424)
    def _scatter_rows(self, per_round_values, count: int) -> np.ndarray:
        """(count, n, 4) arrays, zero except on the rows of the remark gates: row first + j holds per_round_values[t][j]."""
        out = np.zeros((count, self.size, 4), dtype=np.uint64)
        firsts = np.asarray(self.shuffle_remark_constraint_indices(), dtype=np.int64)
        rounds = self.n_iteration_shuffle_scalar_mul
        if len(firsts) and rounds:
            rows = (firsts[:, None] + np.arange(rounds)[None, :]).reshape(-1)
            for t in range(count):
                vals = per_round_values(t)              # (cards, rounds, 4) or (rounds, 4) Montgomery limbs
                out[t, rows] = np.broadcast_to(vals, (len(firsts), rounds, 4)).reshape(-1, 4)
        return out

    def compute_witness_selectors(self) -> np.ndarray:
        cons, codes = self.shuffle_remark_constraints, self._remark_sel_codes
        if cons and all(c is not None for c in codes):            # remark traces: bits and signs only -> table lookup
            table = np.stack([_ZERO, _ONE, mont(-1)])
            arr = np.stack(codes)                                 # (cards, 3, rounds)
            return self._scatter_rows(lambda t: table[arr[:, t, :]], 3)
        return self._scatter_rows(lambda t: np.stack([mont_rows(sel[t]) for _, sel in cons]), 3)

    def compute_anemoi_jive_selectors(self) -> np.ndarray:
        """turbo/mod.rs:285-304 as (4, n, 4) Montgomery limbs."""
        out = np.zeros((4, self.size, 4), dtype=np.uint64)
        firsts = np.asarray(self.anemoi_constraints_indices, dtype=np.int64)
        if len(firsts):
            rows = (firsts[:, None] + np.arange(N_ANEMOI_ROUNDS)[None, :]).reshape(-1)
            kx, ky = self.anemoi_preprocessed_round_keys_x, self.anemoi_preprocessed_round_keys_y
            for t, col in enumerate(([k[0] for k in kx], [k[1] for k in kx], [k[0] for k in ky], [k[1] for k in ky])):
                out[t, rows] = np.broadcast_to(mont_rows(col), (len(firsts), N_ANEMOI_ROUNDS, 4)).reshape(-1, 4)
        return out

    def witness_selector_codes(self):
        """The witness selectors as (3, n) int32 indices into the table (0, 1, -1), or None when some remark gate carries other
        values: what the prover uploads instead of 3 n field elements."""
        codes = self._remark_sel_codes
        if not self.shuffle_remark_constraints or any(c is None for c in codes):
            return None
        firsts = np.asarray(self.shuffle_remark_constraint_indices(), dtype=np.int64)
        rounds = self.n_iteration_shuffle_scalar_mul
        rows = (firsts[:, None] + np.arange(rounds)[None, :]).reshape(-1)
        arr = np.stack(codes)                                     # (cards, 3, rounds) int32
        out = np.zeros((3, self.size), dtype=np.int32)
        for t in range(3):
            out[t, rows] = arr[:, t, :].reshape(-1)
        return out

    def _table_selectors(self, table) -> np.ndarray:
        if not self.shuffle_remark_constraints:
            return np.zeros((12, self.size, 4), dtype=np.uint64)
        return self._scatter_rows(lambda t: mont_rows([table[j][t % 4][t // 4] for j in range(self.n_iteration_shuffle_scalar_mul)]), 12)

    def compute_shuffle_generator_selectors(self) -> np.ndarray:
        return self._table_selectors(self.shuffle_generators)

    def compute_shuffle_public_key_selectors(self) -> np.ndarray:
        return self._table_selectors(self.shuffle_public_keys)

    def compute_q_ecc(self) -> np.ndarray:
        return self._scatter_rows(lambda t: mont_rows([1] * self.n_iteration_shuffle_scalar_mul), 1)[0]

    def verify_witness(self, witness, online_vars) -> None:
        """turbo/mod.rs:1041-1396: the Anemoi and remark equations, the gate equation with the public inputs, the boolean rows.  Raises UzkgeError naming the first failing row.  Host-side check, Python integers."""
        if self.selectors is None:
            raise UzkgeError("call cs.pad() before verify_witness")
        if len(witness) != self.num_vars:
            raise UzkgeError(f"witness len = {len(witness)}, num_vars = {self.num_vars}")
        if not (len(online_vars) == len(self.public_vars_witness_indices) == len(self.public_vars_constraint_indices)):
            raise UzkgeError("wrong number of online variables")
        R, wir = FR_MODULUS, self.wiring
        err = self.check_anemoi_rows(witness, lambda j, row: int(wir[j, row]))
        if err:
            raise UzkgeError(err)
        a_ed = self.edwards_a
        for first, sel in self.shuffle_remark_constraints:
            for r in range(self.n_iteration_shuffle_scalar_mul):
                row = first + r
                a, b, c, d, o = (witness[int(wir[j, row])] for j in range(5))
                an, bn, cn = (witness[int(wir[j, row + 1])] for j in range(3))
                s1, s2, s3 = sel[0][r], sel[1][r], sel[2][r]
                if s1 not in (0, 1) or s2 not in (0, 1) or s3 not in (1, R - 1):
                    raise UzkgeError(f"cs index {first} round {r}: wire selectors out of range")
                w4 = [(1 - s1) * (1 - s2), s1 * (1 - s2), (1 - s1) * s2, s1 * s2]
                e = [0, 0, 0, 0]
                for t in range(4):
                    px, py, pd = self.shuffle_public_keys[r][t]
                    gx, gy, gd = self.shuffle_generators[r][t]
                    e[0] += w4[t] * (s3 * an - s3 * a * py - b * px + a * b * an * pd)
                    e[1] += w4[t] * (s3 * bn + a_ed * a * px - s3 * b * py - a * b * bn * pd)
                    e[2] += w4[t] * (s3 * cn - s3 * c * gy - d * gx + c * d * cn * gd)
                    e[3] += w4[t] * (s3 * o + a_ed * c * gx - s3 * d * gy - c * d * o * gd)
                for t in range(4):
                    if e[t] % R:
                        raise UzkgeError(f"cs index {first} round {r}: equation {t + 1} of shuffle does not hold")
        pub = {}
        for c_i, w_i, v in zip(self.public_vars_constraint_indices, self.public_vars_witness_indices, online_vars):
            if witness[w_i] != v % R:
                raise UzkgeError(f"cs index {c_i}: online var does not match witness")
            pub[c_i] = v % R
        booleans = set(self.boolean_constraint_indices)
        tab = self._sel_table
        for row in range(self.size):
            w = [witness[int(wir[j, row])] for j in range(5)]
            q = [tab[int(self._sel_codes[j, row])] for j in range(N_SELECTORS)]
            g = (q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3] + q[4] * w[0] * w[1] + q[5] * w[2] * w[3] + q[6] + pub.get(row, 0)
                 + q[7] * w[0] * w[1] * w[2] * w[3] * w[4] - q[8] * w[4])
            if g % R:
                raise UzkgeError(f"cs index {row}: gate equation does not hold")
            if row in booleans and any(w[j] not in (0, 1) for j in (1, 2, 3)):
                raise UzkgeError(f"cs index {row}: a boolean-constrained wire is not one or zero")

    def get_witness_array(self) -> np.ndarray:
        """The witness as (num_vars, 4) Montgomery limbs."""
        if self.witness_array is None:
            self.witness_array = mont_rows(self.witness)
        return self.witness_array

    def compute_permutation(self) -> np.ndarray:
        """constraint_system/mod.rs:54-84: every variable's positions (wire-major order) form one cycle, each position pointing to
        the next larger one and the last back to the first.  (The reference's loop is quadratic; this is the same map, sorted.)"""
        v = self.wiring.reshape(-1).astype(np.int64)
        order = np.argsort(v, kind="stable")
        sv = v[order]
        last_of_group = np.ones(len(v), dtype=bool)
        last_of_group[:-1] = sv[1:] != sv[:-1]
        first_of_group = np.ones(len(v), dtype=bool)
        first_of_group[1:] = sv[1:] != sv[:-1]
        group_start = np.maximum.accumulate(np.where(first_of_group, np.arange(len(v)), 0))
        nxt = np.empty(len(v), dtype=np.int64)
        nxt[:-1] = order[1:]
        nxt[last_of_group] = order[group_start[last_of_group]]
        perm = np.empty(len(v), dtype=np.int64)
        perm[order] = nxt
        return perm

    @classmethod
    def synthetic(cls, log_size: int, seed: int = 0xB2000004, layers: int = 8, witness: str = "uniform") -> "TurboCS":
        """a satisfied circuit of 2^log_size gates made of insert_add_gate / insert_mul_gate (turbo/mod.rs:481-522) in
        `layers` layers, every gate reading two outputs of the previous layer (so the copy constraints are not trivial).  The
        witness is evaluated layer by layer on the GPU (gather, pointwise product, sum).  No public inputs.
        witness = "uniform": the free inputs are uniform field elements (every wire value is full width: the worst case for the
        commitments); "bits": the inputs are bits, so all values stay below 2^layers -- the zeros / bits / small integers that
        dominate real circuits' wire vectors."""
        n = 1 << log_size
        cs = cls()
        rng = np.random.default_rng(seed)
        g = n - cs.size                      # gates to add; the two constant gates of `new` come first
        per = [g // layers + (1 if i < g % layers else 0) for i in range(layers)]
        dev = _dev()
        n_inputs = max(per[0], 2)
        if witness == "bits":
            inputs = np.stack([_ZERO, _ONE])[rng.integers(0, 2, n_inputs)]
        elif witness == "uniform":
            inputs = _random_fr(n_inputs, seed ^ 0x5EED)
        else:
            raise ParameterError("witness must be 'uniform' or 'bits'")
        n_vars = 2 + n_inputs + g
        wit = DevVec(n_vars, dev)
        head = np.concatenate([np.stack([_ZERO, _ONE]), inputs])
        wit.t[: 4 * head.shape[0]].copy_(torch.from_numpy(head.view(np.int64).reshape(-1)))
        sel = np.zeros((N_SELECTORS, n), dtype=np.int8)     # codes: 0 -> 0, 1 -> 1
        wir = np.zeros((N_WIRES_PER_GATE, n), dtype=np.uint32)
        sel[6, 1] = 1                                        # constant gates of TurboCS::new: q_c = 0, 1; q_out = 1
        sel[8, :2] = 1
        wir[:, 1] = 1
        prev_lo, prev_n = 2, n_inputs
        row, var = 2, 2 + n_inputs
        ones = np.stack([_ONE, _ONE])
        for cnt in per:
            if cnt == 0:
                continue
            a = (prev_lo + rng.integers(0, prev_n, cnt)).astype(np.uint32)
            b = (prev_lo + rng.integers(0, prev_n, cnt)).astype(np.uint32)
            is_mul = rng.integers(0, 2, cnt).astype(bool)
            out = np.arange(var, var + cnt, dtype=np.uint32)
            wir[0, row:row + cnt], wir[1, row:row + cnt], wir[4, row:row + cnt] = a, b, out
            sel[0, row:row + cnt] = sel[1, row:row + cnt] = ~is_mul
            sel[4, row:row + cnt] = is_mul
            sel[8, row:row + cnt] = 1
            da, db = DevVec(cnt, dev, zero=False), DevVec(cnt, dev, zero=False)
            ia, ib = torch.from_numpy(a.view(np.int32)).to(dev), torch.from_numpy(b.view(np.int32)).to(dev)
            ffi.fr_gather_device(wit.ptr, ia.data_ptr(), cnt, da.ptr)
            ffi.fr_gather_device(wit.ptr, ib.data_ptr(), cnt, db.ptr)
            prod, summ = DevVec(cnt, dev, zero=False), DevVec(cnt, dev, zero=False)
            ffi.fr_mul_device(da.ptr, db.ptr, cnt, prod.ptr)
            ffi.fr_lincomb_device([da.ptr, db.ptr], [cnt, cnt], ones, summ.ptr, cnt)
            m = torch.from_numpy(is_mul).to(dev).repeat_interleave(4)
            wit.t[4 * var: 4 * (var + cnt)].copy_(torch.where(m, prod.t[: 4 * cnt], summ.t[: 4 * cnt]))
            prev_lo, prev_n = var, cnt
            row += cnt
            var += cnt
        torch.cuda.synchronize()
        cs.selectors = np.stack([_ZERO, _ONE])[sel.astype(np.int64)]
        cs.wiring = wir
        cs.size = n
        cs.num_vars = n_vars
        cs.witness = None
        cs.witness_array = wit.numpy(n_vars)
        cs._sel = cs._wir = None
        return cs


def _random_fr(n: int, seed: int) -> np.ndarray:
    """n uniform Montgomery residues < r (numpy PCG64, rejection on the top limb)."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 61) - 1)      # < 2^253 < r: uniform enough for synthetic witnesses
    return out


# ---------------------------------------------------------------------------------------------- parameters
@dataclass
class PlonkVerifierParams:
    """plonk/indexer.rs:155-192 (default features)."""
    cm_q_vec: list
    cm_s_vec: list
    cm_qb: KZGCommitment
    cm_prk_vec: list
    anemoi_generator: int
    anemoi_generator_inv: int
    k: list
    cs_size: int
    public_vars_constraint_indices: list
    lagrange_constants: list
    # `shuffle` feature set (plonk/indexer.rs:163-176)
    cm_q_ecc: KZGCommitment | None = None
    cm_shuffle_generator_vec: list | None = None
    cm_shuffle_public_key_vec: list | None = None
    edwards_a: int = 0


@dataclass
class PlonkProverParams:
    """plonk/indexer.rs:77-139: coefficient forms and coset evaluations, all resident in HBM."""
    q_polys: list
    s_polys: list
    qb_poly: DevVec
    q_prk_polys: list
    verifier_params: PlonkVerifierParams
    group: DevVec
    coset_quotient: DevVec
    l1_coset_evals: DevVec
    z_h_inv_coset_evals: np.ndarray
    q_coset_evals: list
    s_coset_evals: list
    qb_coset_eval: DevVec
    q_prk_coset_evals: list
    sigma_evals: DevVec                 # encode_perm_to_group of the permutation: 5 n values (the evaluation form of s_polys)
    wiring: torch.Tensor                # 5 n uint32 variable indices (extend_witness)
    n: int = 0
    m: int = 0
    factor: int = 0
    root: int = 0
    root_m: int = 0
    workspace: dict = field(default_factory=dict)
    # `shuffle` feature set (plonk/indexer.rs:100-139): None when the parameters were built for the default features
    q_ecc_poly: DevVec | None = None
    q_ecc_coset_eval: DevVec | None = None
    q_shuffle_generator_polys: list | None = None
    q_shuffle_generator_coset_evals: list | None = None
    q_shuffle_public_key_polys: list | None = None
    q_shuffle_public_key_coset_evals: list | None = None

    def get_verifier_params_ref(self) -> PlonkVerifierParams:
        return self.verifier_params


@dataclass
class PlonkProof:
    """plonk/indexer.rs:33-75 (default features)."""
    cm_w_vec: list
    cm_t_vec: list
    cm_z: KZGCommitment
    prk_3_poly_eval_zeta: int
    prk_4_poly_eval_zeta: int
    w_polys_eval_zeta: list
    w_polys_eval_zeta_omega: list
    z_eval_zeta_omega: int
    s_polys_eval_zeta: list
    opening_witness_zeta: KZGCommitment
    opening_witness_zeta_omega: KZGCommitment
    # `shuffle` feature set
    cm_w_sel_vec: list | None = None
    q_ecc_poly_eval_zeta: int | None = None
    w_sel_polys_eval_zeta: list | None = None

    def to_bytes_be(self) -> bytes:
        """PlonkProof::to_bytes_be (plonk/indexer.rs:538-590): what the Solidity verifier of the reference consumes."""
        sc = lambda v: int(v).to_bytes(32, "big")
        pts = lambda cms: b"".join(c.to_transcript_bytes() for c in cms)
        out = pts(self.cm_w_vec) + pts(self.cm_w_sel_vec or []) + pts(self.cm_t_vec) + pts([self.cm_z])
        out += sc(self.prk_3_poly_eval_zeta) + sc(self.prk_4_poly_eval_zeta)
        out += b"".join(sc(v) for v in self.w_polys_eval_zeta + self.w_polys_eval_zeta_omega) + sc(self.z_eval_zeta_omega)
        out += b"".join(sc(v) for v in self.s_polys_eval_zeta)
        if self.cm_w_sel_vec is not None:
            out += sc(self.q_ecc_poly_eval_zeta) + b"".join(sc(v) for v in self.w_sel_polys_eval_zeta)
        return out + pts([self.opening_witness_zeta, self.opening_witness_zeta_omega])


# ---------------------------------------------------------------------------------------------- device helpers
def _root(n: int) -> int:
    return unmont(ffi.fr_root_of_unity(n))


def _ifft(src_ptr: int, n: int, out: DevVec, scratch: DevVec) -> None:
    ffi.ntt_fr_device(src_ptr, out.ptr, scratch.ptr, n, n, True, None)


def _coset_fft(poly: DevVec, m: int, k1: np.ndarray, out: DevVec, scratch: DevVec) -> None:
    ffi.ntt_fr_device(poly.ptr, out.ptr, scratch.ptr, poly.len, m, False, k1)


def _commit_dev(pcs, vecs, overlap=None) -> list:
    """PolyComScheme::commit for device-resident coefficient vectors; independent commitments share one pass when they fit.
    `pcs` is a KZGCommitmentSchemeBN254 (bases on this GPU) or anything with `commit_device` (dist.SplitCommitter: bases split
    over the GPUs of the box).  `overlap`: a callable that enqueues device work NOT depending on these commitments; it is issued
    after the MSMs and before the host waits for them, so the GPU keeps working while the host hashes the transcript."""
    for v in vecs:
        if v.len > pcs.max_degree() + 1:
            raise DegreeError("DegreeError")
    if hasattr(pcs, "commit_device"):
        out = pcs.commit_device(vecs)
        if overlap:
            overlap()
        return out
    dev = vecs[0].t.device
    out = torch.zeros(12 * len(vecs), dtype=torch.int64, device=dev)
    if len(vecs) > 1:   # one call: the engine batches what fits one pass and pipelines the rest
        ffi.msm_g1_batch_device(pcs.handle, [v.ptr for v in vecs], [v.len for v in vecs], out.data_ptr())
    else:
        for i, v in enumerate(vecs):
            ffi.msm_g1_device(pcs.handle, v.ptr, v.len, out.data_ptr() + 96 * i)
    if overlap:
        # read the results back on a side stream that waits for the MSMs only, not for the overlapped work behind them
        done = torch.cuda.Event()
        done.record()
        overlap()
        side = _side_stream(dev)
        side.wait_event(done)
        out.record_stream(side)
        with torch.cuda.stream(side):
            host = out.cpu()
    else:
        host = out.cpu()
    jac = host.numpy().view(np.uint64).reshape(len(vecs), 12)
    return [KZGCommitment(jac[i].copy()) for i in range(len(vecs))]


_SIDE_STREAMS: dict = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


N_BLIND_SLOTS = 3   # the largest hiding degree (TurboCS::get_hiding_degree, z_poly)


class _View:
    """A (pointer, length) window into a DevVec, accepted wherever a coefficient vector is committed."""

    def __init__(self, base: DevVec, first: int, length: int):
        self.t, self.ptr, self.len = base.t, base.at(first), length


def _lagrange_commit_scheme(pcs, lagrange_pcs, n: int, ws: dict):
    """The SRS the `commit` closure of prover_with_lagrange needs (plonk/prover.rs:131-146), as ONE base vector:
    [L_0(tau) G .. L_{n-1}(tau) G | SRS[0..3) | SRS[n..n+3)].  lagrange_pcs.commit(evals) + pcs.apply_blind_factors(cm, blinds, n)
    = C + sum_i b_i (SRS[i] - SRS[n + i]) (kzg_poly_commitment.rs:299-313) is then a single MSM over the scalars
    [evals | b_0 b_1 b_2 | -b_0 -b_1 -b_2] -- no separate scalar multiplications for the blinds."""
    cached = ws.get("lagrange_scheme")
    if cached is not None and cached[0] is lagrange_pcs:
        return cached[1]
    mono = pcs.public_parameter_group_1
    idx = list(range(N_BLIND_SLOTS)) + [n + i for i in range(N_BLIND_SLOTS)]
    pts = np.concatenate([lagrange_pcs.public_parameter_group_1[:n], mono[idx]])
    scheme = KZGCommitmentSchemeBN254(pts, getattr(lagrange_pcs, "window_bits", 0))
    ws["lagrange_scheme"] = (lagrange_pcs, scheme)
    return scheme


def _commit_coefs_lagrange(lag, polys, n: int, scratch: DevVec) -> list:
    """Commit COEFFICIENT vectors of up to n + 3 entries over the Lagrange SRS (helpers.rs:1363-1391, pcs.rs:139-163): with
    f = f_lo + X^n f_hi, f(tau) G = MSM(L_i(tau) G, f_lo on H) + sum_i f_hi[i] SRS[n + i], i.e. one size-n forward transform and
    one MSM over `_lagrange_commit_scheme`'s bases with the scalars [f_lo on H | 0 0 0 | f_hi].  (The reference folds f_hi into
    f_lo and cancels it again through apply_blind_factors; the group element is the same.)"""
    dev = polys[0].t.device
    stride = n + 8
    buf = DevVec(len(polys) * stride, dev)
    for i, p in enumerate(polys):
        if p.len > n + N_BLIND_SLOTS:
            raise DegreeError("DegreeError")
        if p.len == 0:
            continue
        ffi.ntt_fr_device(p.ptr, buf.at(i * stride), scratch.ptr, min(p.len, n), n, False, None)
        hi = p.len - n
        if hi > 0:
            buf.t[4 * (i * stride + n + N_BLIND_SLOTS): 4 * (i * stride + n + N_BLIND_SLOTS + hi)].copy_(p.t[4 * n: 4 * (n + hi)])
    return _commit_dev(lag, [_View(buf, i * stride, n + 2 * N_BLIND_SLOTS) for i in range(len(polys))])


def _set_blind_slots(buf: DevVec, first: int, blinds) -> None:
    """Slots [first, first + 6) <- [b_0 b_1 b_2 | -b_0 -b_1 -b_2] (missing blinds are zero)."""
    b = list(blinds) + [0] * (N_BLIND_SLOTS - len(blinds))
    rows = mont_rows(b + [-x for x in b])
    buf.t[4 * first: 4 * (first + 2 * N_BLIND_SLOTS)].copy_(torch.from_numpy(rows.view(np.int64).reshape(-1)))


def _evals(polys_points, dev) -> list[int]:
    """FpPolynomial::eval for (poly, point) pairs: one batched Horner evaluation on the device (two launches), one small D2H."""
    vals = torch.zeros(4 * len(polys_points), dtype=torch.int64, device=dev)
    points = []
    for _, x in polys_points:
        if x not in points:
            points.append(x)
    if len(points) <= 2 and len(polys_points) <= ffi.EVAL_BATCH_MAX:
        ffi.poly_eval_batch_fr_device([p.ptr for p, _ in polys_points], [max(p.len, 1) for p, _ in polys_points],
                                      [points.index(x) for _, x in polys_points], mont_rows(points), vals.data_ptr())
    else:
        for i, (p, x) in enumerate(polys_points):
            ffi.poly_horner_fr_device(p.ptr, max(p.len, 1), mont(x), 0, vals.data_ptr() + 32 * i)
    rows = vals.cpu().numpy().view(np.uint64).reshape(-1, 4)
    return [unmont(r) for r in rows]


def _add_coefs(poly: DevVec, idx, vals) -> None:
    """FpPolynomial::add_coef_assign for a few coefficients (grows `len` like the reference's resize)."""
    for i in range(0, len(idx), ffi.SPARSE_MAX):
        ffi.fr_add_sparse_device(poly.ptr, idx[i:i + ffi.SPARSE_MAX], mont_rows(vals[i:i + ffi.SPARSE_MAX]))
    poly.len = max(poly.len, max(idx) + 1)
    assert poly.len <= poly.cap


# ---------------------------------------------------------------------------------------------- indexer
def indexer(cs: TurboCS, pcs, shuffle: bool = False, lagrange_pcs=None) -> PlonkProverParams:
    """plonk/indexer.rs:240-536 with permutation = None, verifier_params = None.  shuffle = True builds the parameters of the
    `shuffle` feature set (what zshuffle is compiled with): q_ecc and the 12 + 12 shuffle selector polynomials
    (indexer.rs:414-476) from the circuit's remark gates (zero polynomials when it has none); the public-key selectors start as a
    copy of the generator selectors (indexer.rs:471-476) until refresh_prover_params_public_key loads a key.  The prover then
    also commits the witness-selector polynomials, evaluates terms 12-18 of the quotient and opens q_ecc / w_sel, i.e. produces the
    proof format of the reference's deployed verifier.  lagrange_pcs: when given (and of the circuit's size) every commitment
    is the MSM of the EVALUATIONS over the Lagrange SRS (indexer.rs:284-301, `commit` closure) instead of the coefficients'."""
    if cs.selectors is None:
        raise UzkgeError("call cs.pad() before indexing")
    n, m = cs.size, cs.quot_eval_dom_size()
    factor = m // n
    if n * factor != m:
        raise UzkgeError("SetupError")
    dev = _dev()
    root, root_m = _root(n), _root(m)
    k = choose_ks(ChaChaRng.from_seed(bytes(32)), N_WIRES_PER_GATE)       # indexer.rs:258: fixed seed
    k1 = mont(k[1])
    scratch = DevVec(m, dev, zero=False)
    group = DevVec(n, dev, zero=False)
    ffi.fr_powers_device(mont(root), n, group.ptr)
    coset_quotient = DevVec(m, dev, zero=False)
    ffi.fr_powers_device(mont(root_m), m, coset_quotient.ptr, scale=k1)

    # Step 1: permutation polynomials.  table[c n + i] = k_c w^i, sigma = table[perm]
    perm = cs.compute_permutation()
    table = DevVec(N_WIRES_PER_GATE * n, dev, zero=False)
    for c in range(N_WIRES_PER_GATE):
        ffi.fr_powers_device(mont(root), n, table.at(c * n), scale=mont(k[c]))
    sigma = DevVec(N_WIRES_PER_GATE * n, dev, zero=False)
    d_perm = torch.from_numpy(perm.astype(np.uint32).view(np.int32)).to(dev)
    ffi.fr_gather_device(table.ptr, d_perm.data_ptr(), N_WIRES_PER_GATE * n, sigma.ptr)
    del table

    if lagrange_pcs is not None and lagrange_pcs.max_degree() + 1 != n:
        lagrange_pcs = None                                                # indexer.rs:262-267
    commit_src = {}      # id(coefficient vector) -> what is committed for it: the evaluations on the Lagrange path

    def preprocess(evals):
        """evals: a DevVec or _View of n values on H -> (coefficients, coset evaluations)."""
        coefs = DevVec(n, dev, zero=False)
        _ifft(evals.ptr, n, coefs, scratch)
        coset = DevVec(m, dev, zero=False)
        _coset_fft(coefs, m, k1, coset, scratch)
        commit_src[id(coefs)] = evals if lagrange_pcs is not None else coefs
        return coefs, coset

    def commit_all(polys):
        if lagrange_pcs is not None:
            return _commit_many(lagrange_pcs, [commit_src[id(p)] for p in polys])
        return _commit_many(pcs, polys)

    s_polys, s_coset = [], []
    for i in range(N_WIRES_PER_GATE):
        c, e = preprocess(_View(sigma, i * n, n))
        s_polys.append(c)
        s_coset.append(e)
    # Step 2: selector polynomials
    q_polys, q_coset = [], []
    for i in range(N_SELECTORS):
        c, e = preprocess(DevVec.from_numpy(cs.selectors[i], dev))
        q_polys.append(c)
        q_coset.append(e)
    # Step 3: L1 and Z_H
    l1 = DevVec(n, dev)
    ffi.fr_add_sparse_device(l1.ptr, [0], mont_rows([n]))
    _l1_coefs, l1_coset = preprocess(l1)
    z_h_inv = []
    mult, step = pow(k[1], n, FR_MODULUS), pow(root_m, n, FR_MODULUS)
    for _ in range(factor):
        z_h_inv.append(pow((mult - 1) % FR_MODULUS, -1, FR_MODULUS))
        mult = mult * step % FR_MODULUS
    # Step 4: Lagrange constants (helpers.rs:1170-1179), only for the public-input rows
    lagrange_constants = []
    for ci in cs.public_vars_constraint_indices:
        # prod_{i != j} (w^j - w^i) = n * w^{-j}  (derivative of X^n - 1 at w^j)
        lagrange_constants.append(pow(n * pow(root, -ci, FR_MODULUS) % FR_MODULUS, -1, FR_MODULUS))
    # Step 5: boolean constraints; Step 6: Anemoi round keys (zero polynomials without Anemoi gates)
    zero_poly, zero_coset = DevVec(n, dev), DevVec(m, dev)
    if cs.boolean_constraint_indices:
        qb = DevVec(n, dev)
        idx = list(cs.boolean_constraint_indices)
        for i in range(0, len(idx), ffi.SPARSE_MAX):
            ffi.fr_add_sparse_device(qb.ptr, idx[i:i + ffi.SPARSE_MAX], mont_rows([1] * len(idx[i:i + ffi.SPARSE_MAX])))
        qb_poly, qb_coset = preprocess(qb)
    else:
        qb_poly, qb_coset = zero_poly, zero_coset
    if cs.anemoi_constraints_indices:      # indexer.rs:385-412: the preprocessed round keys on the rows of the Anemoi gates
        prk_evals = cs.compute_anemoi_jive_selectors()
        prk = [preprocess(DevVec.from_numpy(prk_evals[i], dev)) for i in range(4)]
        q_prk_polys, q_prk_coset = [c for c, _ in prk], [e for _, e in prk]
    else:
        q_prk_polys, q_prk_coset = [zero_poly] * 4, [zero_coset] * 4

    commit_src[id(zero_poly)] = zero_poly
    n_fixed = N_SELECTORS + N_WIRES_PER_GATE
    cms = commit_all(q_polys + s_polys + [qb_poly, zero_poly] + (q_prk_polys if cs.anemoi_constraints_indices else []))
    identity = cms[n_fixed + 1]
    vp = PlonkVerifierParams(
        cm_q_vec=cms[:N_SELECTORS], cm_s_vec=cms[N_SELECTORS:n_fixed], cm_qb=cms[n_fixed],
        cm_prk_vec=cms[n_fixed + 2:] if cs.anemoi_constraints_indices else [identity] * 4,
        anemoi_generator=cs.anemoi_generator, anemoi_generator_inv=cs.anemoi_generator_inv, k=k, cs_size=n,
        public_vars_constraint_indices=list(cs.public_vars_constraint_indices), lagrange_constants=lagrange_constants)
    d_wiring = torch.from_numpy(cs.wiring.reshape(-1).view(np.int32)).to(dev)
    extra = {}
    if shuffle:
        # Steps 7-9 (indexer.rs:414-476): q_ecc = 1 on the rows of the remark gates, the 12 generator selectors, and the
        # public-key selectors as a copy of them
        if cs.shuffle_remark_constraints:
            q_ecc_poly, q_ecc_coset = preprocess(DevVec.from_numpy(cs.compute_q_ecc(), dev))
            gen_evals = cs.compute_shuffle_generator_selectors()
            gen = [preprocess(DevVec.from_numpy(gen_evals[i], dev)) for i in range(12)]
            gen_polys, gen_coset = [g[0] for g in gen], [g[1] for g in gen]
            cms2 = commit_all([q_ecc_poly] + gen_polys)
            vp.cm_q_ecc, vp.cm_shuffle_generator_vec = cms2[0], cms2[1:]
        else:
            q_ecc_poly, q_ecc_coset, gen_polys, gen_coset = zero_poly, zero_coset, [zero_poly] * 12, [zero_coset] * 12
            vp.cm_q_ecc, vp.cm_shuffle_generator_vec = identity, [identity] * 12
        vp.cm_shuffle_public_key_vec = list(vp.cm_shuffle_generator_vec)
        vp.edwards_a = cs.edwards_a
        extra = dict(q_ecc_poly=q_ecc_poly, q_ecc_coset_eval=q_ecc_coset, q_shuffle_generator_polys=gen_polys,
                     q_shuffle_generator_coset_evals=gen_coset, q_shuffle_public_key_polys=list(gen_polys),
                     q_shuffle_public_key_coset_evals=list(gen_coset))
    torch.cuda.synchronize()
    return PlonkProverParams(**extra, **dict(
        q_polys=q_polys, s_polys=s_polys, qb_poly=qb_poly, q_prk_polys=q_prk_polys, verifier_params=vp, group=group,
        coset_quotient=coset_quotient, l1_coset_evals=l1_coset, z_h_inv_coset_evals=mont_rows(z_h_inv), q_coset_evals=q_coset,
        s_coset_evals=s_coset, qb_coset_eval=qb_coset, q_prk_coset_evals=q_prk_coset, sigma_evals=sigma, wiring=d_wiring,
        n=n, m=m, factor=factor, root=root, root_m=root_m, workspace={"scratch": scratch}))


def _commit_many(pcs, vecs) -> list:
    return _commit_dev(pcs, vecs)


def refresh_prover_params_public_key(cs: TurboCS, prover_params: PlonkProverParams, pcs, shuffle_pk, lagrange_pcs=None) -> list:
    """shuffle/src/gen_params/params.rs:57-129: load a new joint public key into the circuit and rebuild what depends on it -- the
    12 public-key selector polynomials (ifft over H), their evaluations on the quotient coset (coset fft over 6n points) and
    their commitments (over the Lagrange SRS when one of the circuit's size is given, else over the monomial SRS).  Updates
    `prover_params` in place and returns the 12 commitments.  (The reference reloads both SRS files here; the caller passes the
    resident ones.)"""
    P = prover_params
    if P.q_shuffle_public_key_polys is None:
        raise ParameterError("the parameters were built without the shuffle feature set")
    cs.load_shuffle_remark_parameters(shuffle_pk)
    n, m = cs.size, cs.quot_eval_dom_size()
    if m % n != 0 or n != P.n:
        raise ParameterError("ParameterError")
    if lagrange_pcs is not None and lagrange_pcs.max_degree() + 1 != n:
        lagrange_pcs = None
    dev = _dev()
    scratch = P.workspace.get("scratch") or DevVec(m, dev, zero=False)
    k1 = mont(P.verifier_params.k[1])
    evals = cs.compute_shuffle_public_key_selectors()
    d_evals, polys, cosets = [], [], []
    for i in range(12):
        ev = DevVec.from_numpy(evals[i], dev)
        coefs, coset = DevVec(n, dev, zero=False), DevVec(m, dev, zero=False)
        _ifft(ev.ptr, n, coefs, scratch)
        _coset_fft(coefs, m, k1, coset, scratch)
        d_evals.append(ev)
        polys.append(coefs)
        cosets.append(coset)
    cms = _commit_many(lagrange_pcs, d_evals) if lagrange_pcs is not None else _commit_many(pcs, polys)
    torch.cuda.synchronize()
    P.q_shuffle_public_key_polys, P.q_shuffle_public_key_coset_evals = polys, cosets
    P.verifier_params.cm_shuffle_public_key_vec = cms
    P.workspace.pop("coset_params", None)
    return cms


# ---------------------------------------------------------------------------------------------- the quotient round, coset by coset
# The 6 n points k[1] w_m^p of the quotient domain are the 6 cosets g_j <w_n>, g_j = k[1] w_m^j (p = 6 i + j).  On one coset a
# polynomial of n + 3 coefficients is a size-n coset transform of its folded coefficients (X^n = g_j^n there), the w-shifted point
# of p is the next point of the same coset, and Z_H is the constant g_j^n - 1: the quotient map of a coset needs nothing from the
# other cosets, so the round splits over GPUs by coset (dist.SplitCommitter.quotient_*).
class CosetParams:
    """Everything the quotient map needs on coset j, built from the COEFFICIENT forms of the preprocessed polynomials."""

    def __init__(self, q_polys, s_polys, qb_poly, q_prk_polys, k, n: int, j: int, dev, anemoi=(0, 0)):
        m = 6 * n
        self.anemoi = (mont(anemoi[0]), mont(anemoi[1]))      # generator and its inverse (quotient terms 8-11)
        root_m, root_n = _root(m), _root(n)
        self.n, self.j = n, j
        self.g = k[1] * pow(root_m, j, FR_MODULUS) % FR_MODULUS
        self.g_n = pow(self.g, n, FR_MODULUS)
        self.z_h_inv = pow((self.g_n - 1) % FR_MODULUS, -1, FR_MODULUS)
        self.scratch = DevVec(n, dev, zero=False)
        g_m = mont(self.g)

        def ev(poly):
            out = DevVec(n, dev, zero=False)
            ffi.ntt_fr_device(poly.ptr, out.ptr, self.scratch.ptr, min(poly.len, n), n, False, g_m)
            return out

        cache = {}

        def ev_shared(poly):            # the zero selectors share one buffer in the parameters: keep sharing their evaluations
            if poly.ptr not in cache:
                cache[poly.ptr] = ev(poly)
            return cache[poly.ptr]

        self.q = [ev_shared(p) for p in q_polys]
        self.s = [ev_shared(p) for p in s_polys]
        self.qb = ev_shared(qb_poly)
        self.q_prk = [ev_shared(p) for p in q_prk_polys]
        ones = DevVec(n, dev, zero=False)                       # l1_coefs = 1 + X + ... + X^(n-1)  (indexer.rs:343-346)
        ffi.fr_powers_device(mont(1), n, ones.ptr)
        self.l1 = ev(ones)
        self.coset_quotient = DevVec(n, dev, zero=False)        # g_j w_n^i
        ffi.fr_powers_device(mont(root_n), n, self.coset_quotient.ptr, scale=g_m)
        self.evals = [DevVec(n, dev, zero=(i == 6)) for i in range(7)]   # w0..w4, z, pi on this coset (pi stays zero if never given)
        self.head = DevVec(4, dev, zero=False)

    def eval_folded(self, poly_t, length: int, out: DevVec) -> None:
        """out = the polynomial (`length` <= n + 3 coefficients in `poly_t`) on this coset: fold X^(n + l) = g^n X^l into the head
        coefficients (restored afterwards), then one size-n coset transform."""
        n = self.n
        extra = max(0, length - n)
        ptr = poly_t.data_ptr()
        if extra:
            self.head.t[: 4 * extra].copy_(poly_t[: 4 * extra])
            ffi.fr_lincomb_device([ptr, ptr + 32 * n], [extra, extra], mont_rows([1, self.g_n]), ptr, extra)
        ffi.ntt_fr_device(ptr, out.ptr, self.scratch.ptr, min(length, n), n, False, mont(self.g))
        if extra:
            poly_t[: 4 * extra].copy_(self.head.t[: 4 * extra])

    def quotient(self, polys, k, alpha: int, beta: int, gamma: int, out: DevVec) -> None:
        """polys: [(tensor, length)] for w0..w4, z and pi (pi may be None = zero).  out: t on this coset (n values)."""
        for (t_, ln), e in zip(polys[:6], self.evals[:6]):
            self.eval_folded(t_, ln, e)
        if polys[6] is not None:
            self.eval_folded(polys[6][0], polys[6][1], self.evals[6])
        ffi.plonk_quotient_fr_device(
            [e.ptr for e in self.evals[:5]], [c.ptr for c in self.q], self.evals[6].ptr, self.evals[5].ptr, [c.ptr for c in self.s],
            self.coset_quotient.ptr, self.l1.ptr, self.qb.ptr, [c.ptr for c in self.q_prk], mont_rows(k), mont(alpha), mont(beta),
            mont(gamma), self.anemoi[0], self.anemoi[1], mont_rows([self.z_h_inv]), self.n, 1, out.ptr)


def interleave_cosets(t_cosets: DevVec, n: int, out: DevVec, idx_cache: dict) -> None:
    """out[6 i + j] = t_cosets[j * n + i]: the natural order the 6n coset iFFT expects."""
    dev = out.t.device
    if idx_cache.get("n") != n:
        p = torch.arange(6 * n, device=dev, dtype=torch.int64)
        idx_cache["idx"] = ((p % 6) * n + p // 6).to(torch.int32)
        idx_cache["n"] = n
    ffi.fr_gather_device(t_cosets.ptr, idx_cache["idx"].data_ptr(), 6 * n, out.ptr)


# ---------------------------------------------------------------------------------------------- prover
def hide_polynomial(prng, polynomial: DevVec, hiding_degree: int, zeroing_degree: int) -> list[int]:
    """helpers.rs:139-154: add (b_0 + b_1 X + ...) (X^zeroing_degree - 1)."""
    blinds, idx, vals = [], [], []
    for i in range(hiding_degree):
        blind = fr_rand(prng)
        blinds.append(blind)
        idx += [i, zeroing_degree + i]
        vals += [blind, -blind]
    _add_coefs(polynomial, idx, vals)
    return blinds


def first_lagrange_poly(zeta: int, group_order: int) -> tuple[int, int]:
    """helpers.rs:1412-1423: (Z_H(zeta), L_1(zeta))."""
    z_h = (pow(zeta, group_order, FR_MODULUS) - 1) % FR_MODULUS
    return z_h, z_h * pow((zeta - 1) % FR_MODULUS, -1, FR_MODULUS) % FR_MODULUS


def batch_prove_quotient(transcript: Transcript, polys, evals, point: int, max_degree: int, scratch_h: DevVec, out_q: DevVec):
    """The polynomial part of poly_commit/pcs.rs:107-168 (lagrange_pcs = None): h = sum_j alpha^j (p_j - p_j(x)), out_q = h / (X - x).
    `evals` are the values p_j(x) the caller already holds (the reference re-evaluates them, pcs.rs:126).  Returns the device
    scalar holding the remainder, which the caller checks (PCSProveEvalError) when it next synchronises."""
    init_pcs_batch_eval_transcript(transcript, max_degree, point)
    alpha = transcript.get_challenge_field_elem()
    mults, mult, const = [], 1, 0
    for v in evals:
        mults.append(mult)
        const = (const + mult * v) % FR_MODULUS
        mult = mult * alpha % FR_MODULUS
    hlen = max(p.len for p in polys)
    ffi.fr_lincomb_device([p.ptr for p in polys], [p.len for p in polys], mont_rows(mults), scratch_h.ptr, hlen)
    ffi.fr_add_sparse_device(scratch_h.ptr, [0], mont_rows([-const]))
    rem = torch.zeros(4, dtype=torch.int64, device=scratch_h.t.device)
    ffi.poly_horner_fr_device(scratch_h.ptr, hlen, mont(point), out_q.ptr, rem.data_ptr())
    out_q.len = hlen - 1
    return rem


def batch_prove(pcs, transcript: Transcript, polys, evals, point: int, max_degree: int, scratch_h: DevVec, scratch_q: DevVec) -> KZGCommitment:
    """poly_commit/pcs.rs:107-168: the commitment of the quotient above."""
    rem = batch_prove_quotient(transcript, polys, evals, point, max_degree, scratch_h, scratch_q)
    if rem.cpu().numpy().any():
        raise UzkgeError("PCSProveEvalError")
    return _commit_dev(pcs, [scratch_q])[0]


def prover(prng, transcript: Transcript, pcs, cs: TurboCS, prover_params: PlonkProverParams, w,
           timings: dict | None = None, lagrange_pcs=None, quotient_by_cosets: bool = False, lagrange_all: bool | None = None) -> PlonkProof:
    """plonk/prover.rs:76-394 (`prover` = `prover_with_lagrange` with lagrange_pcs = None).  `w`: the witness, (num_vars, 4)
    Montgomery limbs (numpy) or a DevVec already in HBM.  With a `lagrange_pcs` whose size matches the circuit
    (prover.rs:119-124) the wire and z commitments are MSMs of the EVALUATION vectors against the Lagrange SRS with the blind
    terms as six extra bases (prover.rs:131-146, `_lagrange_commit_scheme`); the quotient pieces and opening proofs are committed in coefficient form -- the same group
    elements as the reference's Lagrange branch (helpers.rs:1363-1391, pcs.rs:139-163), without its extra transforms.
    lagrange_all: also commit the witness-selector polynomials, the quotient pieces and the two opening quotients over the Lagrange
    SRS (one forward transform each, `_commit_coefs_lagrange`) -- what the reference's Lagrange branch does, and the only
    possibility with the bundled production parameters, whose monomial SRS holds tau^i only for i < 2051 and i in [n, n + 3)
    (gen_params/mod.rs:147-171).  None = decide from the SRS: on when bases below n are missing.
    quotient_by_cosets: evaluate the quotient round coset by coset (CosetParams) on this GPU -- the single-GPU form of what
    dist.SplitCommitter distributes; the proof is the same."""
    if cs.is_verifier_only():
        raise UzkgeError("FuncParamsError")
    P = prover_params
    vp = P.verifier_params
    n, m, k = P.n, P.m, vp.k
    dev = P.group.t.device
    ws = P.workspace
    scratch = ws["scratch"]
    k1, k1_inv = mont(k[1]), mont(pow(k[1], -1, FR_MODULUS))
    cap = n + 8
    marks = []
    if lagrange_pcs is not None and (lagrange_pcs.max_degree() + 1 != n or hasattr(pcs, "commit_device")):
        lagrange_pcs = None
    if lagrange_pcs is None:
        lagrange_all = False
    elif lagrange_all is None:
        if "srs_truncated" not in ws:
            ws["srs_truncated"] = not bool(np.asarray(pcs.public_parameter_group_1[:n]).any(axis=1).all())
        lagrange_all = ws["srs_truncated"]

    def mark(name):
        if timings is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    wit = w if isinstance(w, DevVec) else DevVec.from_numpy(w, dev)
    if wit.len != cs.num_vars:
        raise ParameterError("witness length != num_vars")
    n_online = len(cs.public_vars_witness_indices)
    online_rows = None
    if n_online:
        d_online = wit.t.view(-1, 4)[torch.tensor(cs.public_vars_witness_indices, device=dev)]
        online_rows = d_online.cpu().numpy().view(np.uint64)

    def init_transcript():
        """transcript_init_plonk (plonk/transcript.rs:8-31).  Host-only work (hundreds of big-integer conversions and appends for
        the application circuits): deferred until round 1's kernels are enqueued, so the GPU is already busy; it must only precede
        the first commitment appended."""
        online_values = [unmont(r) for r in online_rows] if n_online else []
        transcript_init_plonk(transcript, vp, online_values, P.root)

    # 1. the PI polynomial (helpers.rs:111-131)
    pi = DevVec(n, dev)
    if n_online:
        # evals[row] = the public input constrained at that row (rows are distinct: one constant gate per prepare_pi_variable);
        # the values never leave the device: a row scatter of the witness entries read above
        rows_idx = ws.get("pi_rows")
        if rows_idx is None or rows_idx.numel() != n_online:
            rows_idx = ws["pi_rows"] = torch.tensor(list(vp.public_vars_constraint_indices), device=dev)
        pi.t.view(-1, 4)[rows_idx] = d_online
        _ifft(pi.ptr, n, pi, scratch)

    coset = ws.get("coset")
    if coset is None:
        coset = ws["coset"] = [DevVec(m, dev, zero=False) for _ in range(N_WIRES_PER_GATE + 3)]
    w_coset, pi_coset, z_coset, t_buf = coset[:5], coset[5], coset[6], coset[7]
    multi_gpu = hasattr(pcs, "transform_many")

    # 2. witness polynomials: extend, interpolate, hide, commit
    stride = n + 8          # every evaluation vector is followed by the 6 blind slots of the Lagrange commitment
    ext = ws.get("ext") or DevVec(N_WIRES_PER_GATE * stride, dev)
    ws["ext"] = ext
    w_polys, w_blinds = [], []
    for i in range(N_WIRES_PER_GATE):
        ffi.fr_gather_device(wit.ptr, P.wiring.data_ptr() + 4 * i * n, n, ext.at(i * stride))
        f = DevVec(cap, dev, length=n)
        _ifft(ext.at(i * stride), n, f, scratch)
        w_blinds.append(hide_polynomial(prng, f, cs.get_hiding_degree(i), n))
        w_polys.append(f)
    # the quotient round's coset evaluations of the wire polynomials depend on no challenge: they are enqueued behind the MSMs,
    # so the GPU computes them while the host waits for the commitments and hashes the transcript

    def w_sel_coset_bufs():
        if "w_sel_coset" not in ws:
            ws["w_sel_coset"] = [DevVec(m, dev, zero=False) for _ in range(3)]
        return ws["w_sel_coset"]

    def wire_cosets():
        if not multi_gpu and not quotient_by_cosets:
            for p, c in zip(w_polys, w_coset):
                _coset_fft(p, m, k1, c, scratch)
        if w_sel_polys:                                          # the witness selectors' too (`shuffle` feature set)
            for p, c in zip(w_sel_polys, w_sel_coset_bufs()):
                _coset_fft(p, m, k1, c, scratch)
        init_transcript()                                        # host work, with the GPU busy on the above

    # 3. (`shuffle` feature set) witness-selector polynomials (prover.rs:177-191): the remark gates' bit / sign columns
    # (compute_witness_selectors, turbo/mod.rs:171-191; zero on H without remark gates), hidden with 2 blinds each.  They depend on
    # no challenge and the RNG draws keep the reference's order (wires, then selectors), so they are committed in the same batch
    # as the wires when both go to the same SRS.
    shuffle = P.q_ecc_poly is not None
    w_sel_polys, w_sel_blinds, cm_w_sel_vec = [], [], None
    if shuffle:
        has_remark = bool(cs.shuffle_remark_constraints)
        sel_ev = DevVec(3 * stride, dev)
        if has_remark:
            codes = cs.witness_selector_codes()
            if codes is not None:          # bits and signs: upload 3 n small indices, expand against (0, 1, -1) on the device
                table = ws.get("w_sel_table")
                if table is None:
                    table = ws["w_sel_table"] = DevVec.from_numpy(np.stack([_ZERO, _ONE, mont(-1)]), dev)
                d_codes = torch.from_numpy(codes).to(dev)
                for i in range(3):
                    ffi.fr_gather_device(table.ptr, d_codes.data_ptr() + 4 * i * n, n, sel_ev.at(i * stride))
            else:
                sel_host = cs.compute_witness_selectors()
                for i in range(3):
                    sel_ev.t[4 * i * stride: 4 * (i * stride + n)].copy_(torch.from_numpy(sel_host[i].view(np.int64).reshape(-1)))
        for i in range(3):
            if has_remark:
                f = DevVec(cap, dev, length=n)
                _ifft(sel_ev.at(i * stride), n, f, scratch)
            else:
                f = DevVec(cap, dev, length=1)
            w_sel_blinds.append(hide_polynomial(prng, f, 2, n))
            w_sel_polys.append(f)
    if lagrange_pcs is not None:
        lag = _lagrange_commit_scheme(pcs, lagrange_pcs, n, ws)
        for i in range(N_WIRES_PER_GATE):
            _set_blind_slots(ext, i * stride + n, w_blinds[i])
        wire_vecs, wire_scheme = [_View(ext, i * stride, n + 2 * N_BLIND_SLOTS) for i in range(N_WIRES_PER_GATE)], lag
    else:
        wire_vecs, wire_scheme = w_polys, pcs
    sel_vecs, sel_scheme = w_sel_polys, pcs
    if shuffle and lagrange_all:
        for i in range(3):
            _set_blind_slots(sel_ev, i * stride + n, w_sel_blinds[i])
        sel_vecs, sel_scheme = [_View(sel_ev, i * stride, n + 2 * N_BLIND_SLOTS) for i in range(3)], lag
    if shuffle and sel_scheme is wire_scheme:
        cms = _commit_dev(wire_scheme, wire_vecs + sel_vecs, overlap=wire_cosets)
        cm_w_vec, cm_w_sel_vec = cms[:N_WIRES_PER_GATE], cms[N_WIRES_PER_GATE:]
    else:
        cm_w_vec = _commit_dev(wire_scheme, wire_vecs, overlap=wire_cosets)
        if shuffle:
            cm_w_sel_vec = _commit_dev(sel_scheme, sel_vecs)
    for cm in cm_w_vec + (cm_w_sel_vec or []):
        transcript.append_commitment(cm)
    mark("round1_wires")

    # 4. beta, gamma
    beta = transcript.get_challenge_field_elem()
    transcript.append_single_byte(0x01)
    gamma = transcript.get_challenge_field_elem()

    # 5. z: running product on H, interpolate, hide, commit (helpers.rs:160-220)
    z_poly = DevVec(cap, dev, length=n)
    tmp = ws.get("ztmp") or DevVec(4 * n, dev, zero=False)
    ws["ztmp"] = tmp
    z_ev = DevVec(stride, dev, zero=False)
    ffi.plonk_z_evals_fr_device([ext.at(i * stride) for i in range(N_WIRES_PER_GATE)], [P.sigma_evals.at(i * n) for i in range(N_WIRES_PER_GATE)],
                                P.group.ptr, mont_rows(k), mont(beta), mont(gamma), n, z_ev.ptr, tmp.ptr)
    _ifft(z_ev.ptr, n, z_poly, scratch)
    z_blinds = hide_polynomial(prng, z_poly, 3, n)
    def z_coset_eval():
        if not multi_gpu and not quotient_by_cosets:
            _coset_fft(z_poly, m, k1, z_coset, scratch)

    if lagrange_pcs is not None:
        _set_blind_slots(z_ev, n, z_blinds)
        cm_z = _commit_dev(lag, [_View(z_ev, 0, n + 2 * N_BLIND_SLOTS)], overlap=z_coset_eval)[0]
    else:
        cm_z = _commit_dev(pcs, [z_poly], overlap=z_coset_eval)[0]
    transcript.append_commitment(cm_z)
    mark("round2_z")

    # 6. alpha;  7. t = numerator / Z_H on the coset k[1] <w_m>, split, commit (helpers.rs:223-678, 1323-1408)
    alpha = transcript.get_challenge_field_elem()
    # over GPUs: by coset from 3 GPUs on (measured at 2^22: 2 GPUs 142 ms by coset vs 137 ms with only the six transforms
    # distributed; 8 GPUs 70 ms vs 87 ms)
    dist_cosets = hasattr(pcs, "quotient_by_cosets") and getattr(pcs, "world", 1) >= 3
    dist_cosets = dist_cosets and not vp.anemoi_generator      # the distributed setup does not carry the Anemoi constants
    by_cosets = (quotient_by_cosets or dist_cosets) and P.factor == 6 and P.q_ecc_poly is None
    if by_cosets:
        for p in w_polys + [z_poly]:
            p.t[4 * p.len: 4 * (n + 3)].zero_()
        polys = [(p.t, n + 3) for p in w_polys + [z_poly]] + [(pi.t, n) if n_online else None]
        t_cosets = ws.get("t_cosets") or DevVec(m, dev, zero=False)
        ws["t_cosets"] = t_cosets
        if dist_cosets:
            pcs.quotient_by_cosets(P, polys, k, alpha, beta, gamma, t_cosets)        # the cosets are dealt to the GPUs of the box
        else:
            cps = ws.get("coset_params")
            if cps is None:
                cps = ws["coset_params"] = [CosetParams(P.q_polys, P.s_polys, P.qb_poly, P.q_prk_polys, k, n, j, dev,
                                                       (vp.anemoi_generator, vp.anemoi_generator_inv)) for j in range(6)]
            for j, cp in enumerate(cps):
                cp.quotient(polys, k, alpha, beta, gamma, _View(t_cosets, j * n, n))
        interleave_cosets(t_cosets, n, t_buf, ws.setdefault("interleave", {}))
    elif multi_gpu:
        # several GPUs: the six independent 6n transforms of this round go one per rank (dist.SplitCommitter)
        for p in w_polys + [z_poly]:
            p.t[4 * p.len: 4 * (n + 3)].zero_()
        pcs.transform_many([(p.t, c.t) for p, c in zip(w_polys + [z_poly], w_coset + [z_coset])], n + 3, m, False, k1)
    if by_cosets:
        pass
    elif n_online:
        _coset_fft(pi, m, k1, pi_coset, scratch)
    else:
        pi_coset.t.zero_()
    shuffle_args = None
    if shuffle:
        w_sel_coset = w_sel_coset_bufs()          # filled behind the round-1 commitments (wire_cosets)
        shuffle_args = {"w_sel": [c.ptr for c in w_sel_coset], "q_ecc": P.q_ecc_coset_eval.ptr,
                        "pk": [c.ptr for c in P.q_shuffle_public_key_coset_evals], "gen": [c.ptr for c in P.q_shuffle_generator_coset_evals],
                        "edwards_a": mont(vp.edwards_a)}
    if not by_cosets:
        ffi.plonk_quotient_fr_device(
            [c.ptr for c in w_coset], [c.ptr for c in P.q_coset_evals], pi_coset.ptr, z_coset.ptr, [c.ptr for c in P.s_coset_evals],
            P.coset_quotient.ptr, P.l1_coset_evals.ptr, P.qb_coset_eval.ptr, [c.ptr for c in P.q_prk_coset_evals], mont_rows(k),
            mont(alpha), mont(beta), mont(gamma), mont(vp.anemoi_generator), mont(vp.anemoi_generator_inv), P.z_h_inv_coset_evals,
            m, P.factor, t_buf.ptr, shuffle=shuffle_args)
    ffi.ntt_fr_device(t_buf.ptr, t_buf.ptr, scratch.ptr, m, m, True, k1_inv)
    coefs_len = ffi.fr_trimmed_len_device(t_buf.ptr, m)
    mark("round3_quotient")
    piece = n + 2
    t_polys, prev = [], 0
    for i in range(N_WIRES_PER_GATE):
        start = i * piece
        end = coefs_len if i == N_WIRES_PER_GATE - 1 else (i + 1) * piece
        take = max(0, min(coefs_len, end) - start) if start < coefs_len else 0
        tp = DevVec(max(cap, take + 1), dev, length=take)
        if take:
            tp.t[: 4 * take].copy_(t_buf.t[4 * start: 4 * (start + take)])
        rand = fr_rand(prng)
        if i != N_WIRES_PER_GATE - 1:
            # coefs.resize(n + 1, zero); coefs[n] += rand; coefs[0] -= prev_coef   (helpers.rs:1351-1354)
            _add_coefs(tp, [piece, 0], [rand, -prev])
            tp.len = piece + 1
        else:
            # the last piece: [-prev_coef] when it is empty, else coefs[0] -= prev_coef   (helpers.rs:1355-1361); the buffer is
            # zero-initialised, so both cases are one add at index 0
            _add_coefs(tp, [0], [-prev])
        prev = rand
        t_polys.append(tp)
    cm_t_vec = _commit_coefs_lagrange(lag, t_polys, n, scratch) if lagrange_all else _commit_many(pcs, t_polys)
    for cm in cm_t_vec:
        transcript.append_commitment(cm)
    mark("round3_commit_t")

    # 8. zeta;  9a. openings
    zeta = transcript.get_challenge_field_elem()
    zeta_omega = P.root * zeta % FR_MODULUS
    s_open = P.s_polys[: N_WIRES_PER_GATE - 1]
    pts = ([(p, zeta) for p in w_polys] + [(p, zeta) for p in s_open] + [(P.q_prk_polys[2], zeta), (P.q_prk_polys[3], zeta)]
           + [(z_poly, zeta_omega)] + [(p, zeta_omega) for p in w_polys[:3]])
    if shuffle:
        pts += [(P.q_ecc_poly, zeta)] + [(p, zeta) for p in w_sel_polys]
    ev = _evals(pts, dev)
    w_polys_eval_zeta, s_polys_eval_zeta = ev[0:5], ev[5:9]
    prk_3_poly_eval_zeta, prk_4_poly_eval_zeta, z_eval_zeta_omega = ev[9], ev[10], ev[11]
    w_polys_eval_zeta_omega = ev[12:15]
    q_ecc_poly_eval_zeta, w_sel_polys_eval_zeta = (ev[15], ev[16:19]) if shuffle else (None, None)
    for v in w_polys_eval_zeta + s_polys_eval_zeta + (w_sel_polys_eval_zeta or []):
        transcript.append_challenge(v)
    transcript.append_challenge(prk_3_poly_eval_zeta)
    transcript.append_challenge(prk_4_poly_eval_zeta)
    transcript.append_challenge(z_eval_zeta_omega)
    if shuffle:
        transcript.append_challenge(q_ecc_poly_eval_zeta)
    for v in w_polys_eval_zeta_omega:
        transcript.append_challenge(v)
    # 10. u
    _u = transcript.get_challenge_field_elem()

    # 9b. the linearisation polynomial r (helpers.rs:681-999)
    z_h_eval_zeta, first_lagrange_eval_zeta = first_lagrange_poly(zeta, n)
    a = [pow(alpha, i, FR_MODULUS) for i in range(8)]
    we = w_polys_eval_zeta
    sel_mult = [we[0], we[1], we[2], we[3], we[0] * we[1], we[2] * we[3], 1, we[0] * we[1] * we[2] * we[3] * we[4], -we[4]]
    terms = [(sel_mult[i], P.q_polys[i]) for i in range(N_SELECTORS)]
    z_scalar = alpha
    beta_zeta = beta * zeta % FR_MODULUS
    for i in range(N_WIRES_PER_GATE):
        z_scalar = z_scalar * (we[i] + k[i] * beta_zeta + gamma) % FR_MODULUS
    z_scalar += first_lagrange_eval_zeta * a[2]
    terms.append((z_scalar, z_poly))
    s_last = alpha * z_eval_zeta_omega % FR_MODULUS * beta % FR_MODULUS
    for i in range(N_WIRES_PER_GATE - 1):
        s_last = s_last * (we[i] + beta * s_polys_eval_zeta[i] + gamma) % FR_MODULUS
    terms.append((-s_last, P.s_polys[N_WIRES_PER_GATE - 1]))
    terms.append((we[1] * (we[1] - 1) * a[3] + we[2] * (we[2] - 1) * a[4] + we[3] * (we[3] - 1) * a[5], P.qb_poly))
    terms.append((prk_3_poly_eval_zeta * a[6], P.q_prk_polys[0]))
    terms.append((prk_3_poly_eval_zeta * a[7], P.q_prk_polys[1]))
    if shuffle:
        # 6.-9. the remark-gate parts (helpers.rs:747-983): per selector combination c, over the public-key / generator selector
        # polynomials x_c, y_c, dxy_c
        wsl, wo, ed_a = w_sel_polys_eval_zeta, w_polys_eval_zeta_omega, vp.edwards_a
        ah = [pow(alpha, i, FR_MODULUS) for i in range(14)]
        sel = [((1 - wsl[0]) * (1 - wsl[1]) + q_ecc_poly_eval_zeta - 1) % FR_MODULUS, wsl[0] * (1 - wsl[1]) % FR_MODULUS,
               (1 - wsl[0]) * wsl[1] % FR_MODULUS, wsl[0] * wsl[1] % FR_MODULUS]
        pk, gen = P.q_shuffle_public_key_polys, P.q_shuffle_generator_polys
        for c in range(4):
            terms += [(ah[10] * sel[c] * we[0] * we[1] * wo[0], pk[8 + c]), (-ah[10] * sel[c] * wsl[2] * we[0], pk[4 + c]),
                      (-ah[10] * sel[c] * we[1], pk[c]),
                      (-ah[11] * sel[c] * we[0] * we[1] * wo[1], pk[8 + c]), (ah[11] * sel[c] * we[0] * ed_a, pk[c]),
                      (-ah[11] * sel[c] * wsl[2] * we[1], pk[4 + c]),
                      (ah[12] * sel[c] * we[2] * we[3] * wo[2], gen[8 + c]), (-ah[12] * sel[c] * wsl[2] * we[2], gen[4 + c]),
                      (-ah[12] * sel[c] * we[3], gen[c]),
                      (-ah[13] * sel[c] * we[2] * we[3] * we[4], gen[8 + c]), (ah[13] * sel[c] * we[2] * ed_a, gen[c]),
                      (-ah[13] * sel[c] * wsl[2] * we[3], gen[4 + c])]
    zfactor = pow(zeta, piece, FR_MODULUS)
    exponent = z_h_eval_zeta
    for tp in t_polys:
        terms.append((-exponent, tp))
        exponent = exponent * zfactor % FR_MODULUS
    merged: dict = {}            # the same polynomial may appear in several terms (the shared zero selector, the 3 uses of x_c ...)
    for sc_, p_ in terms:
        key = p_.ptr
        merged[key] = ((merged[key][0] + sc_) % FR_MODULUS, p_) if key in merged else (sc_ % FR_MODULUS, p_)
    terms = list(merged.values())
    rlen = max(p.len for _, p in terms)
    r_poly = DevVec(max(cap, rlen), dev, zero=False, length=rlen)
    first = True
    for i in range(0, len(terms), ffi.LINCOMB_MAX - 1):
        chunk = terms[i:i + ffi.LINCOMB_MAX - 1]
        ptrs, lens, scs = [p.ptr for _, p in chunk], [p.len for _, p in chunk], [s_ for s_, _ in chunk]
        if not first:            # accumulate onto the partial sum
            ptrs, lens, scs = ptrs + [r_poly.ptr], lens + [rlen], scs + [1]
        ffi.fr_lincomb_device(ptrs, lens, mont_rows(scs), r_poly.ptr, rlen)
        first = False
    r_eval_zeta = _evals([(r_poly, zeta)], dev)[0]
    mark("round4_evals_r")

    polys_to_open = w_polys + s_open + [P.q_prk_polys[2], P.q_prk_polys[3]]
    evals_to_open = w_polys_eval_zeta + s_polys_eval_zeta + [prk_3_poly_eval_zeta, prk_4_poly_eval_zeta]
    if shuffle:
        polys_to_open += [P.q_ecc_poly] + w_sel_polys
        evals_to_open += [q_ecc_poly_eval_zeta] + w_sel_polys_eval_zeta
    polys_to_open.append(r_poly)
    evals_to_open.append(r_eval_zeta)
    # the first opening proof does not enter the transcript before the second is built (prover.rs:359-381), so the two
    # quotients are committed in one batch
    hmax = max(p.len for p in polys_to_open + [z_poly])
    sh, q1, q2 = DevVec(hmax, dev, zero=False), DevVec(hmax, dev, zero=False), DevVec(hmax, dev, zero=False)
    rem1 = batch_prove_quotient(transcript, polys_to_open, evals_to_open, zeta, n + 2, sh, q1)
    rem2 = batch_prove_quotient(transcript, [z_poly] + w_polys[:3], [z_eval_zeta_omega] + w_polys_eval_zeta_omega, zeta_omega, n + 2, sh, q2)
    if lagrange_all:
        opening_witness_zeta, opening_witness_zeta_omega = _commit_coefs_lagrange(lag, [q1, q2], n, scratch)
    else:
        opening_witness_zeta, opening_witness_zeta_omega = _commit_dev(pcs, [q1, q2])
    if torch.cat([rem1, rem2]).cpu().numpy().any():
        raise UzkgeError("PCSProveEvalError")
    mark("round5_openings")
    if timings is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            timings[name] = timings.get(name, 0.0) + e0.elapsed_time(e1)
    return PlonkProof(
        cm_w_vec=cm_w_vec, cm_t_vec=cm_t_vec, cm_z=cm_z, prk_3_poly_eval_zeta=prk_3_poly_eval_zeta,
        prk_4_poly_eval_zeta=prk_4_poly_eval_zeta, w_polys_eval_zeta=w_polys_eval_zeta, w_polys_eval_zeta_omega=w_polys_eval_zeta_omega,
        z_eval_zeta_omega=z_eval_zeta_omega, s_polys_eval_zeta=s_polys_eval_zeta, opening_witness_zeta=opening_witness_zeta,
        opening_witness_zeta_omega=opening_witness_zeta_omega, cm_w_sel_vec=cm_w_sel_vec, q_ecc_poly_eval_zeta=q_ecc_poly_eval_zeta,
        w_sel_polys_eval_zeta=w_sel_polys_eval_zeta)
