"""Host mirror of the reference's Anemoi-Jive primitives over BN254's scalar field: what zmatchmaking's circuit builder evaluates to
fill the Anemoi gates of TurboCS.  Python integers; nothing here touches the GPU.

  AnemoiJive (trait), anemoi_permutation, eval_variable_length_hash(_with_trace), eval_stream_cipher(_with_trace)
                                       /root/reference/uzkge/src/anemoi/mod.rs:18-379
  MDSMatrix<F, 2>::permute_in_place    /root/reference/uzkge/src/anemoi/mds/mod.rs:28-61
  AnemoiVLHTrace, AnemoiStreamCipherTrace
                                       /root/reference/uzkge/src/anemoi/traces.rs
  AnemoiJive254 (the BN254 instance)   /root/reference/uzkge/src/anemoi/bn254/mod.rs:6-378

The instance's constants are DERIVED here from the Anemoi specification (g = 5 the field's multiplicative generator, alpha = 5,
delta = 1 / g, round constants from the first two hundred decimals of pi:
C[r][i] = g (pi0^r)^2 + (pi0^r + pi1^i)^alpha,  D[r][i] = g (pi1^i)^2 + (pi0^r + pi1^i)^alpha + delta), not copied; the tests
compare every derived constant and the reference's known answers (anemoi/tests.rs) through tests/golden/anemoi_bn254.json.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .rng import FR_MODULUS as R

N_ANEMOI_ROUNDS = 14
_PI_0 = 1415926535897932384626433832795028841971693993751058209749445923078164062862089986280348253421170679
_PI_1 = 8214808651328230664709384460955058223172535940812848111745028410270193852110555964462294895493038196


_G, _ALPHA = 5, 5
_DELTA = pow(_G, -1, R)
_ROUND_KEYS_X = [[(_G * pow(_PI_0, 2 * r, R) + pow(pow(_PI_0, r, R) + pow(_PI_1, i, R), _ALPHA, R)) % R for i in range(2)]
                 for r in range(N_ANEMOI_ROUNDS)]
_ROUND_KEYS_Y = [[(_G * pow(_PI_1, 2 * i, R) + pow(pow(_PI_0, r, R) + pow(_PI_1, i, R), _ALPHA, R) + _DELTA) % R for i in range(2)]
                 for r in range(N_ANEMOI_ROUNDS)]


@dataclass
class AnemoiVLHTrace:
    """traces.rs:6-18 (N = 2): states are ([x0, x1], [y0, y1])."""
    input: list = field(default_factory=list)
    before_permutation: list = field(default_factory=list)
    intermediate_values_before_constant_additions: list = field(default_factory=list)   # per permutation ([rounds][2], [rounds][2])
    after_permutation: list = field(default_factory=list)
    output: int = 0


@dataclass
class AnemoiStreamCipherTrace:
    """traces.rs: the same with a list of outputs."""
    input: list = field(default_factory=list)
    before_permutation: list = field(default_factory=list)
    intermediate_values_before_constant_additions: list = field(default_factory=list)
    after_permutation: list = field(default_factory=list)
    output: list = field(default_factory=list)


class AnemoiJive254:
    """The 2-column, 14-round Anemoi-Jive instance over BN254 Fr (anemoi/bn254/mod.rs)."""
    N = 2
    NUM_ROUNDS = N_ANEMOI_ROUNDS
    ALPHA = 5
    GENERATOR = 5
    GENERATOR_INV = pow(5, -1, R)
    GENERATOR_SQUARE_PLUS_ONE = 26
    ALPHA_INV = pow(5, -1, R - 1)
    MDS_MATRIX = [[1, 5], [5, 26]]
    ROUND_KEYS_X = _ROUND_KEYS_X
    ROUND_KEYS_Y = _ROUND_KEYS_Y

    # ---- the linear layer: MDS on x, MDS on the word-rotated y, then y += x, x += y (mds/mod.rs:43-60, mod.rs:135-139)
    @classmethod
    def _linear(cls, x, y):
        m = cls.MDS_MATRIX
        nx = [(m[i][0] * x[0] + m[i][1] * x[1]) % R for i in range(2)]
        ny = [(m[i][0] * y[1] + m[i][1] * y[0]) % R for i in range(2)]
        ny = [(ny[i] + nx[i]) % R for i in range(2)]
        nx = [(nx[i] + ny[i]) % R for i in range(2)]
        return nx, ny

    @classmethod
    def preprocessed_round_keys(cls):
        """PREPROCESSED_ROUND_KEYS_{X,Y}: the round keys pushed through the linear layer (what the gate equations use)."""
        px, py = [], []
        for r in range(cls.NUM_ROUNDS):
            a, b = cls._linear(cls.ROUND_KEYS_X[r], cls.ROUND_KEYS_Y[r])
            px.append(a)
            py.append(b)
        return px, py

    @classmethod
    def _permutation(cls, x, y, trace=None):
        """mod.rs:347-377; with `trace` the per-round states of mod.rs:128-160."""
        g, g_inv = cls.GENERATOR, cls.GENERATOR_INV
        if trace is not None:
            trace.before_permutation.append((list(x), list(y)))
        ix, iy = [], []
        for r in range(cls.NUM_ROUNDS):
            x = [(x[i] + cls.ROUND_KEYS_X[r][i]) % R for i in range(2)]
            y = [(y[i] + cls.ROUND_KEYS_Y[r][i]) % R for i in range(2)]
            x, y = cls._linear(x, y)
            for i in range(2):
                x[i] = (x[i] - g * y[i] * y[i]) % R
                y[i] = (y[i] - pow(x[i], cls.ALPHA_INV, R)) % R
                x[i] = (x[i] + g * y[i] * y[i] + g_inv) % R
            ix.append(list(x))
            iy.append(list(y))
        x, y = cls._linear(x, y)
        if trace is not None:
            trace.intermediate_values_before_constant_additions.append((ix, iy))
            trace.after_permutation.append((list(x), list(y)))
        return x, y

    @classmethod
    def anemoi_permutation(cls, x, y):
        return cls._permutation(list(x), list(y))

    @staticmethod
    def _pad(values):
        """mod.rs:56-69: the sponge's padding to multiples of 2 N - 1 = 3; returns (padded input, sigma)."""
        inp = [v % R for v in values]
        if len(inp) % 3 == 0 and inp:
            return inp, 1
        inp.append(1)
        if len(inp) % 3:
            inp += [0] * (3 - len(inp) % 3)
        return inp, 0

    @classmethod
    def _absorb(cls, values, trace=None):
        inp, sigma = cls._pad(values)
        x, y = [0, 0], [0, 0]
        for c in range(0, len(inp), 3):
            x = [(x[0] + inp[c]) % R, (x[1] + inp[c + 1]) % R]
            y = [(y[0] + inp[c + 2]) % R, y[1]]
            x, y = cls._permutation(x, y, trace)
        y[1] = (y[1] + sigma) % R
        return x, y

    @classmethod
    def eval_variable_length_hash(cls, values) -> int:
        return cls._absorb(values)[0][0]

    @classmethod
    def eval_variable_length_hash_with_trace(cls, values) -> AnemoiVLHTrace:
        trace = AnemoiVLHTrace(input=[v % R for v in values])
        x, _ = cls._absorb(values, trace)
        trace.output = x[0]
        return trace

    @classmethod
    def _squeeze(cls, x, y, output_len: int, trace=None) -> list:
        """mod.rs:203-230."""
        if output_len <= 2:
            return list(x[:output_len])
        if output_len <= 3:
            return list(x) + list(y[:output_len - 2])
        out = list(x) + [y[0]]
        squeezing_times, remaining = output_len // 3 - 1, output_len % 3
        for _ in range(squeezing_times):
            x, y = cls._permutation(x, y, trace)
            out += list(x) + [y[0]]
        if remaining:
            x, y = cls._permutation(x, y, trace)
            out += (list(x) + list(y))[:remaining]
        return out

    @classmethod
    def eval_stream_cipher(cls, values, output_len: int) -> list:
        x, y = cls._absorb(values)
        return cls._squeeze(x, y, output_len)

    @classmethod
    def eval_stream_cipher_with_trace(cls, values, output_len: int) -> AnemoiStreamCipherTrace:
        trace = AnemoiStreamCipherTrace(input=[v % R for v in values])
        x, y = cls._absorb(values, trace)
        trace.output = cls._squeeze(x, y, output_len, trace)
        return trace


class AnemoiGates:
    """The Anemoi methods of TurboCS (mixed into plonk.TurboCS): /root/reference/uzkge/src/plonk/constraint_system/anemoi/mod.rs:8-534,
    turbo/mod.rs:285-308, 899-924.  One permutation = 14 gates with all-zero selectors whose rows carry the preprocessed round keys in
    the q_prk polynomials (quotient terms 8-11), followed by the linear-layer output gates."""

    def _init_anemoi(self) -> None:
        self.anemoi_generator = 0
        self.anemoi_generator_inv = 0
        self.anemoi_preprocessed_round_keys_x = None
        self.anemoi_preprocessed_round_keys_y = None
        self.anemoi_mds = None
        self.anemoi_constraints_indices: list = []

    def load_anemoi_parameters(self, params=AnemoiJive254) -> None:
        """turbo/mod.rs:917-924."""
        self.anemoi_preprocessed_round_keys_x, self.anemoi_preprocessed_round_keys_y = params.preprocessed_round_keys()
        self.anemoi_generator, self.anemoi_generator_inv = params.GENERATOR, params.GENERATOR_INV
        self.anemoi_mds = params.MDS_MATRIX

    def attach_anemoi_jive_constraints_to_gate(self) -> None:
        if not self.anemoi_generator:
            raise ValueError("load_anemoi_parameters first")
        self.anemoi_constraints_indices.append(self.size - 1)

    def anemoi_permutation_round(self, input_var, output_var, intermediate_val, checksum=None, salt=None):
        """anemoi/mod.rs:10-195.  input_var = ([x0, x1], [y0, y1]); output_var likewise with None for outputs nobody reads;
        intermediate_val = (x[rounds][2], y[rounds][2])."""
        m = self.anemoi_mds
        iv = ([], [])
        for r in range(N_ANEMOI_ROUNDS):
            x0, x1 = self.new_variable(intermediate_val[0][r][0]), self.new_variable(intermediate_val[0][r][1])
            y0, y1 = self.new_variable(intermediate_val[1][r][0]), self.new_variable(intermediate_val[1][r][1])
            iv[0].append([x0, x1])
            iv[1].append([y0, y1])
        first = [input_var[0][0], input_var[0][1], input_var[1][0], input_var[1][1], iv[1][0][1]]
        if salt is not None:
            self._push_gate((0, 0, 0, 1), (0, 0), -salt, 0, 0, first)
        else:
            self._push_gate((0, 0, 0, 0), (0, 0), 0, 0, 0, first)
        self.attach_anemoi_jive_constraints_to_gate()
        for r in range(1, N_ANEMOI_ROUNDS):
            self._push_gate((0, 0, 0, 0), (0, 0), 0, 0, 0, [iv[0][r - 1][0], iv[0][r - 1][1], iv[1][r - 1][0], iv[1][r - 1][1], iv[1][r][1]])
        last = [iv[0][13][0], iv[0][13][1], iv[1][13][0], iv[1][13][1]]
        if output_var[0][0] is not None:
            self._push_gate((2 * m[0][0], 2 * m[0][1], m[0][1], m[0][0]), (0, 0), 0, 0, 1, last + [output_var[0][0]])
        if output_var[0][1] is not None:
            self._push_gate((2 * m[1][0], 2 * m[1][1], m[1][1], m[1][0]), (0, 0), 0, 0, 1, last + [output_var[0][1]])
        if output_var[1][0] is not None:
            self._push_gate((m[0][0], m[0][1], m[0][1], m[0][0]), (0, 0), 0, 0, 1, last + [output_var[1][0]])
        if output_var[1][1] is not None:
            self._push_gate((m[1][0], m[1][1], m[1][1], m[1][0]), (0, 0), 0, 0, 1, last + [output_var[1][1]])
        if checksum is not None:
            var = self.new_variable(checksum)
            s0, s1 = m[0][0] + m[1][0], m[0][1] + m[1][1]
            self._push_gate((3 * s0, 3 * s1, 2 * s1, 2 * s0), (0, 0), 0, 0, 1, last + [var])
            return var
        return None

    @staticmethod
    def _pad_vars(vars_, one_var, zero_var):
        """The sponge padding on variable indices; returns (padded, sigma variable)."""
        v = list(vars_)
        if len(v) % 3 == 0 and v:
            return v, one_var
        v.append(one_var)
        if len(v) % 3:
            v += [zero_var] * (3 - len(v) % 3)
        return v, zero_var

    def _state_vars(self, state):
        return ([self.new_variable(state[0][0]), self.new_variable(state[0][1])], [self.new_variable(state[1][0]), self.new_variable(state[1][1])])

    def anemoi_variable_length_hash(self, trace: AnemoiVLHTrace, input_var, output_var: int) -> None:
        """anemoi/mod.rs:198-313."""
        if len(input_var) != len(trace.input):
            raise ValueError("one variable per input element expected")
        zero_var = self.zero_var()
        padded, _ = self._pad_vars(input_var, self.one_var(), zero_var)
        chunks = [padded[i:i + 3] for i in range(0, len(padded), 3)]
        if len(chunks) != len(trace.before_permutation):
            raise ValueError("trace does not match the input length")
        ivs = trace.intermediate_values_before_constant_additions
        x_var, y_var = [chunks[0][0], chunks[0][1]], [chunks[0][2], zero_var]
        only_out = ([output_var, None], [None, None])
        if len(chunks) == 1:
            self.anemoi_permutation_round((x_var, y_var), only_out, ivs[0])
            return
        new_x, new_y = self._state_vars(trace.after_permutation[0])
        self.anemoi_permutation_round((x_var, y_var), (list(new_x), list(new_y)), ivs[0])
        for rr in range(1, len(chunks)):
            x_var = [self.add(new_x[0], chunks[rr][0]), self.add(new_x[1], chunks[rr][1])]
            y_var = [self.add(new_y[0], chunks[rr][2]), new_y[1]]
            if rr == len(chunks) - 1:
                self.anemoi_permutation_round((x_var, y_var), only_out, ivs[rr])
            else:
                new_x, new_y = self._state_vars(trace.after_permutation[rr])
                self.anemoi_permutation_round((x_var, y_var), (list(new_x), list(new_y)), ivs[rr])

    def anemoi_stream_cipher(self, trace: AnemoiStreamCipherTrace, input_var, output_var) -> None:
        """anemoi/mod.rs:316-533."""
        if len(input_var) != len(trace.input) or len(output_var) != len(trace.output):
            raise ValueError("one variable per input / output element expected")
        zero_var = self.zero_var()
        outs = list(output_var)
        if len(outs) % 3:
            outs += [None] * (3 - len(outs) % 3)
        out_chunks = [outs[i:i + 3] for i in range(0, len(outs), 3)]
        padded, sigma_var = self._pad_vars(input_var, self.one_var(), zero_var)
        if len(padded) + len(outs) - 3 != 3 * len(trace.before_permutation):
            raise ValueError("trace does not match the input / output lengths")
        in_chunks = [padded[i:i + 3] for i in range(0, len(padded), 3)]
        n_in, n_out = len(in_chunks), len(out_chunks)
        ivs = trace.intermediate_values_before_constant_additions
        x_var, y_var = [in_chunks[0][0], in_chunks[0][1]], [in_chunks[0][2], zero_var]
        as_out = lambda c: ([c[0], c[1]], [c[2], None])
        if n_in == 1:
            self.anemoi_permutation_round((x_var, y_var), as_out(out_chunks[0]), ivs[0])
            if n_out == 1:
                return
            new_x, new_y = self._state_vars(trace.after_permutation[0])
            new_y[1] = self.add(new_y[1], sigma_var)
            first_squeeze = 1
        else:
            new_x, new_y = self._state_vars(trace.after_permutation[0])
            self.anemoi_permutation_round((x_var, y_var), (list(new_x), list(new_y)), ivs[0])
            for rr in range(1, n_in - 1):
                x_var = [self.add(new_x[0], in_chunks[rr][0]), self.add(new_x[1], in_chunks[rr][1])]
                y_var = [self.add(new_y[0], in_chunks[rr][2]), new_y[1]]
                new_x, new_y = self._state_vars(trace.after_permutation[rr])
                self.anemoi_permutation_round((x_var, y_var), (list(new_x), list(new_y)), ivs[rr])
            x_var = [self.add(new_x[0], in_chunks[n_in - 1][0]), self.add(new_x[1], in_chunks[n_in - 1][1])]
            y_var = [self.add(new_y[0], in_chunks[n_in - 1][2]), new_y[1]]
            if n_out > 1:
                new_x, new_y = self._state_vars(trace.after_permutation[n_in - 1])
                new_y[1] = self.add(new_y[1], sigma_var)
            self.anemoi_permutation_round((x_var, y_var), as_out(out_chunks[0]), ivs[n_in - 1])
            first_squeeze = n_in
        # the squeezing rounds
        for rr in range(1, n_out):
            x_var, y_var = list(new_x), list(new_y)
            t = rr - 1 + first_squeeze
            if rr != n_out - 1:
                new_x, new_y = self._state_vars(trace.after_permutation[t])
            self.anemoi_permutation_round((x_var, y_var), as_out(out_chunks[rr]), ivs[t])

    def compute_anemoi_jive_selectors_int(self) -> list:
        """turbo/mod.rs:285-304."""
        polys = [[0] * self.size for _ in range(4)]
        for first in self.anemoi_constraints_indices:
            for j in range(N_ANEMOI_ROUNDS):
                polys[0][first + j] = self.anemoi_preprocessed_round_keys_x[j][0]
                polys[1][first + j] = self.anemoi_preprocessed_round_keys_x[j][1]
                polys[2][first + j] = self.anemoi_preprocessed_round_keys_y[j][0]
                polys[3][first + j] = self.anemoi_preprocessed_round_keys_y[j][1]
        return polys

    def check_anemoi_rows(self, witness, wire_of) -> str | None:
        """The Anemoi part of verify_witness (turbo/mod.rs:1060-1146); wire_of(j, row) -> variable index.  Returns an error text."""
        g, g_inv = self.anemoi_generator, self.anemoi_generator_inv
        g2 = (g * g + 1) % R
        for first in self.anemoi_constraints_indices:
            for r in range(N_ANEMOI_ROUNDS):
                a, b, c, d, o = (witness[wire_of(j, first + r)] for j in range(5))
                an, bn, cn, dn = (witness[wire_of(j, first + r + 1)] for j in range(4))
                if o != dn:
                    return f"cs index {first} round {r}: the output wire does not equal the fourth wire of the next constraint"
                pa, pb = self.anemoi_preprocessed_round_keys_x[r]
                pc, pd = self.anemoi_preprocessed_round_keys_y[r]
                da, cb = a + d, b + c
                d2a, c2b = da + a, cb + b
                t1, t2 = (da + g * cb + pc) % R, (g * da + g2 * cb + pd) % R
                if (pow(t1 - cn, 5, R) + g * t1 * t1 - (d2a + g * c2b + pa)) % R:
                    return f"cs index {first} round {r}: the first equation of anemoi does not hold"
                if (pow(t2 - dn, 5, R) + g * t2 * t2 - (g * d2a + g2 * c2b + pb)) % R:
                    return f"cs index {first} round {r}: the second equation of anemoi does not hold"
                if (pow(t1 - cn, 5, R) + g * cn * cn + g_inv - an) % R:
                    return f"cs index {first} round {r}: the third equation of anemoi does not hold"
                if (pow(t2 - dn, 5, R) + g * dn * dn + g_inv - bn) % R:
                    return f"cs index {first} round {r}: the fourth equation of anemoi does not hold"
        return None
