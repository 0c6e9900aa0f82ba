"""Host mirror of zmatchmaking's circuit (BASELINE's second application, 2^13 gates): a commit-and-reveal Fisher-Yates
shuffle of N = 50 inputs driven by an Anemoi stream cipher.  Python integers; the circuit and witness built here go through
plonk.indexer / plonk.prover on the GPU.

  Matchmaking::generate_constraints    /root/reference/matchmaking/src/matchmaking.rs:42-229
  build_cs, transcript labels, N       /root/reference/matchmaking/src/build_cs.rs:19-67
"""
from __future__ import annotations

from .anemoi import AnemoiJive254
from .rng import FR_MODULUS as R

N = 50
PLONK_PROOF_TRANSCRIPT = b"Plonk Matchmaking Proof"


def _sum_chunks(cs, vars_, boolean: bool) -> int:
    zero_var, s = cs.zero_var(), cs.zero_var()
    for i in range(0, len(vars_), 3):
        c = vars_[i:i + 3]
        if len(c) == 3:
            s = cs.linear_combine([s, c[0], c[1], c[2]], 1, 1, 1, 1)
        elif len(c) == 2:
            s = cs.linear_combine([s, c[0], c[1], zero_var], 1, 1, 1, 0)
        else:
            s = cs.linear_combine([s, c[0], zero_var, zero_var], 1, 1, 0, 0)
        if boolean:
            cs.attach_boolean_constraint_to_gate()
    return s


def generate_constraints(cs, input_vars, committed_input_var: int, committed_output_var: int, committed_trace, random_number_var: int,
                         params=AnemoiJive254) -> list:
    """matchmaking.rs:42-229.  Returns the output variables: the inputs after the shuffle whose i-th swap index is
    stream_cipher(seed, random number)[i - 1] mod (i + 1)."""
    n = len(input_vars)
    if n <= 2:
        raise ValueError("N > 2 expected")
    index_vars = [cs.zero_var(), cs.one_var()]
    for i in range(2, n):
        v = cs.new_variable(i)
        cs.insert_constant_gate(v, i)
        index_vars.append(v)
    cs.anemoi_variable_length_hash(committed_trace, [committed_input_var], committed_output_var)
    sc_trace = params.eval_stream_cipher_with_trace([committed_trace.input[0], cs.witness[random_number_var]], n - 1)
    sc_out_vars = [cs.new_variable(x) for x in sc_trace.output]
    cs.anemoi_stream_cipher(sc_trace, [committed_input_var, random_number_var], sc_out_vars)

    output_vars = list(input_vars)
    zero_var = cs.zero_var()
    for i in range(1, n):
        value = sc_trace.output[i - 1]
        quotient, remainder = divmod(value, i + 1)
        n_var = cs.new_variable(value)
        quotient_var, remainder_var = cs.new_variable(quotient), cs.new_variable(remainder)
        cs._push_gate((i + 1, 1, 0, 0), (0, 0), 0, 0, 1, [quotient_var, remainder_var, zero_var, zero_var, n_var])
        bits_vars = [cs.new_variable(1 if j == remainder else 0) for j in range(i + 1)]
        cs.insert_constant_gate(_sum_chunks(cs, bits_vars, True), 1)             # exactly one bit is set
        for j in range(len(bits_vars)):                                          # index_j * bit_j - remainder * bit_j = 0
            cs._push_gate((0, 0, 0, 0), (1, -1), 0, 0, 0, [index_vars[j], bits_vars[j], remainder_var, bits_vars[j], zero_var])
        output_i_var = output_vars[i]
        products = [cs.mul(b, o) for b, o in zip(bits_vars, output_vars)]
        output_vars[i] = _sum_chunks(cs, products, False)
        for j in range(i):
            output_vars[j] = cs.select(output_vars[j], output_i_var, bits_vars[j])
    return output_vars


def build_cs(cs, inputs, committed_seed: int, random_number: int, params=AnemoiJive254):
    """matchmaking/src/build_cs.rs:26-67 on an empty TurboCS.  Returns (cs, output variables).  Public inputs, in order: the
    inputs, the outputs, the random number, the seed's commitment."""
    cs.load_anemoi_parameters(params)
    input_vars = [cs.new_variable(v) for v in inputs]
    random_number_var = cs.new_variable(random_number)
    committed_trace = params.eval_variable_length_hash_with_trace([committed_seed])
    committed_input_var = cs.new_variable(committed_seed)
    committed_output_var = cs.new_variable(committed_trace.output)
    output_vars = generate_constraints(cs, input_vars, committed_input_var, committed_output_var, committed_trace, random_number_var, params)
    for v in input_vars:
        cs.prepare_pi_variable(v)
    for v in output_vars:
        cs.prepare_pi_variable(v)
    cs.prepare_pi_variable(random_number_var)
    cs.prepare_pi_variable(committed_output_var)
    cs.pad()
    return cs, output_vars
