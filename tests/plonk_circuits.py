"""Shared test circuits: the same sequence of TurboCS calls applied to the product's TurboCS and to the oracle's."""
import random

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def build_circuit(cs, n_gates: int, seed: int, n_public: int = 1, n_boolean: int = 1):
    """Random addition / multiplication gates over a growing pool of variables, `n_public` public inputs and `n_boolean`
    boolean gates with the boolean selector attached; padded to a power of two.  Returns cs."""
    rnd = random.Random(seed)
    pool = [cs.new_variable(rnd.randrange(FR)) for _ in range(4)]
    for _ in range(n_public):
        pub = cs.new_variable(rnd.randrange(FR))
        cs.prepare_pi_variable(pub)
        pool.append(pub)
    for g in range(n_gates):
        a, b = rnd.choice(pool), rnd.choice(pool)
        pool.append(cs.add(a, b) if rnd.random() < 0.5 else cs.mul(a, b))
    for _ in range(n_boolean):
        # a boolean gate with the qb selector: wires 1..3 must be 0 / 1 (w1 = bit, w2 = w3 = variable 0)
        bit = cs.new_variable(rnd.randrange(2))
        cs.insert_mul_gate(cs.one_var() if hasattr(cs, "one_var") else 1, bit, bit)
        cs.attach_boolean_constraint_to_gate()
    cs.pad()
    return cs


def shuffle_inputs(n_cards: int, seed: int):
    """Deterministic inputs of a zshuffle circuit (shuffle/src/build_cs.rs:26-56): joint public key, input deck, remark bits and the
    permutation, as plain integers / tuples, drawn with Python's generator and the oracle's curve arithmetic."""
    from oracle import babyjubjub as bj

    rnd = random.Random(seed)
    sk = rnd.randrange(1, bj.ORDER)
    pk = bj.mul(sk, bj.GEN)
    cards, messages = [], []
    for _ in range(n_cards):
        msg = bj.mul(rnd.randrange(1, bj.ORDER), bj.GEN)
        cards.append(bj.encrypt(rnd.randrange(1, bj.ORDER), msg, pk))
        messages.append(msg)
    bits = [[[rnd.random() < 0.5 for _ in range(3)] for _ in range(bj.NUM_ITERATIONS)] for _ in range(n_cards)]
    perm = list(range(n_cards))
    rnd.shuffle(perm)
    return {"sk": sk, "pk": pk, "cards": cards, "messages": messages, "bits": bits, "perm": perm}


def build_shuffle_circuit(cs, inp):
    """build_cs (shuffle/src/build_cs.rs:26-56) on the product's TurboCS or on the oracle's, from the same inputs.  Returns
    (cs, output card variables)."""
    n = len(inp["cards"])
    matrix = [[1 if inp["perm"][i] == j else 0 for j in range(n)] for i in range(n)]
    cs.load_shuffle_remark_parameters(inp["pk"])
    remarked = []
    if hasattr(cs, "_init_shuffle"):         # the product's host mirror
        from uzkge_b200 import shuffle as sh

        pks = sh.BabyJubjubShuffle.crate_public_keys(inp["pk"])
        for card, bits in zip(inp["cards"], inp["bits"]):
            c = sh.Ciphertext(card[0], card[1])
            trace = sh.BabyJubjubShuffle.eval_remark_with_trace(c, bits, inp["pk"], pks)
            var = cs.new_card_variable(c)
            cs.prepare_pi_card_variable(var)
            remarked.append(cs.eval_card_remark(trace, var))
        out = cs.shuffle_card(remarked, sh.Permutation(matrix))
    else:
        from oracle import babyjubjub as bj

        for card, bits in zip(inp["cards"], inp["bits"]):
            fb, iv = bj.remark_trace(card, bits, inp["pk"])
            var = cs.new_card_variable(card)
            cs.prepare_pi_card_variable(var)
            remarked.append(cs.eval_card_remark(fb, iv, var))
        out = cs.shuffle_card(remarked, matrix)
    for cv in out:
        cs.prepare_pi_card_variable(cv)
    cs.pad()
    return cs, out


def transplant(cs):
    """A frozen circuit of the product's TurboCS (uzkge_b200/plonk.py) as the restatement's TurboCS (oracle/plonk_prover.py): same
    selectors, wiring, witness and gate annotations.  Used where the circuit builder is pinned by reference data (the bundled
    verifier keys) rather than by a second implementation."""
    from oracle import plonk_prover as pp

    o = pp.TurboCS()
    n = cs.size
    tab = cs._sel_table
    o.selectors = [[tab[int(c)] for c in cs._sel_codes[j]] for j in range(9)]
    o.wiring = [[int(v) for v in cs.wiring[j]] for j in range(5)]
    o.size, o.num_vars, o.witness = n, cs.num_vars, list(cs.witness)
    o.public_vars_constraint_indices = list(cs.public_vars_constraint_indices)
    o.public_vars_witness_indices = list(cs.public_vars_witness_indices)
    o.boolean_constraint_indices = list(cs.boolean_constraint_indices)
    o.anemoi_constraints_indices = list(cs.anemoi_constraints_indices)
    o.anemoi_generator, o.anemoi_generator_inv = cs.anemoi_generator, cs.anemoi_generator_inv
    if cs.anemoi_preprocessed_round_keys_x is not None:
        o.anemoi_prk = (cs.anemoi_preprocessed_round_keys_x, cs.anemoi_preprocessed_round_keys_y)
    return o
