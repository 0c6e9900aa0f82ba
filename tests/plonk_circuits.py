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
