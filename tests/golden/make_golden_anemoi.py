"""Extract the reference's Anemoi-Jive parameters and known answers into a data fixture (run HERE, where /root/reference exists).

    python tests/golden/make_golden_anemoi.py

Sources: /root/reference/uzkge/src/anemoi/bn254/mod.rs:13-377 (generator, its inverse, round keys, preprocessed round keys, MDS
matrix, alpha inverse) and /root/reference/uzkge/src/anemoi/tests.rs:10-21, 210-237 (the sponge and stream-cipher known answers
for the input [1, 2, 3, 4]).  Output tests/golden/anemoi_bn254.json: data only (the round-key tables as SHA-256 digests plus their first row).
"""
import hashlib
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/uzkge/src/anemoi"


def main():
    src = open(f"{REF}/bn254/mod.rs").read()

    def table(name):
        body = src[src.index("const " + name + ":"):]
        body = body[: body.index("];\n") + 2]
        vals = re.findall(r'MontFp!\(\s*"(\d+)"\s*\)', body)
        return [vals[i:i + 2] for i in range(0, len(vals), 2)]

    out = {}
    for name in ("ROUND_KEYS_X", "ROUND_KEYS_Y", "PREPROCESSED_ROUND_KEYS_X", "PREPROCESSED_ROUND_KEYS_Y"):
        t = table(name)
        assert len(t) == 14
        # a known answer for the 28 values of each table without carrying them: SHA-256 of "a,b;a,b;..." plus the first row
        out[name.lower()] = {"sha256": hashlib.sha256(";".join(",".join(row) for row in t).encode()).hexdigest(), "first": t[0]}
    out["mds_matrix"] = table("MDS_MATRIX")
    assert len(out["mds_matrix"]) == 2
    out["generator"] = re.search(r'const GENERATOR: Fr = MontFp!\("(\d+)"\)', src).group(1)
    out["generator_inv"] = re.search(r'const GENERATOR_INV: Fr =\s*MontFp!\("(\d+)"\)', src).group(1)
    limbs = re.findall(r"(\d+)u64", src[src.index("fn get_alpha_inv"):])
    out["alpha_inv"] = str(sum(int(v) << (64 * i) for i, v in enumerate(limbs)))
    tests = open(f"{REF}/tests.rs").read()
    kat = tests[tests.index("fn test_eval_stream_cipher()"):]
    kat = kat[kat.index("let expect"):]
    kat = kat[: kat.index("];")]
    out["stream_cipher_1234"] = re.findall(r'MontFp!\(\s*"(\d+)"\s*\)', kat)
    vlh = tests[tests.index("fn test_anemoi_variable_length_hash()"):]
    out["variable_length_hash_1234"] = re.search(r'MontFp!\(\s*"(\d+)"\s*\)', vlh).group(1)
    assert len(out["stream_cipher_1234"]) == 7
    with open(os.path.join(HERE, "anemoi_bn254.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
