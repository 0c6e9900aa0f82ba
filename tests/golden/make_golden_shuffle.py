"""Extract the reference's preprocessed Baby Jubjub generator tables into a data fixture (run HERE, where /root/reference exists).

    python tests/golden/make_golden_shuffle.py

Source: /root/reference/uzkge/src/shuffle/babyjubjub.rs:24-3566 -- `get_preprocessed_generators_{x,y,dxy}`: for each of the 84
iterations of the remark gate's scalar multiplication the four points (j + 1) * 16^i * G (affine x, y and d * x * y), i.e. the
output of `Remark::crate_generators` (uzkge/src/shuffle/remark.rs:39-60).  They pin the twisted Edwards arithmetic, the curve
constants a and d and the generator of every restatement in this repository (tests/test_shuffle_host.py).
Output tests/golden/babyjubjub_generators.json: the SHA-256 digest of each table (decimal entries joined by "," within a row and
";" between rows) plus the first and the last row -- a known answer for all 1008 values without carrying them.
"""
import hashlib
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open("/root/reference/uzkge/src/shuffle/babyjubjub.rs").read()
    parts = re.split(r"fn get_preprocessed_generators_(\w+)\(\)", src)
    out = {}
    for name, body in zip(parts[1::2], parts[2::2]):
        body = body.split("#[test]")[0]
        segs = re.findall(r'vec!\[((?:\s*MontFp!\(\s*"\d+"\s*\),?)+)\s*\]', body)
        table = [re.findall(r'"(\d+)"', s) for s in segs]
        assert len(table) == 84 and all(len(r) == 4 for r in table), name
        digest = hashlib.sha256(";".join(",".join(row) for row in table).encode()).hexdigest()
        out[name] = {"sha256": digest, "rows": 84, "first": table[0], "last": table[-1]}
    assert sorted(out) == ["dxy", "x", "y"]
    with open(os.path.join(HERE, "babyjubjub_generators.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
