"""Generate tests/golden/* from the reference's own fixtures (run HERE, where /root/reference exists).

    python tests/golden/make_golden.py

Sources (all under /root/reference):
* uzkge/parameters/lagrange-srs-{4096,16384}.bin, srs-padding.bin  -- bundled SRS files
  (format: uzkge/src/poly_commit/kzg_poly_commitment.rs:207-264).  Known answer:
  MSM(lagrange_srs_n, [w_n^(i*j)]_i) == srs_padding[j]  because sum_i w^(ij) L_i(tau) = tau^j.
* contracts/solidity/contracts/shuffle/VerifierKey_{20,52}.sol      -- domain generator, k[0..5], cs_size
  (written by uzkge/src/gen_params/solidity.rs:17-146).
* .../VerifierKeyExtra{1,2}_{20,52}.sol                             -- w^idx and w^idx / n lists.

Outputs are data only (canonical integers, little-endian u64 limbs); no reference source is copied.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import bn254 as o  # noqa: E402

REF = "/root/reference"


def limbs(points):
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, P in enumerate(points):
        if P is None:
            continue
        out[i, :4] = o.int_to_limbs(P[0])
        out[i, 4:] = o.int_to_limbs(P[1])
    return out


def main():
    # ---- SRS fixtures (canonical, NON-Montgomery affine coordinates; identity = zeros)
    for n in (4096, 16384):
        raw = open(f"{REF}/uzkge/parameters/lagrange-srs-{n}.bin", "rb").read()
        pts = o.parse_srs_g1(raw)
        assert len(pts) == n and all(o.g1_is_on_curve(P) for P in pts)
        np.save(os.path.join(HERE, f"lagrange_srs_{n}.npy"), limbs(pts))
    raw = open(f"{REF}/uzkge/parameters/srs-padding.bin", "rb").read()
    pts = o.parse_srs_g1(raw)
    assert all(o.g1_is_on_curve(P) for P in pts)
    assert pts[0] == o.G1_GEN
    np.save(os.path.join(HERE, "srs_padding_head.npy"), limbs(pts[:64]))
    # tau^n .. tau^(n+2) for n = 4096, 8192, 16384: what load_srs_params puts at [n, n + 3) (uzkge/src/gen_params/mod.rs:147-171)
    assert len(pts) == 2060
    np.save(os.path.join(HERE, "srs_padding_tail.npy"), limbs(pts[2051:2060]))

    # ---- domain KATs from the generated Solidity verifier keys
    kat = {}
    for cards in (20, 52):
        txt = open(f"{REF}/contracts/solidity/contracts/shuffle/VerifierKey_{cards}.sol").read()
        words = {int(a, 16): v for a, v in re.findall(r"mstore\(add\(vk, (0x[0-9a-f]+)\), (0x[0-9a-f]+|\d+)\)", txt)}
        cs_size = int(words[0xC20], 0)
        entry = {
            "cs_size": cs_size,
            "root": words[0xC00],
            "k": [words[0xB40 + 0x20 * i] for i in range(5)],
        }
        for name, key in (("Extra1", "PI_POLY_INDICES_LOC"), ("Extra2", "PI_POLY_LAGRANGE_LOC")):
            t = open(f"{REF}/contracts/solidity/contracts/shuffle/VerifierKey{name}_{cards}.sol").read()
            vals = re.findall(key + r"\[(\d+)\] = (0x[0-9a-f]+);", t)
            entry[key] = [v for _, v in sorted(vals, key=lambda x: int(x[0]))]
        kat[str(cards)] = entry
    json.dump(kat, open(os.path.join(HERE, "domain_kat.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
