"""Extract zmatchmaking's bundled verifier key into a data fixture (run HERE, where /root/reference exists).

    python tests/golden/make_golden_matchmaking.py

Source: /root/reference/matchmaking/parameters/vk-specific.bin = bincode(VerifierParamsSplitSpecific { shrunk_cs: TurboCS<Fr>,
verifier_params: PlonkVerifierParams }) (uzkge/src/gen_params/mod.rs:85-92; field order: plonk/constraint_system/turbo/mod.rs:30-93,
plonk/indexer.rs:153-191; `ark_serialize` fields are length-prefixed byte strings holding arkworks' compressed encoding,
utils/serialization.rs:6-49).  G1 commitments are decompressed (x, flags in the top two bits: 0x80 = y is the larger root,
0x40 = infinity).  The file predates the `shuffle` feature set (8 selector commitments, no shuffle fields).  Also writes the Lagrange SRS of the circuit's size (uzkge/parameters/lagrange-srs-8192.bin).
Outputs tests/golden/matchmaking_vk.json, tests/golden/lagrange_srs_8192.npy: data only.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import bn254 as o  # noqa: E402

P = o.FQ


class Reader:
    def __init__(self, raw):
        self.raw, self.pos = raw, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.raw, self.pos)[0]
        self.pos += 8
        return v

    def take(self, n):
        b = self.raw[self.pos:self.pos + n]
        assert len(b) == n
        self.pos += n
        return b

    def blob(self):
        return self.take(self.u64())

    def usizes(self):
        return [self.u64() for _ in range(self.u64())]


def fr(b):
    assert len(b) == 32
    return int.from_bytes(b, "little")


def fr_vec(b):
    n = struct.unpack_from("<Q", b, 0)[0]
    assert len(b) == 8 + 32 * n
    return [fr(b[8 + 32 * i: 40 + 32 * i]) for i in range(n)]


def g1(b):
    assert len(b) == 32
    v = int.from_bytes(b, "little")
    neg, inf = bool(v >> 255), bool((v >> 254) & 1)
    x = v & ((1 << 254) - 1)
    if inf:
        return None
    y = pow((x * x * x + 3) % P, (P + 1) // 4, P)
    assert y * y % P == (x * x * x + 3) % P
    if (y > P - y) != neg:
        y = P - y
    return (x, y)


def main():
    r = Reader(open("/root/reference/matchmaking/parameters/vk-specific.bin", "rb").read())
    cs = {}
    assert r.blob() == bytes(8)                                  # selectors: empty
    assert [r.usizes() for _ in range(5)] == [[]] * 5            # wiring
    cs["edwards_a"] = fr(r.blob())
    for _ in range(6):
        assert r.blob() == bytes(8)                              # shuffle tables: empty
    assert len(r.blob()) == 28 * 32 and len(r.blob()) == 28 * 32  # Anemoi round keys (zeroed by shrink_to_verifier_only)
    r.blob(), r.blob()                                           # anemoi generator / inverse (zeroed)
    assert r.usizes() == []
    cs["n_iteration_shuffle_scalar_mul"], cs["num_vars"], cs["size"] = r.u64(), r.u64(), r.u64()
    assert r.usizes() == [] and r.usizes() == [] and r.usizes() == []
    assert r.blob() == bytes(8)
    assert r.take(1) == b"\x01"                                  # verifier_only
    assert r.blob() == bytes(8)                                  # witness
    cms = lambda: [g1(r.blob()) for _ in range(r.u64())]
    # The bundled file predates the `shuffle` feature set: 8 selector commitments (no q_ecc column), no cm_q_ecc / cm_shuffle_* /
    # edwards_a fields -- PlonkVerifierParams as it is compiled without `features = ["shuffle"]`.
    vp = {"cm_q_vec": cms(), "cm_s_vec": cms(), "cm_qb": g1(r.blob()), "cm_prk_vec": cms()}
    vp["anemoi_generator"], vp["anemoi_generator_inv"] = fr(r.blob()), fr(r.blob())
    vp["k"] = fr_vec(r.blob())
    vp["cs_size"] = r.u64()
    vp["public_vars_constraint_indices"] = r.usizes()
    vp["lagrange_constants"] = fr_vec(r.blob())
    assert r.pos == len(r.raw), (r.pos, len(r.raw))
    assert len(vp["cm_q_vec"]) == 8 and len(vp["cm_s_vec"]) == 5 and len(vp["cm_prk_vec"]) == 4 and vp["anemoi_generator"] == 5
    assert all(c is None or o.g1_is_on_curve(c) for key in ("cm_q_vec", "cm_s_vec", "cm_prk_vec") for c in vp[key])
    enc = lambda v: (None if v is None else [hex(v[0]), hex(v[1])]) if (v is None or isinstance(v, tuple)) else v
    out = {"shrunk_cs": cs}
    for key, v in vp.items():
        if isinstance(v, list) and v and (v[0] is None or isinstance(v[0], tuple)):
            out[key] = [enc(c) for c in v]
        elif v is None or isinstance(v, tuple):
            out[key] = enc(v)
        elif isinstance(v, list):
            out[key] = [hex(x) if x > (1 << 32) else x for x in v]
        else:
            out[key] = hex(v) if v > (1 << 32) else v
    json.dump(out, open(os.path.join(HERE, "matchmaking_vk.json"), "w"), indent=0)

    sys.path.insert(0, HERE)
    import make_golden as mg

    pts = o.parse_srs_g1(open("/root/reference/uzkge/parameters/lagrange-srs-8192.bin", "rb").read())
    assert len(pts) == 8192
    np.save(os.path.join(HERE, "lagrange_srs_8192.npy"), mg.limbs(pts))
    print({k: (len(v) if isinstance(v, list) else v) for k, v in out.items() if k != "shrunk_cs"}, out["shrunk_cs"])


if __name__ == "__main__":
    main()
