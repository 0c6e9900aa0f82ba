"""Extract the reference's golden shuffle proofs into data fixtures (run HERE, where /root/reference exists).

    python tests/golden/make_golden_proof.py

Sources (all under /root/reference/contracts/solidity):
* test/plonk_{20,52}.js                        -- a PlonK proof (PlonkProof::to_bytes_be, uzkge/src/plonk/indexer.rs:538-590, `shuffle`
                                                  feature), the public inputs (input deck, output deck) and the 12 public-key
                                                  commitments the Solidity verifier is called with and must accept
* contracts/shuffle/VerifierKey_{20,52}.sol    -- the verifier parameters (uzkge/src/gen_params/solidity.rs)
* contracts/verifier/PlonkVerifier.sol:2025-2040 -- the two G2 elements of the KZG pairing check (EIP-197 order x1, x0, y1, y0)
Outputs tests/golden/plonk_{20,52}_golden.json: data only.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/contracts/solidity"


def main():
    sol = open(f"{REF}/contracts/verifier/PlonkVerifier.sol").read()
    tail = sol[sol.rindex("left_first_x)"):]
    g2 = re.findall(r"mstore\(add\(mload\(0x40\), (0x[0-9a-f]+)\), (0x[0-9a-f]{64})\)", tail)
    g2 = {int(a, 16): int(v, 16) for a, v in g2}
    tau_h = [g2[0x40], g2[0x60], g2[0x80], g2[0xA0]]
    h = [g2[0x100], g2[0x120], g2[0x140], g2[0x160]]
    for cards in (20, 52):
        js = open(f"{REF}/test/plonk_{cards}.js").read()
        hexes = re.findall(r'"(0x[0-9a-fA-F]+)"', js)
        proof = hexes[0]
        n_pi = 8 * cards
        vals = [int(x, 16) for x in hexes[1:]]
        assert len(vals) == n_pi + 24, (cards, len(vals))
        vk_txt = open(f"{REF}/contracts/shuffle/VerifierKey_{cards}.sol").read()
        words = {int(a, 16): int(v, 0) for a, v in re.findall(r"mstore\(add\(vk, (0x[0-9a-f]+)\), (0x[0-9a-f]+|\d+)\)", vk_txt)}
        out = {
            "proof": proof,
            "public_inputs": [hex(v) for v in vals[:n_pi]],
            "public_key_commitments": [hex(v) for v in vals[n_pi:]],
            "vk_words": {hex(k): hex(v) for k, v in sorted(words.items())},
            "g2_tau_h_eip197": [hex(v) for v in tau_h],
            "g2_h_eip197": [hex(v) for v in h],
            "n_cards": cards,
        }
        json.dump(out, open(os.path.join(HERE, f"plonk_{cards}_golden.json"), "w"), indent=0)
        print(cards, len(proof) // 2 - 1, "proof bytes;", n_pi, "public inputs")


if __name__ == "__main__":
    main()
