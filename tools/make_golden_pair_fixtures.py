#!/usr/bin/env python3
"""Trims the reference's consistent golden (proof, verification key) pairs to their first three queries and stores them under
tests/golden/pair_*.json, so the CPU suite can run both verifiers of this repo on proofs boojum itself made without
/root/reference (which is absent on the GPU box).  Query indexes are drawn from the transcript one after the other after the
final FRI monomials, so the first K queries of a proof verify under the same proof config with K queries.
Pairs: /root/reference/{proof.json + vk.json (MainVM, 2^20, cap 32), compression_{1..4}_{proof,vk}.json}, and the pairs under
test_proofs/ + setup/ whose setup openings end in the VK's cap (tools/golden_transcript.py): base-layer 4, 8, 13 and all node proofs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
K = 3
PAIRS = ([("mainvm", "proof.json", "vk.json", "base_1")]
         + [(f"compression_{m}", f"compression_{m}_proof.json", f"compression_{m}_vk.json", f"compression_{m}") for m in (1, 2, 3, 4)]
         + [(f"base_{t}", f"test_proofs/base_layer/basic_circuit_proof_{t}_0.json", f"setup/base_layer/vk_{t}.json", f"base_{t}") for t in (8, 13)]
         + [("node_3", "test_proofs/recursion_layer/node_layer_proof_3_0_0.json", "setup/recursion_layer/vk_node.json", "recursion")])


def inner(d, key):
    return d if key in d else list(d.values())[0]


def main():
    for name, p, v, kind in PAIRS:
        pr = inner(json.load(open(os.path.join(REF, p))), "proof_config")
        vk = inner(json.load(open(os.path.join(REF, v))), "setup_merkle_tree_cap")
        pr = dict(pr)
        pr["queries_per_fri_repetition"] = pr["queries_per_fri_repetition"][:K]
        out = {"source": [p, v], "kind": kind, "n_queries": K, "vk": vk, "proof": pr}
        path = os.path.join(ROOT, "tests", "golden", f"pair_{name}.json")
        json.dump(out, open(path, "w"), separators=(",", ":"))
        print(path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
