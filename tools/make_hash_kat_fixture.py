#!/usr/bin/env python3
"""Extracts hash-agnostic Poseidon2 known-answer material from the reference's golden proofs.

Source: /root/reference/test_proofs/base_layer/basic_circuit_proof_1_0.json (MainVM, 2^21 leaves, cap 16).
  * `node`: the 31 distinct top-of-path siblings of the witness oracle and its 16-entry cap.  Whatever the tree hasher is,
    two of those siblings that sit under the same cap entry satisfy cap[j] = H_node(left, right)  (one permutation).
  * `leaf`: the 16 distinct 8-element leaves of the last FRI oracle (16 leaves, empty path) and its cap:
    every H_leaf(8 elements) is a cap entry (one permutation).
Run here (needs /root/reference); the output tests/golden/poseidon2_kat.json travels with the repo.
"""
import json, os

REF = "/root/reference/test_proofs/base_layer/basic_circuit_proof_1_0.json"


def main():
    pr = json.load(open(REF))["MainVM"]
    qs = pr["queries_per_fri_repetition"]
    tops, leaves = [], []
    for q in qs:
        t = q["witness_query"]["proof"][-1]
        if t not in tops:
            tops.append(t)
        l = q["fri_queries"][-1]["leaf_elements"]
        assert q["fri_queries"][-1]["proof"] == []
        if l not in leaves:
            leaves.append(l)
    out = {
        "source": "test_proofs/base_layer/basic_circuit_proof_1_0.json",
        "node": {"top_siblings": tops, "cap": pr["witness_oracle_cap"]},
        "leaf": {"leaves": leaves, "cap": pr["fri_intermediate_oracles_caps"][-1]},
    }
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "poseidon2_kat.json")
    json.dump(out, open(dst, "w"))
    print(dst, len(tops), "top siblings,", len(leaves), "leaves")


if __name__ == "__main__":
    main()
