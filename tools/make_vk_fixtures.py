#!/usr/bin/env python3
"""Builds tests/golden/vk_shapes.json from the reference's verification keys and golden proofs (run in the build
container; /root/reference does not exist on the GPU box).

Per circuit: the VK's `fixed_parameters` (geometry, lookup parameters, selector tree, public-input locations -- the input of
geometry.geometry_from_vk), its `setup_merkle_tree_cap`, and the oracle widths OBSERVED in a golden proof of that circuit
(leaf lengths of the four trace oracles, Merkle path length, counts of values_at_z / z_omega / 0, FRI leaf sizes and path
lengths, number of queries) so the tests can check the column-count formulas of zkgpu_num_*_cols and the folding schedule
against the reference's own artefacts."""
import glob
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vk_shapes.json")


def inner(d):
    if "fixed_parameters" in d or "proof_config" in d:
        return None, d
    (name, body), = d.items()
    return name, body


def proof_shape(path):
    name, pr = inner(json.load(open(path)))
    q = pr["queries_per_fri_repetition"][0]
    return {
        "source": os.path.relpath(path, REF), "variant": name, "proof_config": pr["proof_config"],
        "W": len(q["witness_query"]["leaf_elements"]), "S2": len(q["stage_2_query"]["leaf_elements"]),
        "Q": len(q["quotient_query"]["leaf_elements"]), "S": len(q["setup_query"]["leaf_elements"]),
        "path_len": len(q["witness_query"]["proof"]), "n_queries": len(pr["queries_per_fri_repetition"]),
        "values_at_z": len(pr["values_at_z"]), "values_at_z_omega": len(pr["values_at_z_omega"]), "values_at_0": len(pr["values_at_0"]),
        "fri": [[len(f["leaf_elements"]), len(f["proof"])] for f in q["fri_queries"]],
        "final_fri_monomials": [len(pr["final_fri_monomials"][0]), len(pr["final_fri_monomials"][1])],
        "cap_len": len(pr["witness_oracle_cap"]), "n_public_inputs": len(pr["public_inputs"]),
    }


def main():
    out = {"base": {}, "recursion": {}, "compression": {}, "aux": {}}
    # EIP-4844 circuit (prove_eip4844_circuit, src/prover_utils.rs:763-808): a VK only, the reference holds no proof of it
    name, vk = inner(json.load(open(f"{REF}/setup/aux_layer/eip4844_vk.json")))
    out["aux"]["eip4844"] = {"variant": name or "EIP4844", "fixed_parameters": vk["fixed_parameters"],
                             "setup_merkle_tree_cap": vk["setup_merkle_tree_cap"], "proof_shapes": []}
    for t in range(1, 14):
        name, vk = inner(json.load(open(f"{REF}/setup/base_layer/vk_{t}.json")))
        proofs = sorted(glob.glob(f"{REF}/test_proofs/base_layer/basic_circuit_proof_{t}_*.json"))
        # basic_circuit_proof_2_0.json is stale (domain 2^15 test geometry): shape recorded, flagged
        shapes = [proof_shape(p) for p in proofs]
        out["base"][str(t)] = {"variant": name, "fixed_parameters": vk["fixed_parameters"], "setup_merkle_tree_cap": vk["setup_merkle_tree_cap"],
                               "proof_shapes": shapes}
    rec = {"scheduler": ("vk_1.json", ["scheduler_proof.json"]), "leaf_3": ("vk_3.json", ["leaf_layer_proof_3_0.json"]),
           "node": ("vk_node.json", ["node_layer_proof_3_0_0.json"])}
    for key, (vkf, proofs) in rec.items():
        name, vk = inner(json.load(open(f"{REF}/setup/recursion_layer/{vkf}")))
        shapes = [proof_shape(f"{REF}/test_proofs/recursion_layer/{p}") for p in proofs if os.path.exists(f"{REF}/test_proofs/recursion_layer/{p}")]
        out["recursion"][key] = {"variant": name, "fixed_parameters": vk["fixed_parameters"], "setup_merkle_tree_cap": vk["setup_merkle_tree_cap"],
                                 "proof_shapes": shapes}
    # compression layer (aux_layer/compression_modes/mode_N.rs): the one-shot VK/proof pairs at the reference root
    # (src/proof_compression/mod.rs:45-86) and the wrapper-facing mode-1 pair under setup/ and test_proofs/
    comp = {"1": ("compression_1_vk.json", "compression_1_proof.json", 1), "2": ("compression_2_vk.json", "compression_2_proof.json", 2),
            "3": ("compression_3_vk.json", "compression_3_proof.json", 3), "4": ("compression_4_vk.json", "compression_4_proof.json", 4),
            "2_for_wrapper": ("compression_2_for_wrapper_vk.json", "compression_2_for_wrapper_proof.json", 2),
            "1_for_wrapper": ("setup/aux_layer/compression_for_wrapper_vk_1.json", "test_proofs/aux_layer/compression_for_wrapper_proof_1.json", 1)}
    for key, (vkf, prf, mode) in comp.items():
        name, vk = inner(json.load(open(f"{REF}/{vkf}")))
        out["compression"][key] = {"variant": name or f"CompressionMode{mode}Circuit", "mode": mode, "fixed_parameters": vk["fixed_parameters"],
                                   "setup_merkle_tree_cap": vk["setup_merkle_tree_cap"], "proof_shapes": [proof_shape(f"{REF}/{prf}")]}
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()


def trimmed_proofs(n_queries=2):
    """tests/golden/proof_<name>_2q.json: golden proofs cut to their first two queries (the file format test needs the
    structure, not all 100 openings)."""
    picks = {"mainvm_1_0": "test_proofs/base_layer/basic_circuit_proof_1_0.json",
             "node_3_0_0": "test_proofs/recursion_layer/node_layer_proof_3_0_0.json"}
    for name, rel in picks.items():
        d = json.load(open(os.path.join(REF, rel)))
        (variant, body), = d.items()
        body["queries_per_fri_repetition"] = body["queries_per_fri_repetition"][:n_queries]
        out = os.path.join(os.path.dirname(OUT), f"proof_{name}_{n_queries}q.json")
        with open(out, "w") as f:
            json.dump({variant: body}, f, separators=(",", ":"))
        print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    trimmed_proofs()
