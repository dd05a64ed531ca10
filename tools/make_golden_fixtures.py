#!/usr/bin/env python3
"""Builds tests/golden/*.json from the reference's golden proofs (run in the build container, where /root/reference
exists; the fixtures travel to the GPU box, the reference does not).

fri_chain_<name>.json: for a golden proof, the FRI part of the first N queries (leaf elements of every FRI oracle),
`final_fri_monomials`, and -- recovered WITHOUT the hash by tools/golden_fri_chain.py (polynomial gcd over GF(p^2))
-- every folding challenge and each query's leaf index in every oracle.  All 100 queries of the proof are consistent
with the recovered values, which pins: Ext2 = F[u]/(u^2-7), split leaf serialisation [c0.., c1..], LDE domain
7*<omega>, bit-reversed enumeration, un-normalised fold with c, c^2, c^4 per oracle, folding schedule.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_fri_chain import analyse  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PROOFS = {
    "mainvm_1_0": "test_proofs/base_layer/basic_circuit_proof_1_0.json",
    "ram_8_0": "test_proofs/base_layer/basic_circuit_proof_8_0.json",
    "node_3_0_0": "test_proofs/recursion_layer/node_layer_proof_3_0_0.json",
    # compression layer: LDE 32 (2^16), LDE 512 (2^13), and the wrapper-facing variants (LDE 512; LDE 2 at 2^16), final
    # polynomial of ONE monomial for the first three.  Modes 3 and 4 (last oracle folds by 8 onto a constant) are degenerate for
    # this hash-free recovery (tools/golden_fri_chain.py).
    "compression_1": "compression_1_proof.json",
    "compression_2": "compression_2_proof.json",
    "compression_2_for_wrapper": "compression_2_for_wrapper_proof.json",
    "compression_1_for_wrapper": "test_proofs/aux_layer/compression_for_wrapper_proof_1.json",
}
N_QUERIES = 16


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, rel in PROOFS.items():
        path = os.path.join(REF, rel)
        res = analyse(path)
        pr = json.load(open(path))
        if "proof_config" not in pr:
            pr = pr[list(pr.keys())[0]]
        Q = pr["queries_per_fri_repetition"]
        fx = {
            "source": rel,
            "proof_config": pr["proof_config"],
            "schedule": res["schedule"],
            "log_domains": res["log_domains"],
            "challenges": res["challenges"],
            "all_queries_consistent": all(all(i is not None for i in lv) for lv in res["leaf_indexes"]),
            "n_queries_in_proof": len(Q),
            "final_fri_monomials": pr["final_fri_monomials"],
            "queries": [
                {"leaf_indexes": [res["leaf_indexes"][k][q] for k in range(len(res["schedule"]))],
                 "fri_leaves": [fq["leaf_elements"] for fq in Q[q]["fri_queries"]]}
                for q in range(min(N_QUERIES, len(Q)))
            ],
        }
        with open(os.path.join(OUT, f"fri_chain_{name}.json"), "w") as f:
            json.dump(fx, f)
        print(name, "ok", fx["all_queries_consistent"])


# ---- DEEP structure: phi, z and the query positions recovered hash-free (tools/golden_deep.py) for lookup-free circuits
DEEP = {   # name -> (proof, VK with the public-input locations, key of the circuit in tests/golden/vk_shapes.json)
    "node_3_0_0": ("test_proofs/recursion_layer/node_layer_proof_3_0_0.json", "setup/recursion_layer/vk_node.json", ["recursion", "node"]),
    "compression_1": ("compression_1_proof.json", "compression_1_vk.json", ["compression", "1"]),
    # NOT reproduced by this structure (no common root of the three-query system, every hypothesis tried): compression mode 2
    # (compression_2_proof.json, compression_2_for_wrapper_proof.json -- no specialised boolean column, BoundedBoolean gate) and
    # the base-layer circuits with lookups (basic_circuit_proof_1_0.json).  See DESIGN.md section 5.
    "compression_1_for_wrapper": ("test_proofs/aux_layer/compression_for_wrapper_proof_1.json", "setup/aux_layer/compression_for_wrapper_vk_1.json",
                                  ["compression", "1_for_wrapper"]),
}
N_DEEP_QUERIES = 6


def deep_fixtures():
    import golden_deep
    from golden_fri_chain import P, omega, brev
    for name, (rel, vk_rel, shape_key) in DEEP.items():
        vk = json.load(open(os.path.join(REF, vk_rel)))
        if "fixed_parameters" not in vk:
            vk = vk[list(vk.keys())[0]]
        fp = vk["fixed_parameters"]
        pil = fp["public_inputs_locations"]
        par = fp["parameters"]
        # + the specialised boolean column, except compression mode 2 (BoundedBoolean gate on general-purpose columns instead)
        n_perm = par["num_columns_under_copy_permutation"] + (0 if shape_key[1].startswith("2") else 1)
        n_const = par["num_constant_columns"] + fp["extra_constant_polys_for_selectors"]

        def order(w, s, s2, qq):
            b = lambda v: (v, 0)
            sig, con = s[:n_perm], s[n_perm:n_perm + n_const]
            s2e = [(s2[2 * e], s2[2 * e + 1]) for e in range(len(s2) // 2)]
            qe = [(qq[2 * e], qq[2 * e + 1]) for e in range(len(qq) // 2)]
            return [b(v) for v in w] + [b(v) for v in con] + [b(v) for v in sig] + s2e + qe, s2e[0], []

        fx_path = os.path.join(OUT, f"fri_chain_{name}.json")
        res, _ = golden_deep.solve(os.path.join(REF, rel), fx_path, order=order, pi_locs=pil)
        assert len(res) == 1 and res[0]["consistent"], name
        r = res[0]
        fx = json.load(open(fx_path))
        pr = json.load(open(os.path.join(REF, rel)))
        if "proof_config" not in pr:
            pr = pr[list(pr.keys())[0]]
        log_dom = fx["log_domains"][0]
        qs = []
        for q in range(min(N_DEEP_QUERIES, len(fx["queries"]))):
            Q = pr["queries_per_fri_repetition"][q]
            pos = r["positions"][q]
            idx = (fx["queries"][q]["leaf_indexes"][0] << 3) + pos
            fl = Q["fri_queries"][0]["leaf_elements"]
            qs.append({"lde_index": idx, "x": 7 * pow(omega(log_dom), brev(idx, log_dom), P) % P,
                       "witness": Q["witness_query"]["leaf_elements"], "setup": Q["setup_query"]["leaf_elements"],
                       "stage_2": Q["stage_2_query"]["leaf_elements"], "quotient": Q["quotient_query"]["leaf_elements"],
                       "fri_base_value": [fl[pos], fl[len(fl) // 2 + pos]]})
        out = {"source": rel, "shape_key": shape_key, "phi": list(r["phi"]), "z": list(r["z"]),
               "all_fixture_queries_consistent": r["consistent"], "positions_in_fri_leaf": r["positions"],
               "public_inputs": pr["public_inputs"], "values_at_z": [c["coeffs"] for c in pr["values_at_z"]],
               "values_at_z_omega": [c["coeffs"] for c in pr["values_at_z_omega"]], "values_at_0": [c["coeffs"] for c in pr["values_at_0"]],
               "queries": qs}
        with open(os.path.join(OUT, f"deep_{name}.json"), "w") as f:
            json.dump(out, f)
        print("deep", name, "ok")


if __name__ == "__main__":
    if "--deep-only" not in sys.argv:
        main()
    deep_fixtures()
