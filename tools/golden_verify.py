#!/usr/bin/env python3
"""Runs BOTH verifiers of this repo -- the oracle's (oracle/prover.c orc_verify) and the product's (zkgpu_verify, CPU like the
reference's) -- on the reference's own golden proof / verification-key pairs: the acceptance test of north_star
("every produced proof ... verifies against the reference's own verification keys") turned around, so that every convention the
prover shares with the verifiers is checked against proofs boojum itself made.

  python tools/golden_verify.py                      # all consistent pairs found under /root/reference
  python tools/golden_verify.py proof.json vk.json mainvm|compression_N|base_T|recursion"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zkevm_test_harness_b200 import geometry as G, proof_format as PF
from tests import oracle_lib

REF = "/root/reference"
PAIRS = ([("proof.json", "vk.json", "base_1")] + [(f"compression_{m}_proof.json", f"compression_{m}_vk.json", f"compression_{m}") for m in (1, 2, 3, 4)]
         + [(f"test_proofs/base_layer/basic_circuit_proof_{t}_0.json", f"setup/base_layer/vk_{t}.json", f"base_{t}") for t in (4, 8, 13)]
         + [(f"test_proofs/recursion_layer/node_layer_proof_{t}_0_0.json", "setup/recursion_layer/vk_node.json", "recursion") for t in range(3, 16)])


def load_pair(proof_path, vk_path, kind):
    _, vk = G.load_vk_json(vk_path)
    if kind.startswith("compression_"):
        mode = int(kind.split("_")[1])
        geo = G.geometry_from_vk(vk, G.COMPRESSION_GATE_ORDER[mode], has_boolean_col=1 if mode == 1 else 0)
    elif kind.startswith("base_"):
        geo = G.geometry_from_vk(vk, G.BASE_LAYER_GATE_ORDER[int(kind.split("_")[1])])
    elif kind == "recursion":
        geo = G.geometry_from_vk(vk, G.RECURSION_GATE_ORDER)
    else:
        raise ValueError(kind)
    d = json.load(open(proof_path))
    flat, _ = PF.proof_from_dict(d)
    inner = d if "proof_config" in d else list(d.values())[0]
    pc = inner["proof_config"]
    cfg = G.make_proof_config(geo.log_n, pc["fri_lde_factor"], pc["merkle_tree_cap_size"], pc["security_level"], pc["pow_bits"])
    assert cfg.n_queries == len(inner["queries_per_fri_repetition"]), (cfg.n_queries, len(inner["queries_per_fri_repetition"]))
    cap = np.array(vk["setup_merkle_tree_cap"], dtype=np.uint64).reshape(-1)
    return geo, cfg, cap, flat


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    skip = "--skip-quotient-identity" in sys.argv
    pairs = [(args[0], args[1], args[2])] if len(args) == 3 else [(os.path.join(REF, p), os.path.join(REF, v), k) for p, v, k in PAIRS]
    orc = oracle_lib.load()
    product = None
    try:
        from era_zkevm_test_harness_b200 import prover_utils as PU
        product = PU
    except Exception as e:  # noqa: BLE001
        print("product library not loadable:", e)
    rc = 0
    for p, v, k in pairs:
        geo, cfg, cap, flat = load_pair(p, v, k)
        ok, msg = orc.verify(geo, cfg, cap, flat, skip_quotient_identity=skip)
        print(f"{os.path.basename(p):34s} oracle verifier: {'ACCEPT' if ok else 'reject: ' + msg}")
        rc |= not ok
        if product:
            ok, msg = product.verify_proof(geo, cfg, cap, flat, skip_quotient_identity=skip)
            print(f"{'':34s} zkgpu_verify:    {'ACCEPT' if ok else 'reject: ' + msg}")
            rc |= not ok
    return rc


if __name__ == "__main__":
    sys.exit(main())
