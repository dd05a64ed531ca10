#!/usr/bin/env python3
"""Rebuilds the setup Merkle cap (the hash part of the verification key) a golden proof was made with, from the proof alone:
boojum's OracleQuery stores leaf + path but not the index, so the index of every query is recovered from the WITNESS opening
(whose cap is in the proof) by trying all left/right patterns (oracle/primitives.c orc_merkle_find_index), and the setup opening of
the same query, hashed along the same index, ends in one entry of the setup cap.  100 queries cover all 16 entries with
probability 0.97.  Needed because most golden proofs under test_proofs/ are older than the VKs under setup/ (DESIGN.md section 5).
Usage: python tools/golden_recover_vk_cap.py proof.json [out_vk.json template_vk.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib
from tools.golden_transcript import merkle_root_from_path


def recover(pr, orc):
    cap_w = np.array(pr["witness_oracle_cap"], dtype=np.uint64)
    n_cap = len(cap_w)
    setup_cap = [None] * n_cap
    idxs = []
    for q in pr["queries_per_fri_repetition"]:
        idx = orc.merkle_find_index(q["witness_query"]["leaf_elements"], q["witness_query"]["proof"], cap_w)
        assert idx is not None, "witness opening fits no index"
        idxs.append(idx)
        root, top = merkle_root_from_path(q["setup_query"]["leaf_elements"], q["setup_query"]["proof"], idx)
        if setup_cap[top] is None: setup_cap[top] = list(root)
        else: assert setup_cap[top] == list(root), "two setup openings disagree on a cap entry"
    return setup_cap, idxs


if __name__ == "__main__":
    orc = oracle_lib.load()
    pr = json.load(open(sys.argv[1]))
    if "proof_config" not in pr: pr = list(pr.values())[0]
    cap, idxs = recover(pr, orc)
    missing = [i for i, c in enumerate(cap) if c is None]
    print("recovered", len(cap) - len(missing), "of", len(cap), "cap entries; missing", missing)
    if len(sys.argv) > 3 and not missing:
        vk = json.load(open(sys.argv[3]))
        inner = vk if "setup_merkle_tree_cap" in vk else list(vk.values())[0]
        inner["setup_merkle_tree_cap"] = cap
        json.dump(vk, open(sys.argv[2], "w"))
        print("wrote", sys.argv[2])
