#!/usr/bin/env python3
"""Caches the CPU oracle's proof of the headline configuration -- MainVM-shaped circuit, trace 2^20, lde 2, cap 16, 100 queries,
trace seed 1 (oracle/synth.c) -- as tests/golden/oracle_proof_mainvm_2pow20_seed1.npy (93 069 u64, 745 KB).

oracle/prover.c needs ~8 minutes on 8 cores (16 GB) for this size, too long for the GPU suite, so the proof is produced here once
and the GPU test (tests/test_gpu_prover.py::test_full_size_mainvm_proof_equals_oracle) compares every u64 of the CUDA prover's
output with it.  Re-run after ANY change of the oracle's conventions.  Also used for compression mode 1 at its reference size
(2^16 x LDE 32): tests/golden/oracle_proof_compression_1_seed1.npy."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zkevm_test_harness_b200 import geometry as G  # noqa: E402
from tests import oracle_lib  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    orc = oracle_lib.load()
    which = sys.argv[1:] or ["mainvm", "compression_1"]
    if "mainvm" in which:
        geo, cfg = G.mainvm_like_geometry(20), G.base_layer_proof_config(20)
        wit, setup = orc.synth_trace(geo, seed=1)
        t = time.time()
        proof = orc.prove(geo, cfg, wit, setup)
        print("mainvm 2^20:", time.time() - t, "s", proof.size, "u64")
        np.save(os.path.join(OUT, "oracle_proof_mainvm_2pow20_seed1.npy"), proof)
    if "compression_1" in which:
        fixture = json.load(open(os.path.join(OUT, "vk_shapes.json")))
        geo, cfg = [(g, c) for k, g, c, _ in G.compression_geometries_from_fixture(fixture) if k == "compression_1"][0]
        wit, setup = orc.synth_trace(geo, seed=1)
        t = time.time()
        proof = orc.prove(geo, cfg, wit, setup)
        print("compression mode 1 (2^16 x 32):", time.time() - t, "s", proof.size, "u64")
        np.save(os.path.join(OUT, "oracle_proof_compression_1_seed1.npy"), proof)


if __name__ == "__main__":
    main()
