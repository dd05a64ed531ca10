#!/usr/bin/env python3
"""Proves ONE circuit of the fixture table (tests/golden/vk_shapes.json) at its reference size with a synthetic trace:
  python tools/prove_one.py compression_3 [--reps 3] [--log-n N]
keys: base_1_MainVM ... base_13_L1MessagesHasher, recursion_{scheduler,leaf_3,node}, compression_{1,2,3,4}.
Prints ms per proof (CUDA events, witness resident in HBM).  Used under ncu for per-circuit launch lists."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zkevm_test_harness_b200 import GpuContext, geometry as G, prover_utils as PU  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("key")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=None)
    a = ap.parse_args()
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "vk_shapes.json")))
    table = {k: (g, G.base_layer_proof_config(g.log_n)) for k, g, _ in G.circuit_geometries_from_fixture(fx)}
    table.update({k: (g, c) for k, g, c, _ in G.compression_geometries_from_fixture(fx)})
    geo, cfg = table[a.key]
    if a.log_n is not None and a.log_n != geo.log_n:
        geo = geo.scaled(a.log_n)
        cfg = G.make_proof_config(a.log_n, 1 << cfg.log_lde, cfg.cap_size, security_level=cfg.n_queries * cfg.log_lde)
    ctx = GpuContext(0)
    wit, setup = PU.synth_trace(geo, seed=3)
    sd = PU.create_setup_data(ctx, geo, cfg, setup)
    d_wit = torch.from_numpy(wit.view(np.int64)).cuda()
    proof = PU.prove_circuit(ctx, sd, d_wit)   # warm-up (tables, arena)
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        proof = PU.prove_circuit(ctx, sd, d_wit)
    e1.record()
    torch.cuda.synchronize()
    ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
    print(json.dumps({"circuit": a.key, "log_n": geo.log_n, "lde": 1 << cfg.log_lde, "cap": cfg.cap_size, "W": geo.n_witness,
                      "ms_per_proof": round(e0.elapsed_time(e1) / a.reps, 2), "launches_per_proof": (ctx.kernel_launches - l0) // a.reps,
                      "verified": bool(ok), "msg": msg}))
    sd.close()
    ctx.close()


if __name__ == "__main__":
    main()
