#!/usr/bin/env python3
"""Proves every circuit type of the reference (13 base-layer circuits, scheduler, leaf, node; geometry from the VK fixtures)
at the full trace length 2^20 on one GPU: setup time, ms per proof (device-resident witness, CUDA events, 1 warm-up + 2 timed),
proof size, CPU-verifier verdict.  BASELINE configs 3 and 4 at full size.  Output: one JSON line per circuit."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zkevm_test_harness_b200 import GpuContext, geometry as G, prover_utils as PU  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    only = sys.argv[2:] if len(sys.argv) > 2 else None
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "vk_shapes.json")))
    ctx = GpuContext(0)
    for key, geo, _ in G.circuit_geometries_from_fixture(fx):
        if only and not any(o in key for o in only):
            continue
        g = geo.scaled(log_n) if log_n != geo.log_n else geo
        cfg = G.base_layer_proof_config(log_n)
        t0 = time.time()
        wit, setup = PU.synth_trace(g, seed=77)
        t_synth = time.time() - t0
        t0 = time.time()
        sd = PU.create_setup_data(ctx, g, cfg, setup)
        torch.cuda.synchronize()
        t_setup = time.time() - t0
        del setup
        d_wit = torch.from_numpy(wit.view(np.int64)).cuda()
        proof = PU.prove_circuit(ctx, sd, d_wit)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            proof = PU.prove_circuit(ctx, sd, d_wit)
        e1.record()
        torch.cuda.synchronize()
        ok, msg = PU.verify_proof(g, cfg, sd.vk_cap, proof)
        print(json.dumps({"circuit": key, "log_n": log_n, "W": g.n_witness, "S2": g.n_stage2, "S": g.n_setup, "ms_per_proof": round(e0.elapsed_time(e1) / 2, 2),
                          "setup_s": round(t_setup, 2), "synth_trace_cpu_s": round(t_synth, 1), "proof_bytes": int(proof.size) * 8, "verified": bool(ok),
                          "msg": msg}), flush=True)
        sd.close()
        del d_wit, wit
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
