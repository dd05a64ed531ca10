#!/usr/bin/env python3
"""One synthetic block through the whole proving workflow on N GPUs (BASELINE configs 3, 4 and 5):
base layer (13 circuit types) -> leaf -> node -> scheduler -> compression modes 1..4, every circuit at its reference size
(base / recursion 2^20, compression 2^16 / 2^13 / 2^12 / 2^15 with LDE 32 / 512 / 1024 / 2048), jobs farmed over the ranks,
one gather per stage (block.py).  Witnesses are synthetic satisfying traces (Rust synthesis is out of scope), generated on the
host and excluded from `prove_seconds`; every proof is checked by the CPU verifier on rank 0 unless --no-verify.

  python tools/bench_block.py [--log-n 20] [--instances 1] [--only-types 1,8] [--compression 1,2,3,4] [--out DIR]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_block.py ...
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zkevm_test_harness_b200 import GpuContext, block as B, prover_utils as PU  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--compression-log-n", type=int, default=None, help="override the trace length of the compression circuits")
    ap.add_argument("--instances", type=int, default=1, help="instances per base circuit type")
    ap.add_argument("--only-types", default="", help="comma-separated base circuit types (default: all 13)")
    ap.add_argument("--compression", default="1,2,3,4")
    ap.add_argument("--max-resident", type=int, default=4)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-verify", action="store_true")
    a = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    fixture = json.load(open(os.path.join(ROOT, "tests", "golden", "vk_shapes.json")))
    table, base_keys, leaf_keys, node_key, sched_key = B.circuit_table(fixture, log_n=a.log_n, compression_log_n=a.compression_log_n)
    types = [int(t) for t in a.only_types.split(",") if t] or sorted(base_keys)
    modes = tuple(int(m) for m in a.compression.split(",") if m)
    plan = B.plan_block({t: a.instances for t in types}, base_keys, leaf_keys, node_key, sched_key, compression_modes=modes)
    ctx = GpuContext(local)
    prover = B.GpuBlockProver(ctx, table, max_resident=a.max_resident)
    per_job = []

    def prove(job, seed):
        t0 = time.time()
        p = prover.prove(job, seed)
        per_job.append((job.file, time.time() - t0))
        return p

    caps = {}

    def verify(job, proof):
        key = job.geometry_key
        if key not in caps:   # the VK of a circuit type: from the resident setup if this rank built it, else build it once
            caps[key] = prover.vk_caps.get(key)
            if caps[key] is None:
                caps[key] = prover.setup(key).vk_cap.copy()
        geo, cfg = table[key]
        ok, msg = PU.verify_proof(geo, cfg, caps[key], proof)
        if not ok:
            print("verification FAILED:", job.file, msg, file=sys.stderr)
        return ok

    t0 = time.time()
    res = B.prove_block(plan, prove, block_seed=1, out_dir=a.out if rank == 0 else None,
                        verify=None if (a.no_verify or rank != 0) else verify, device=device if world > 1 else None,
                        prefetch=prover.prefetch)
    torch.cuda.synchronize()
    wall = time.time() - t0
    secs = torch.tensor([prover.seconds["prove"], prover.seconds["synth_trace"], prover.seconds["setup"]], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"workload": "one synthetic block: base -> leaf -> node -> scheduler -> compression", "n_gpus": world, "log_n": a.log_n,
                "base_types": types, "instances_per_type": a.instances, "compression_modes": list(modes), "n_proofs": plan.n_jobs,
                "prove_seconds_max_rank": round(float(secs[0]), 3), "synth_trace_cpu_seconds_max_rank": round(float(secs[1]), 1),
                "setup_seconds_max_rank": round(float(secs[2]), 2), "wall_seconds": round(wall, 1),
                "proofs_per_prove_second": round(plan.n_jobs / float(secs[0]), 2),
                "stages": [{**s, "seconds": round(s["seconds"], 2)} for s in res["stages"]],
                "rank0_prove_ms": prover.prove_ms, "rank0_jobs_ms_incl_trace_wait_and_setup": {f: round(1e3 * t, 1) for f, t in per_job}, "verified": not a.no_verify}
        print(json.dumps(line), flush=True)
    prover.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
