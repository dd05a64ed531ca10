#!/usr/bin/env python3
"""bench.py -- base-layer proofs/sec at trace 2^20 (MainVM shape) on N x B200, with the NTT roofline and a CPU baseline.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--log-n 20]

One step = one proof of one MainVM-shaped circuit instance (W=156, S2=58, Q=16, S=167 columns, 2^20 rows, lde 2,
cap 16, 100 queries) over a synthetic satisfying trace.  `value` is measured with the witness resident in HBM
(zkgpu_prove_device), `e2e` through the public host-buffer calls with pinned host witnesses, H2D and proof D2H inside the timed
region: staged (zkgpu_witness_stage of witness k+1 while zkgpu_prove_staged proves witness k -- the reference proves its circuits
in a loop) as the headline, and the single call zkgpu_prove beside it.  Multi-GPU: independent circuit instances, one per rank per step (weak scaling, no data-path collective);
NCCL is used once per step to gather the finished proofs to rank 0, as north_star asks.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--cpu-log-n", type=int, default=16, help="trace length of the bounded CPU sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
WORKLOAD = "MainVM-shaped base-layer circuit (W156/S2 58/Q16/S167, 11 gates incl. flattened Poseidon2, lookup 3x8), trace 2^{log_n}, lde 2, cap 16, 100 queries"


def _geometry_module():
    """geometry.py alone (ctypes structs, pure Python): the CPU arm must not import the package, whose __init__ pulls in torch
    and the binding of libzkgpu.so."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("zk_geometry_only", os.path.join(ROOT, "era_zkevm_test_harness_b200", "geometry.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle_with_all_cores():
    """liboracle.so with its OpenMP team set to every core this process may run on -- explicitly, because
    torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which would serialise the CPU prover."""
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from tests import oracle_lib
    oracle = oracle_lib.load()
    return oracle, oracle.set_threads(cores)


def cpu_prove_once(log_n, seed=1):
    """One whole proof of the MainVM-shaped circuit at trace 2^log_n by the CPU restatement (oracle/prover.c, OpenMP), on a trace
    made by oracle/synth.c.  Nothing of the product (libzkgpu.so, torch) is loaded on this path.  kind = "port": the reference's
    Rust prover (boojum) cannot be built in this image (no Rust toolchain, un-vendored git dependencies)."""
    G = _geometry_module()
    oracle, threads = _oracle_with_all_cores()
    geo = G.mainvm_like_geometry(log_n)
    cfg = G.base_layer_proof_config(log_n)
    wit, setup = oracle.synth_trace(geo, seed=seed)
    t0 = time.time()
    proof = oracle.prove(geo, cfg, wit, setup)
    return time.time() - t0, threads, int(proof.size)


def cpu_baseline(log_n_full, log_n_sample):
    """Bounded sample for the GPU arm's line: one whole CPU proof of the same geometry and proof config on a SHORTER trace
    (2^log_n_sample rows, ~10-30 s), reported both as measured and scaled by N log N to the metric's trace length.  The
    un-extrapolated number is what `bench.py --impl reference` measures: one full 2^20 proof."""
    dt, threads, n_u64 = cpu_prove_once(log_n_sample)
    scale = ((1 << log_n_full) * log_n_full) / ((1 << log_n_sample) * log_n_sample)
    return {"value": 1.0 / (dt * scale), "unit": "proofs/s", "cores": threads, "kind": "port", "extrapolated": log_n_sample != log_n_full,
            "sample": f"one whole proof (oracle/prover.c, OpenMP x{threads}) of the MainVM geometry at trace 2^{log_n_sample} took {dt:.2f} s"
                      + (f"; scaled by N*log2(N) x{scale:.1f} to trace 2^{log_n_full}" if log_n_sample != log_n_full else ""),
            "sample_seconds": dt, "sample_value_unscaled": 1.0 / dt, "sample_log_n": log_n_sample, "proof_u64": n_u64}


def run_reference(args):
    """CPU arm: ONE full-size proof (trace 2^log_n, the GPU arm's exact geometry and proof config) by the CPU restatement on all
    host cores -- measured, not extrapolated.  A full proof takes minutes on the host, so --steps/--warmup are not repeated:
    steps_measured = 1, warmup_measured = 0 (the page cache and the OpenMP team are warmed by the trace generator).  Under
    torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads, n_u64 = cpu_prove_once(args.log_n)
    value = 1.0 / sec
    cb = {"value": value, "unit": "proofs/s", "cores": threads, "kind": "port", "extrapolated": False,
          "sample": f"one whole proof (oracle/prover.c, OpenMP x{threads}) of the MainVM geometry at trace 2^{args.log_n} took {sec:.1f} s",
          "sample_seconds": sec, "proof_u64": n_u64}
    print(json.dumps({
        "impl": "reference", "metric": "base_layer_proofs_per_sec_trace_2^20", "value": value, "unit": "proofs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_measured": 1, "warmup_measured": 0, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(log_n=args.log_n) + "; one instance per GPU per step",
                   "note": "CPU restatement of the prover (oracle port, not boojum) on all host cores of rank 0; one full-size proof, no extrapolation"},
        "cpu_baseline": cb, "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ntt_traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum of the two launches (pass A, pass B) of one forward coset NTT batch of 156
    columns x 2^20, read from the newest committed `ncu --set full` export profiles/r*_prims_raw.csv (tools/ncu_capture.sh over
    tools/profile_kernels.py primitives: the same ctx.ntt_forward call this bench times).  Returns (bytes, source) or (None, why)."""
    import csv
    import glob
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prims_raw.csv")), reverse=True):
        try:
            rows = list(csv.reader(open(path)))
            idx = {h: i for i, h in enumerate(rows[0])}
            units = rows[1]
            fwd = [r for r in rows[2:] if "ntt1024" in r[idx["Kernel Name"]]][:2]      # first captured batch: forward pass A + pass B
            if len(fwd) < 2 or any(r[idx["Grid Size"]].replace(" ", "") != "(128,156,1)" for r in fwd):
                continue
            total = 0.0
            for r in fwd:
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(r[idx[key]].replace(",", "")) * scale[units[idx[key]]]
            return total, os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, "no profiles/r*_prims_raw.csv with an ntt1024 batch of 156 columns"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, device_index):
        self.proc = None
        self.lines = []
        self.device_index = device_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from era_zkevm_test_harness_b200 import GpuContext, farm, geometry as G, prover_utils as PU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    host_cores = farm.bind_host_to_gpu(local_rank) if world > 1 else 0   # pinned witness buffers on the GPU's NUMA node
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = GpuContext(local_rank)
    dev = ctx.device

    log_n = args.log_n
    geo = G.mainvm_like_geometry(log_n)
    cfg = G.base_layer_proof_config(log_n)
    n = 1 << log_n
    # every rank proves its own circuit instance (different seed = different witness of the same circuit type)
    wit, setup = PU.synth_trace(geo, seed=100 + rank, pinned=True)
    sd = PU.create_setup_data(ctx, geo, cfg, setup)
    del setup
    d_wit = torch.from_numpy(wit.view(np.int64)).to(dev)
    n_proof = PU.proof_size_u64(geo, cfg)
    proof = torch.empty(n_proof, dtype=torch.int64).pin_memory()
    proof_np = proof.numpy().view(np.uint64)
    from era_zkevm_test_harness_b200 import farm

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_proofs():
        # rank r proved instance r of this step; one NCCL gather of the fixed-size proof buffers to rank 0
        if distributed:
            return farm.gather_proofs({rank: proof_np}, n_proof, world, device=dev)
        return [proof_np]

    def step_device():
        PU.prove_circuit(ctx, sd, d_wit, proof_out=proof_np)
        gather_proofs()

    def step_e2e():
        PU.prove_circuit(ctx, sd, wit, proof_out=proof_np)   # pinned host witness -> H2D -> prove -> proof D2H
        gather_proofs()

    # the reference proves its circuits in a loop, so circuit k+1's witness is known while circuit k is proven: its upload is
    # started (zkgpu_witness_stage, second staging slot) before proof k runs.  Every step's H2D copy is issued inside the timed
    # region; the first step of a run uploads its own witness without overlap, the last one stages nothing.
    pipe = {"k": 0, "staged": False, "left": 0}

    def step_e2e_staged():
        k = pipe["k"]
        if not pipe["staged"]:
            PU.stage_witness(ctx, sd, wit, k % 2)
        pipe["left"] -= 1
        if pipe["left"] > 0:
            PU.stage_witness(ctx, sd, wit, (k + 1) % 2)   # the same pinned buffer stands in for the next circuit's witness
            pipe["staged"] = True
        else:
            pipe["staged"] = False
        PU.prove_staged(ctx, sd, k % 2, proof_out=proof_np)
        gather_proofs()
        pipe["k"] = k + 1

    per_step = []

    def timed(fn, steps, warmup, begin=None):
        if begin:
            begin(warmup)
        for _ in range(warmup):
            fn()
        barrier()
        if begin:
            begin(steps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e0.record()
        for i in range(steps):
            fn()
            marks[i].record()
        e1.record()
        barrier()
        per_step.append([round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(steps)])
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ctx.kernel_launches - l0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e_single, _ = timed(step_e2e, e2e_steps, 1)
    ms_e2e, _ = timed(step_e2e_staged, args.steps, 1, begin=lambda n: pipe.update(staged=False, left=n))

    # ---- sanity: the last proof verifies (rank 0; CPU verifier)
    verified = None
    if rank == 0:
        verified, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof_np)
        if not verified:
            raise SystemExit("bench: produced proof does not verify: " + msg)

    # ---- NTT roofline, measured live (rank 0): forward coset NTT of the W witness columns, 2 launches per transform
    roof = None
    extra = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        W = geo.n_witness
        x = d_wit
        out = torch.empty_like(x)
        for _ in range(3):
            ctx.ntt_forward(x, log_n, 7, out=out)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.ntt_forward(x, log_n, 7, out=out)   # 1.3 GB in + 1.3 GB out per call: larger than L2
        e1.record()
        torch.cuda.synchronize()
        ntt_ms = e0.elapsed_time(e1) / reps
        alg_bytes = 16.0 * n * W                      # 16*n bytes per size-n NTT per column
        achieved = alg_bytes / (ntt_ms * 1e-3) / 1e9
        traffic, traffic_src = ntt_traffic_from_profiles() if (log_n == 20 and W == 156) else (None, "not the captured shape")
        roof = {"bound": "hbm", "kernel": "zk::ntt1024_kernel (pass A strided columns + pass B rows = 2 launches per batched 2^20 coset NTT)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_transform_batch": alg_bytes, "ms_per_transform_batch": ntt_ms,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "note": "integer-issue-bound, not HBM-bound: no 64-bit multiplier on sm_100a (ncu: alu pipe 67-75 % busy, dram 18-27 %)"}
        # Poseidon2 throughput of the leaf kernel (the ALU-bound part): 2^(log_n+1) leaves x W columns
        lde = torch.empty((W, 2 * n), dtype=torch.int64, device=dev)
        lde[:, :n] = x; lde[:, n:] = out
        tree = torch.empty((4 * n - 16, 4), dtype=torch.int64, device=dev)
        ctx.merkle_build(lde, 2 * n, 1, 16, tree=tree)
        torch.cuda.synchronize()
        e0.record(); ctx.merkle_build(lde, 2 * n, 1, 16, tree=tree); e1.record()
        torch.cuda.synchronize()
        mk_ms = e0.elapsed_time(e1)
        perms = 2 * n * ((W + 7) // 8) + 2 * n
        extra = {"ntt_GBps": achieved, "merkle_commit_ms_W_cols": mk_ms, "poseidon2_perms_per_s": perms / mk_ms * 1e3}
        del lde, tree, out

    cb = None
    if rank == 0 and not args.skip_cpu_baseline and world == 1:
        cb = cpu_baseline(log_n, args.cpu_log_n)

    if rank == 0:
        sec_per_step = ms_dev / 1e3 / args.steps
        value = world / sec_per_step
        e2e_value = world / (ms_e2e / 1e3 / args.steps)
        e2e_single = world / (ms_e2e_single / 1e3 / e2e_steps)
        line = {
            "metric": "base_layer_proofs_per_sec_trace_2^20", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(log_n=log_n) + "; one instance per GPU per step",
                       "l2": "inputs larger than L2 (1.3 GB witness, 13 GB setup cosets per proof)", "proof_bytes": n_proof * 8,
                       "proof_verified_by_cpu_verifier": verified},
            "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": int(wit.nbytes) * world, "d2h_bytes_per_step": n_proof * 8 * world,
                    "host_cores_bound_per_rank": host_cores,
                    "steps": args.steps,
                    "mode": "staged: the upload of witness k+1 (zkgpu_witness_stage) overlaps proof k (zkgpu_prove_staged); every copy inside the timed region",
                    "single_call_value": e2e_single, "single_call_steps": e2e_steps,
                    "single_call_mode": "zkgpu_prove: chunked upload overlapped with the NTTs of the same proof"},
            "gpu_launches": launches, "ms_each_step": per_step, "clocks": clocks, "roofline": roof, "cpu_baseline": cb, "kernels": extra,
        }
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
