"""Block proving workflow over the GPU farm: the host-side mirror of the proving loops of the reference's `basic_test`.

Reference (`/root/reference/src/tests/complex_tests/mod.rs`):
  * base layer   :316-410   every circuit instance the scheduler emitted, proven independently
  * leaf layer   :560-660   per base circuit type, chunks of RECURSION_ARITY (32) base proofs -> one leaf proof each
                            (leaf circuit type = base type + 2, `base_circuit_type_into_recursive_leaf_circuit_type`)
  * node layer   :796-946   per circuit type, depth by depth, chunks of 32 -> one node proof, until one proof is left
  * scheduler    :1083-1140 one proof over the 13 per-type aggregates
  * compression  `src/proof_wrapper_utils/compression.rs:40-140`  modes 1..N chained on the scheduler proof
File names are the reference's (`src/data_source/local_file_data_source.rs:562-640`).

What runs where: the dependency structure, the job -> rank assignment, the per-stage gather to rank 0 and the file sink
are here; a stage's jobs are independent, so ranks never exchange trace data (farm.py).  What the reference does BETWEEN
the stages -- synthesising the recursive verifier circuit over the child proofs -- is host Rust and out of scope
(DESIGN.md section 0), so the witness of a recursion-layer job is a synthetic satisfying trace of that circuit type's
geometry whose seed is derived from the flat child proofs: stage k+1 still cannot start before stage k is gathered.

The prover is a callback (`prove(job, seed) -> flat proof u64[]`): the GPU prover in production (`GpuBlockProver`), the
CPU oracle in the gloo tests.
"""
import hashlib
import os
import time
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Tuple

import numpy as np

from . import farm
from . import geometry as G
from . import proof_format
from . import prover_utils as PU

RECURSION_ARITY = 32       # circuit_definitions RECURSION_ARITY (mod.rs:466, :679)
SCHEDULER_TYPE = 1         # ZkSyncRecursionLayerStorageType::SchedulerCircuit
NODE_TYPE = 2              # ZkSyncRecursionLayerStorageType::NodeLayerCircuit


def leaf_type_for_base(base_type: int) -> int:
    """base_circuit_type_into_recursive_leaf_circuit_type: VM (1) -> LeafLayerCircuitForMainVM (3), ..."""
    return base_type + 2


@dataclass
class Job:
    stage: str                 # "base" | "leaf" | "node" | "scheduler" | "compression"
    geometry_key: str          # key into the geometry table ("base_1_MainVM", "recursion_leaf_3", "recursion_node", ...)
    numeric_type: int          # the number in the reference's file name
    index: int
    depth: int = 0
    children: Tuple[str, ...] = ()   # file names of the proofs this job aggregates
    file: str = ""
    variant: str = ""          # enum variant used as the JSON tag


@dataclass
class BlockPlan:
    stages: List[Tuple[str, List[Job]]] = field(default_factory=list)

    @property
    def n_jobs(self):
        return sum(len(j) for _, j in self.stages)


def _chunks(n, k):
    return [(i, min(i + k, n)) for i in range(0, n, k)]


def plan_block(base_instances: Dict[int, int], base_keys: Dict[int, str], leaf_keys: Dict[int, str], node_key: str,
               scheduler_key: str, arity: int = RECURSION_ARITY, compression_modes=(1, 2, 3, 4)) -> BlockPlan:
    """base_instances: {base circuit type: instance count} as the circuit scheduler emitted them.
    *_keys map circuit types to geometry keys.  Returns the stage list in dependency order."""
    plan = BlockPlan()
    base_jobs, per_type_files = [], {}
    for t in sorted(base_instances):
        files = []
        for i in range(base_instances[t]):
            f = f"base_layer/basic_circuit_proof_{t}_{i}.json"
            base_jobs.append(Job("base", base_keys[t], t, i, file=f, variant=base_keys[t].split("_", 2)[2]))
            files.append(f)
        per_type_files[t] = files
    plan.stages.append(("base", base_jobs))

    leaf_jobs, current = [], {}
    for t, files in per_type_files.items():
        lt = leaf_type_for_base(t)
        outs = []
        for i, (a, b) in enumerate(_chunks(len(files), arity)):
            f = f"recursion_layer/leaf_layer_proof_{lt}_{i}.json"
            leaf_jobs.append(Job("leaf", leaf_keys[lt], lt, i, children=tuple(files[a:b]), file=f, variant="LeafLayerCircuit"))
            outs.append(f)
        current[lt] = outs
    plan.stages.append(("leaf", leaf_jobs))

    depth = 0
    while True:  # the reference always runs depth 0 (mod.rs:812-946), then continues while more than one proof is left
        node_jobs, nxt = [], {}
        for lt, files in current.items():
            outs = []
            for i, (a, b) in enumerate(_chunks(len(files), arity)):
                f = f"recursion_layer/node_layer_proof_{lt}_{depth}_{i}.json"
                node_jobs.append(Job("node", node_key, lt, i, depth=depth, children=tuple(files[a:b]), file=f, variant="NodeLayerCircuit"))
                outs.append(f)
            nxt[lt] = outs
        plan.stages.append((f"node_depth_{depth}", node_jobs))
        current = nxt
        depth += 1
        if all(len(v) == 1 for v in current.values()):
            break

    sched_children = tuple(current[lt][0] for lt in sorted(current))
    plan.stages.append(("scheduler", [Job("scheduler", scheduler_key, SCHEDULER_TYPE, 0, children=sched_children,
                                          file="recursion_layer/scheduler_proof.json", variant="SchedulerCircuit")]))
    prev = "recursion_layer/scheduler_proof.json"
    for m in compression_modes:
        f = f"aux_layer/compression_proof_{m}.json"
        plan.stages.append((f"compression_{m}", [Job("compression", f"compression_{m}", m, 0, children=(prev,), file=f,
                                                     variant=f"CompressionMode{m}Circuit")]))
        prev = f
    return plan


SYNTHETIC_NOTE = {
    "synthetic": True,
    "why": "Artefacts of this directory are NOT interchangeable with the reference's files of the same names: (1) hashing, Merkle "
           "trees, transcript, DEEP and FRI are pinned on the reference's golden proofs, but the gate polynomials / quotient term "
           "order of the un-vendored boojum crate are not (DESIGN.md section 5), so the reference verifier would reject the quotient "
           "identity; (2) witnesses and setup columns are synthetic satisfying traces of each circuit's geometry, and "
           "recursion / compression jobs do not verify their children in-circuit (the recursive-verifier synthesis is host Rust).",
}


def synthetic_root(out_dir):
    """Everything this package writes in the reference's file layout goes under <out_dir>/synthetic/, next to a manifest saying
    why it must not be mistaken for reference-compatible proofs or verification keys."""
    root = os.path.join(out_dir, "synthetic")
    os.makedirs(root, exist_ok=True)
    manifest = os.path.join(root, "SYNTHETIC.json")
    if not os.path.exists(manifest):
        import json
        with open(manifest, "w") as f:
            json.dump(SYNTHETIC_NOTE, f, indent=2)
    return root


def seed_for(job: Job, proofs: Dict[str, np.ndarray], block_seed: int) -> int:
    """Seed of a job's synthetic trace: the block seed and the job identity for base jobs, a digest of the flat child proofs
    for aggregation jobs (so the data dependency between stages is real)."""
    h = hashlib.blake2b(digest_size=8)
    h.update(f"{block_seed}|{job.file}".encode())
    for c in job.children:
        h.update(np.ascontiguousarray(proofs[c], dtype=np.uint64).tobytes())
    return int.from_bytes(h.digest(), "little") >> 1


def assign_jobs(jobs: List[Job], world: int, rank: int) -> List[int]:
    """Jobs of one stage, ordered by circuit type (as the plan emits them), dealt round-robin: few distinct types per rank."""
    return farm.assign_instances(len(jobs), world, rank)


def gather_stage(local: Dict[int, np.ndarray], n_jobs: int, device=None, group=None):
    """Proofs of one stage have different lengths per circuit type: pad to the stage maximum, one gather (farm.gather_proofs),
    trim on rank 0.  Slot layout: [length, proof...]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mx = max((p.size for p in local.values()), default=0)
    if world > 1:
        t = torch.tensor([mx], dtype=torch.int64, device=device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        mx = int(t.item())
    padded = {}
    for i, p in local.items():
        buf = np.zeros(mx + 1, dtype=np.uint64)
        buf[0] = p.size
        buf[1:1 + p.size] = p
        padded[i] = buf
    out = farm.gather_proofs(padded, mx + 1, n_jobs, device=device, group=group)
    if out is None:
        return None
    return [o[1:1 + int(o[0])].copy() for o in out]


def prove_block(plan: BlockPlan, prove: Callable[[Job, int], np.ndarray], block_seed: int = 0, out_dir: str = None,
                verify: Callable[[Job, np.ndarray], bool] = None, device=None, group=None, security_level=None,
                prefetch: Callable[[List[Tuple[Job, int]]], None] = None):
    """Runs the plan stage by stage.  Every rank calls this with the same plan.  Returns on rank 0 a dict
    {"proofs": {file: flat proof}, "stages": [{name, jobs, seconds}]}; other ranks get {"proofs": None, ...}.
    Rank 0 holds the gathered proofs of every finished stage and broadcasts the child proofs a stage needs for its seeds."""
    import torch
    import torch.distributed as dist
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    rank = dist.get_rank(group) if multi else 0
    proofs: Dict[str, np.ndarray] = {}
    report = []
    for name, jobs in plan.stages:
        t0 = time.time()
        # seeds are computed where the child proofs live (rank 0) and broadcast: 8 bytes per job instead of the proofs
        seeds = [seed_for(j, proofs, block_seed) for j in jobs] if rank == 0 else [0] * len(jobs)
        if multi:
            st = torch.tensor(seeds, dtype=torch.int64, device=device if device is not None else "cpu")
            dist.broadcast(st, src=0, group=group)
            seeds = [int(x) for x in st.cpu().tolist()]
        local = {}
        mine = assign_jobs(jobs, world, rank)
        if prefetch is not None:   # lets the prover start producing this rank's witnesses on host threads
            prefetch([(jobs[i], seeds[i]) for i in mine])
        for i in mine:
            local[i] = np.ascontiguousarray(prove(jobs[i], seeds[i]), dtype=np.uint64)
        gathered = gather_stage(local, len(jobs), device=device, group=group)
        failed = ""
        if rank == 0:
            for j, p in zip(jobs, gathered):
                if verify is not None and not verify(j, p):
                    failed = j.file
                    break
                proofs[j.file] = p
                if out_dir is not None:
                    path = os.path.join(synthetic_root(out_dir), j.file)
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    proof_format.save_proof_json(path, p, j.variant, security_level=security_level)
        if multi:   # every rank learns about a failed verification before anyone raises: nobody is left in the next collective
            flag = torch.tensor([1 if failed else 0], dtype=torch.int64, device=device if device is not None else "cpu")
            dist.broadcast(flag, src=0, group=group)
            if int(flag.item()):
                raise RuntimeError(f"block: a proof of stage {name} does not verify" + (f" ({failed})" if failed else ""))
        elif failed:
            raise RuntimeError(f"block: proof {failed} does not verify")
        report.append({"stage": name, "jobs": len(jobs), "seconds": time.time() - t0})
    return {"proofs": proofs if rank == 0 else None, "stages": report}


class GpuBlockProver:
    """prove/verify callbacks over one GpuContext with an LRU of resident setups (a 2^20 setup is ~13 GB of HBM: cosets,
    monomials, values, tree; `max_resident` bounds how many circuit types stay on the GPU).  Setup columns of a circuit type
    come from `setup_source(geometry_key) -> (Geometry, ProofConfig, setup_cols)`; the witness of a job from
    `witness_source(job, geo, seed) -> witness columns` (synthetic traces by default)."""

    def __init__(self, ctx, circuits: Dict[str, Tuple[G.Geometry, G.ProofConfig]], max_resident: int = 6, setup_seed: int = 77,
                 host_threads: int = 8):
        self.ctx, self.circuits, self.max_resident, self.setup_seed = ctx, circuits, max_resident, setup_seed
        self.resident: "OrderedDict[str, PU.SetupData]" = OrderedDict()
        self.vk_caps: Dict[str, np.ndarray] = {}
        self.seconds = {"synth_trace": 0.0, "setup": 0.0, "prove": 0.0}
        self.pool = ThreadPoolExecutor(max_workers=host_threads) if host_threads > 1 else None
        # witness generation takes a pinned buffer each: two generator threads keep one buffer free for the proof in flight
        self.gen_pool = ThreadPoolExecutor(max_workers=2) if host_threads > 1 else None
        self.pinned, self.staged, self.order, self.next_slot = None, {}, [], 0
        self.pending = {}   # (geometry key, witness seed) -> future of (pinned buffer, witness columns)
        self.pending_setup = {}   # geometry key -> future of the setup columns
        self.prove_ms = {}        # file -> wall ms of the prove call alone (host witness in, proof out)

    def _trace(self, key, witness_seed):
        # setup_seed fixes the circuit TYPE (its setup columns / VK), witness_seed the instance (zkgpu_synth_trace_instance)
        geo, _ = self.circuits[key]
        t0 = time.time()
        out = PU.synth_trace(geo, seed=self.setup_seed, witness_seed=witness_seed)
        self.seconds["synth_trace"] += time.time() - t0
        return out

    def setup(self, key) -> PU.SetupData:
        if key in self.resident:
            self.resident.move_to_end(key)
            return self.resident[key]
        while len(self.resident) >= self.max_resident:
            _, old = self.resident.popitem(last=False)
            old.close()
        geo, cfg = self.circuits[key]
        fut = self.pending_setup.pop(key, None)
        if fut is not None:
            t0 = time.time()
            setup_cols = fut.result()
            self.seconds["synth_trace"] += time.time() - t0
        else:
            _, setup_cols = self._trace(key, self.setup_seed)
        t0 = time.time()
        sd = PU.create_setup_data(self.ctx, geo, cfg, setup_cols)
        self.seconds["setup"] += time.time() - t0
        self.resident[key] = sd
        self.vk_caps[key] = sd.vk_cap.copy()
        return sd

    def _pinned(self):
        if self.pinned is None:   # three buffers of the widest witness: one being proven, one staged, one being generated
            words = max(g.n_witness << g.log_n for g, _ in self.circuits.values())
            self.pinned = PU.PinnedPool(3, words)
        return self.pinned

    def _generate(self, geo, seed):
        buf = self._pinned().take()
        try:
            wit, _ = PU.synth_trace(geo, seed=self.setup_seed, witness_seed=seed, out_witness=buf)
        except Exception:
            self.pinned.give(buf)
            raise
        return buf, wit

    def prefetch(self, jobs_and_seeds):
        """starts generating the witnesses of the given jobs, in proving order, on host threads into pinned buffers
        (zkgpu_synth_trace_instance releases the GIL); the order is also the look-ahead order of `prove`"""
        self.order = [(job.geometry_key, seed) for job, seed in jobs_and_seeds]
        if self.pool is None:
            return
        for job, seed in jobs_and_seeds:
            geo = self.circuits[job.geometry_key][0]
            if job.geometry_key not in self.resident and job.geometry_key not in self.pending_setup:
                self.pending_setup[job.geometry_key] = self.pool.submit(
                    lambda g=geo: PU.synth_trace(g, seed=self.setup_seed, witness_seed=self.setup_seed)[1])
            self.pending[(job.geometry_key, seed)] = self.gen_pool.submit(self._generate, geo, seed)

    def _stage(self, key, seed, slot, wait):
        """uploads the witness of (key, seed) into a staging slot if its generation is finished (or `wait`); -> staged entry"""
        fut = self.pending.get((key, seed))
        if fut is None or not (wait or fut.done()) or key not in self.resident:
            return None
        t0 = time.time()
        buf, wit = fut.result()
        self.seconds["synth_trace"] += time.time() - t0   # only the time the GPU actually waited
        del self.pending[(key, seed)]
        PU.stage_witness(self.ctx, self.resident[key], wit, slot)
        self.staged[(key, seed)] = (slot, buf)
        return self.staged[(key, seed)]

    def prove(self, job: Job, seed: int) -> np.ndarray:
        """Pinned witness -> zkgpu_witness_stage -> zkgpu_prove_staged; the NEXT job's witness (prefetch order) is staged into the
        other slot before this proof starts, so its upload runs under this proof (the bench's staged mode, bench.py)."""
        key = job.geometry_key
        sd = self.setup(key)
        me = (key, seed)
        if me not in self.staged:
            if me not in self.pending:     # not prefetched: generate now, on this thread
                from concurrent.futures import Future
                self.pending[me] = Future()
                self.pending[me].set_result(self._generate(self.circuits[key][0], seed))
            self._stage(key, seed, self.next_slot, wait=True)
            self.next_slot ^= 1
        slot, buf = self.staged.pop(me)
        if me in self.order:               # look ahead: stage the following job if its witness and setup are ready
            i = self.order.index(me)
            if i + 1 < len(self.order) and self.order[i + 1] not in self.staged:
                if self._stage(*self.order[i + 1], slot ^ 1, wait=False) is not None:
                    self.next_slot = slot
        t0 = time.time()
        proof = PU.prove_staged(self.ctx, sd, slot)
        dt = time.time() - t0
        self.pinned.give(buf)
        self.seconds["prove"] += dt
        self.prove_ms[job.file] = round(1e3 * dt, 1)
        return proof

    def close(self):
        if self.gen_pool is not None:
            self.gen_pool.shutdown(wait=True, cancel_futures=True)
        if self.pool is not None:
            self.pool.shutdown(wait=True, cancel_futures=True)
        for sd in self.resident.values():
            sd.close()
        self.resident.clear()
        if self.pinned is not None:
            self.pinned.close()
            self.pinned = None


def circuit_table(fixture, log_n=None, compression_log_n=None):
    """{geometry key: (Geometry, ProofConfig)} for every job kind of a block: the 13 base circuits, 13 leaf circuits
    (all share the leaf geometry of the fixture), node, scheduler, compression modes 1..4 (geometry and proof config from
    compression_N_vk.json, geometry.compression_geometries_from_fixture).  Also returns the key maps plan_block needs."""
    table, base_keys, leaf_keys = {}, {}, {}
    rec = {}
    for key, geo, _ in G.circuit_geometries_from_fixture(fixture):
        g = geo.scaled(log_n) if (log_n is not None and log_n != geo.log_n) else geo
        cfg = G.base_layer_proof_config(g.log_n)
        if key.startswith("base_"):
            t = int(key.split("_")[1])
            base_keys[t] = key
            table[key] = (g, cfg)
        else:
            rec[key] = (g, cfg)
    leaf_src = next(k for k in rec if "leaf" in k)
    node_key = next(k for k in rec if "node" in k)
    sched_key = next(k for k in rec if "scheduler" in k)
    table[node_key], table[sched_key] = rec[node_key], rec[sched_key]
    for t in base_keys:
        lt = leaf_type_for_base(t)
        k = f"recursion_leaf_{lt}"
        table[k] = rec[leaf_src]
        leaf_keys[lt] = k
    for key, geo, cfg, _ in G.compression_geometries_from_fixture(fixture):
        if "for_wrapper" in key:
            continue
        if compression_log_n is not None and compression_log_n != geo.log_n:
            geo = geo.scaled(compression_log_n)
            cfg = G.make_proof_config(compression_log_n, 1 << cfg.log_lde, cfg.cap_size, security_level=cfg.n_queries * cfg.log_lde)
        table[key] = (geo, cfg)
    return table, base_keys, leaf_keys, node_key, sched_key
