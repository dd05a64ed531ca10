"""Block workflow (block.py): plan shape against the reference's file naming and aggregation arity, and a world_size-2 gloo
run of a whole small block -- base -> leaf -> node -> scheduler -> compression -- with the CPU oracle standing in for the GPU
prover (tests only), every gathered proof checked by the verifier on rank 0."""
import json
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from era_zkevm_test_harness_b200 import block as B
from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import proof_format
from era_zkevm_test_harness_b200 import prover_utils as PU

FIXTURE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))


def _plan(base_instances, log_n=6):
    table, base_keys, leaf_keys, node_key, sched_key = B.circuit_table(FIXTURE, log_n=log_n, compression_log_n=log_n)
    return table, B.plan_block(base_instances, base_keys, leaf_keys, node_key, sched_key)


def test_plan_follows_reference_names_and_arity():
    table, plan = _plan({1: 70, 8: 33, 13: 1})
    names = [n for n, _ in plan.stages]
    assert names == ["base", "leaf", "node_depth_0", "scheduler", "compression_1", "compression_2", "compression_3", "compression_4"]
    stages = dict(plan.stages)
    assert len(stages["base"]) == 104 and stages["base"][0].file == "base_layer/basic_circuit_proof_1_0.json"
    leaf = stages["leaf"]
    # 70 MainVM proofs -> 3 leaf proofs of type 3 (32 + 32 + 6); 33 RAM -> 2 of type 10; 1 -> 1 of type 15
    assert [(j.numeric_type, len(j.children)) for j in leaf] == [(3, 32), (3, 32), (3, 6), (10, 32), (10, 1), (15, 1)]
    assert leaf[0].file == "recursion_layer/leaf_layer_proof_3_0.json"
    node = stages["node_depth_0"]
    assert [j.file for j in node] == [f"recursion_layer/node_layer_proof_{t}_0_0.json" for t in (3, 10, 15)]
    assert len(node[0].children) == 3
    assert stages["scheduler"][0].children == tuple(j.file for j in node)
    assert stages["compression_1"][0].children == ("recursion_layer/scheduler_proof.json",)
    assert all(j.geometry_key in table for _, js in plan.stages for j in js)
    # every file name of the reference's own test_proofs tree that this block would produce is spelled the same way
    assert os.path.basename(stages["base"][0].file) == "basic_circuit_proof_1_0.json"


def test_deep_node_tree():
    _, plan = _plan({1: 32 * 33})          # 1056 base -> 33 leaf -> 2 node (depth 0) -> 1 node (depth 1)
    names = [n for n, _ in plan.stages]
    assert names[:4] == ["base", "leaf", "node_depth_0", "node_depth_1"]
    stages = dict(plan.stages)
    assert len(stages["leaf"]) == 33 and len(stages["node_depth_0"]) == 2 and len(stages["node_depth_1"]) == 1
    assert stages["node_depth_1"][0].file == "recursion_layer/node_layer_proof_3_1_0.json"


def test_instances_share_a_setup_but_not_a_witness():
    geo = G.small_test_geometry(log_n=6, n_copy=16, lookup=True)
    w1, s1 = PU.synth_trace(geo, seed=7, witness_seed=100)
    w2, s2 = PU.synth_trace(geo, seed=7, witness_seed=200)
    assert (s1 == s2).all() and not (w1 == w2).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_lib
    oracle = oracle_lib.load()
    table, base_keys, leaf_keys, node_key, sched_key = B.circuit_table(FIXTURE, log_n=6, compression_log_n=6)
    # small security level so the CPU oracle finishes in seconds; compression modes keep their LDE factors / cap sizes
    def cfg_of(key):
        g, c = table[key]
        if key.startswith("compression_"):
            return c
        return G.make_proof_config(g.log_n, 2, 16, security_level=4)
    plan = B.plan_block({1: 5, 8: 2}, base_keys, leaf_keys, node_key, sched_key, arity=2, compression_modes=(1,))
    caps = {}

    def setup_cols(key):
        return PU.synth_trace(table[key][0], seed=77, witness_seed=77)[1]

    def prove(job, seed):
        geo = table[job.geometry_key][0]
        wit, setup = PU.synth_trace(geo, seed=77, witness_seed=seed)
        return oracle.prove(geo, cfg_of(job.geometry_key), wit, setup)

    def verify(job, proof):
        key = job.geometry_key
        if key not in caps:
            caps[key] = oracle.setup_cap(table[key][0], cfg_of(key), setup_cols(key))
        ok, msg = PU.verify_proof(table[key][0], cfg_of(key), caps[key], proof)
        assert ok, (job.file, msg)
        return ok

    res = B.prove_block(plan, prove, block_seed=5, out_dir=out_dir if rank == 0 else None, verify=verify if rank == 0 else None,
                        security_level=4)
    if rank == 0:
        assert len(res["proofs"]) == plan.n_jobs
        syn = os.path.join(out_dir, "synthetic")   # the reference's layout, under synthetic/ with a manifest (block.synthetic_root)
        assert os.path.exists(os.path.join(syn, "SYNTHETIC.json"))
        files = sorted(os.path.relpath(os.path.join(d, f), syn) for d, _, fs in os.walk(syn) for f in fs)
        assert "recursion_layer/scheduler_proof.json" in files and "aux_layer/compression_proof_1.json" in files
        assert "base_layer/basic_circuit_proof_1_4.json" in files and "recursion_layer/node_layer_proof_3_1_0.json" in files
        # the file is the reference's externally tagged JSON and loads back to the same flat proof
        flat, variant = proof_format.load_proof_json(os.path.join(syn, "base_layer/basic_circuit_proof_8_1.json"))
        assert variant == "RAMPermutation" and (flat == res["proofs"]["base_layer/basic_circuit_proof_8_1.json"]).all()
        # two instances of one circuit type: same VK, different proofs
        a, b = res["proofs"]["base_layer/basic_circuit_proof_1_0.json"], res["proofs"]["base_layer/basic_circuit_proof_1_1.json"]
        assert a.size == b.size and not (a == b).all()
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert res["proofs"] is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_block(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
