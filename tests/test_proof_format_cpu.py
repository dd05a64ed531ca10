"""The reference's on-disk proof format (serde-JSON of boojum's Proof behind an externally tagged enum,
src/data_source/local_file_data_source.rs:51-55): golden files load into the flat buffer and come back identical; proofs
produced by this framework serialise to the same structure."""
import json
import os

import numpy as np
import pytest

from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import proof_format as PF
from era_zkevm_test_harness_b200 import prover_utils as PU

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["proof_mainvm_1_0_2q.json", "proof_node_3_0_0_2q.json"])
def test_golden_proof_round_trips_through_flat_buffer(name):
    d = json.load(open(os.path.join(GOLDEN, name)))
    flat, variant = PF.proof_from_dict(d)
    assert variant in ("MainVM", "NodeLayerCircuit")
    hd = PF.parse_header(flat)
    assert hd["log_n"] == 20 and hd["schedule"] == [3, 3, 3, 3, 3, 2] and hd["cap"] == 16
    back = PF.proof_to_dict(flat, variant, security_level=d[variant]["proof_config"]["security_level"])
    assert back == d
    # the flat length is what the C ABI computes for that circuit, up to the number of queries kept in the fixture
    fx = json.load(open(os.path.join(GOLDEN, "vk_shapes.json")))
    entry = fx["base"]["1"] if variant == "MainVM" else fx["recursion"]["node"]
    geo = G.geometry_from_vk(entry, G.BASE_LAYER_GATE_ORDER[1] if variant == "MainVM" else G.RECURSION_GATE_ORDER)
    cfg = G.base_layer_proof_config(20)
    cfg.n_queries = len(d[variant]["queries_per_fri_repetition"])
    assert flat.size == PU.proof_size_u64(geo, cfg)


def test_produced_proof_serialises_like_the_reference(oracle, tmp_path):
    geo = G.small_test_geometry(8, 16, True)
    cfg = G.make_proof_config(8, 2, 4, security_level=12)
    wit, setup = PU.synth_trace(geo, seed=9)
    proof = oracle.prove(geo, cfg, wit, setup)
    path = tmp_path / "basic_circuit_proof_1_0.json"
    PF.save_proof_json(path, proof, "MainVM", security_level=12)
    d = json.load(open(path))
    golden = json.load(open(os.path.join(GOLDEN, "proof_mainvm_1_0_2q.json")))
    assert list(d) == ["MainVM"] and list(d["MainVM"]) == list(golden["MainVM"])          # same fields, same order
    assert list(d["MainVM"]["queries_per_fri_repetition"][0]) == list(golden["MainVM"]["queries_per_fri_repetition"][0])
    flat, variant = PF.load_proof_json(path)
    assert variant == "MainVM" and (flat == proof).all()
    ok, msg = PU.verify_proof(geo, cfg, oracle.setup_cap(geo, cfg, setup), flat)
    assert ok, msg
