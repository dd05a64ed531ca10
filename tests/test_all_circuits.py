"""All 16 circuit types of the reference (geometry + gate set read from its verification keys, tests/golden/vk_shapes.json)
on a scaled trace: CPU -- the oracle's proof verifies; GPU -- libzkgpu's proof is byte-identical to the oracle's and verifies.
Covers BASELINE configs 3 and 4 (full base-layer set; leaf / node / scheduler recursion circuits) at parity-test size."""
import json
import os

import numpy as np
import pytest

from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import prover_utils as PU

FIXTURE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
CIRCUITS = [(k, g) for k, g, _ in G.circuit_geometries_from_fixture(FIXTURE)]
IDS = [c[0] for c in CIRCUITS]


@pytest.mark.parametrize("key,geo", CIRCUITS, ids=IDS)
def test_oracle_proves_and_verifies_every_circuit(oracle, key, geo):
    g = geo.scaled(8)
    cfg = G.make_proof_config(8, 2, 16, security_level=6)
    wit, setup = PU.synth_trace(g, seed=5)
    proof = oracle.prove(g, cfg, wit, setup)
    ok, msg = PU.verify_proof(g, cfg, oracle.setup_cap(g, cfg, setup), proof)
    assert ok, msg


@pytest.mark.gpu
@pytest.mark.parametrize("key,geo", CIRCUITS, ids=IDS)
def test_gpu_proof_of_every_circuit_is_bit_identical_to_oracle(gpu, oracle, key, geo):
    log_n = 11 if key.startswith("base_1_") or key.startswith("recursion") else 10
    g = geo.scaled(log_n)
    cfg = G.make_proof_config(log_n, 2, 16, security_level=12)
    wit, setup = PU.synth_trace(g, seed=31)
    sd = PU.create_setup_data(gpu, g, cfg, setup)
    assert (sd.vk_cap == oracle.setup_cap(g, cfg, setup)).all()
    proof = PU.prove_circuit(gpu, sd, wit)
    ref = oracle.prove(g, cfg, wit, setup)
    diff = np.nonzero(proof != ref)[0]
    assert diff.size == 0, f"first differing u64 at {int(diff[0])}"
    ok, msg = PU.verify_proof(g, cfg, sd.vk_cap, proof)
    assert ok, msg
    sd.close()
