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
# compression modes 1-4 with their own proof configs (LDE 32 / 512 / 1024 / 2048, cap 16 / 256) on a scaled trace
COMPRESSION = [(k, g, c) for k, g, c, _ in G.compression_geometries_from_fixture(FIXTURE) if "for_wrapper" not in k]


@pytest.mark.parametrize("key,geo", CIRCUITS, ids=IDS)
def test_oracle_proves_and_verifies_every_circuit(oracle, key, geo):
    g = geo.scaled(8)
    cfg = G.make_proof_config(8, 2, 16, security_level=6)
    wit, setup = PU.synth_trace(g, seed=5)
    proof = oracle.prove(g, cfg, wit, setup)
    cap = oracle.setup_cap(g, cfg, setup)
    ok, msg = PU.verify_proof(g, cfg, cap, proof)
    assert ok, msg
    ok, msg = oracle.verify(g, cfg, cap, proof)   # two independent acceptance checks: product verifier and oracle verifier
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
    ok, msg = oracle.verify(g, cfg, sd.vk_cap, proof)
    assert ok, msg
    sd.close()


def _compression_case(geo, cfg, log_n):
    g = geo.scaled(log_n)
    c = G.make_proof_config(log_n, 1 << cfg.log_lde, cfg.cap_size, security_level=2 * cfg.log_lde)   # 2 queries
    return g, c


@pytest.mark.parametrize("key,geo,cfg", COMPRESSION, ids=[c[0] for c in COMPRESSION])
def test_oracle_proves_and_verifies_compression_circuits(oracle, key, geo, cfg):
    g, c = _compression_case(geo, cfg, 5)
    wit, setup = PU.synth_trace(g, seed=9)
    assert wit.shape[0] == g.n_witness and g.n_witness == g.n_perm + g.n_witness_plain
    proof = oracle.prove(g, c, wit, setup)
    cap = oracle.setup_cap(g, c, setup)
    ok, msg = PU.verify_proof(g, c, cap, proof)
    assert ok, msg
    ok, msg = oracle.verify(g, c, cap, proof)
    assert ok, msg
    # a plain witness cell is covered by the gate relations: corrupting one breaks the quotient identity
    if g.n_witness_plain:
        wit2 = wit.copy()
        wit2[g.n_perm + 3, 0] ^= 1   # row 0 carries gate 0..; pick the row of the flattened Poseidon2 gate below
        p2 = [i for i in range(g.n_gates) if g.gates[i].kind == G.GATE_POSEIDON2_FLATTENED][0]
        wit2 = wit.copy()
        wit2[g.n_perm + 3, p2] ^= 1   # synthetic traces put gate i on rows r = i mod n_gates
        bad = oracle.prove(g, c, wit2, setup)
        ok2, msg2 = PU.verify_proof(g, c, cap, bad)
        assert not ok2 and "quotient" in msg2
        ok3, msg3 = oracle.verify(g, c, cap, bad)
        assert not ok3 and "quotient" in msg3


@pytest.mark.gpu
@pytest.mark.parametrize("key,geo,cfg", COMPRESSION, ids=[c[0] for c in COMPRESSION])
def test_gpu_proof_of_compression_circuits_is_bit_identical_to_oracle(gpu, oracle, key, geo, cfg):
    g, c = _compression_case(geo, cfg, 6)
    wit, setup = PU.synth_trace(g, seed=13)
    sd = PU.create_setup_data(gpu, g, c, setup)
    assert (sd.vk_cap == oracle.setup_cap(g, c, setup)).all()
    proof = PU.prove_circuit(gpu, sd, wit)
    ref = oracle.prove(g, c, wit, setup)
    diff = np.nonzero(proof != ref)[0]
    assert diff.size == 0, f"first differing u64 at {int(diff[0])}"
    ok, msg = PU.verify_proof(g, c, sd.vk_cap, proof)
    assert ok, msg
    sd.close()
