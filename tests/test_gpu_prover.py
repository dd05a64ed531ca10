"""GPU parity of the whole proving path through the C ABI: proofs are byte-identical to the CPU oracle's
(oracle/prover.c) on the same seeded trace, and verify under the CPU verifier."""
import numpy as np
import pytest
import torch

from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import prover_utils as PU

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    d = np.nonzero(a != b)[0]
    return None if d.size == 0 else int(d[0])


CASES = {
    "small_lookup": lambda: (G.small_test_geometry(8, 16, True), G.make_proof_config(8, 2, 4, security_level=12)),
    "small_nolookup_lde4": lambda: (G.small_test_geometry(7, 20, False), G.make_proof_config(7, 4, 8, security_level=10)),
    "small_two_pass_ntt": lambda: (G.small_test_geometry(12, 16, True), G.make_proof_config(12, 2, 16, security_level=10)),
    "mainvm_gates_2^9": lambda: (G.mainvm_like_geometry(9), G.make_proof_config(9, 2, 16, security_level=8)),
    "mainvm_gates_2^12": lambda: (G.mainvm_like_geometry(12), G.make_proof_config(12, 2, 16, security_level=20)),
}


@pytest.mark.parametrize("case", list(CASES))
def test_gpu_proof_is_bit_identical_to_oracle(gpu, oracle, case):
    geo, cfg = CASES[case]()
    wit, setup = PU.synth_trace(geo, seed=11)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    assert (sd.vk_cap == oracle.setup_cap(geo, cfg, setup)).all(), "verification key cap differs from the oracle"
    proof = PU.prove_circuit(gpu, sd, wit)
    ref = oracle.prove(geo, cfg, wit, setup)
    assert proof.size == ref.size
    assert _first_diff(proof, ref) is None, f"first differing u64 at {_first_diff(proof, ref)}"
    ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
    assert ok, msg
    # device-resident witness entry point gives the same bytes
    d_wit = torch.from_numpy(wit.view(np.int64)).to(gpu.device)
    assert (PU.prove_circuit(gpu, sd, d_wit) == proof).all()
    sd.close()


def test_unsatisfied_trace_gives_rejected_proof(gpu, oracle):
    """like boojum, the prover does not check satisfiability: a broken trace still yields a (bit-identical) proof, which the
    verifier rejects at the quotient identity"""
    geo, cfg = CASES["small_lookup"]()
    wit, setup = PU.synth_trace(geo, seed=2)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    wit[3, 5] ^= np.uint64(1)
    proof = PU.prove_circuit(gpu, sd, wit)
    assert (proof == oracle.prove(geo, cfg, wit, setup)).all()
    ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
    assert not ok and "quotient" in msg
    sd.close()


def test_full_size_mainvm_proof_verifies(gpu):
    """BASELINE config 2 shape: MainVM geometry (W=156, S2=58, Q=16, S=167), trace 2^20, lde 2, cap 16, 100 queries.
    The oracle needs minutes at this size, so the check is the size-independent one: the proof verifies (quotient
    identity at z, 100 x 4 Merkle openings, DEEP consistency, full FRI chain, final polynomial)."""
    geo = G.mainvm_like_geometry(20)
    cfg = G.base_layer_proof_config(20)
    wit, setup = PU.synth_trace(geo, seed=20)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    proof = PU.prove_circuit(gpu, sd, wit)
    ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
    assert ok, msg
    bad = proof.copy()
    bad[32 + geo.n_public_inputs + 5] ^= np.uint64(1)
    ok, _ = PU.verify_proof(geo, cfg, sd.vk_cap, bad)
    assert not ok
    sd.close()
