"""CPU: the oracle's whole-prover restatement produces proofs the product's CPU verifier accepts; corrupted proofs are
rejected (mirrors the reference's negative tests, src/tests/complex_tests/wrapper_negative_tests.rs:112-206)."""
import numpy as np
import pytest

from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import prover_utils as PU


def _prove_small(oracle, geo, cfg, seed=1):
    wit, setup = PU.synth_trace(geo, seed)
    vk_cap = oracle.setup_cap(geo, cfg, setup)
    proof = oracle.prove(geo, cfg, wit, setup)
    return wit, setup, vk_cap, proof


@pytest.fixture(scope="module")
def small(oracle):
    geo = G.small_test_geometry(log_n=8, n_copy=16, lookup=True)
    cfg = G.make_proof_config(8, 2, 4, security_level=12)
    return (geo, cfg) + _prove_small(oracle, geo, cfg)


def test_shapes_match_reference_tables():
    g = G.mainvm_like_geometry(20)
    # SURVEY.md section 8a row 1 (measured from the golden proofs): W=156, S2=58, Q=16, S=167, values_at_z 360 / 1 / 9
    assert (g.n_witness, g.n_stage2, g.n_quotient, g.n_setup) == (156, 58, 16, 167)
    cfg = G.base_layer_proof_config(20)
    assert cfg.n_queries == 100 and list(cfg.fri_schedule[:cfg.n_fri_oracles]) == [3, 3, 3, 3, 3, 2]
    cols = PU.num_columns(g)
    assert cols == dict(witness=156, permuted=155, setup=167, stage2=58, quotient=16)


def test_oracle_proof_verifies(oracle, small):
    geo, cfg, wit, setup, vk_cap, proof = small
    assert proof.size == PU.proof_size_u64(geo, cfg)
    ok, msg = PU.verify_proof(geo, cfg, vk_cap, proof)
    assert ok, msg
    ok, msg = oracle.verify(geo, cfg, vk_cap, proof)     # the oracle's own, independently written verifier
    assert ok, msg


def test_no_lookup_geometry_verifies(oracle):
    geo = G.small_test_geometry(log_n=7, n_copy=20, lookup=False)
    cfg = G.make_proof_config(7, 4, 8, security_level=10)
    _, _, vk_cap, proof = _prove_small(oracle, geo, cfg, seed=3)
    ok, msg = PU.verify_proof(geo, cfg, vk_cap, proof)
    assert ok, msg


def test_unsatisfied_witness_is_rejected(oracle, small, capfd):
    geo, cfg, wit, setup, vk_cap, _ = small
    bad = wit.copy()
    bad[5, 17] ^= np.uint64(1)  # break one gate relation / copy constraint
    proof = oracle.prove(geo, cfg, bad, setup)
    ok, msg = PU.verify_proof(geo, cfg, vk_cap, proof)
    assert not ok
    ok, msg = oracle.verify(geo, cfg, vk_cap, proof)
    assert not ok and "quotient" in msg


@pytest.mark.parametrize("what", ["public_input", "witness_cap", "values_at_z", "fri_leaf", "final_monomial", "vk_cap", "query_leaf"])
def test_corrupted_proof_is_rejected(oracle, small, what):
    geo, cfg, wit, setup, vk_cap, proof = small
    p = proof.copy()
    cap = vk_cap.copy()
    n_pi, c4 = geo.n_public_inputs, cfg.cap_size * 4
    off_pi = 32
    off_wcap = off_pi + n_pi
    n_final = int(p[14])
    off_final = off_wcap + 3 * c4
    off_at_z = off_final + 2 * n_final
    if what == "public_input":
        p[off_pi] = (int(p[off_pi]) + 1) % PU_P
    elif what == "witness_cap":
        p[off_wcap + 1] ^= np.uint64(1)
    elif what == "values_at_z":
        p[off_at_z + 6] = (int(p[off_at_z + 6]) + 1) % PU_P
    elif what == "final_monomial":
        p[off_final] = (int(p[off_final]) + 1) % PU_P
    elif what == "vk_cap":
        cap[0, 0] ^= np.uint64(1)
    elif what == "fri_leaf":
        p[p.size - 2 - 4 * 0 - 3] ^= np.uint64(1)  # inside the last query's last FRI leaf / path
    elif what == "query_leaf":
        n_at = int(p[10]) + int(p[11]) + int(p[12])
        off_q = off_at_z + 2 * n_at
        for k in range(cfg.n_fri_oracles):
            leaves = (1 << (geo.log_n + cfg.log_lde)) >> sum(cfg.fri_schedule[:k + 1])
            off_q += min(cfg.cap_size, leaves) * 4
        p[off_q + 3] ^= np.uint64(1)
    ok, msg = PU.verify_proof(geo, cfg, cap, p)
    assert not ok and msg
    ok2, msg2 = oracle.verify(geo, cfg, cap, p)          # both verifiers reject every corruption of wrapper_negative_tests.rs:112-206
    assert not ok2 and msg2


PU_P = (1 << 64) - (1 << 32) + 1


def test_malformed_inputs_raise(small):
    geo, cfg, wit, setup, vk_cap, proof = small
    ok, msg = PU.verify_proof(geo, cfg, vk_cap, proof[:-1])
    assert not ok and "length" in msg
    bad = G.Geometry.from_buffer_copy(bytes(geo))
    bad.quotient_degree = 3
    with pytest.raises(Exception):
        PU.proof_size_u64(bad, cfg)


@pytest.mark.parametrize("log_n", [9])
def test_mainvm_gate_set_small_trace(oracle, log_n):
    """all 11 MainVM gates incl. the degree-7 flattened Poseidon2 gate, 130 copy columns, lookup 3x8, on a short trace"""
    geo = G.mainvm_like_geometry(log_n)
    cfg = G.make_proof_config(log_n, 2, 16, security_level=8)
    _, _, vk_cap, proof = _prove_small(oracle, geo, cfg, seed=5)
    ok, msg = PU.verify_proof(geo, cfg, vk_cap, proof)
    assert ok, msg


@pytest.mark.parametrize("mode,queries", [(1, 16), (2, 9), (3, 8), (4, 8)])
def test_compression_mode_proof_configs(oracle, mode, queries):
    """proof configs of the compression chain (aux_layer/compression_modes/mode_N.rs): query counts as in the reference's
    one-shot compression proofs (16/9/8/8), and the oracle prover + CPU verifier handle LDE 32..2048, cap 256"""
    assert G.compression_layer_proof_config(mode).n_queries == queries
    log_n = 6
    geo = G.small_test_geometry(log_n, 16, lookup=(mode == 1))
    cfg = G.compression_layer_proof_config(mode, log_n)
    wit, setup = PU.synth_trace(geo, seed=mode)
    proof = oracle.prove(geo, cfg, wit, setup)
    ok, msg = PU.verify_proof(geo, cfg, oracle.setup_cap(geo, cfg, setup), proof)
    assert ok, msg


@pytest.mark.parametrize("kind", ["mainvm", "small_nolookup", "compression_1", "compression_4", "storage_application"])
def test_oracle_trace_generator_matches_product_generator(oracle, kind):
    """oracle/synth.c (used by bench.py's CPU reference arm, which must not load libzkgpu.so) and the product's host-side
    zkgpu_synth_trace produce the same satisfying trace, bit for bit, for both seeding modes."""
    import json, os
    fixture = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
    if kind == "mainvm":
        geo = G.mainvm_like_geometry(9)
    elif kind == "small_nolookup":
        geo = G.small_test_geometry(log_n=7, lookup=False)
    elif kind == "storage_application":
        geo = [g for k, g, _ in G.circuit_geometries_from_fixture(fixture) if k.startswith("base_10_")][0]
        geo.log_n, geo.table_len = 8, min(geo.table_len, 1 << 8)
    else:
        geo = [g for k, g, _c, _ in G.compression_geometries_from_fixture(fixture) if k == kind][0]
        geo.log_n = 7
    for i in range(geo.n_public_inputs):
        geo.pi_row[i] = 5
    for seed, wseed in ((3, None), (3, 11)):
        w0, s0 = PU.synth_trace(geo, seed=seed, witness_seed=wseed)
        w1, s1 = oracle.synth_trace(geo, seed=seed, witness_seed=wseed)
        assert np.array_equal(w0, w1) and np.array_equal(s0, s1)
