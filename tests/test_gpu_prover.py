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
    "compression1_cfg_lde32": lambda: (G.small_test_geometry(9, 16, True), G.compression_layer_proof_config(1, 9)),
    "compression4_cfg_lde2048_cap256": lambda: (G.small_test_geometry(6, 16, False), G.compression_layer_proof_config(4, 6)),
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


def test_full_size_mainvm_proof_equals_oracle(gpu):
    """The headline configuration end to end, every u64: MainVM geometry, trace 2^20, lde 2, cap 16, 100 queries, trace seed 1.
    The CPU oracle's proof of exactly this input is cached in tests/golden/oracle_proof_mainvm_2pow20_seed1.npy
    (tools/make_oracle_fullsize_fixture.py: ~8 min of oracle/prover.c on 8 cores); the CUDA prover must reproduce all 93 069
    words -- this is the only size that runs ntt1024_kernel, the 17 GiB arena, size_t-wide indices and the staged upload."""
    import os
    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_proof_mainvm_2pow20_seed1.npy"))
    geo = G.mainvm_like_geometry(20)
    cfg = G.base_layer_proof_config(20)
    wit, setup = PU.synth_trace(geo, seed=1, pinned=True)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    proof = PU.prove_circuit(gpu, sd, wit)
    assert proof.size == ref.size
    assert _first_diff(proof, ref) is None, f"first differing u64 at {_first_diff(proof, ref)}"
    # the staged (double-buffered upload) entry point and the device-resident one give the same bytes
    PU.stage_witness(gpu, sd, wit, 0)
    assert (PU.prove_staged(gpu, sd, 0) == ref).all()
    d_wit = torch.from_numpy(wit.view(np.int64)).to(gpu.device)
    assert (PU.prove_circuit(gpu, sd, d_wit) == ref).all()
    sd.close()


def test_compression_mode_1_reference_size_equals_oracle(gpu):
    """Compression mode 1 at its reference size (2^16 rows x LDE 32, cap 16, 16 queries; plain witness columns, FRI schedule
    3,3,3,3,3,1): every u64 of the CUDA proof equals the cached oracle proof (tests/golden/oracle_proof_compression_1_seed1.npy)."""
    import json
    import os
    golden = os.path.join(os.path.dirname(__file__), "golden")
    ref = np.load(os.path.join(golden, "oracle_proof_compression_1_seed1.npy"))
    fixture = json.load(open(os.path.join(golden, "vk_shapes.json")))
    geo, cfg = [(g, c) for k, g, c, _ in G.compression_geometries_from_fixture(fixture) if k == "compression_1"][0]
    wit, setup = PU.synth_trace(geo, seed=1)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    proof = PU.prove_circuit(gpu, sd, wit)
    assert proof.size == ref.size
    assert _first_diff(proof, ref) is None, f"first differing u64 at {_first_diff(proof, ref)}"
    ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, proof)
    assert ok, msg
    sd.close()


def test_prove_from_variables_matches_column_path(gpu, oracle):
    """the reference's hand-off: variable values + per-type variable maps (DenseVariablesCopyHint) instead of materialised
    columns.  A random assignment of cells to variables (with repeated variables = copy constraints and placeholders for zero
    cells) must give the same proof bytes as proving the columns directly."""
    geo, cfg = CASES["mainvm_gates_2^9"]()
    wit, setup = PU.synth_trace(geo, seed=17)
    n, npm = 1 << geo.log_n, geo.n_perm
    cells = wit[:npm].reshape(-1)
    # variables = distinct cell values (equal cells share a variable, like copy-constrained cells do); zero cells -> placeholder
    uniq, inverse = np.unique(cells, return_inverse=True)
    rng = np.random.default_rng(5)
    perm = rng.permutation(uniq.size)
    values = np.empty_like(uniq)
    values[perm] = uniq
    maps = perm[inverse].astype(np.uint32)
    maps[cells == 0] = np.uint32(0xFFFFFFFF)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    PU.set_variable_maps(gpu, sd, maps.reshape(npm, n))
    mult = wit[geo.n_witness - 1] if geo.lookup_reps else None
    got = PU.prove_from_variables(gpu, sd, values, mult)
    ref = PU.prove_circuit(gpu, sd, wit)
    assert (got == ref).all()
    assert (got == oracle.prove(geo, cfg, wit, setup)).all()
    sd.close()


def test_prove_from_hints_with_plain_witness_columns(gpu, oracle):
    """Compression-mode-1 shape (52 copy + boolean column + 78 plain witness columns): both hints of the reference's hand-off --
    DenseVariablesCopyHint for the copy-permuted columns, DenseWitnessCopyHint for the plain witness columns -- give the proof
    of the column path, bit for bit; an index past the value array is an error, not a silent zero."""
    import json
    import os
    fixture = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
    geo, cfg = [(g, c) for k, g, c, _ in G.compression_geometries_from_fixture(fixture) if k == "compression_1"][0]
    geo = geo.scaled(8)
    cfg = G.make_proof_config(8, 1 << cfg.log_lde, cfg.cap_size, security_level=4 * cfg.log_lde)
    wit, setup = PU.synth_trace(geo, seed=23)
    n, npm, npl = 1 << geo.log_n, geo.n_perm, geo.n_witness_plain
    assert npl == 78 and wit.shape[0] == npm + npl
    rng = np.random.default_rng(9)

    def to_maps(cells):
        uniq, inverse = np.unique(cells, return_inverse=True)
        perm = rng.permutation(uniq.size)
        values = np.empty_like(uniq)
        values[perm] = uniq
        maps = perm[inverse].astype(np.uint32)
        maps[cells == 0] = np.uint32(0xFFFFFFFF)
        return values, maps

    var_values, var_maps = to_maps(wit[:npm].reshape(-1))
    wit_values, wit_maps = to_maps(wit[npm:npm + npl].reshape(-1))
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    PU.set_variable_maps(gpu, sd, var_maps.reshape(npm, n))
    PU.set_witness_maps(gpu, sd, wit_maps.reshape(npl, n))
    got = PU.prove_from_hints(gpu, sd, var_values, wit_values)
    assert (got == PU.prove_circuit(gpu, sd, wit)).all()
    assert (got == oracle.prove(geo, cfg, wit, setup)).all()
    from era_zkevm_test_harness_b200 import ZkGpuError
    with pytest.raises(ZkGpuError, match="n_wits"):
        PU.prove_from_hints(gpu, sd, var_values, wit_values[: int(wit_maps[wit_maps != 0xFFFFFFFF].max())])
    with pytest.raises(ZkGpuError, match="n_vars"):
        PU.prove_from_hints(gpu, sd, var_values[: int(var_maps[var_maps != 0xFFFFFFFF].max())], wit_values)
    with pytest.raises(ZkGpuError, match="plain witness"):
        PU.prove_from_variables(gpu, sd, var_values)
    with pytest.raises(ValueError):
        PU.prove_circuit(gpu, sd, wit, proof_out=np.empty(16, dtype=np.uint64))
    sd.close()


def test_compute_setups_writes_reference_style_vk_files(gpu, oracle, tmp_path):
    """compute_setups mirror: one VK JSON per circuit type under the reference's file names; the cap in each file is the
    oracle's commitment of the same setup columns, the geometry read back from the file is the one that was committed."""
    import json
    import os
    from era_zkevm_test_harness_b200 import compute_setups as CS
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
    paths = CS.generate_base_layer_vks(gpu, str(tmp_path), fx, log_n=9)
    paths.update(CS.generate_recursive_layer_vks(gpu, str(tmp_path), fx, log_n=9))
    # the reference's file names, under <root>/synthetic/ with a manifest: synthetic setup columns + unpinned gate polynomials
    assert os.path.exists(tmp_path / "synthetic" / "SYNTHETIC.json")
    assert sorted(os.listdir(tmp_path / "synthetic" / "base_layer")) == sorted(f"vk_{t}.json" for t in range(1, 14))
    assert sorted(os.listdir(tmp_path / "synthetic" / "recursion_layer")) == ["vk_1.json", "vk_3.json", "vk_node.json"]
    for t in (1, 8, 10):
        (variant, vk), = json.load(open(paths[t])).items()
        assert variant == fx["base"][str(t)]["variant"] and vk["fixed_parameters"]["domain_size"] == 512
        geo = G.geometry_from_vk(vk, G.BASE_LAYER_GATE_ORDER[t])
        cfg = G.base_layer_proof_config(9)
        setup = PU.synth_trace(geo, seed=0x5E7)[1]
        assert (np.array(vk["setup_merkle_tree_cap"], dtype=np.uint64) == oracle.setup_cap(geo, cfg, setup)).all()


def test_recursive_layer_vks_and_proofs_self_check(gpu, oracle, tmp_path):
    """generate_recursive_layer_vks_and_proofs (compute_setups.rs:439-586): each recursion-layer circuit type is set up, proven
    once with a placeholder witness and the proof verified against the fresh VK before anything is written; the stored proof
    loads back and is accepted by BOTH verifiers; a witness that does not fit the setup aborts the run."""
    import json
    import os
    from era_zkevm_test_harness_b200 import compute_setups as CS, proof_format
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
    out = CS.generate_recursive_layer_vks_and_proofs(gpu, str(tmp_path), fx, log_n=9)
    assert set(out) == {"scheduler", "leaf_3", "node"}
    for key, rec in out.items():
        (variant, vk), = json.load(open(rec["vk"])).items()
        geo = G.geometry_from_vk(vk, G.RECURSION_GATE_ORDER)
        cfg = G.recursion_layer_proof_config(9)
        flat, _ = proof_format.load_proof_json(rec["proof"])
        cap = np.array(vk["setup_merkle_tree_cap"], dtype=np.uint64)
        ok, msg = PU.verify_proof(geo, cfg, cap, flat)
        assert ok, msg
        ok, msg = oracle.verify(geo, cfg, cap, flat)
        assert ok, msg

    def broken(geo):
        w = PU.synth_trace(geo, seed=0x5E7)[0]
        w[2, 3] ^= np.uint64(1)
        return w
    with pytest.raises(RuntimeError, match="does not verify"):
        CS.generate_recursive_layer_vks_and_proofs(gpu, str(tmp_path / "bad"), fx, log_n=9, witness_source=broken)


def test_staged_witness_upload_gives_the_same_proofs(gpu):
    """zkgpu_witness_stage / zkgpu_prove_staged: two instances of one circuit type, the second uploaded while the first is
    proven; proofs equal the single-call ones, slots can be reused, an empty slot is an error."""
    geo = G.mainvm_like_geometry(11)
    cfg = G.make_proof_config(11, 2, 16, security_level=10)
    wit_a, setup = PU.synth_trace(geo, seed=21, witness_seed=1, pinned=True)
    wit_b, _ = PU.synth_trace(geo, seed=21, witness_seed=2, pinned=True)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    ref_a, ref_b = PU.prove_circuit(gpu, sd, wit_a).copy(), PU.prove_circuit(gpu, sd, wit_b).copy()
    assert not (ref_a == ref_b).all()
    with pytest.raises(Exception):
        PU.prove_staged(gpu, sd, 1)
    PU.stage_witness(gpu, sd, wit_a, 0)
    PU.stage_witness(gpu, sd, wit_b, 1)
    got_a = PU.prove_staged(gpu, sd, 0).copy()
    PU.stage_witness(gpu, sd, wit_b, 0)          # slot 0 is free again: its proof has been read
    got_b = PU.prove_staged(gpu, sd, 1).copy()
    got_b2 = PU.prove_staged(gpu, sd, 0).copy()
    assert (got_a == ref_a).all() and (got_b == ref_b).all() and (got_b2 == ref_b).all()
    for p in (got_a, got_b):
        ok, msg = PU.verify_proof(geo, cfg, sd.vk_cap, p)
        assert ok, msg
    sd.close()


def test_pinned_host_buffers_from_the_library(gpu):
    """zkgpu_host_alloc gives page-locked memory the prover can read asynchronously; a proof from it equals one from numpy memory"""
    import ctypes
    geo = G.small_test_geometry(8, 16, True)
    cfg = G.make_proof_config(8, 2, 4, security_level=8)
    wit, setup = PU.synth_trace(geo, seed=4)
    sd = PU.create_setup_data(gpu, geo, cfg, setup)
    ptr = gpu.lib.zkgpu_host_alloc(wit.nbytes)
    assert ptr
    pinned = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64)), shape=wit.shape)
    pinned[:] = wit
    assert (PU.prove_circuit(gpu, sd, pinned) == PU.prove_circuit(gpu, sd, wit)).all()
    del pinned
    gpu.lib.zkgpu_host_free(ptr)
    sd.close()


def test_gpu_block_prover_staged_path_matches_oracle(gpu, oracle, tmp_path):
    """The block workflow on the GPU (block.GpuBlockProver: witnesses generated into pinned buffers on host threads, staged upload
    with look-ahead, LRU of resident setups) on a small synthetic block -- base (two types, several instances) -> leaf -> node ->
    scheduler -> compression mode 1: every proof equals the oracle's proof of the same trace word for word, is accepted by both
    verifiers, and lands under <out>/synthetic/ in the reference's file layout."""
    import json
    import os
    from era_zkevm_test_harness_b200 import block as B, proof_format
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
    table, base_keys, leaf_keys, node_key, sched_key = B.circuit_table(fx, log_n=8, compression_log_n=6)
    table = {k: (g, c if k.startswith("compression_") else G.make_proof_config(g.log_n, 2, 16, security_level=6)) for k, (g, c) in table.items()}
    plan = B.plan_block({1: 3, 8: 2}, base_keys, leaf_keys, node_key, sched_key, arity=2, compression_modes=(1,))
    prover = B.GpuBlockProver(gpu, table, max_resident=3, setup_seed=77, host_threads=4)
    checked = []

    def verify(job, proof):
        geo, cfg = table[job.geometry_key]
        cap = prover.vk_caps.get(job.geometry_key)
        if cap is None:
            cap = prover.setup(job.geometry_key).vk_cap.copy()
        ok1, _ = PU.verify_proof(geo, cfg, cap, proof)
        ok2, _ = oracle.verify(geo, cfg, cap, proof)
        checked.append(job.file)
        return ok1 and ok2

    seeds = {}

    def prove(job, seed):
        seeds[job.file] = seed
        return prover.prove(job, seed)

    res = B.prove_block(plan, prove, block_seed=3, out_dir=str(tmp_path), verify=verify, prefetch=prover.prefetch, security_level=6)
    assert len(res["proofs"]) == plan.n_jobs == len(checked)
    for _, jobs in plan.stages:
        for job in jobs[:2]:     # the oracle on the same trace: same setup seed, the job's witness seed
            geo, cfg = table[job.geometry_key]
            wit, setup = PU.synth_trace(geo, seed=77, witness_seed=seeds[job.file])
            assert (res["proofs"][job.file] == oracle.prove(geo, cfg, wit, setup)).all(), job.file
    flat, variant = proof_format.load_proof_json(str(tmp_path / "synthetic" / "recursion_layer" / "scheduler_proof.json"))
    assert variant == "SchedulerCircuit" and (flat == res["proofs"]["recursion_layer/scheduler_proof.json"]).all()
    assert os.path.exists(tmp_path / "synthetic" / "SYNTHETIC.json")
    prover.close()
