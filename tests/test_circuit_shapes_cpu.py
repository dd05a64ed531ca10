"""Every circuit type of the reference (13 base-layer circuits, scheduler, leaf, node, compression modes 1-4): the geometry read from its
verification key (tests/golden/vk_shapes.json <- setup/**/vk_*.json) must give, through the column-count formulas of the
C ABI (zkgpu_num_*_cols, zkgpu_proof_size_u64) and the folding-schedule rule, exactly the oracle widths, opening counts,
Merkle path lengths and FRI leaf shapes observed in the reference's own golden proofs of that circuit."""
import json
import math
import os

import pytest

from era_zkevm_test_harness_b200 import geometry as G
from era_zkevm_test_harness_b200 import prover_utils as PU

FIXTURE = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vk_shapes.json")))
CIRCUITS = list(G.circuit_geometries_from_fixture(FIXTURE))
# compression modes 1-4 (+ the wrapper-facing variants): geometry with plain witness columns from compression_N_vk.json
CIRCUITS += [(k, g, e) for k, g, _cfg, e in G.compression_geometries_from_fixture(FIXTURE)]


@pytest.mark.parametrize("key,geo,entry", CIRCUITS, ids=[c[0] for c in CIRCUITS])
def test_shapes_match_golden_proofs(key, geo, entry):
    cols = PU.num_columns(geo)
    assert cols["witness"] == geo.n_witness and cols["setup"] == geo.n_setup and cols["stage2"] == geo.n_stage2
    if key == "aux_eip4844":
        # the reference holds a VK but no proof of the EIP-4844 circuit: check the counts the geometry implies
        # (eip4844/mod.rs:43-56: 60 copy columns, 8 constant columns, lookup 3 x 20; no specialised boolean column)
        assert (geo.n_copy, geo.lookup_width, geo.lookup_reps, geo.has_boolean_col) == (60, 3, 20, 0)
        assert (geo.n_witness, geo.n_setup, geo.n_stage2) == (60 + 60 + 1, 120 + 11 + 4, 2 * (15 + 20 + 1))
        assert [(geo.gates[i].n_consts) for i in range(geo.n_gates)] == [8, 0, 2, 4, 0, 1, 0, 0]
        return
    assert entry["proof_shapes"], "no golden proof for this circuit"
    for sh in entry["proof_shapes"]:
        pc = sh["proof_config"]
        log_lde = int(math.log2(pc["fri_lde_factor"]))
        log_n = sh["path_len"] + int(math.log2(pc["merkle_tree_cap_size"])) - log_lde
        if log_n != geo.log_n:
            # test_proofs/base_layer/basic_circuit_proof_2_0.json was produced with a 2^15 test geometry
            assert key.startswith("base_2_") and log_n == 15
            g = geo.scaled(log_n)
        else:
            g = geo
        assert (sh["W"], sh["S2"], sh["Q"], sh["S"]) == (g.n_witness, g.n_stage2, g.n_quotient, g.n_setup)
        assert sh["values_at_z"] == g.n_witness + g.n_setup + g.n_stage2 // 2 + g.n_quotient // 2
        assert sh["values_at_z_omega"] == 1
        assert sh["values_at_0"] == (g.lookup_reps + 1 if g.lookup_reps else 0)
        assert sh["n_public_inputs"] == g.n_public_inputs
        cfg = G.make_proof_config(log_n, pc["fri_lde_factor"], pc["merkle_tree_cap_size"], pc["security_level"], pc["pow_bits"])
        assert cfg.n_queries == sh["n_queries"]
        sched = list(cfg.fri_schedule[: cfg.n_fri_oracles])
        assert [leaf // 2 for leaf, _ in sh["fri"]] == [1 << s for s in sched]
        log_dom, cap_log = log_n + log_lde, int(math.log2(pc["merkle_tree_cap_size"]))
        for (leaf, path), s in zip(sh["fri"], sched):
            assert path == max(log_dom - s - cap_log, 0)
            log_dom -= s
        assert sh["final_fri_monomials"] == [1 << (log_dom - log_lde)] * 2
        # flat proof buffer: header + every field of Proof<F,H,EXT>
        per_query = sh["W"] + sh["S2"] + sh["Q"] + sh["S"] + 4 * 4 * sh["path_len"] + sum(leaf + 4 * path for leaf, path in sh["fri"])
        caps = 4 * sh["cap_len"] * 3 + sum(4 * min(sh["cap_len"], 1 << (lf_dom)) for lf_dom in
                                            [log_n + log_lde - sum(sched[:k + 1]) for k in range(len(sched))])
        expect = 32 + g.n_public_inputs + caps + sum(sh["final_fri_monomials"]) + 2 * (sh["values_at_z"] + 1 + sh["values_at_0"]) \
            + sh["n_queries"] * per_query + 1
        assert PU.proof_size_u64(g, cfg) == expect


def test_gate_sets_cover_every_reference_gate_index():
    for key, geo, entry in CIRCUITS:
        kinds = [geo.gates[i].kind for i in range(geo.n_gates)]
        assert G.GATE_CONSTANTS_ALLOCATOR in kinds and G.GATE_FMA in kinds, key
        if key.startswith("compression"):
            assert G.GATE_CONDITIONAL_SWAP4 in kinds and G.GATE_FMA_EXT in kinds
            assert (G.GATE_POSEIDON2_FLATTENED in kinds) == (geo.n_witness_plain > 0)   # 130 cells = copy + plain columns
            if geo.n_witness_plain:
                assert geo.n_copy + geo.n_witness_plain == 130
            else:
                assert G.GATE_MATMUL12_EXTERNAL in kinds and G.GATE_MATMUL12_INNER in kinds and G.GATE_NONLINEARITY7 in kinds
        if key.startswith("recursion"):
            assert G.GATE_FMA_EXT in kinds and G.GATE_POSEIDON2_FLATTENED in kinds
