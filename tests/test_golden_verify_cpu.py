"""Both verifiers of this repo -- the oracle's (oracle/prover.c orc_verify) and the product's (zkgpu_verify, CPU like the
reference's) -- on proofs made by the reference itself against the reference's own verification keys (tests/golden/pair_*.json,
tools/make_golden_pair_fixtures.py: MainVM 2^20, compression modes 1-4, base-layer RAMPermutation and L1MessagesHasher with
lookups, a node-layer proof; first three queries of each).

What this pins on boojum's own output: the Poseidon2 permutation and sponge framing, leaf and node hashing of all four trace
oracles and every FRI oracle, the Merkle cap convention, the Fiat-Shamir transcript (absorb order, ONE padding, 8 challenges
per squeeze), the query-index derivation, the lookup sum check, the DEEP combination, every FRI fold and the final polynomial.
The ONE check that is switched off is the quotient identity at z: the gate polynomials / term order of the un-vendored boojum
crate are not pinned yet, and the strict xfail below turns into a failure the day they are."""
import glob, json, os
import numpy as np
import pytest
from era_zkevm_test_harness_b200 import geometry as G, proof_format as PF, prover_utils as PU
from tests import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "pair_*.json")))
IDS = [os.path.basename(f)[5:-5] for f in FIXTURES]


def load_pair(path):
    d = json.load(open(path))
    kind, vk, pr = d["kind"], d["vk"], d["proof"]
    if kind.startswith("compression_"):
        mode = int(kind.split("_")[1])
        geo = G.geometry_from_vk(vk, G.COMPRESSION_GATE_ORDER[mode], has_boolean_col=1 if mode == 1 else 0)
    elif kind.startswith("base_"):
        geo = G.geometry_from_vk(vk, G.BASE_LAYER_GATE_ORDER[int(kind.split("_")[1])])
    else:
        geo = G.geometry_from_vk(vk, G.RECURSION_GATE_ORDER)
    pc = pr["proof_config"]
    log_lde = pc["fri_lde_factor"].bit_length() - 1
    cfg = G.make_proof_config(geo.log_n, pc["fri_lde_factor"], pc["merkle_tree_cap_size"], security_level=d["n_queries"] * log_lde)
    assert cfg.n_queries == d["n_queries"]
    flat, _ = PF.proof_from_dict(pr)
    return geo, cfg, np.array(vk["setup_merkle_tree_cap"], dtype=np.uint64).reshape(-1), flat


def test_fixture_set():
    assert {"mainvm", "compression_1", "compression_2", "compression_3", "compression_4", "base_8", "base_13", "node_3"} <= set(IDS)


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_reference_proofs_pass_every_check_but_the_quotient_identity(path):
    geo, cfg, cap, flat = load_pair(path)
    oracle = oracle_lib.load()
    ok, msg = oracle.verify(geo, cfg, cap, flat, skip_quotient_identity=True)
    assert ok, "oracle verifier: " + msg
    ok, msg = PU.verify_proof(geo, cfg, cap, flat, skip_quotient_identity=True)
    assert ok, "zkgpu_verify: " + msg


@pytest.mark.parametrize("path", FIXTURES[:3], ids=IDS[:3])
def test_corrupted_reference_proofs_are_rejected(path):
    """the accepted proofs are not accepted by accident: one flipped word anywhere past the header is caught"""
    geo, cfg, cap, flat = load_pair(path)
    oracle = oracle_lib.load()
    rng = np.random.default_rng(5)
    hdr = 32 + geo.n_public_inputs
    for pos in [hdr + 1, hdr + 3 * cfg.cap_size * 4 + 5, flat.size // 2, flat.size - 9] + list(rng.integers(hdr, flat.size - 1, 6)):
        bad = flat.copy(); bad[pos] ^= np.uint64(2)
        assert not oracle.verify(geo, cfg, cap, bad, skip_quotient_identity=True)[0], pos
        assert not PU.verify_proof(geo, cfg, cap, bad, skip_quotient_identity=True)[0], pos
    bad_cap = cap.copy(); bad_cap[7] ^= np.uint64(1)
    assert not oracle.verify(geo, cfg, bad_cap, flat, skip_quotient_identity=True)[0]
    assert not PU.verify_proof(geo, cfg, bad_cap, flat, skip_quotient_identity=True)[0]


@pytest.mark.xfail(reason="gate polynomials / quotient term order of the un-vendored boojum crate: not pinned (DESIGN.md section 5)", strict=True)
def test_reference_proof_passes_the_quotient_identity():
    geo, cfg, cap, flat = load_pair(os.path.join(HERE, "golden", "pair_node_3.json"))
    ok, msg = oracle_lib.load().verify(geo, cfg, cap, flat)
    assert ok, msg


DEEP_FOR_PAIR = {"compression_1": "deep_compression_1", "compression_2": "deep_compression_2", "node_3": "deep_node_3_0_0",
                 "base_8": "deep_ram_8_0", "base_13": "deep_l1_messages_hasher_13_0"}
FRI_FOR_PAIR = {"compression_1": "fri_chain_compression_1", "compression_2": "fri_chain_compression_2", "node_3": "fri_chain_node_3_0_0",
                "base_8": "fri_chain_ram_8_0"}


@pytest.mark.parametrize("pair", sorted(DEEP_FOR_PAIR))
def test_transcript_challenges_equal_the_hash_free_recoveries(pair):
    """Two independent routes to the same numbers: z, the DEEP challenge and the FRI fold challenges replayed through the pinned
    Poseidon2 transcript (tools/golden_transcript.py) against the values tools/golden_deep.py / golden_fri_chain.py had recovered
    from the same proofs algebraically, without any hash (tests/golden/deep_*.json, fri_chain_*.json)."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from tools.golden_transcript import Transcript, flat_cap
    d = json.load(open(os.path.join(HERE, "golden", f"pair_{pair}.json")))
    vk, pr = d["vk"], d["proof"]
    deep = json.load(open(os.path.join(HERE, "golden", DEEP_FOR_PAIR[pair] + ".json")))
    tr = Transcript(n_chal=8)
    tr.absorb(flat_cap(vk["setup_merkle_tree_cap"])); tr.absorb(pr["public_inputs"]); tr.absorb(flat_cap(pr["witness_oracle_cap"]))
    for _ in range(8 if pr["values_at_0"] else 4): tr.challenge()             # beta, gamma (+ lookup beta, gamma)
    tr.absorb(flat_cap(pr["stage_2_oracle_cap"])); tr.challenge(); tr.challenge()   # alpha
    tr.absorb(flat_cap(pr["quotient_oracle_cap"]))
    assert tr.challenge_ext() == deep["z"]
    for key in ("values_at_z", "values_at_z_omega", "values_at_0"):
        tr.absorb([x for e in pr[key] for x in e["coeffs"]])
    assert tr.challenge_ext() == deep["phi"]
    if pair in FRI_FOR_PAIR:
        fri = json.load(open(os.path.join(HERE, "golden", FRI_FOR_PAIR[pair] + ".json")))
        got = []
        for cap in [pr["fri_base_oracle_cap"]] + pr["fri_intermediate_oracles_caps"]:
            tr.absorb(flat_cap(cap)); got.append(tr.challenge_ext())
        assert got == fri["challenges"]
