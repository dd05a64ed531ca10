"""Poseidon2 pin against the reference's golden proofs -- hash-agnostic known answers (tools/make_hash_kat_fixture.py).

These two tests are the mechanical definition of "Poseidon2 parity pinned" (DESIGN.md section 5).  They were strict xfails
until the round-constant table was identified (tools/gen_poseidon_constants.py: the plonky2 table of the Crandall-prime era);
they are hard gates now: the oracle's permutation, sponge framing and node hash reproduce the reference's digests.
"""
import json, os
import numpy as np
import pytest
from tests import oracle_lib

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "poseidon2_kat.json")))


def test_fixture_shape():
    assert len(KAT["node"]["cap"]) == 16 and len(KAT["leaf"]["cap"]) == 16
    assert 16 < len(KAT["node"]["top_siblings"]) <= 32          # two siblings under each of 16 cap entries
    assert len(KAT["leaf"]["leaves"]) == 16 and all(len(l) == 8 for l in KAT["leaf"]["leaves"])


def test_node_hash_reproduces_golden_cap():
    """Some ordered pair of top-of-path siblings must hash to a cap entry (reference tree: boojum MerkleTreeWithCap)."""
    orc = oracle_lib.load()
    cap = {tuple(c) for c in KAT["node"]["cap"]}
    tops = KAT["node"]["top_siblings"]
    hits = sum(tuple(int(x) for x in orc.hash_node(a, b)) in cap for a in tops for b in tops if a != b)
    assert hits >= len(tops) // 2 - 1


def test_leaf_hash_reproduces_golden_cap():
    """Every leaf of the last FRI oracle (16 leaves, empty path) must hash to a cap entry."""
    orc = oracle_lib.load()
    cap = {tuple(c) for c in KAT["leaf"]["cap"]}
    assert all(tuple(int(x) for x in orc.hash_leaf(np.array(l, dtype=np.uint64))) in cap for l in KAT["leaf"]["leaves"])
