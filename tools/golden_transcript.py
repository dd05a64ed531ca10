#!/usr/bin/env python3
"""Replays the hash-dependent parts of the reference's golden proofs with the oracle's Poseidon2 (pinned, tests/test_hash_pin_cpu.py):
Merkle paths of every query against the caps in the proof / verification key, and the Fiat-Shamir transcript, whose outputs are
compared with the values recovered hash-free by tools/golden_deep.py and tools/golden_fri_chain.py (z, the DEEP challenge, the FRI
challenges, the query indexes).  Usage: python tools/golden_transcript.py <proof.json> <vk.json> [deep_fixture.json] [fri_fixture.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib

P = (1 << 64) - (1 << 32) + 1
orc = oracle_lib.load()


def u64(a):
    return np.array(a, dtype=np.uint64)


def merkle_root_from_path(leaf, path, idx):
    h = orc.hash_leaf(u64(leaf))
    for sib in path:
        sib = u64(sib)
        h = orc.hash_node(h, sib) if (idx & 1) == 0 else orc.hash_node(sib, h)
        idx >>= 1
    return tuple(int(x) for x in h), idx


def check_query_paths(q, caps, idx, log=print):
    ok = True
    for name, cap in caps.items():
        oq = q[name]
        root, top = merkle_root_from_path(oq["leaf_elements"], oq["proof"], idx)
        good = root == tuple(cap[top]) if top < len(cap) else False
        ok &= good
        if not good:
            log(f"  {name}: path does not reach cap[{top}]")
    return ok


class Transcript:
    """oracle/prover.c tr_* restated for the replay (so hypotheses can be varied here before the C code changes).
    Pinned on the golden proofs: overwrite mode, state kept across squeezes, and every squeeze that follows witnessed elements
    appends a ONE to them before the zero fill (the tree hasher does not)."""
    def __init__(self, rate=8, n_chal=4, pad_zero=True):
        self.st = np.zeros(12, dtype=np.uint64); self.buf = []; self.pos = n_chal; self.n_chal = n_chal; self.rate = rate

    def absorb(self, vals):
        self.buf += [int(v) for v in vals]

    def challenge(self):
        if self.buf:
            self.buf.append(1)
            for i in range(0, len(self.buf), self.rate):
                blk = self.buf[i:i + self.rate]
                blk += [0] * (self.rate - len(blk))
                self.st[:self.rate] = u64(blk)
                self.st = orc.permute(self.st.reshape(1, 12)).reshape(12)
            self.buf = []; self.pos = 0
        elif self.pos == self.n_chal:
            self.st = orc.permute(self.st.reshape(1, 12)).reshape(12); self.pos = 0
        v = int(self.st[self.pos]); self.pos += 1
        return v

    def challenge_ext(self):
        return [self.challenge(), self.challenge()]


def flat_cap(cap):
    return [x for c in cap for x in c]


def main():
    proof_path, vk_path = sys.argv[1], sys.argv[2]
    pr = json.load(open(proof_path))
    vk = json.load(open(vk_path))
    if "proof_config" not in pr: pr = list(pr.values())[0]       # typed wrapper {"CircuitName": {...}}
    if "setup_merkle_tree_cap" not in vk: vk = list(vk.values())[0]
    deep = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
    fri = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else None
    caps = {"witness_query": pr["witness_oracle_cap"], "stage_2_query": pr["stage_2_oracle_cap"], "quotient_query": pr["quotient_oracle_cap"],
            "setup_query": vk["setup_merkle_tree_cap"]}
    if deep:
        for fq in deep["queries"]:
            q = [q for q in pr["queries_per_fri_repetition"] if q["witness_query"]["leaf_elements"][:8] == fq["witness"][:8]][0]
            print("query at lde index", fq["lde_index"], "paths ok:", check_query_paths(q, caps, fq["lde_index"]))
    tr = Transcript(n_chal=int(os.environ.get('N_CHAL', 4)))
    tr.absorb(flat_cap(vk["setup_merkle_tree_cap"])); tr.absorb(pr["public_inputs"]); tr.absorb(flat_cap(pr["witness_oracle_cap"]))
    beta, gamma = tr.challenge_ext(), tr.challenge_ext()
    lookups = len(pr["values_at_0"]) > 0
    if lookups: lbeta, lgamma = tr.challenge_ext(), tr.challenge_ext()
    tr.absorb(flat_cap(pr["stage_2_oracle_cap"])); alpha = tr.challenge_ext()
    tr.absorb(flat_cap(pr["quotient_oracle_cap"])); z = tr.challenge_ext()
    print("z (transcript)", z, " z (recovered)", deep and deep["z"])
    def ext_list(l): return [x for e in l for x in e["coeffs"]]
    tr.absorb(ext_list(pr["values_at_z"])); tr.absorb(ext_list(pr["values_at_z_omega"])); tr.absorb(ext_list(pr["values_at_0"]))
    phi = tr.challenge_ext()
    print("phi (transcript)", phi, " phi (recovered)", deep and deep["phi"])
    chs = []
    for cap in [pr["fri_base_oracle_cap"]] + pr["fri_intermediate_oracles_caps"]:
        tr.absorb(flat_cap(cap)); chs.append(tr.challenge_ext())
    print("fri challenges (transcript)", chs[:2], " (recovered)", fri and fri["challenges"][:2])
    fin = pr["final_fri_monomials"]
    if os.environ.get("FIN_INTERLEAVED"):
        tr.absorb([x for a, b in zip(fin[0], fin[1]) for x in (a, b)])
    else:
        tr.absorb(fin[0]); tr.absorb(fin[1])
    log_n = int(vk["fixed_parameters"]["domain_size"]).bit_length() - 1
    log_lde = int(pr["proof_config"]["fri_lde_factor"]).bit_length() - 1
    bits_needed = log_n + log_lde
    take = 64 - bits_needed
    avail = []
    n_ok = 0
    fri_caps = [pr["fri_base_oracle_cap"]] + pr["fri_intermediate_oracles_caps"]
    for qi, q in enumerate(pr["queries_per_fri_repetition"]):
        while len(avail) < bits_needed:
            c = tr.challenge()
            avail += [(c >> b) & 1 for b in range(take)]
        bits, avail = avail[:bits_needed], avail[bits_needed:]
        idx = sum(b << i for i, b in enumerate(bits))
        ok = check_query_paths(q, caps, idx, log=(print if os.environ.get('VERBOSE') else (lambda m: None)))
        if os.environ.get('VERBOSE'): print('query', qi, 'idx', idx, 'base oracles ok', ok)
        # FRI oracles: leaf index = idx >> (sum of folding steps so far + this step)
        fidx = idx
        for k, fq in enumerate(q["fri_queries"]):
            step = (len(fq["leaf_elements"]) // 2).bit_length() - 1
            fidx >>= step
            root, top = merkle_root_from_path(fq["leaf_elements"], fq["proof"], fidx)
            ok &= top < len(fri_caps[k]) and root == tuple(fri_caps[k][top])
        n_ok += ok
    print("queries whose transcript-derived index opens all oracles:", n_ok, "of", len(pr["queries_per_fri_repetition"]))


if __name__ == "__main__":
    main()
