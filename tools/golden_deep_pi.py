#!/usr/bin/env python3
"""DEEP recovery for a golden proof whose public-input ROW is not the VK's (tools/deep/deep_solve_pi.c): finds phi, z AND the row
hash-free from four queries, then hands the row to tools/golden_deep.py, which verifies every query of the fixture.

Why: the golden base-layer proofs of circuit types 1, 5, 6, 7, 9, 11, 12 and the scheduler proof were produced with an older
circuit layout than the verification keys under setup/ (like the known-stale basic_circuit_proof_2_0.json): everything in their DEEP
relation matches the reference structure except the row the four public inputs are opened at.
Usage: python tools/golden_deep_pi.py <proof.json> <vk.json> <fri_chain fixture>"""
import json, os, struct, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_deep
from golden_fri_chain import P, omega
from make_golden_fixtures import reference_order

HERE = os.path.dirname(os.path.abspath(__file__))


def log_in_subgroup(r, log_n):
    """row with omega_{2^log_n}^row = r (Pohlig-Hellman in the 2-group), or None"""
    w = omega(log_n)
    if pow(r, 1 << log_n, P) != 1:
        return None
    row, winv = 0, pow(w, P - 2, P)
    cur = r
    for k in range(log_n):
        # bit k of row: (cur)^(2^(log_n-1-k)) == -1 ?
        if pow(cur, 1 << (log_n - 1 - k), P) != 1:
            row |= 1 << k
            cur = cur * pow(winv, 1 << k, P) % P
    return row if pow(w, row, P) == r else None


def find_row(proof_path, vk_path, fx_path, queries=(0, 1, 2, 3)):
    pr = json.load(open(proof_path))
    if "proof_config" not in pr:
        pr = pr[list(pr.keys())[0]]
    vk = json.load(open(vk_path))
    while "fixed_parameters" not in vk:
        vk = vk[list(vk.keys())[0]]
    fp = vk["fixed_parameters"]
    fx = json.load(open(fx_path))
    lde = pr["proof_config"]["fri_lde_factor"]
    log_n = fx["log_domains"][0] - (lde.bit_length() - 1)
    order = reference_order(fp, 1)
    polys, n0 = golden_deep.build(pr, fx, list(queries), order, log_n, None)       # E0: no public-input terms
    cols = [c for c, _ in fp["public_inputs_locations"]]
    path = "/tmp/deep_pi_in.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<4Q", len(queries), 8, n0, omega(log_n)))
        for q in queries:
            m0 = fx["queries"][q]["leaf_indexes"][0]
            w = pr["queries_per_fri_repetition"][q]["witness_query"]["leaf_elements"]
            for j in range(8):
                x = golden_deep.polys_for(pr, q, m0, j, fx["log_domains"][0], order, None, log_n)[0]
                d = [(w[c] - pr["public_inputs"][t]) % P for t, c in enumerate(cols)]
                f.write(struct.pack("<5Q", x, *d))
                for p in polys[(q, j)]:
                    f.write(struct.pack("<%dQ" % (2 * n0), *[v for co in p for v in co]))
    exe = "/tmp/deep_solve_pi"
    subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-o", exe, os.path.join(HERE, "deep", "deep_solve_pi.c")])
    out = subprocess.run([exe, path], capture_output=True, text=True)
    sys.stderr.write(out.stderr)
    rows = []
    for l in out.stdout.splitlines():
        if l.startswith("HIT"):
            t = l.split()
            r0, r1 = int(t[9]), int(t[10])
            row = log_in_subgroup(r0, log_n) if r1 == 0 else None
            print(l, "-> row", row)
            if row is not None:
                rows.append(row)
    return sorted(set(rows)), fp


if __name__ == "__main__":
    rows, fp = find_row(sys.argv[1], sys.argv[2], sys.argv[3])
    print("candidate rows:", rows, "VK row:", fp["public_inputs_locations"][0][1])
    for row in rows:
        pil = [[c, row] for c, _ in fp["public_inputs_locations"]]
        res, raw = golden_deep.solve(sys.argv[1], sys.argv[3], order=reference_order(fp, 1), pi_locs=pil)
        print("row", row, [(r["consistent"], r["phi"], r["z"]) for r in res])
