#!/usr/bin/env python3
"""Hash-free recovery of the DEEP challenge phi and the opening point z from a golden proof, and with them a check of the
DEEP (quotening) structure this framework restates (oracle/prover.c "DEEP polynomial"):

  h(x) = sum_{i < n_z} phi^i (F_i(x) - F_i(z)) / (x - z)  +  phi^{n_z} (Z(x) - Z(z w)) / (x - z w)
         + sum_j phi^{n_z + 1 + j} (A_j(x) - A_j(0)) / x

         + sum_t phi^{..} (w_col_t(x) - pi_t) / (x - w^row_t)            (public inputs, `pi_locs`)

with F_i in a caller-chosen order (`order`: the pairing of leaf elements with `values_at_z` is the hypothesis under test; the
order confirmed on golden proofs is [witness leaf][constants][sigmas][stage-2 Ext2 polys][quotient Ext2 polys]), Z = stage-2
poly 0, A_j = the lookup polys (stage-2 polys C.. ), h = the base FRI oracle.  Known per query without the hash: the leaf of every trace
oracle at the query point, the 8 values of the FRI base-oracle leaf and its index (tools/golden_fri_chain.py), all openings.
Unknown: phi, z (Ext2) and the position j of the query point inside its FRI leaf.  Multiplying out gives
E_{q,j}(phi, z) = a + b z + c z^2 = 0 with polynomials a, b, c in phi; tools/deep/deep_solve.c eliminates z (resultant of two
quadratics) and finds phi as a common root over three queries; this script then verifies EVERY query of the fixture.
Usage: python tools/golden_deep.py <proof.json> <fri_chain fixture> [out.json]"""
import json, os, struct, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_fri_chain import P, omega, ea, es, em, esc, einv, brev, Z as ZERO

HERE = os.path.dirname(os.path.abspath(__file__))


def ext(c):
    return (c["coeffs"][0], c["coeffs"][1])


def polys_for(pr, q, m0, j, log_dom, order, pi_locs=None, log_n=None, pi_first=False, tail=None):
    Q = pr["queries_per_fri_repetition"][q]
    idx = (m0 << 3) + j
    x = 7 * pow(omega(log_dom), brev(idx, log_dom), P) % P
    fl = Q["fri_queries"][0]["leaf_elements"]; hh = len(fl) // 2
    h = (fl[j], fl[hh + j])
    w = Q["witness_query"]["leaf_elements"]; s = Q["setup_query"]["leaf_elements"]
    s2 = Q["stage_2_query"]["leaf_elements"]; qq = Q["quotient_query"]["leaf_elements"]
    s2e = [(s2[2 * e], s2[2 * e + 1]) for e in range(len(s2) // 2)]
    qe = [(qq[2 * e], qq[2 * e + 1]) for e in range(len(qq) // 2)]
    if callable(order):    # order(w, s, s2, q) -> (F list of Ext2 values, index of Z in F-space stage-2 list, lookup polys list)
        F, zpoly, lookups = order(w, s, s2, qq)
    else:
        parts = {"w": [(v, 0) for v in w], "s": [(v, 0) for v in s], "2": s2e, "q": qe}
        F = sum((parts[k] for k in order), [])
        zpoly, lookups = s2e[0], None
    V = [ext(c) for c in pr["values_at_z"]]
    assert len(F) == len(V), (len(F), len(V))
    n_z = len(V)
    at0 = [ext(c) for c in pr["values_at_0"]]
    zw_val = ext(pr["values_at_z_omega"][0])
    C = len(s2e) - len(at0)            # lookup polys are the last ones of stage 2
    if lookups is None:
        lookups = s2e[C:]
    n = n_z + 1 + len(at0) + (len(pi_locs) if pi_locs else 0)
    T = [es(F[i], V[i]) for i in range(n_z)] + [ZERO] * (n - n_z)
    n_pi = len(pi_locs) if pi_locs else 0
    # order of the three groups after the openings at z: "w" = z*omega (1 term), "0" = openings at 0, "p" = public inputs
    tail = tail or ("w0p" if not pi_first else "wp0")
    off, pos_of = n_z, {}
    for ch in tail:
        pos_of[ch] = off
        off += {"w": 1, "0": len(at0), "p": n_pi}[ch]
    U = [ZERO] * n; U[pos_of["w"]] = es(zpoly, zw_val)
    xinv = pow(x, P - 2, P)
    M = [ZERO] * n
    base0, base_pi = pos_of["0"], pos_of["p"]
    for t in range(len(at0)):
        M[base0 + t] = esc(es(lookups[t], at0[t]), xinv)
    if pi_locs:   # public inputs as openings of variable columns at w^row: (w_col(x) - value) / (x - w^row)
        om_n = omega(log_n)
        for t, (col, row) in enumerate(pi_locs):
            den = (x - pow(om_n, row, P)) % P
            M[base_pi + t] = esc(es((w[col], 0), (pr["public_inputs"][t], 0)), pow(den, P - 2, P))
    M[0] = es(M[0], h)
    return x, T, U, M, n


def build(pr, fx, qs, order, log_n, pi_locs=None, pi_first=False, tail=None):
    log_dom = fx["log_domains"][0]
    om = omega(log_n)
    out = {}
    for q in qs:
        m0 = fx["queries"][q]["leaf_indexes"][0]
        for j in range(8):
            x, T, U, M, n = polys_for(pr, q, m0, j, log_dom, order, pi_locs, log_n, pi_first, tail)
            a = [ea(ea(esc(T[i], x), esc(U[i], x)), esc(M[i], x * x % P)) for i in range(n)]
            b = [es(es(esc(T[i], (P - om) % P), U[i]), esc(M[i], x * (1 + om) % P)) for i in range(n)]
            c = [esc(M[i], om) for i in range(n)]
            out[(q, j)] = (a, b, c)
    return out, n


def peval(p, v):
    r = ZERO
    for co in reversed(p):
        r = ea(em(r, v), co)
    return r


def solve(proof_path, fx_path, order=("w", "s", "2", "q"), pi_locs=None, pi_first=False, triple=(0, 1, 2)):
    pr = json.load(open(proof_path))
    if "proof_config" not in pr:
        pr = pr[list(pr.keys())[0]]
    fx = json.load(open(fx_path))
    lde = pr["proof_config"]["fri_lde_factor"]
    log_n = fx["log_domains"][0] - (lde.bit_length() - 1)
    polys, n = build(pr, fx, list(triple), order, log_n, pi_locs, pi_first)
    path = "/tmp/deep_in.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<3Q", 3, 8, n))
        for q in triple:
            for j in range(8):
                for p in polys[(q, j)]:
                    f.write(struct.pack("<%dQ" % (2 * n), *[v for co in p for v in co]))
    exe = "/tmp/deep_solve"
    subprocess.check_call(["gcc", "-O3", "-fopenmp", "-o", exe, os.path.join(HERE, "deep", "deep_solve.c")])
    out = subprocess.run([exe, path], capture_output=True, text=True).stdout
    hits = [l.split() for l in out.splitlines() if l.startswith("HIT") and "deg 1" in l]
    results = []
    for hshit in hits:
        j0, j1 = int(hshit[1]), int(hshit[2])
        phi = (int(hshit[6]), int(hshit[7]))
        a1, b1, c1 = [peval(p, phi) for p in polys[(triple[0], j0)]]
        a2, b2, c2 = [peval(p, phi) for p in polys[(triple[1], j1)]]
        den = es(em(c2, b1), em(c1, b2))
        if den == ZERO:
            continue
        z = em(es(em(c1, a2), em(c2, a1)), einv(den))
        # verify every query of the fixture
        allp, _ = build(pr, fx, range(len(fx["queries"])), order, log_n, pi_locs, pi_first)
        pos = []
        for q in range(len(fx["queries"])):
            hit = None
            for j in range(8):
                a, b, c = [peval(p, phi) for p in allp[(q, j)]]
                if ea(ea(a, em(b, z)), em(c, em(z, z))) == ZERO:
                    hit = j
            pos.append(hit)
        results.append({"phi": phi, "z": z, "positions": pos, "consistent": all(p is not None for p in pos)})
    return results, out


if __name__ == "__main__":
    res, raw = solve(sys.argv[1], sys.argv[2])
    print(raw.strip().splitlines()[-1])
    for r in res:
        print(r)
    if len(sys.argv) > 3 and res:
        json.dump(res[0], open(sys.argv[3], "w"))
