#!/usr/bin/env python3
"""The quotient identity at z of a golden proof, evaluated term group by term group over GF(p^2) in plain Python, with the
conventions that cannot be read off the proof alone (term order, signs, variable order inside gates, non-residues) exposed as
knobs.  Used to pin those conventions on the reference's own proofs; the pinned set is what oracle/gates.h,
oracle/prover.c and the CUDA kernels implement.  Usage: python tools/golden_quotient.py proof.json vk.json kind"""
import json, os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.golden_transcript import Transcript, flat_cap
from era_zkevm_test_harness_b200 import geometry as G

P = (1 << 64) - (1 << 32) + 1
NR = 7  # u^2 = 7


def e(a, b=0): return (a % P, b % P)
def eadd(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def esub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def emul(x, y): return ((x[0] * y[0] + NR * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def eneg(x): return ((-x[0]) % P, (-x[1]) % P)
def einv(x):
    n = (x[0] * x[0] - NR * x[1] * x[1]) % P
    ni = pow(n, P - 2, P)
    return (x[0] * ni % P, (-x[1]) * ni % P)
def epow(x, k):
    r = e(1)
    while k:
        if k & 1: r = emul(r, x)
        x = emul(x, x); k >>= 1
    return r
ONE, ZERO = e(1), e(0)


def non_residues(num, domain_size, mode="boojum"):
    """k_0 = 1, then boojum's make_non_residues: successive smallest quadratic non-residues whose cosets k*H are new"""
    if mode == "pow7":
        return [pow(7, i, P) for i in range(num)]
    if mode == "ints":      # 1, 2, 3, ... skipping values that fall into an earlier coset
        out, cur, seen = [], 0, set()
        while len(out) < num:
            cur += 1
            t = pow(cur, domain_size, P)
            if t in seen: continue
            seen.add(t); out.append(cur)
        return out
    if mode == "qnr_all":   # every quadratic non-residue, no leading 1
        out, cur, seen = [], 1, set()
        while len(out) < num:
            cur += 1
            if pow(cur, (P - 1) // 2, P) != P - 1: continue
            t = pow(cur, domain_size, P)
            if t in seen: continue
            seen.add(t); out.append(cur)
        return out
    out, cur, seen = [1], 1, {1}
    while len(out) < num:
        cur += 1
        if pow(cur, (P - 1) // 2, P) != P - 1: continue
        t = pow(cur, domain_size, P)
        if t in seen: continue
        seen.add(t); out.append(cur)
    return out


from tools.golden_gates import eval_gate   # noqa: E402  (gate library with variant knobs)


def load(proof_path, vk_path, kind, n_chal=8):
    pr = json.load(open(proof_path)); vk = json.load(open(vk_path))
    if "proof_config" not in pr: pr = list(pr.values())[0]
    if "setup_merkle_tree_cap" not in vk: vk = list(vk.values())[0]
    if kind.startswith("compression_"):
        mode = int(kind.split("_")[1]); order = G.COMPRESSION_GATE_ORDER[mode]; has_bool = 1 if mode == 1 else 0
    elif kind.startswith("base_"):
        order = G.BASE_LAYER_GATE_ORDER[int(kind.split("_")[1])]; has_bool = 1
    else:
        order = G.RECURSION_GATE_ORDER; has_bool = 1
    fp = vk["fixed_parameters"]; par = fp["parameters"]
    gates = []; G._walk_selector_tree(fp["selectors_placement"], [], gates); gates.sort()
    lp = fp["lookup_parameters"]
    LW, LR = (0, 0) if lp == "NoLookup" else (list(lp.values())[0]["width"], list(lp.values())[0]["num_repetitions"])
    c = dict(N=fp["domain_size"], n_copy=par["num_columns_under_copy_permutation"], n_plain=par["num_witness_columns"],
             n_const=par["num_constant_columns"] + fp["extra_constant_polys_for_selectors"] + (1 if LR else 0), LW=LW, LR=LR,
             QD=fp["quotient_degree"], has_bool=has_bool, table_id_col=(fp["table_ids_column_idxes"][0] if LR else 0),
             gates=[(order[i], nc, deg, path) for i, nc, deg, path in gates], pis=fp["public_inputs_locations"])
    c["NP"] = c["n_copy"] + has_bool + LW * LR
    c["C"] = -(-c["NP"] // c["QD"])
    tr = Transcript(n_chal=n_chal)
    tr.absorb(flat_cap(vk["setup_merkle_tree_cap"])); tr.absorb(pr["public_inputs"]); tr.absorb(flat_cap(pr["witness_oracle_cap"]))
    ch = {}
    l1 = [tr.challenge() for _ in range(8)]
    ch["lanes1"] = l1
    ch["beta"], ch["gamma"] = (l1[0], l1[1]), (l1[2], l1[3])
    if LR: ch["lbeta"], ch["lgamma"] = (l1[4], l1[5]), (l1[6], l1[7])
    tr.absorb(flat_cap(pr["stage_2_oracle_cap"]))
    lanes = [tr.challenge() for _ in range(8)]
    a0, a1 = [int(x) for x in os.environ.get("ALPHA_LANES", "0,1").split(",")]
    ch["alpha"] = (lanes[a0], lanes[a1]); ch["alpha_lanes"] = lanes
    tr.absorb(flat_cap(pr["quotient_oracle_cap"])); ch["z"] = tuple(tr.challenge_ext())
    az = [tuple(x["coeffs"]) for x in pr["values_at_z"]]
    o = {}
    k = 0
    def take(n):
        nonlocal k
        r = az[k:k + n]; k += n; return r
    o["perm"] = take(c["NP"]); o["plain"] = take(c["n_plain"]); o["const"] = take(c["n_const"]); o["sigma"] = take(c["NP"])
    o["z_and_partial"] = take(c["C"]); o["mult"] = take(1 if LR else 0); o["A"] = take(LR); o["B"] = take(1 if LR else 0)
    o["tables"] = take(LW + 1 if LR else 0); o["q"] = take(c["QD"])
    assert k == len(az), (k, len(az))
    o["z_omega"] = tuple(pr["values_at_z_omega"][0]["coeffs"])
    o["at_0"] = [tuple(x["coeffs"]) for x in pr["values_at_0"]]
    return c, ch, o, pr


def rhs(c, ch, o):
    z = ch["z"]; zn = epow(z, c["N"])
    q = ZERO; zp = ONE
    for qi in o["q"]:
        q = eadd(q, emul(zp, qi)); zp = emul(zp, zn)
    return emul(q, esub(zn, ONE))


def copy_perm_terms(c, ch, o, nr_mode="boojum"):
    z, N = ch["z"], c["N"]
    zn = epow(z, N)
    # boojum: "unnormalized_l1_inverse_at_z" = (z^n - 1) / (z - 1), WITHOUT the 1/n of the Lagrange polynomial
    l0 = emul(esub(zn, ONE), einv(esub(z, ONE)))
    if os.environ.get("L0_NORMALIZED"): l0 = emul(l0, einv(e(N)))
    terms = [emul(esub(o["z_and_partial"][0], ONE), l0)]
    ks = non_residues(c["NP"], N, nr_mode)
    for j in range(c["C"]):
        num, den = ONE, ONE
        for i in range(j * c["QD"], min((j + 1) * c["QD"], c["NP"])):
            w = o["perm"][i]
            num = emul(num, eadd(eadd(w, emul(ch["beta"], emul(e(ks[i]), z))), ch["gamma"]))
            den = emul(den, eadd(eadd(w, emul(ch["beta"], o["sigma"][i])), ch["gamma"]))
        prev = o["z_and_partial"][j]
        cur = o["z_and_partial"][j + 1] if j + 1 < c["C"] else o["z_omega"]
        if os.environ.get("CP_SWAP"): num, den = den, num
        terms.append(esub(emul(cur, den), emul(prev, num)))
    return terms


def lookup_terms(c, ch, o, swap=False):
    if not c["LR"]: return []
    LW = c["LW"]; lb, lg = (ch["lgamma"], ch["lbeta"]) if swap else (ch["lbeta"], ch["lgamma"])
    gp = [ONE]
    for _ in range(LW): gp.append(emul(gp[-1], lg))
    lw = o["perm"][c["n_copy"] + c["has_bool"]:] if not os.environ.get("LOOKUP_COLS_FIRST") else o["perm"][c["n_copy"]:c["n_copy"] + LW * c["LR"]]
    tid = emul(gp[LW], o["const"][c["table_id_col"]])
    t = []
    for i in range(c["LR"]):
        den = eadd(lb, tid)
        for j in range(LW): den = eadd(den, emul(gp[j], lw[i * LW + j]))
        t.append(esub(emul(o["A"][i], den), ONE))
    den = lb
    for j in range(LW + 1): den = eadd(den, emul(gp[j], o["tables"][j]))
    t.append(esub(emul(o["B"][0], den), o["mult"][0]))
    return t


def gate_terms(c, o, variants=None):
    """-> list per gate of (name, [relation values], selector)"""
    cells = o["perm"][:c["n_copy"]] + o["plain"]
    out = []
    for gi, (name, nc, deg, path) in enumerate(c["gates"]):
        sel = ONE
        for b, bit in enumerate(path):
            sel = emul(sel, o["const"][b] if bit else esub(ONE, o["const"][b]))
        rel = eval_gate(name, c, cells, o["const"][len(path):], nc, (variants or {}).get(name, 0))
        out.append((name, rel, sel))
    return out


def combine(groups, alpha):
    """groups: list of lists of Ext2 terms (already multiplied by selectors), consecutive powers of alpha"""
    acc = ZERO; ap = ONE
    for gterms in groups:
        for t in gterms:
            acc = eadd(acc, emul(ap, t)); ap = emul(ap, alpha)
    return acc


def check(c, ch, o, order="lookup,spec,gates,cp", variants=None, nr_mode="boojum", lookup_swap=False):
    gt = gate_terms(c, o, variants)
    parts = {"gates": [[emul(r, sel) for r in rel] for _, rel, sel in gt],
             "spec": [[esub(emul(o["perm"][c["n_copy"]], o["perm"][c["n_copy"]]), o["perm"][c["n_copy"]])]] if c["has_bool"] else [],
             "lookup": [lookup_terms(c, ch, o, lookup_swap)], "cp": [copy_perm_terms(c, ch, o, nr_mode)]}
    groups = []
    for name in order.split(","): groups += parts[name]
    return combine(groups, ch["alpha"]) == rhs(c, ch, o)


if __name__ == "__main__":
    c, ch, o, pr = load(sys.argv[1], sys.argv[2], sys.argv[3])
    print("gates:", [(n, nc, d) for n, nc, d, _ in c["gates"]])
    for order in ("lookup,spec,gates,cp", "gates,spec,lookup,cp"):
        for nr in ("boojum", "pow7"):
            print(order, nr, check(c, ch, o, order, nr_mode=nr))
