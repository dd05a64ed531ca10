#!/usr/bin/env python3
"""Several golden proofs of ONE circuit (the 13 node-layer proofs share vk_node.json) give a linear system over GF(p^2) in one
unknown scalar per term group of the quotient identity: sum_g c_g * (sum_r alpha^(off_g + r) sel_g R_{g,r}(z)) = q(z) Z_H(z).
A correct convention set solves it with every c_g = 1; a sign error shows as c_g = -1, a wrong gate polynomial makes it
inconsistent.  This localises errors that the single yes/no of the identity cannot."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.golden_quotient import *


def solve(A, b):
    """Gaussian elimination over GF(p^2); A: rows x cols (rows >= cols). Returns (solution, max residual is zero?)"""
    n, m = len(A), len(A[0])
    M = [list(r) + [bb] for r, bb in zip(A, b)]
    piv = []
    r = 0
    for col in range(m):
        pr_ = next((i for i in range(r, n) if M[i][col] != ZERO), None)
        if pr_ is None: continue
        M[r], M[pr_] = M[pr_], M[r]
        inv = einv(M[r][col])
        M[r] = [emul(x, inv) for x in M[r]]
        for i in range(n):
            if i != r and M[i][col] != ZERO:
                f = M[i][col]
                M[i] = [esub(x, emul(f, y)) for x, y in zip(M[i], M[r])]
        piv.append(col); r += 1
    sol = [None] * m
    for i, col in enumerate(piv): sol[col] = M[i][m]
    consistent = all(M[i][m] == ZERO for i in range(r, n))
    return sol, consistent, len(piv)


def group_columns(c, ch, o, order, variants, nr_mode, split=()):
    """-> (labels, values) one Ext2 value per group with its alpha offset applied"""
    gt = gate_terms(c, o, variants)
    parts = {"gates": [(name, [emul(r, sel) for r in rel]) for name, rel, sel in gt],
             "spec": [("bool", [esub(emul(o["perm"][c["n_copy"]], o["perm"][c["n_copy"]]), o["perm"][c["n_copy"]])])] if c["has_bool"] else [],
             "lookup": [("lookup", lookup_terms(c, ch, o))] if c["LR"] else [], "cp": [("cp", copy_perm_terms(c, ch, o, nr_mode))]}
    labels, vals = [], []
    ap = ONE
    for name in order.split(","):
        for label, terms in parts[name]:
            if not terms: continue
            if label in split:
                for i, t in enumerate(terms):
                    labels.append(f"{label}[{i}]"); vals.append(emul(ap, t)); ap = emul(ap, ch["alpha"])
                continue
            acc = ZERO
            for t in terms:
                acc = eadd(acc, emul(ap, t)); ap = emul(ap, ch["alpha"])
            labels.append(label); vals.append(acc)
    return labels, vals


def main():
    R = "/root/reference"
    kind = sys.argv[1] if len(sys.argv) > 1 else "node"
    order = os.environ.get("ORDER", "lookup,spec,gates,cp")
    nr = os.environ.get("NR", "boojum")
    split = tuple(os.environ.get("SPLIT", "").split(",")) if os.environ.get("SPLIT") else ()
    variants = eval(os.environ.get("VARIANTS", "{}"))
    A, b = [], []
    for t in range(3, 16):
        c, ch, o, pr = load(f"{R}/test_proofs/recursion_layer/node_layer_proof_{t}_0_0.json", f"{R}/setup/recursion_layer/vk_node.json", "recursion")
        labels, vals = group_columns(c, ch, o, order, variants, nr, split)
        rv = rhs(c, ch, o)
        sc = os.environ.get("RHS_SCALE")
        if sc == "alpha": rv = emul(rv, ch["alpha"])
        if sc == "alphainv": rv = emul(rv, einv(ch["alpha"]))
        if sc and sc.startswith("rev"):   # reversed powers: multiply every group by alpha^(n-1) and invert alpha -- emulate by conj trick
            pass
        A.append(vals); b.append(rv)
    sol, ok, rank = solve(A, b)
    print("unknowns", len(labels), "equations", len(A), "rank", rank, "consistent:", ok)
    for l, s in zip(labels, sol):
        tag = "= 1" if s == ONE else "= -1" if s == eneg(ONE) else ""
        print(f"  {l:24s} {s} {tag}")


if __name__ == "__main__":
    main()
