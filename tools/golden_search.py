#!/usr/bin/env python3
"""Brute force over the local conventions of the quotient identity of ONE golden proof: for every term group (gate, boolean column,
lookup, copy permutation) a list of candidate values (sign, variable layout, relation order); the identity picks one per group."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.golden_quotient import *
from tools import golden_gates as GG

ALTS = {"Selection": 2, "ParallelSelection4": 4, "ZeroCheck": 4, "ZeroCheckWitness": 2, "UIntXAdd": 3, "ConditionalSwap4": 2, "Poseidon2Flattened": 2}


def group_sum(terms, ap, alpha):
    acc = ZERO
    for t in terms:
        acc = eadd(acc, emul(ap, t)); ap = emul(ap, alpha)
    return acc, ap


def candidates(c, ch, o, order):
    """-> list of (label, [(variant description, value)])"""
    alpha = ch["alpha"]
    cells = o["perm"][:c["n_copy"]] + o["plain"]
    groups = []
    ap = ONE
    for part in order.split(","):
        if part == "gates":
            for name, nc, deg, path in c["gates"]:
                sel = ONE
                for b, bit in enumerate(path): sel = emul(sel, o["const"][b] if bit else esub(ONE, o["const"][b]))
                cands = []; nxt = ap
                for alt in range(ALTS.get(name, 1)):
                    rel = GG.eval_gate(name, c, cells, o["const"][len(path):], nc, alt << 1)
                    if not rel: continue
                    v, nxt_ = group_sum([emul(r, sel) for r in rel], ap, alpha)
                    if alt == 0: nxt = nxt_      # the alpha offset of the following groups follows variant 0 of every gate
                    cands += [((name, alt, "+"), v), ((name, alt, "-"), eneg(v))]
                if cands: groups.append((name, cands)); ap = nxt
        elif part == "spec" and c["has_bool"]:
            b = o["perm"][c["n_copy"]] if not os.environ.get("LOOKUP_COLS_FIRST") else o["perm"][c["NP"] - 1]
            v, ap = group_sum([esub(emul(b, b), b)], ap, alpha)
            groups.append(("bool", [(("bool", "+"), v), (("bool", "-"), eneg(v))]))
        elif part == "lookup" and c["LR"]:
            cands = []
            for swap in (False, True):
                v, nxt = group_sum(lookup_terms(c, ch, o, swap), ap, alpha)
                cands += [(("lookup", swap, "+"), v), (("lookup", swap, "-"), eneg(v))]
            ap = nxt; groups.append(("lookup", cands))
        elif part == "cp":
            cands = []
            for nr in ("boojum", "pow7"):
                for l0n in (0, 1):
                    if l0n: os.environ["L0_NORMALIZED"] = "1"
                    else: os.environ.pop("L0_NORMALIZED", None)
                    t = copy_perm_terms(c, ch, o, nr)
                    v, nxt = group_sum(t, ap, alpha)
                    cands += [(("cp", nr, l0n, "+"), v), (("cp", nr, l0n, "-"), eneg(v))]
            os.environ.pop("L0_NORMALIZED", None)
            ap = nxt; groups.append(("cp", cands))
    return groups


def search(groups, target):
    # meet in the middle over the two halves of the group list
    h = len(groups) // 2
    def sums(gs):
        out = {(): ZERO}
        res = [((), ZERO)]
        for _, cands in gs:
            res = [(d + (desc,), eadd(acc, v)) for d, acc in res for desc, v in cands]
        return res
    left = sums(groups[:h]); right = sums(groups[h:])
    table = {}
    for d, v in left: table.setdefault(v, []).append(d)
    hits = []
    for d, v in right:
        need = esub(target, v)
        for dl in table.get(need, []): hits.append(dl + d)
    return hits, len(left), len(right)


if __name__ == "__main__":
    c, ch, o, pr = load(sys.argv[1], sys.argv[2], sys.argv[3])
    target = rhs(c, ch, o)
    for order in ("lookup,spec,gates,cp", "gates,spec,lookup,cp", "spec,lookup,gates,cp", "lookup,gates,spec,cp"):
        groups = candidates(c, ch, o, order)
        hits, nl, nr_ = search(groups, target)
        print(order, "combos", nl, "x", nr_, "hits:", len(hits))
        for h in hits: print("   ", h)
