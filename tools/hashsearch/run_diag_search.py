#!/usr/bin/env python3
"""Writes the KAT input of tools/hashsearch/diag_search.cu (mid-states after the first four full rounds for the four framings:
initial external layer yes/no x rate-first/capacity-first input placement; the 64 cap elements) and a self-test target."""
import json, os, struct, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_poseidon_constants import round_constants
P = (1 << 64) - (1 << 32) + 1
RC = round_constants()
M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]
SH = [4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12]


def ext(s):
    o = [0] * 12
    for b in range(3):
        x = s[4 * b:4 * b + 4]
        for i in range(4):
            o[4 * b + i] = sum(M4[i][j] * x[j] for j in range(4)) % P
    t = [(o[i] + o[4 + i] + o[8 + i]) % P for i in range(4)]
    return [(o[i] + t[i % 4]) % P for i in range(12)]


def first_half(s, init):
    if init:
        s = ext(s)
    for r in range(4):
        s = ext([pow((s[i] + RC[12 * r + i]) % P, 7, P) for i in range(12)])
    return s


def rest(s, sh):
    for r in range(4, 26):
        s[0] = pow((s[0] + RC[12 * r]) % P, 7, P)
        sm = sum(s) % P
        s = [(s[i] * (1 << sh[i]) + sm) % P for i in range(12)]
    for r in range(26, 30):
        s = ext([pow((s[i] + RC[12 * r + i]) % P, 7, P) for i in range(12)])
    return s


def main():
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon2_kat.json")))
    leaf = kat["leaf"]["leaves"][0]
    targets = sorted(set(x for c in kat["leaf"]["cap"] for x in c))
    mids = [first_half(list(inp), init) for init in (1, 0) for inp in (leaf + [0] * 4, [0] * 4 + leaf)]
    if "--selftest" in sys.argv:   # plant the output of the recollected diagonal as a 65th target: the search must report it
        targets = targets[:63] + [rest(list(mids[0]), SH)[5]]
    out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "/tmp/diag_kat.bin"
    with open(out, "wb") as f:
        f.write(struct.pack("<2Q", len(mids), len(targets)))
        for m in mids:
            f.write(struct.pack("<12Q", *m))
        f.write(struct.pack("<%dQ" % len(targets), *targets))
    print(out, len(mids), "mid-states,", len(targets), "targets")


if __name__ == "__main__":
    main()
