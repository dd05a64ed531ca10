"""Gate constraint polynomials over GF(p^2) for tools/golden_quotient.py, with variant knobs (bit 0 of `variant` flips the sign of
every relation of the gate; higher bits select alternative variable layouts / relation orders).  Variant 0 of every gate is the
convention pinned on the golden proofs and implemented by oracle/gates.h and csrc/gates.cuh."""
P = (1 << 64) - (1 << 32) + 1
NR = 7
def e(a, b=0): return (a % P, b % P)
def eadd(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def esub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def emul(x, y): return ((x[0] * y[0] + NR * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def eneg(x): return ((-x[0]) % P, (-x[1]) % P)
def escale(x, k): return (x[0] * k % P, x[1] * k % P)
ONE, ZERO = e(1), e(0)
def epow7(x):
    x2 = emul(x, x); x4 = emul(x2, x2)
    return emul(emul(x4, x2), x)

WIDTH = {"ConstantsAllocator": 1, "FmaBaseNoConst": 4, "Reduction4": 5, "Selection": 4, "ParallelSelection4": 13, "ZeroCheck": 3,
         "UIntXAdd": 5, "DotProduct4": 9, "U8x4FMA": 26, "Poseidon2Flattened": 130, "FmaExt": 8, "U32TriAddCarryAsChunk": 5,
         "BoundedBoolean": 1, "MatMul12External": 24, "MatMul12Inner": 24, "Nonlinearity7": 2, "ConditionalSwap4": 17,
         "ZeroCheckWitness": 2, "BooleanAllColumns": 1, "PublicInput": 0, "Nop": 0}

M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]
SH = [4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12]
def p2_external(s):
    return [sum_e([escale(s[j], M4[i % 4][j % 4] * (2 if i // 4 == j // 4 else 1)) for j in range(12)]) for i in range(12)]
def p2_internal(s):
    t = sum_e(s)
    return [eadd(escale(s[i], 1 << SH[i]), t) for i in range(12)]
def sum_e(xs):
    a = ZERO
    for x in xs: a = eadd(a, x)
    return a

_RC = None
def rc():
    global _RC
    if _RC is None:
        import os, sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
        from gen_poseidon_constants import round_constants
        _RC = round_constants()
    return _RC


def instances(name, c, n_consts):
    w = WIDTH[name]
    if not w: return 0
    if name == "ConstantsAllocator": return n_consts
    if name == "Poseidon2Flattened": return (c["n_copy"] + c["n_plain"]) // 130
    if name == "BoundedBoolean": return min(c["n_copy"], 10)
    if name == "ZeroCheckWitness": return min(c["n_copy"] // 2, c["n_plain"])
    return c["n_copy"] // w


def eval_gate(name, c, v, k, n_consts, variant=0):
    inst = instances(name, c, n_consts)
    out = []
    alt = variant >> 1
    if name == "ConstantsAllocator":
        for t in range(inst): out.append(esub(v[t], k[t]))
    elif name == "FmaBaseNoConst":
        for t in range(inst):
            x = v[4 * t:4 * t + 4]
            kq, kl = (k[1], k[0]) if alt & 1 else (k[0], k[1])
            out.append(esub(eadd(emul(emul(kq, x[0]), x[1]), emul(kl, x[2])), x[3]))
    elif name == "Reduction4":
        for t in range(inst):
            x = v[5 * t:5 * t + 5]
            out.append(esub(sum_e([emul(k[i], x[i]) for i in range(4)]), x[4]))
    elif name == "Selection":
        for t in range(inst):
            x = v[4 * t:4 * t + 4]
            if alt & 1 == 0: a, b, s, r = x          # boojum: a, b, selector, result
            else: s, a, b, r = x
            if alt & 2: a, b = b, a
            out.append(esub(eadd(emul(s, a), emul(esub(ONE, s), b)), r))
    elif name == "ParallelSelection4":
        for t in range(inst):
            x = v[13 * t:13 * t + 13]
            if alt & 3 == 0:
                s = x[0]; tri = [(x[1 + 3 * i], x[2 + 3 * i], x[3 + 3 * i]) for i in range(4)]
            elif alt & 3 == 1:
                s = x[0]; tri = [(x[1 + i], x[5 + i], x[9 + i]) for i in range(4)]
            elif alt & 3 == 2:
                s = x[12]; tri = [(x[3 * i], x[3 * i + 1], x[3 * i + 2]) for i in range(4)]
            else:
                s = x[8]; tri = [(x[i], x[4 + i], x[9 + i]) for i in range(4)]
            if alt & 4: tri = [(b, a, r) for a, b, r in tri]
            for a, b, r in tri: out.append(esub(eadd(emul(s, a), emul(esub(ONE, s), b)), r))
    elif name == "ZeroCheck":
        for t in range(inst):
            x = v[3 * t:3 * t + 3]
            if alt & 1: var, inv, flag = x
            else: var, flag, inv = x
            r1 = emul(var, flag)
            r2 = esub(emul(var, inv), esub(ONE, flag))
            out += [r2, r1] if alt & 2 else [r1, r2]
    elif name == "ZeroCheckWitness":
        for t in range(inst):
            var, flag = v[2 * t], v[2 * t + 1]; inv = v[c["n_copy"] + t]
            r1 = emul(var, flag)
            r2 = esub(emul(var, inv), esub(ONE, flag))
            out += [r2, r1] if alt & 2 else [r1, r2]
    elif name == "UIntXAdd":
        for t in range(inst):
            a, b, cin, cc, cout = v[5 * t:5 * t + 5]
            r1 = esub(eadd(eadd(a, b), cin), eadd(cc, emul(k[0], cout)))
            r2 = esub(emul(cout, cout), cout)
            out += [r1] if alt == 0 else [r1, r2] if alt == 1 else [r2, r1]
    elif name == "U32TriAddCarryAsChunk":
        for t in range(inst):
            x = v[5 * t:5 * t + 5]
            out.append(esub(eadd(eadd(x[0], x[1]), x[2]), eadd(x[3], escale(x[4], 1 << 32))))
    elif name in ("BoundedBoolean", "BooleanAllColumns"):
        for t in range(inst): out.append(esub(emul(v[t], v[t]), v[t]))
    elif name in ("MatMul12External", "MatMul12Inner"):
        for t in range(inst):
            x = v[24 * t:24 * t + 24]
            s = p2_external(x[:12]) if name == "MatMul12External" else p2_internal(x[:12])
            for i in range(12): out.append(esub(s[i], x[12 + i]))
    elif name == "Nonlinearity7":
        for t in range(inst): out.append(esub(epow7(eadd(v[2 * t], k[0])), v[2 * t + 1]))
    elif name == "ConditionalSwap4":
        for t in range(inst):
            x = v[17 * t:17 * t + 17]
            if alt == 0: a, b, s, ra, rb = x[0:4], x[4:8], x[8], x[9:13], x[13:17]
            else: s, a, b, ra, rb = x[0], x[1:5], x[5:9], x[9:13], x[13:17]
            for i in range(4):
                out.append(esub(eadd(emul(s, b[i]), emul(esub(ONE, s), a[i])), ra[i]))
                out.append(esub(eadd(emul(s, a[i]), emul(esub(ONE, s), b[i])), rb[i]))
    elif name == "DotProduct4":
        for t in range(inst):
            x = v[9 * t:9 * t + 9]
            out.append(esub(sum_e([emul(x[2 * i], x[2 * i + 1]) for i in range(4)]), x[8]))
    elif name == "FmaExt":
        for t in range(inst):
            x = v[8 * t:8 * t + 8]
            def ext_from(c0, c1):   # an Ext2 element whose coordinates are themselves GF(p^2) values: c0 + u*c1, u^2 = 7
                return (c0, c1)
            def xmul(a, b):          # (a0 + a1 u)(b0 + b1 u)
                return (eadd(emul(a[0], b[0]), escale(emul(a[1], b[1]), 7)), eadd(emul(a[0], b[1]), emul(a[1], b[0])))
            kq, kl = (k[0], k[1]), (k[2], k[3])
            a, b, cc, d = (x[0], x[1]), (x[2], x[3]), (x[4], x[5]), (x[6], x[7])
            q = xmul(kq, xmul(a, b)); l = xmul(kl, cc)
            out.append(esub(eadd(q[0], l[0]), d[0])); out.append(esub(eadd(q[1], l[1]), d[1]))
    elif name == "U8x4FMA":
        for t in range(inst):
            x = v[26 * t:26 * t + 26]
            r = ZERO
            for i in range(4):
                for j in range(4): r = eadd(r, escale(emul(x[i], x[4 + j]), 1 << (8 * (i + j))))
            for i in range(4):
                sh = 1 << (8 * i)
                r = eadd(r, escale(eadd(x[8 + i], x[12 + i]), sh))
                r = esub(r, escale(x[16 + i], sh))
                r = esub(r, escale(x[20 + i], sh << 32))
            out.append(r)
    elif name == "Poseidon2Flattened" and alt >= 2:
        # design B: the new variables of a FULL round are the round's OUTPUT (after the external matrix), so the last twelve
        # cells are the permutation output; partial rounds: alt 2 -> the S-box output of lane 0, alt 3 -> lane 0 after the matrix
        if inst:
            RC = rc()
            s = p2_external(list(v[:12])); col = 12; r = 0
            def full():
                nonlocal s, col, r
                sb = p2_external([epow7(eadd(s[i], e(RC[12 * r + i]))) for i in range(12)])
                for i in range(12): out.append(esub(sb[i], v[col + i]))
                s = list(v[col:col + 12]); col += 12; r += 1
            for q in range(4): full()
            for q in range(22):
                sb = epow7(eadd(s[0], e(RC[12 * r])))
                if alt == 2:
                    out.append(esub(sb, v[col])); s[0] = v[col]; s = p2_internal(s)
                else:
                    t = p2_internal([sb] + s[1:])
                    out.append(esub(t[0], v[col])); s = [v[col]] + t[1:]
                col += 1; r += 1
            for q in range(4): full()
    elif name == "Poseidon2Flattened":
        if inst:
            RC = rc()
            s = p2_external(list(v[:12])); col = 12; r = 0
            for q in range(4):
                for i in range(12):
                    out.append(esub(epow7(eadd(s[i], e(RC[12 * r + i]))), v[col + i])); s[i] = v[col + i]
                col += 12; r += 1; s = p2_external(s)
            for q in range(22):
                out.append(esub(epow7(eadd(s[0], e(RC[12 * r]))), v[col])); s[0] = v[col]
                col += 1; r += 1; s = p2_internal(s)
            for q in range(4):
                if alt & 1 and q == 3:
                    # the last round's new variables are the OUTPUT state (after the final external matrix)
                    sb = p2_external([epow7(eadd(s[i], e(RC[12 * r + i]))) for i in range(12)])
                    for i in range(12): out.append(esub(sb[i], v[col + i]))
                    break
                for i in range(12):
                    out.append(esub(epow7(eadd(s[i], e(RC[12 * r + i]))), v[col + i])); s[i] = v[col + i]
                col += 12; r += 1; s = p2_external(s)
    elif name in ("PublicInput", "Nop"):
        pass
    else:
        raise KeyError(name)
    if variant & 1: out = [eneg(x) for x in out]
    return out
