"""CPU tests of the oracle (oracle/*.c): self-consistency, definitions, and the reference's golden vectors."""
import glob
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import P, rand_field

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def brev(x, b):
    return int(format(x, "0%db" % b)[::-1], 2) if b else 0


def test_field_constants(oracle):
    # 7 generates F_p^*, G = 7^((p-1)/2^32) is the 2^32-th root boojum uses (verified numerically)
    assert oracle.pow(7, (P - 1) >> 32) == 0x185629DCDA58878C
    assert oracle.omega(20) == 3511170319078647661
    assert oracle.omega(21) == 17654865857378133588
    assert oracle.pow(oracle.omega(6), 1) == pow(2, 39, P)  # omega_64 = 2^39 (shift-only twiddles)
    assert pow(7, (P - 1) // 2, P) == P - 1  # 7 is a quadratic non-residue -> u^2 = 7 defines Ext2


def test_field_vectors_against_python_ints(oracle):
    rng = np.random.default_rng(1)
    a, b = rand_field(rng, 4096), rand_field(rng, 4096)
    a[:4] = [0, 1, P - 1, P - 2]; b[:4] = [P - 1, P - 1, P - 1, 2]
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    assert [int(x) for x in oracle.vec("orc_gl_mul_vec", a, b)] == [x * y % P for x, y in zip(ai, bi)]
    assert [int(x) for x in oracle.vec("orc_gl_add_vec", a, b)] == [(x + y) % P for x, y in zip(ai, bi)]
    assert [int(x) for x in oracle.vec("orc_gl_sub_vec", a, b)] == [(x - y) % P for x, y in zip(ai, bi)]
    inv = oracle.vec("orc_gl_inv_vec", a[1:])
    assert all(int(x) * int(y) % P == 1 for x, y in zip(a[1:], inv))
    e = oracle.vec("orc_gl2_mul_vec", a, b)
    for i in range(0, 64, 2):
        c0 = (ai[i] * bi[i] + 7 * ai[i + 1] * bi[i + 1]) % P
        c1 = (ai[i] * bi[i + 1] + ai[i + 1] * bi[i]) % P
        assert (int(e[i]), int(e[i + 1])) == (c0, c1)


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_ntt_definition_and_roundtrip(oracle, log_n):
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    a = rand_field(rng, n)
    f = oracle.ntt(a)
    w = oracle.omega(log_n)
    naive = [sum(int(a[i]) * pow(w, i * k, P) for i in range(n)) % P for k in range(n)]
    assert [int(x) for x in f] == naive
    assert (oracle.ntt(f, inverse=True) == a).all()


def test_lde_is_evaluation_on_bitreversed_cosets(oracle):
    log_n, log_lde = 6, 2
    n = 1 << log_n
    rng = np.random.default_rng(7)
    vals = rand_field(rng, (3, n))
    mono, lde = oracle.lde(vals, log_lde)
    w = oracle.omega(log_n)
    for col in range(3):
        coeffs = [int(x) for x in mono[col]]
        # monomials interpolate the values on H (natural order)
        for i in (0, 1, 17, n - 1):
            x = pow(w, i, P)
            assert sum(c * pow(x, k, P) for k, c in enumerate(coeffs)) % P == int(vals[col, i])
        # coset c of the LDE: shift 7*omega_{4n}^bitrev(c), position j <-> omega_n^bitrev(j)
        wl = oracle.omega(log_n + log_lde)
        for c in range(1 << log_lde):
            shift = 7 * pow(wl, brev(c, log_lde), P) % P
            assert shift == oracle.coset_shift(log_n, log_lde, c)
            for j in (0, 1, 2, 3, n - 1):
                x = shift * pow(w, brev(j, log_n), P) % P
                assert sum(cf * pow(x, k, P) for k, cf in enumerate(coeffs)) % P == int(lde[col, c * n + j])
    # x and -x adjacent inside a coset (the FRI pairing the golden proofs show)
    assert pow(w, brev(1, log_n), P) == P - 1


def test_poseidon2_structure(oracle):
    # permutation is a bijection-like map: distinct inputs -> distinct outputs, deterministic, zero state not fixed
    z = oracle.permute(np.zeros(12, dtype=np.uint64))[0]
    assert z.any()
    s = np.arange(12, dtype=np.uint64)
    assert (oracle.permute(s) == oracle.permute(s.copy())).all()
    assert (oracle.permute(s)[0] != z).any()
    # sponge framing: overwrite mode, zero padding of the tail, node = one permutation of left||right||0000
    els = np.arange(1, 12, dtype=np.uint64)
    st = np.zeros(12, dtype=np.uint64)
    st[:8] = els[:8]
    st = oracle.permute(st)[0]
    st[:8] = 0
    st[:3] = els[8:]
    st = oracle.permute(st)[0]
    assert (oracle.hash_leaf(els) == st[:4]).all()
    l, r = np.array([1, 2, 3, 4], dtype=np.uint64), np.array([5, 6, 7, 8], dtype=np.uint64)
    assert (oracle.hash_node(l, r) == oracle.permute(np.array([1, 2, 3, 4, 5, 6, 7, 8, 0, 0, 0, 0], dtype=np.uint64))[0][:4]).all()


@pytest.mark.parametrize("n_cols,epl,cap", [(1, 1, 1), (7, 1, 4), (8, 1, 16), (9, 1, 16), (2, 8, 4), (2, 4, 64)])
def test_merkle_paths_verify(oracle, n_cols, epl, cap):
    n_leaves = 64
    rng = np.random.default_rng(n_cols * 100 + epl)
    cols = rand_field(rng, (n_cols, n_leaves * epl))
    tree = oracle.merkle_build(cols, n_leaves, epl, cap)
    cap_digests = tree[2 * n_leaves - 2 * cap:]
    assert cap_digests.shape[0] == cap
    for idx in (0, 1, 31, 63):
        leaf = np.concatenate([cols[c, idx * epl:(idx + 1) * epl] for c in range(n_cols)])
        path = oracle.merkle_path(tree, n_leaves, cap, idx)
        assert oracle.merkle_verify(leaf, path, cap_digests, idx)
        bad = leaf.copy(); bad[0] ^= np.uint64(1)
        assert not oracle.merkle_verify(bad, path, cap_digests, idx)
        if path.shape[0]:
            assert not oracle.merkle_verify(leaf, path, cap_digests, idx ^ 1)


def test_fri_fold_array_matches_leaf_fold(oracle):
    log_dom = 9
    rng = np.random.default_rng(3)
    c0, c1 = rand_field(rng, 1 << log_dom), rand_field(rng, 1 << log_dom)
    ch = (123456789, 987654321)
    shift = oracle.pow(7, 4)
    a0, a1 = c0, c1
    c, s, ld = ch, shift, log_dom
    for _ in range(3):
        a0, a1 = oracle.fri_fold(a0, a1, ld, s, c)
        c = ((c[0] * c[0] + 7 * c[1] * c[1]) % P, 2 * c[0] * c[1] % P)
        s = s * s % P; ld -= 1
    for m in (0, 5, 63):
        v = oracle.fri_fold_leaf(c0[8 * m:8 * m + 8], c1[8 * m:8 * m + 8], log_dom, shift, 8 * m, ch)
        assert v == (int(a0[m]), int(a1[m]))


def test_fri_fold_degree_reduction(oracle):
    # folding evaluations of a degree < d polynomial gives evaluations of a degree < d/2 polynomial on the squared domain
    log_dom, d = 8, 32
    n = 1 << log_dom
    rng = np.random.default_rng(5)
    coef = np.zeros(n, dtype=np.uint64); coef[:d] = rand_field(rng, d)
    ev = oracle.coset_evals_bitrev(coef, 7)
    z = np.zeros(n, dtype=np.uint64)
    o0, o1 = oracle.fri_fold(ev, z, log_dom, 7, (5, 0))
    back = oracle.ntt(np.array([o0[brev(i, log_dom - 1)] for i in range(n // 2)], dtype=np.uint64), inverse=True)
    # undo the coset shift 49 = 7^2
    inv49 = pow(49, P - 2, P)
    mono = [int(back[i]) * pow(inv49, i, P) % P for i in range(n // 2)]
    assert all(m == 0 for m in mono[d // 2:])
    # f'(y) = 2*f_even(y) + c*2*f_odd(y)
    assert mono[0] == (2 * int(coef[0]) + 5 * 2 * int(coef[1])) % P
    assert not o1.any()


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLDEN, "fri_chain_*.json"))))
def test_golden_fri_chain(oracle, fixture):
    """Reference golden proofs (test_proofs/**): every FRI leaf of a query folds, with the proof's challenges, onto the
    right element of the next oracle's leaf, and the last one onto the final polynomial (bit-exact)."""
    fx = json.load(open(fixture))
    assert fx["all_queries_consistent"]
    sched, logd, chall = fx["schedule"], fx["log_domains"], fx["challenges"]
    mon = fx["final_fri_monomials"]
    for q in fx["queries"]:
        for k, s in enumerate(sched):
            le = q["fri_leaves"][k]
            h = len(le) // 2
            assert h == 1 << s
            m = q["leaf_indexes"][k]
            shift = oracle.pow(7, 1 << (logd[0] - logd[k]))
            got = oracle.fri_fold_leaf(le[:h], le[h:], logd[k], shift, m << s, chall[k])
            if k + 1 < len(sched):
                nxt = q["fri_leaves"][k + 1]
                hn = len(nxt) // 2
                assert q["leaf_indexes"][k + 1] == m >> sched[k + 1]
                pos = m & (hn - 1)
                assert got == (nxt[pos], nxt[hn + pos])
            else:
                ld = logd[-1]
                x = oracle.pow(7, 1 << (logd[0] - ld)) * pow(oracle.omega(ld), brev(m, ld), P) % P
                assert got == oracle.eval_ext_poly_at_base(mon[0], mon[1], x)


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLDEN, "deep_*.json"))))
def test_golden_deep_structure(oracle, fixture):
    """Reference golden proofs: with phi and z recovered WITHOUT the hash (tools/golden_deep.py), the oracle's DEEP combination
    of the four trace-oracle leaves of a query -- openings paired in the order of the proof's values_at_z (variables, plain
    witness, constants, sigmas, grand product + partial products, lookup multiplicities, lookup A polys and B, lookup table
    columns, quotient), the z*omega opening of the grand product, the lookup polys opened at 0, the public inputs opened at
    omega^row -- reproduces the value the proof holds in its FRI base oracle, bit for bit.  Covers lookup-free circuits (node,
    compression modes 1 and 2) and five base-layer circuit types with lookups of width 1, 3 and 4."""
    from era_zkevm_test_harness_b200 import geometry as G
    fx = json.load(open(fixture))
    assert fx["all_fixture_queries_consistent"]
    shapes = json.load(open(os.path.join(GOLDEN, "vk_shapes.json")))
    entry = shapes[fx["shape_key"][0]][fx["shape_key"][1]]
    if fx["shape_key"][0] == "compression":
        mode = entry["mode"]
        geo = G.geometry_from_vk(entry, G.COMPRESSION_GATE_ORDER[mode], has_boolean_col=1 if mode == 1 else 0)
    elif fx["shape_key"][0] == "base":
        geo = G.geometry_from_vk(entry, G.BASE_LAYER_GATE_ORDER[int(fx["shape_key"][1])])
        assert geo.lookup_reps > 0 and len(fx["values_at_0"]) == geo.lookup_reps + 1
    else:
        geo = G.geometry_from_vk(entry, G.RECURSION_GATE_ORDER)
    assert len(fx["values_at_z"]) == geo.n_witness + geo.n_setup + geo.n_stage2 // 2 + geo.n_quotient // 2
    # a stale golden proof (made with an older circuit layout than the VK) opens its public inputs on its own row, recovered
    # hash-free by tools/golden_deep_pi.py: the MainVM proof sits on row 1041222, the VK says 1033357
    for i, row in enumerate(fx.get("public_input_rows", [])):
        geo.pi_row[i] = row
    for q in fx["queries"]:
        assert len(q["witness"]) == geo.n_witness and len(q["setup"]) == geo.n_setup
        got = oracle.deep_at_point(geo, q["witness"], q["setup"], q["stage_2"], q["quotient"], fx["values_at_z"], fx["values_at_z_omega"][0],
                                   fx["values_at_0"], fx["public_inputs"], q["x"], fx["z"], fx["phi"])
        assert got == tuple(q["fri_base_value"])


def test_golden_lookup_sum_check_at_zero():
    """values_at_0 of every golden base-layer proof (fixture: the raw openings): the log-derivative lookup identity
    sum_i A_i(0) = B(0) holds with the openings in the order [A_0 .. A_{reps-1}, B] -- the order and the check of the verifier
    (csrc/host.cu "lookup sum check") -- without going through the hash."""
    fx = json.load(open(os.path.join(GOLDEN, "values_at_0.json")))
    assert len(fx) >= 18
    for name, vals in fx.items():
        assert len(vals) >= 2, name
        s0 = sum(v[0] for v in vals[:-1]) % P
        s1 = sum(v[1] for v in vals[:-1]) % P
        assert [s0, s1] == vals[-1], name


def test_copy_permutation_non_residues(oracle):
    """k_0 = 1, then the successive quadratic non-residues whose cosets k * H are new (boojum make_non_residues, recalled):
    the list is the same for every trace length used here, every k_i (i > 0) is a non-residue and the cosets are disjoint."""
    P = (1 << 64) - (1 << 32) + 1
    want = [1, 7, 11, 13, 14, 19, 21, 22, 26, 28, 31, 33, 35, 37, 38, 39, 42, 43, 44, 47]
    for log_n in (8, 16, 20):
        k = [int(x) for x in oracle.copy_permutation_non_residues(200, log_n)]
        assert k[:20] == want
        assert all(pow(x, (P - 1) // 2, P) == P - 1 for x in k[1:])
        assert len({pow(x, 1 << log_n, P) for x in k}) == len(k)
