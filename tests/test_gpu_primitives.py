"""GPU parity tests (through the C ABI) of the hand-written kernels against the CPU oracle: bit-exact."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from era_zkevm_test_harness_b200.context import to_device_u64, to_numpy_u64
from tests.oracle_lib import P, rand_field

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def brev(x, b):
    return int(format(x, "0%db" % b)[::-1], 2) if b else 0


@pytest.mark.parametrize("log_n", [0, 1, 3, 5, 8, 10, 11, 12, 13, 16])
@pytest.mark.parametrize("shift", [1, 7])
def test_ntt_forward_matches_oracle(gpu, oracle, log_n, shift):
    rng = np.random.default_rng(log_n)
    n_cols = 3 if log_n >= 12 else 5
    a = rand_field(rng, (n_cols, 1 << log_n))
    a[0, :min(4, 1 << log_n)] = [0, 1, P - 1, P - 2][:min(4, 1 << log_n)]
    got = to_numpy_u64(gpu.ntt_forward(to_device_u64(a, gpu.device), log_n, shift))
    assert (got == oracle.coset_evals_bitrev(a, shift)).all()


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 11, 12, 15, 16])
def test_ntt_inverse_matches_oracle(gpu, oracle, log_n):
    rng = np.random.default_rng(100 + log_n)
    a = rand_field(rng, (3, 1 << log_n))
    got = to_numpy_u64(gpu.ntt_inverse(to_device_u64(a, gpu.device), log_n))
    assert (got == oracle.ntt(a, inverse=True)).all()


def test_ntt_full_size_roundtrip_and_spot_values(gpu, oracle):
    """trace length 2^20 (BASELINE config 2): iNTT then coset NTT; one column checked in full against the oracle,
    all columns through the size-independent round-trip / linearity properties."""
    log_n, n_cols = 20, 6
    n = 1 << log_n
    rng = np.random.default_rng(20)
    vals = rand_field(rng, (n_cols, n))
    d_vals = to_device_u64(vals, gpu.device)
    mono = gpu.ntt_inverse(d_vals, log_n)
    ev = gpu.ntt_forward(mono, log_n, 1)  # evaluations on H, bit-reversed
    ev_np = to_numpy_u64(ev)
    idx = np.array([brev(i, log_n) for i in range(0, n, 4099)])
    assert (ev_np[:, idx] == vals[:, ::4099]).all()
    # full bit-reversal check on the device
    perm = torch.from_numpy(np.array([brev(i, 10) for i in range(1024)], dtype=np.int64)).to(gpu.device)
    i = torch.arange(n, device=gpu.device)
    full_perm = (perm[i & 1023] << 10) | perm[i >> 10]
    assert torch.equal(ev[:, full_perm], d_vals)
    # one column against the oracle, bit for bit, on a coset
    shift = oracle.coset_shift(log_n, 1, 1)
    got = to_numpy_u64(gpu.ntt_forward(mono[:1], log_n, shift))
    assert (got[0] == oracle.coset_evals_bitrev(to_numpy_u64(mono[:1])[0], shift)).all()
    # linearity: NTT(a + b) == NTT(a) + NTT(b)
    s = to_device_u64(oracle.vec("orc_gl_add_vec", to_numpy_u64(mono[0]), to_numpy_u64(mono[1])), gpu.device).reshape(1, n)
    lhs = to_numpy_u64(gpu.ntt_forward(s, log_n, 7))[0]
    two = to_numpy_u64(gpu.ntt_forward(mono[:2], log_n, 7))
    assert (lhs == oracle.vec("orc_gl_add_vec", two[0], two[1])).all()


def test_ntt_2pow20_fast_path_inverse_inplace_and_lde(gpu, oracle):
    """the 2^20 fast path (csrc/ntt1024.cu): inverse against the oracle bit for bit, forward in place == out of place,
    and the LDE entry point (monomials + two cosets) against the oracle for a column with edge values."""
    log_n, n = 20, 1 << 20
    rng = np.random.default_rng(2020)
    vals = rand_field(rng, (9, n))      # 9 columns: more members than one CTA tile row, odd batch
    vals[0, :4] = [0, 1, P - 1, P - 2]
    vals[1, :] = 0
    vals[2, :] = P - 1
    d_vals = to_device_u64(vals, gpu.device)
    mono = gpu.ntt_inverse(d_vals, log_n)
    mono_np = to_numpy_u64(mono)
    for c in (0, 1, 2, 8):
        assert (mono_np[c] == oracle.ntt(vals[c], inverse=True)).all(), c
    ref = gpu.ntt_forward(mono, log_n, 7)
    x = mono.clone()
    gpu.ntt_forward(x, log_n, 7, out=x)
    assert torch.equal(x, ref)
    m2, lde = gpu.lde(d_vals[:3], log_n, 1)
    o_mono, o_lde = oracle.lde(vals[:3], 1)
    assert (to_numpy_u64(m2) == o_mono).all()
    assert (to_numpy_u64(lde) == o_lde).all()


# the last three: compression-layer LDE factors (all cosets of an LDE go through one launch per NTT pass, grid.z = coset)
@pytest.mark.parametrize("log_n,log_lde,n_cols", [(4, 1, 2), (10, 1, 9), (12, 3, 3), (13, 1, 4), (5, 11, 3), (12, 6, 2), (13, 5, 2)])
def test_lde_matches_oracle(gpu, oracle, log_n, log_lde, n_cols):
    rng = np.random.default_rng(log_n * 10 + log_lde)
    vals = rand_field(rng, (n_cols, 1 << log_n))
    mono, lde = gpu.lde(to_device_u64(vals, gpu.device), log_n, log_lde)
    o_mono, o_lde = oracle.lde(vals, log_lde)
    assert (to_numpy_u64(mono) == o_mono).all()
    assert (to_numpy_u64(lde) == o_lde).all()


def test_poseidon2_permutation_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(2)
    st = rand_field(rng, (1000, 12))
    st[0] = 0
    st[1] = P - 1
    got = to_numpy_u64(gpu.poseidon2_permute(to_device_u64(st, gpu.device)))
    assert (got == oracle.permute(st)).all()


@pytest.mark.parametrize("n_cols,n_leaves,epl,cap", [(1, 2, 1, 1), (1, 64, 1, 64), (7, 256, 1, 16), (8, 256, 1, 16), (9, 512, 1, 16),
                                                     (156, 1024, 1, 16), (167, 512, 1, 32), (2, 128, 8, 16), (2, 16, 4, 16), (2, 64, 1, 256 // 4)])
def test_merkle_tree_matches_oracle(gpu, oracle, n_cols, n_leaves, epl, cap):
    """ragged leaf widths (not multiples of the rate 8), FRI-style multi-element leaves, cap == leaves edge case"""
    rng = np.random.default_rng(n_cols + n_leaves)
    cols = rand_field(rng, (n_cols, n_leaves * epl))
    tree = to_numpy_u64(gpu.merkle_build(to_device_u64(cols, gpu.device), n_leaves, epl, cap))
    assert (tree == oracle.merkle_build(cols, n_leaves, epl, cap)).all()


def test_merkle_full_size_property(gpu, oracle):
    """2^21 leaves x 16 columns (the quotient oracle of a 2^20 proof): random Merkle paths taken from the GPU tree verify
    under the oracle's verifier against the GPU cap."""
    n_leaves, n_cols, cap = 1 << 21, 16, 16
    rng = np.random.default_rng(21)
    cols = rand_field(rng, (n_cols, n_leaves))
    tree = to_numpy_u64(gpu.merkle_build(to_device_u64(cols, gpu.device), n_leaves, 1, cap))
    capd = tree[2 * n_leaves - 2 * cap:]
    for idx in (0, 1, 12345, n_leaves - 1, 1 << 20):
        path = oracle.merkle_path(tree, n_leaves, cap, idx)
        assert oracle.merkle_verify(cols[:, idx], path, capd, idx)


@pytest.mark.parametrize("log_dom", [1, 4, 12, 21])
def test_fri_fold_matches_oracle(gpu, oracle, log_dom):
    rng = np.random.default_rng(log_dom)
    n = 1 << log_dom
    c0, c1 = rand_field(rng, n), rand_field(rng, n)
    ch = (int(rand_field(rng, 1)[0]), int(rand_field(rng, 1)[0]))
    shift = oracle.pow(7, 1 << 3)
    g0, g1 = gpu.fri_fold(to_device_u64(c0, gpu.device), to_device_u64(c1, gpu.device), log_dom, shift, ch)
    o0, o1 = oracle.fri_fold(c0, c1, log_dom, shift, ch)
    assert (to_numpy_u64(g0) == o0).all() and (to_numpy_u64(g1) == o1).all()


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLDEN, "fri_chain_*.json"))))
def test_fri_fold_kernel_on_reference_golden_proofs(gpu, oracle, fixture):
    """The GPU fold kernel, fed the FRI leaves of the reference's golden proofs at their true positions in the LDE domain,
    reproduces the next oracle's leaf element / the final polynomial value."""
    fx = json.load(open(fixture))
    sched, logd, chall = fx["schedule"], fx["log_domains"], fx["challenges"]
    mon = fx["final_fri_monomials"]
    for q in fx["queries"][:4]:
        for k, s in enumerate(sched):
            le = q["fri_leaves"][k]
            h = len(le) // 2
            m = q["leaf_indexes"][k]
            n = 1 << logd[k]
            c0 = np.zeros(n, dtype=np.uint64); c1 = np.zeros(n, dtype=np.uint64)
            c0[m * h:(m + 1) * h] = le[:h]; c1[m * h:(m + 1) * h] = le[h:]
            d0, d1 = to_device_u64(c0, gpu.device), to_device_u64(c1, gpu.device)
            c = tuple(chall[k]); ld = logd[k]; shift = oracle.pow(7, 1 << (logd[0] - ld))
            for _ in range(s):
                d0, d1 = gpu.fri_fold(d0, d1, ld, shift, c)
                c = ((c[0] * c[0] + 7 * c[1] * c[1]) % P, 2 * c[0] * c[1] % P)
                shift = shift * shift % P; ld -= 1
            got = (int(to_numpy_u64(d0)[m]), int(to_numpy_u64(d1)[m]))
            if k + 1 < len(sched):
                nxt = q["fri_leaves"][k + 1]; hn = len(nxt) // 2; pos = m & (hn - 1)
                assert got == (nxt[pos], nxt[hn + pos])
            else:
                x = oracle.pow(7, 1 << (logd[0] - logd[-1])) * pow(oracle.omega(logd[-1]), brev(m, logd[-1]), P) % P
                assert got == oracle.eval_ext_poly_at_base(mon[0], mon[1], x)


def test_commit_columns_host_matches_oracle(gpu, oracle):
    log_n, log_lde, n_cols, cap = 10, 1, 11, 16
    rng = np.random.default_rng(9)
    vals = rand_field(rng, (n_cols, 1 << log_n))
    h = torch.from_numpy(vals.view(np.int64)).pin_memory()
    cap_gpu = gpu.commit_columns_host(h, log_n, log_lde, cap).numpy().view(np.uint64)
    _, lde = oracle.lde(vals, log_lde)
    n_leaves = 1 << (log_n + log_lde)
    tree = oracle.merkle_build(lde, n_leaves, 1, cap)
    assert (cap_gpu == tree[2 * n_leaves - 2 * cap:]).all()


@pytest.mark.parametrize("pair", ["mainvm", "compression_1", "base_13", "node_3"])
def test_hash_kernels_reproduce_the_reference_digests(gpu, pair):
    """The CUDA leaf- and node-hash kernels against the REFERENCE's own digests (not only against the oracle): the four trace-oracle
    leaves and the first FRI leaf of golden queries (tests/golden/pair_*.json) are hashed on the GPU, walked up their golden Merkle
    paths with the GPU node hash, and must land on the cap entries of the proof / verification key."""
    d = json.load(open(os.path.join(GOLDEN, f"pair_{pair}.json")))
    vk, pr = d["vk"], d["proof"]
    caps = {"witness_query": pr["witness_oracle_cap"], "stage_2_query": pr["stage_2_oracle_cap"], "quotient_query": pr["quotient_oracle_cap"],
            "setup_query": vk["setup_merkle_tree_cap"]}
    queries = pr["queries_per_fri_repetition"]

    def gpu_leaf_hashes(leaves, epl=1):
        # leaves: K lists of equal length L = n_cols * epl; column c of leaf i holds elements [c*epl, (c+1)*epl) of the leaf
        k, length = len(leaves), len(leaves[0])
        n_cols, n_leaves = length // epl, 4
        cols = np.zeros((n_cols, n_leaves * epl), dtype=np.uint64)
        for i, leaf in enumerate(leaves):
            cols[:, i * epl:(i + 1) * epl] = np.array(leaf, dtype=np.uint64).reshape(n_cols, epl)
        tree = to_numpy_u64(gpu.merkle_build(to_device_u64(cols, gpu.device), n_leaves, epl, n_leaves))
        return tree[:k]

    def gpu_node(left, right):
        cols = np.zeros((1, 8), dtype=np.uint64)
        # a 2-leaf tree with cap 1 hashes its two leaf DIGESTS into the root: feed the digests as precomputed leaves is not exposed,
        # so use the 8-element leaf form, which is one permutation of [left, right, 0, 0, 0, 0] -- the node hash itself
        cols[0, :4], cols[0, 4:] = left, right
        return to_numpy_u64(gpu.merkle_build(to_device_u64(cols.reshape(8, 1).copy(), gpu.device), 1, 1, 1))[0]

    # index of every query from the witness opening (the proof stores none)
    from tests import oracle_lib
    orc = oracle_lib.load()
    for name, cap in caps.items():
        digs = gpu_leaf_hashes([q[name]["leaf_elements"] for q in queries])
        for q, h in zip(queries, digs):
            idx = orc.merkle_find_index(q["witness_query"]["leaf_elements"], q["witness_query"]["proof"], np.array(pr["witness_oracle_cap"], dtype=np.uint64))
            assert idx is not None
            for sib in q[name]["proof"]:
                sib = np.array(sib, dtype=np.uint64)
                h = gpu_node(h, sib) if (idx & 1) == 0 else gpu_node(sib, h)
                idx >>= 1
            assert [int(x) for x in h] == cap[idx], (pair, name)
