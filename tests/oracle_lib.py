"""ctypes view of oracle/liboracle.so (CPU restatement of the reference algorithm -- the checker, test-only)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
P = (1 << 64) - (1 << 32) + 1
c_u64 = ctypes.c_uint64
vp = ctypes.c_void_p
sz = ctypes.c_size_t
ci = ctypes.c_int


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_gl_pow.restype = c_u64; L.orc_gl_pow.argtypes = [c_u64, c_u64]
        L.orc_gl_omega.restype = c_u64; L.orc_gl_omega.argtypes = [ci]
        L.orc_lde_coset_shift.restype = c_u64; L.orc_lde_coset_shift.argtypes = [ci, ci, ctypes.c_uint32]
        L.orc_ntt.argtypes = [vp, ci, ci]
        L.orc_bitrev.argtypes = [vp, ci]
        L.orc_coset_evals_bitrev.argtypes = [vp, ci, c_u64, vp]
        L.orc_lde_from_values.argtypes = [vp, ci, ci, vp, vp]
        L.orc_poseidon2_permute.argtypes = [vp]
        L.orc_hash_leaf.argtypes = [vp, sz, vp]
        L.orc_hash_node.argtypes = [vp, vp, vp]
        L.orc_merkle_build.argtypes = [vp, sz, sz, sz, sz, sz, vp]
        L.orc_merkle_path.argtypes = [vp, sz, sz, sz, vp]
        L.orc_merkle_verify.restype = ci; L.orc_merkle_verify.argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_merkle_find_index.restype = sz; L.orc_merkle_find_index.argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_copy_permutation_non_residues.argtypes = [vp, ctypes.c_uint32, ci]
        L.orc_fri_fold.argtypes = [vp, vp, ci, c_u64, vp, vp, vp]
        L.orc_fri_fold_leaf.argtypes = [vp, vp, sz, ci, c_u64, sz, vp, vp]
        L.orc_eval_ext_poly_at_base.argtypes = [vp, vp, sz, c_u64, vp]
        for nm in ("orc_gl_mul_vec", "orc_gl_add_vec", "orc_gl_sub_vec", "orc_gl2_mul_vec"):
            getattr(L, nm).argtypes = [vp, vp, vp, sz]
        L.orc_gl_inv_vec.argtypes = [vp, vp, sz]
        L.orc_proof_size_u64.restype = sz; L.orc_proof_size_u64.argtypes = [vp, vp]
        L.orc_setup_cap.argtypes = [vp, vp, vp, vp]
        L.orc_prove.restype = ctypes.c_long; L.orc_prove.argtypes = [vp, vp, vp, vp, vp, sz]
        L.orc_deep_at_point.argtypes = [vp] * 9 + [c_u64, vp, vp, vp]
        for nm in ("orc_num_witness_cols", "orc_num_setup_cols", "orc_num_stage2_cols"):
            getattr(L, nm).restype = ctypes.c_uint32; getattr(L, nm).argtypes = [vp]
        L.orc_gl2_inv_vec.argtypes = [vp, vp, sz]
        L.orc_verify.restype = ci; L.orc_verify.argtypes = [vp, vp, vp, vp, sz, ctypes.c_char_p, sz]
        L.orc_verify_ex.restype = ci; L.orc_verify_ex.argtypes = [vp, vp, vp, vp, sz, ctypes.c_uint, ctypes.c_char_p, sz]
        L.orc_set_threads.restype = ci; L.orc_set_threads.argtypes = [ci]
        L.orc_synth_trace.restype = ci; L.orc_synth_trace.argtypes = [vp, c_u64, c_u64, ci, vp, vp]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(vp)

    def omega(self, log_n):
        return int(self.lib.orc_gl_omega(log_n))

    def pow(self, b, e):
        return int(self.lib.orc_gl_pow(b, e))

    def coset_shift(self, log_n, log_lde, c):
        return int(self.lib.orc_lde_coset_shift(log_n, log_lde, c))

    def ntt(self, a, inverse=False):
        a = np.array(a, dtype=np.uint64)
        log_n = int(a.shape[-1]).bit_length() - 1
        flat = a.reshape(-1, a.shape[-1])
        for row in flat:
            self.lib.orc_ntt(self._p(row), log_n, 1 if inverse else 0)
        return a

    def coset_evals_bitrev(self, mono, shift):
        mono = np.ascontiguousarray(mono, dtype=np.uint64)
        out = np.empty_like(mono)
        log_n = int(mono.shape[-1]).bit_length() - 1
        m2, o2 = mono.reshape(-1, mono.shape[-1]), out.reshape(-1, mono.shape[-1])
        for i in range(m2.shape[0]):
            self.lib.orc_coset_evals_bitrev(self._p(m2[i]), log_n, shift, self._p(o2[i]))
        return out

    def lde(self, vals, log_lde):
        vals = np.ascontiguousarray(vals, dtype=np.uint64)
        n_cols, n = vals.shape
        log_n = n.bit_length() - 1
        out = np.empty((n_cols, n << log_lde), dtype=np.uint64)
        mono = np.empty_like(vals)
        for i in range(n_cols):
            self.lib.orc_lde_from_values(self._p(vals[i]), log_n, log_lde, self._p(out[i]), self._p(mono[i]))
        return mono, out

    def permute(self, states):
        s = np.array(states, dtype=np.uint64).reshape(-1, 12)
        for row in s:
            self.lib.orc_poseidon2_permute(self._p(row))
        return s

    def hash_leaf(self, els):
        els = np.ascontiguousarray(els, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_hash_leaf(self._p(els), els.size, self._p(out))
        return out

    def hash_node(self, l, r):
        l = np.ascontiguousarray(l, dtype=np.uint64); r = np.ascontiguousarray(r, dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        self.lib.orc_hash_node(self._p(l), self._p(r), self._p(out))
        return out

    def merkle_build(self, cols, n_leaves, elems_per_leaf, cap_size):
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        n_cols, stride = cols.shape
        tree = np.empty((2 * n_leaves - cap_size, 4), dtype=np.uint64)
        self.lib.orc_merkle_build(self._p(cols), stride, n_cols, n_leaves, elems_per_leaf, cap_size, self._p(tree))
        return tree

    def merkle_path(self, tree, n_leaves, cap_size, idx):
        k = (n_leaves // cap_size).bit_length() - 1
        out = np.empty((k, 4), dtype=np.uint64)
        self.lib.orc_merkle_path(self._p(tree), n_leaves, cap_size, idx, self._p(out))
        return out

    def merkle_verify(self, leaf, path, cap, idx):
        leaf = np.ascontiguousarray(leaf, dtype=np.uint64); path = np.ascontiguousarray(path, dtype=np.uint64)
        cap = np.ascontiguousarray(cap, dtype=np.uint64)
        return bool(self.lib.orc_merkle_verify(self._p(leaf), leaf.size, self._p(path), path.shape[0], self._p(cap), idx))

    def merkle_find_index(self, leaf, path, cap):
        """index of an opening whose index is not stored (oracle/primitives.c orc_merkle_find_index); None when no pattern fits"""
        leaf = np.ascontiguousarray(leaf, dtype=np.uint64); path = np.ascontiguousarray(path, dtype=np.uint64).reshape(-1)
        cap = np.ascontiguousarray(cap, dtype=np.uint64).reshape(-1)
        r = self.lib.orc_merkle_find_index(self._p(leaf), leaf.size, self._p(path), path.size // 4, self._p(cap), cap.size // 4)
        return None if r == (1 << 64) - 1 else int(r)

    def copy_permutation_non_residues(self, n, log_n):
        k = np.zeros(n, dtype=np.uint64)
        self.lib.orc_copy_permutation_non_residues(self._p(k), n, log_n)
        return k

    def fri_fold(self, c0, c1, log_dom, shift, ch):
        c0 = np.ascontiguousarray(c0, dtype=np.uint64); c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        half = 1 << (log_dom - 1)
        o0 = np.empty(half, dtype=np.uint64); o1 = np.empty(half, dtype=np.uint64)
        chv = np.array(ch, dtype=np.uint64)
        self.lib.orc_fri_fold(self._p(c0), self._p(c1), log_dom, shift, self._p(chv), self._p(o0), self._p(o1))
        return o0, o1

    def fri_fold_leaf(self, c0, c1, log_dom, shift, base_idx, ch):
        c0 = np.ascontiguousarray(c0, dtype=np.uint64); c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        out = np.empty(2, dtype=np.uint64); chv = np.array(ch, dtype=np.uint64)
        self.lib.orc_fri_fold_leaf(self._p(c0), self._p(c1), c0.size, log_dom, shift, base_idx, self._p(chv), self._p(out))
        return int(out[0]), int(out[1])

    def eval_ext_poly_at_base(self, c0, c1, x):
        c0 = np.ascontiguousarray(c0, dtype=np.uint64); c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        out = np.empty(2, dtype=np.uint64)
        self.lib.orc_eval_ext_poly_at_base(self._p(c0), self._p(c1), c0.size, x, self._p(out))
        return int(out[0]), int(out[1])

    # ---- whole-prover restatement (oracle/prover.c); geo/cfg are ctypes structs of include/zkgpu.h
    def setup_cap(self, geo, cfg, setup_cols):
        setup_cols = np.ascontiguousarray(setup_cols, dtype=np.uint64)
        cap = np.empty((cfg.cap_size, 4), dtype=np.uint64)
        self.lib.orc_setup_cap(ctypes.byref(geo), ctypes.byref(cfg), self._p(setup_cols), self._p(cap))
        return cap

    def prove(self, geo, cfg, wit_cols, setup_cols):
        wit_cols = np.ascontiguousarray(wit_cols, dtype=np.uint64); setup_cols = np.ascontiguousarray(setup_cols, dtype=np.uint64)
        n = int(self.lib.orc_proof_size_u64(ctypes.byref(geo), ctypes.byref(cfg)))
        proof = np.zeros(n, dtype=np.uint64)
        w = self.lib.orc_prove(ctypes.byref(geo), ctypes.byref(cfg), self._p(wit_cols), self._p(setup_cols), self._p(proof), n)
        assert w == n, (w, n)
        return proof

    def verify(self, geo, cfg, vk_cap, proof, skip_quotient_identity=False):
        """The oracle's own verifier (oracle/prover.c orc_verify), independent of the product's zkgpu_verify: (ok, message)."""
        cap = np.ascontiguousarray(vk_cap, dtype=np.uint64); pr = np.ascontiguousarray(proof, dtype=np.uint64)
        buf = ctypes.create_string_buffer(256)
        rc = self.lib.orc_verify_ex(ctypes.byref(geo), ctypes.byref(cfg), self._p(cap), self._p(pr), pr.size,
                                    1 if skip_quotient_identity else 0, buf, 256)
        return rc == 0, buf.value.decode()

    def set_threads(self, n):
        return int(self.lib.orc_set_threads(int(n)))

    def synth_trace(self, geo, seed=0, witness_seed=None):
        """Synthetic satisfying trace (oracle/synth.c): (witness_cols [W,n], setup_cols [S,n]), the generator of
        era_zkevm_test_harness_b200.prover_utils.synth_trace restated in C without the product library."""
        n = 1 << geo.log_n
        wit = np.empty((int(self.lib.orc_num_witness_cols(ctypes.byref(geo))), n), dtype=np.uint64)
        setup = np.empty((int(self.lib.orc_num_setup_cols(ctypes.byref(geo))), n), dtype=np.uint64)
        if witness_seed is None:
            self.lib.orc_synth_trace(ctypes.byref(geo), seed, seed, 0, self._p(wit), self._p(setup))
        else:
            self.lib.orc_synth_trace(ctypes.byref(geo), seed, witness_seed, 1, self._p(wit), self._p(setup))
        return wit, setup

    def deep_at_point(self, geo, wl, sl, l2, lq, at_z, at_zw, at_0, pi, x, z, phi):
        """DEEP combination at one LDE point (oracle/prover.c deep_point): leaves of the four trace oracles at x, openings in
        the proof's order"""
        a = [np.ascontiguousarray(v, dtype=np.uint64).reshape(-1) for v in (wl, sl, l2, lq, at_z, at_zw, at_0 if len(at_0) else [0], pi if len(pi) else [0])]
        zz, pp, out = np.array(z, dtype=np.uint64), np.array(phi, dtype=np.uint64), np.zeros(2, dtype=np.uint64)
        self.lib.orc_deep_at_point(ctypes.byref(geo), *[self._p(v) for v in a], c_u64(int(x)), self._p(zz), self._p(pp), self._p(out))
        return (int(out[0]), int(out[1]))

    def vec(self, name, *arrs):
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in arrs]
        out = np.empty_like(arrs[0])
        n = arrs[0].size if "gl2" not in name else arrs[0].size // 2
        getattr(self.lib, name)(*[self._p(a) for a in arrs], self._p(out), n)
        return out


_ORACLE = None


def load():
    global _ORACLE
    if _ORACLE is None:
        build()
        _ORACLE = Oracle(ctypes.CDLL(LIB))
    return _ORACLE


def rand_field(rng, shape):
    """uniform-ish canonical field elements (rejection of the 2^-32 tail is irrelevant: reduce mod p)"""
    a = rng.integers(0, 1 << 63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
    return np.where(a >= np.uint64(P), a - np.uint64(P), a)


# ---------------------------------------------------------------------------------------------------------------------
# hostcheck: the HOST halves of the product's device headers (glx.cuh lazy arithmetic, poseidon2_core.cuh) compiled with
# g++ so the CPU suite can check them against the oracle without a GPU.
HOSTCHECK_DIR = os.path.join(ROOT, "tests", "hostcheck")
HOSTCHECK_LIB = os.path.join(HOSTCHECK_DIR, "libhostcheck.so")


def build_hostcheck(force=False):
    src = os.path.join(HOSTCHECK_DIR, "hostcheck.cpp")
    csrc = os.path.join(ROOT, "era_zkevm_test_harness_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if force or not os.path.exists(HOSTCHECK_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOSTCHECK_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I/usr/local/cuda/include",
                               "-o", HOSTCHECK_LIB, src])
    return ctypes.CDLL(HOSTCHECK_LIB)
