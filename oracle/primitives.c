/* oracle/primitives.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; never on the product path).
 *
 * CPU restatement of the arithmetic kernels behind the reference's prover call
 *   cs.prove_from_precomputations::<EXT, TR, H, POW>(..)   /root/reference/src/prover_utils.rs:338-348, :533-543
 *   cs.get_full_setup(worker, lde, cap)                     /root/reference/src/prover_utils.rs:185-186
 * The algorithm lives in the third-party crate `boojum` (github.com/matter-labs/era-boojum, branch `main`, NOT
 * vendored, no Cargo.lock in the reference: /root/reference/kzg/Cargo.toml:16, circuit_definitions/Cargo.toml:17),
 * so every function below restates boojum's published algorithm and is anchored on the reference's call sites
 * and golden artefacts (setup and test_proofs JSON trees under /root/reference).
 *
 * PARITY STATUS
 *   - field / extension / roots of unity / coset shift / bit-reversed LDE enumeration / un-normalised FRI fold with
 *     c, c^2, c^4: PINNED hash-free against golden proofs (tests/test_golden_fri.py on tests/golden fixtures).
 *   - Poseidon2 permutation, sponge framing, leaf / node hashing, Merkle caps: PINNED on the golden proofs.  Round constants =
 *     the plonky2 table of the Crandall-prime era (tools/gen_poseidon_constants.py), M4 external matrix, 2^s internal
 *     diagonal, zero-padded overwrite sponge: every Merkle path of every query of 21 golden (proof, VK) pairs verifies
 *     (tests/test_hash_pin_cpu.py, tests/test_golden_verify_cpu.py, tools/golden_verify.py).
 */
#include "gl64.h"
#include "poseidon2_consts.h"
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* OpenMP team size of every orc_* call; bench.py sets it explicitly (torchrun exports OMP_NUM_THREADS=1). Returns the team size. */
#ifdef _OPENMP
#include <omp.h>
EXPORT int orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
EXPORT int orc_set_threads(int n) { (void)n; return 1; }
#endif

/* ------------------------------------------------------------------ field vectors (for parity tests) */
EXPORT void orc_gl_mul_vec(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = gl_mul(a[i], b[i]); }
EXPORT void orc_gl_add_vec(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = gl_add(a[i], b[i]); }
EXPORT void orc_gl_sub_vec(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = gl_sub(a[i], b[i]); }
EXPORT void orc_gl_inv_vec(const uint64_t *a, uint64_t *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = gl_inv(a[i]); }
EXPORT uint64_t orc_gl_pow(uint64_t b, uint64_t e) { return gl_pow(b, e); }
EXPORT uint64_t orc_gl_omega(int log_n) { return gl_omega(log_n); }
EXPORT void orc_gl2_mul_vec(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        gl2 r = gl2_mul(gl2_make(a[2 * i], a[2 * i + 1]), gl2_make(b[2 * i], b[2 * i + 1]));
        o[2 * i] = r.c0; o[2 * i + 1] = r.c1;
    }
}
EXPORT void orc_gl2_inv_vec(const uint64_t *a, uint64_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        gl2 r = gl2_inv(gl2_make(a[2 * i], a[2 * i + 1]));
        o[2 * i] = r.c0; o[2 * i + 1] = r.c1;
    }
}

/* ------------------------------------------------------------------ NTT */
EXPORT void orc_bitrev(uint64_t *a, int log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev32((uint32_t)i, log_n);
        if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
}

/* natural order in, natural order out.  forward: X[k] = sum_i x[i] w^(ik), w = omega_{2^log_n};
 * inverse: x[i] = n^-1 sum_k X[k] w^(-ik). */
EXPORT void orc_ntt(uint64_t *a, int log_n, int inverse) {
    size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    uint64_t w = gl_omega(log_n);
    if (inverse) w = gl_inv(w);
    uint64_t *tw = (uint64_t *)malloc(sizeof(uint64_t) * (n / 2 ? n / 2 : 1));
    tw[0] = 1;
    for (size_t i = 1; i < n / 2; i++) tw[i] = gl_mul(tw[i - 1], w);
    orc_bitrev(a, log_n);
    for (int s = 1; s <= log_n; s++) {
        size_t m = (size_t)1 << s, half = m >> 1, step = n >> s;
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < half; j++) {
                uint64_t t = gl_mul(a[k + j + half], tw[j * step]);
                uint64_t u = a[k + j];
                a[k + j] = gl_add(u, t);
                a[k + j + half] = gl_sub(u, t);
            }
    }
    if (inverse) {
        uint64_t ninv = gl_inv((uint64_t)n % GL_P);
        for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv);
    }
    free(tw);
}

/* Evaluate a polynomial (monomial coefficients, natural order, length n) on the coset shift*<omega_n>, returning the
 * values in BIT-REVERSED enumeration: out[j] = f(shift * omega_n^bitrev(j)).  This is the storage order of every
 * committed oracle (x and -x adjacent; confirmed on golden FRI leaves). */
EXPORT void orc_coset_evals_bitrev(const uint64_t *mono, int log_n, uint64_t shift, uint64_t *out) {
    size_t n = (size_t)1 << log_n;
    uint64_t s = 1;
    for (size_t i = 0; i < n; i++) { out[i] = gl_mul(mono[i], s); s = gl_mul(s, shift); }
    orc_ntt(out, log_n, 0);
    orc_bitrev(out, log_n);
}

/* coset shift of coset c of an LDE by factor 2^log_lde over a trace domain 2^log_n:
 * 7 * omega_{lde*n}^bitrev(c)  (coset index in the TOP bits of the bit-reversed LDE index) */
EXPORT uint64_t orc_lde_coset_shift(int log_n, int log_lde, uint32_t c) {
    uint64_t w = gl_omega(log_n + log_lde);
    return gl_mul(GL_GEN, gl_pow(w, bitrev32(c, log_lde)));
}

/* values on H (natural order) -> LDE by 2^log_lde, coset-major, each coset bit-reversed.
 * out has (n << log_lde) elements. mono_out (optional, n elements) receives the monomial form. */
EXPORT void orc_lde_from_values(const uint64_t *vals, int log_n, int log_lde, uint64_t *out, uint64_t *mono_out) {
    size_t n = (size_t)1 << log_n;
    uint64_t *mono = (uint64_t *)malloc(sizeof(uint64_t) * n);
    memcpy(mono, vals, sizeof(uint64_t) * n);
    orc_ntt(mono, log_n, 1);
    for (uint32_t c = 0; c < (1u << log_lde); c++)
        orc_coset_evals_bitrev(mono, log_n, orc_lde_coset_shift(log_n, log_lde, c), out + (size_t)c * n);
    if (mono_out) memcpy(mono_out, mono, sizeof(uint64_t) * n);
    free(mono);
}

/* ------------------------------------------------------------------ Poseidon2 (width 12, rate 8, x^7, 4+22+4) */
static const int P2_SHIFTS[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12}; /* M_I = J + diag(2^s) */

static inline uint64_t p2_sbox(uint64_t x) {
    uint64_t x2 = gl_sqr(x), x4 = gl_sqr(x2);
    return gl_mul(gl_mul(x4, x2), x);
}
static void p2_external(uint64_t *s) { /* circ(2*M4, M4, M4), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] */
    for (int b = 0; b < 3; b++) {
        uint64_t *x = s + 4 * b;
        uint64_t t0 = gl_add(x[0], x[1]), t1 = gl_add(x[2], x[3]);
        uint64_t t2 = gl_add(gl_add(x[1], x[1]), t1), t3 = gl_add(gl_add(x[3], x[3]), t0);
        uint64_t t1_4 = gl_add(gl_add(t1, t1), gl_add(t1, t1)), t0_4 = gl_add(gl_add(t0, t0), gl_add(t0, t0));
        uint64_t t4 = gl_add(t1_4, t3), t5 = gl_add(t0_4, t2);
        uint64_t t6 = gl_add(t3, t5), t7 = gl_add(t2, t4);
        x[0] = t6; x[1] = t5; x[2] = t7; x[3] = t4;
    }
    for (int i = 0; i < 4; i++) {
        uint64_t t = gl_add(gl_add(s[i], s[4 + i]), s[8 + i]);
        s[i] = gl_add(s[i], t); s[4 + i] = gl_add(s[4 + i], t); s[8 + i] = gl_add(s[8 + i], t);
    }
}
static void p2_internal(uint64_t *s) {
    uint64_t sum = 0;
    for (int i = 0; i < 12; i++) sum = gl_add(sum, s[i]);
    for (int i = 0; i < 12; i++) s[i] = gl_add(gl_mul(s[i], (uint64_t)1 << P2_SHIFTS[i]), sum);
}
EXPORT void orc_poseidon2_permute(uint64_t *s) {
    p2_external(s);
    int r = 0;
    for (int k = 0; k < 4; k++, r++) {
        for (int i = 0; i < 12; i++) s[i] = p2_sbox(gl_add(s[i], ORC_P2_RC[12 * r + i]));
        p2_external(s);
    }
    for (int k = 0; k < 22; k++, r++) {
        s[0] = p2_sbox(gl_add(s[0], ORC_P2_RC[12 * r]));
        p2_internal(s);
    }
    for (int k = 0; k < 4; k++, r++) {
        for (int i = 0; i < 12; i++) s[i] = p2_sbox(gl_add(s[i], ORC_P2_RC[12 * r + i]));
        p2_external(s);
    }
}

/* sponge, rate 8, "overwrite" absorption (GoldilocksPoseidon2Sponge<AbsorptionModeOverwrite>,
 * /root/reference/src/prover_utils.rs:43): chunks of 8 replace lanes 0..7, the last partial chunk is zero padded,
 * digest = lanes 0..3. */
EXPORT void orc_hash_leaf(const uint64_t *els, size_t n, uint64_t out[4]) {
    uint64_t st[12] = {0};
    size_t i = 0;
    while (i < n) {
        size_t take = n - i < 8 ? n - i : 8;
        for (size_t k = 0; k < 8; k++) st[k] = k < take ? els[i + k] : 0;
        orc_poseidon2_permute(st);
        i += take;
    }
    memcpy(out, st, 32);
}
EXPORT void orc_hash_node(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    orc_poseidon2_permute(st);
    memcpy(out, st, 32);
}

/* ------------------------------------------------------------------ Merkle tree with cap */
/* cols: n_cols columns, column c at cols + c*col_stride, each holding n_leaves*elems_per_leaf values.
 * leaf i = for each column c, the elems_per_leaf consecutive values starting at i*elems_per_leaf (column-major
 * concatenation: trace oracles use elems_per_leaf = 1; an FRI oracle stores (c0 column, c1 column) with
 * elems_per_leaf = 2^fold so that leaf_elements = [c0 x 2^k, c1 x 2^k] as in the golden proofs).
 * tree_out: levels concatenated from the leaf hashes (n_leaves digests) down to the cap level (cap_size digests):
 * total (2*n_leaves - cap_size) digests of 4 u64.  n_leaves and cap_size are powers of two, cap_size <= n_leaves. */
EXPORT void orc_merkle_build(const uint64_t *cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf,
                             size_t cap_size, uint64_t *tree_out) {
    size_t leaf_len = n_cols * elems_per_leaf;
#pragma omp parallel
    {
        uint64_t *buf = (uint64_t *)malloc(sizeof(uint64_t) * (leaf_len ? leaf_len : 1));
#pragma omp for schedule(static)
        for (size_t i = 0; i < n_leaves; i++) {
            for (size_t c = 0; c < n_cols; c++)
                for (size_t e = 0; e < elems_per_leaf; e++) buf[c * elems_per_leaf + e] = cols[c * col_stride + i * elems_per_leaf + e];
            orc_hash_leaf(buf, leaf_len, tree_out + 4 * i);
        }
        free(buf);
    }
    uint64_t *prev = tree_out;
    size_t width = n_leaves;
    while (width > cap_size) {
        uint64_t *next = prev + 4 * width;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < width / 2; i++) orc_hash_node(prev + 8 * i, prev + 8 * i + 4, next + 4 * i);
        prev = next;
        width >>= 1;
    }
}
/* offset (in digests) of the cap inside tree_out */
EXPORT size_t orc_merkle_cap_offset(size_t n_leaves, size_t cap_size) { return 2 * n_leaves - 2 * cap_size; }

/* Merkle path of leaf idx: siblings leaf->root, (log2(n_leaves/cap_size)) digests */
EXPORT void orc_merkle_path(const uint64_t *tree, size_t n_leaves, size_t cap_size, size_t idx, uint64_t *path_out) {
    const uint64_t *lvl = tree;
    size_t width = n_leaves, k = 0;
    while (width > cap_size) {
        memcpy(path_out + 4 * k, lvl + 4 * (idx ^ 1), 32);
        lvl += 4 * width;
        width >>= 1; idx >>= 1; k++;
    }
}
EXPORT int orc_merkle_verify(const uint64_t *leaf_els, size_t leaf_len, const uint64_t *path, size_t path_len, const uint64_t *cap,
                             size_t idx) {
    uint64_t cur[4];
    orc_hash_leaf(leaf_els, leaf_len, cur);
    for (size_t k = 0; k < path_len; k++) {
        if (idx & 1) orc_hash_node(path + 4 * k, cur, cur);
        else orc_hash_node(cur, path + 4 * k, cur);
        idx >>= 1;
    }
    return memcmp(cur, cap + 4 * idx, 32) == 0;
}

/* Recovers the leaf index of a Merkle opening whose index is not stored (boojum's OracleQuery holds leaf + path only): all
 * 2^path_len left/right patterns share prefixes, so the candidate roots cost 2^(path_len+1) node hashes; the pattern whose root
 * is a cap entry gives idx = (cap position << path_len) | pattern.  Used by tools/golden_recover_vk_cap.py to rebuild the setup
 * caps (= verification keys) the stale golden proofs were made with.  Returns the index or (size_t)-1. */
EXPORT size_t orc_merkle_find_index(const uint64_t *leaf_els, size_t leaf_len, const uint64_t *path, size_t path_len, const uint64_t *cap,
                                    size_t cap_size) {
    size_t n = 1;
    uint64_t *cur = (uint64_t *)malloc(32), *nxt;
    orc_hash_leaf(leaf_els, leaf_len, cur);
    for (size_t k = 0; k < path_len; k++) {   /* candidate j at level k -> 2j (we are the left child... bit k = 0) and 2j+1 */
        nxt = (uint64_t *)malloc(64 * n);
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < n; j++) {
            orc_hash_node(cur + 4 * j, path + 4 * k, nxt + 4 * j);            /* bit k of the index = 0 */
            orc_hash_node(path + 4 * k, cur + 4 * j, nxt + 4 * (n + j));      /* bit k of the index = 1 */
        }
        free(cur); cur = nxt; n *= 2;                                          /* candidate c: its index bits are c's bits, bit k = c >> k & 1 */
    }
    size_t found = (size_t)-1;
    for (size_t c = 0; c < n && found == (size_t)-1; c++)
        for (size_t t = 0; t < cap_size; t++)
            if (memcmp(cur + 4 * c, cap + 4 * t, 32) == 0) { found = (t << path_len) | c; break; }
    free(cur);
    return found;
}

/* the copy-permutation non-residue table of gl64.h, exported for the CPU suite */
EXPORT void orc_copy_permutation_non_residues(uint64_t *k, uint32_t n, int log_n) { gl_copy_permutation_non_residues(k, n, log_n); }

/* ------------------------------------------------------------------ FRI folding */
/* One un-normalised fold step of an Ext2 oracle stored split (c0[], c1[]) in bit-reversed enumeration.
 * Domain before the step: shift * <omega_{2^log_dom}>, point at index i = shift * omega^bitrev(i).
 *   out[m] = (f[2m] + f[2m+1]) + c * (f[2m] - f[2m+1]) / x_{2m}
 * (no 1/2 factor -- confirmed on the golden proofs; a fold-by-2^k oracle applies this k times with c, c^2, c^4..). */
EXPORT void orc_fri_fold(const uint64_t *in_c0, const uint64_t *in_c1, int log_dom, uint64_t shift, const uint64_t ch[2],
                         uint64_t *out_c0, uint64_t *out_c1) {
    size_t half = (size_t)1 << (log_dom - 1);
    uint64_t w_inv = gl_inv(gl_omega(log_dom)), s_inv = gl_inv(shift);
    gl2 c = gl2_make(ch[0], ch[1]);
#pragma omp parallel for schedule(static)
    for (size_t m = 0; m < half; m++) {
        /* bitrev_{log_dom}(2m) = bitrev_{log_dom-1}(m) */
        uint64_t xinv = gl_mul(s_inv, gl_pow(w_inv, bitrev32((uint32_t)m, log_dom - 1)));
        gl2 a = gl2_make(in_c0[2 * m], in_c1[2 * m]), b = gl2_make(in_c0[2 * m + 1], in_c1[2 * m + 1]);
        gl2 sum = gl2_add(a, b), dif = gl2_mul_base(gl2_sub(a, b), xinv);
        gl2 r = gl2_add(sum, gl2_mul(dif, c));
        out_c0[m] = r.c0; out_c1[m] = r.c1;
    }
}

/* Fold one FRI leaf (n = 2^k Ext2 values at consecutive indices base_idx.. of the domain shift*<omega_{2^log_dom}>,
 * split storage) all the way down to a single value, using c, c^2, c^4, .. for the successive steps.
 * This is what the verifier does per query and what the golden fixtures (tests/golden/fri_chain_*.json) pin. */
EXPORT void orc_fri_fold_leaf(const uint64_t *c0, const uint64_t *c1, size_t n, int log_dom, uint64_t shift, size_t base_idx,
                              const uint64_t ch[2], uint64_t out[2]) {
    gl2 cur[64];
    gl2 c = gl2_make(ch[0], ch[1]);
    for (size_t i = 0; i < n; i++) cur[i] = gl2_make(c0[i], c1[i]);
    while (n > 1) {
        uint64_t w_inv = gl_inv(gl_omega(log_dom)), s_inv = gl_inv(shift);
        for (size_t k = 0; k < n / 2; k++) {
            size_t idx = base_idx + 2 * k;
            uint64_t xinv = gl_mul(s_inv, gl_pow(w_inv, bitrev32((uint32_t)idx, log_dom)));
            gl2 a = cur[2 * k], b = cur[2 * k + 1];
            cur[k] = gl2_add(gl2_add(a, b), gl2_mul(gl2_mul_base(gl2_sub(a, b), xinv), c));
        }
        n >>= 1; log_dom -= 1; shift = gl_sqr(shift); base_idx >>= 1; c = gl2_sqr(c);
    }
    out[0] = cur[0].c0; out[1] = cur[0].c1;
}
/* evaluate an Ext2-coefficient polynomial (split storage, ascending degree) at a base-field point */
EXPORT void orc_eval_ext_poly_at_base(const uint64_t *c0, const uint64_t *c1, size_t n, uint64_t x, uint64_t out[2]) {
    gl2 r = gl2_make(0, 0);
    for (size_t i = n; i-- > 0;) r = gl2_add(gl2_mul_base(r, x), gl2_make(c0[i], c1[i]));
    out[0] = r.c0; out[1] = r.c1;
}
