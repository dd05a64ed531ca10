/* oracle/prover.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; never on the product path).
 *
 * Sequential CPU restatement of the whole proving pipeline behind
 *   cs.prove_from_precomputations::<GoldilocksExt2, Poseidon2 transcript, Poseidon2 sponge, NoPow>(..)
 *     (/root/reference/src/prover_utils.rs:338-348; recursion :533-543) and cs.get_full_setup (:185-186).
 * The algorithm is boojum's (un-vendored dependency, see primitives.c header).  Stage order, oracle shapes, opening
 * counts, folding schedule and query layout follow what the reference's golden proofs show (SURVEY.md section 8a,
 * Appendix A); conventions the goldens cannot show without the hash (alpha-power term order, transcript framing,
 * query-index bit extraction, DEEP term order) are this framework's own and are documented in DESIGN.md -- the
 * product must match THIS file bit for bit, and both verify under the product's CPU verifier.
 *
 * Stages: witness commit -> (beta, gamma[, lookup beta, gamma]) -> stage 2 (copy-permutation grand product with
 * partial products in chunks of quotient_degree; logUp lookup polys A_i, B) -> alpha -> quotient on quotient_degree
 * cosets, split in quotient_degree chunks -> z -> openings at z, z*omega, 0 -> DEEP challenge -> DEEP poly -> FRI
 * (commit / challenge / fold per schedule) -> final monomials -> query indexes -> openings.
 */
#include "gates.h"
#include "poseidon2_consts.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

void orc_ntt(uint64_t *a, int log_n, int inverse);
void orc_bitrev(uint64_t *a, int log_n);
void orc_coset_evals_bitrev(const uint64_t *mono, int log_n, uint64_t shift, uint64_t *out);
uint64_t orc_lde_coset_shift(int log_n, int log_lde, uint32_t c);
void orc_poseidon2_permute(uint64_t *s);
void orc_merkle_build(const uint64_t *cols, size_t col_stride, size_t n_cols, size_t n_leaves, size_t elems_per_leaf, size_t cap_size,
                      uint64_t *tree_out);
void orc_merkle_path(const uint64_t *tree, size_t n_leaves, size_t cap_size, size_t idx, uint64_t *path_out);
int orc_merkle_verify(const uint64_t *leaf_els, size_t leaf_len, const uint64_t *path, size_t path_len, const uint64_t *cap, size_t idx);
void orc_fri_fold_leaf(const uint64_t *c0, const uint64_t *c1, size_t n, int log_dom, uint64_t shift, size_t base_idx, const uint64_t ch[2],
                       uint64_t out[2]);
void orc_eval_ext_poly_at_base(const uint64_t *c0, const uint64_t *c1, size_t n, uint64_t x, uint64_t out[2]);
void orc_fri_fold(const uint64_t *in_c0, const uint64_t *in_c1, int log_dom, uint64_t shift, const uint64_t ch[2], uint64_t *out_c0,
                  uint64_t *out_c1);

/* ------------------------------------------------------------------ geometry helpers */
static uint32_t n_lookup_cols(const zkgpu_geometry *g) { return g->lookup_width * g->lookup_reps; }
static uint32_t n_perm(const zkgpu_geometry *g) { return g->n_copy + (g->has_boolean_col ? 1 : 0) + n_lookup_cols(g); }
static uint32_t n_wit(const zkgpu_geometry *g) { return n_perm(g) + g->n_witness_plain + (g->lookup_reps ? 1 : 0); }
static uint32_t n_setup(const zkgpu_geometry *g) { return n_perm(g) + g->n_const_cols + (g->lookup_reps ? g->lookup_width + 1 : 0); }
static uint32_t n_chunks(const zkgpu_geometry *g) { return (n_perm(g) + g->quotient_degree - 1) / g->quotient_degree; }
static uint32_t n_s2_ext(const zkgpu_geometry *g) { return n_chunks(g) + g->lookup_reps + (g->lookup_reps ? 1 : 0); }
static uint32_t ilog2(size_t x) { uint32_t r = 0; while (((size_t)1 << r) < x) r++; return r; }

EXPORT uint32_t orc_num_witness_cols(const zkgpu_geometry *g) { return n_wit(g); }
EXPORT uint32_t orc_num_setup_cols(const zkgpu_geometry *g) { return n_setup(g); }
EXPORT uint32_t orc_num_stage2_cols(const zkgpu_geometry *g) { return 2 * n_s2_ext(g); }

/* ------------------------------------------------------------------ DEEP combination at one LDE point
 * h(x) = sum_i phi^i (F_i(x) - F_i(z)) / (x - z) + phi^n (Z(x) - Z(z w)) / (x - z w) + sum_j phi^.. (A_j(x) - A_j(0)) / x
 *        + sum_t phi^.. (w_col_t(x) - pi_t) / (x - w^row_t)
 * F_i in the order of the reference's values_at_z (see the openings section of orc_prove); wl/sl/l2/lq = the four trace-oracle
 * leaves at x (stage-2 and quotient Ext2 polys as adjacent (c0, c1) columns).  Pinned hash-free on golden proofs
 * (tests/golden/deep_*.json): node, compression modes 1 and 2, and base-layer circuits with lookups of width 1, 3 and 4. */
typedef struct { int kind; uint32_t idx; } open_src; /* kind 0: witness column, 1: setup column, 2: stage-2 Ext2 poly, 3: quotient Ext2 poly */
static uint32_t opening_sources(const zkgpu_geometry *g, open_src *src) {
    const uint32_t NP = n_perm(g), W = n_wit(g), S = n_setup(g), C = n_chunks(g), E2 = n_s2_ext(g), QD = g->quotient_degree;
    uint32_t k = 0;
    for (uint32_t i = 0; i < W - (g->lookup_reps ? 1 : 0); i++) src[k++] = (open_src){0, i};
    for (uint32_t i = 0; i < g->n_const_cols; i++) src[k++] = (open_src){1, NP + i};
    for (uint32_t i = 0; i < NP; i++) src[k++] = (open_src){1, i};
    for (uint32_t i = 0; i < C; i++) src[k++] = (open_src){2, i};
    if (g->lookup_reps) src[k++] = (open_src){0, W - 1};
    for (uint32_t i = C; i < E2; i++) src[k++] = (open_src){2, i};
    for (uint32_t i = NP + g->n_const_cols; i < S; i++) src[k++] = (open_src){1, i};
    for (uint32_t i = 0; i < QD; i++) src[k++] = (open_src){3, i};
    return k;
}
static gl2 deep_point(const zkgpu_geometry *g, const open_src *src, uint32_t n_at_z, uint32_t n_at_0, const uint64_t *wl, const uint64_t *sl,
                      const uint64_t *l2, const uint64_t *lq, const gl2 *phip, gl2 sum_at_z, gl2 at_zw, const gl2 *at_0,
                      const uint64_t *pi_values, const uint64_t *pi_root, uint64_t x, gl2 z, gl2 zw) {
    const uint32_t C = n_chunks(g);
    gl2 s = gl2_make(0, 0);
    for (uint32_t i = 0; i < n_at_z; i++) {
        const uint32_t e = src[i].idx;
        switch (src[i].kind) {
        case 0: s = gl2_add(s, gl2_mul_base(phip[i], wl[e])); break;
        case 1: s = gl2_add(s, gl2_mul_base(phip[i], sl[e])); break;
        case 2: s = gl2_add(s, gl2_mul(phip[i], gl2_make(l2[2 * e], l2[2 * e + 1]))); break;
        default: s = gl2_add(s, gl2_mul(phip[i], gl2_make(lq[2 * e], lq[2 * e + 1]))); break;
        }
    }
    uint32_t k = n_at_z;
    gl2 xe = gl2_make(x, 0);
    gl2 h = gl2_mul(gl2_sub(s, sum_at_z), gl2_inv(gl2_sub(xe, z)));
    h = gl2_add(h, gl2_mul(gl2_mul(phip[k++], gl2_sub(gl2_make(l2[0], l2[1]), at_zw)), gl2_inv(gl2_sub(xe, zw))));
    uint64_t xinv = gl_inv(x);
    for (uint32_t i = 0; i < n_at_0; i++) {
        gl2 a = gl2_make(l2[2 * (C + i)], l2[2 * (C + i) + 1]);
        h = gl2_add(h, gl2_mul(phip[k++], gl2_mul_base(gl2_sub(a, at_0[i]), xinv)));
    }
    for (uint32_t i = 0; i < g->n_public_inputs; i++) {
        uint64_t num = gl_sub(wl[g->pi_col[i]], pi_values[i]);
        h = gl2_add(h, gl2_mul_base(phip[k++], gl_mul(num, gl_inv(gl_sub(x, pi_root[i])))));
    }
    return h;
}
/* test entry point: the DEEP value at one point from the four leaves and the proof's openings (at_z in the proof's order) */
EXPORT void orc_deep_at_point(const zkgpu_geometry *g, const uint64_t *wl, const uint64_t *sl, const uint64_t *l2, const uint64_t *lq,
                              const uint64_t *at_z, const uint64_t *at_zw, const uint64_t *at_0, const uint64_t *pi_values, uint64_t x,
                              const uint64_t *z2, const uint64_t *phi2, uint64_t *out2) {
    const uint32_t n_at_z = n_wit(g) + n_setup(g) + n_s2_ext(g) + g->quotient_degree, n_at_0 = g->lookup_reps ? g->lookup_reps + 1 : 0;
    open_src *src = (open_src *)malloc(sizeof(open_src) * n_at_z);
    opening_sources(g, src);
    const uint32_t n_deep = n_at_z + 1 + n_at_0 + g->n_public_inputs;
    gl2 *phip = (gl2 *)malloc(sizeof(gl2) * n_deep), phi = gl2_make(phi2[0], phi2[1]), z = gl2_make(z2[0], z2[1]);
    phip[0] = gl2_make(1, 0);
    for (uint32_t i = 1; i < n_deep; i++) phip[i] = gl2_mul(phip[i - 1], phi);
    gl2 sum_at_z = gl2_make(0, 0);
    for (uint32_t i = 0; i < n_at_z; i++) sum_at_z = gl2_add(sum_at_z, gl2_mul(phip[i], gl2_make(at_z[2 * i], at_z[2 * i + 1])));
    uint64_t pi_root[ZKGPU_MAX_PUBLIC_INPUTS];
    for (uint32_t i = 0; i < g->n_public_inputs; i++) pi_root[i] = gl_pow(gl_omega(g->log_n), g->pi_row[i]);
    gl2 h = deep_point(g, src, n_at_z, n_at_0, wl, sl, l2, lq, phip, sum_at_z, gl2_make(at_zw[0], at_zw[1]), (const gl2 *)at_0, pi_values, pi_root,
                       x, z, gl2_mul_base(z, gl_omega(g->log_n)));
    out2[0] = h.c0; out2[1] = h.c1;
    free(src); free(phip);
}

/* ------------------------------------------------------------------ transcript (Poseidon2 sponge, rate 8, overwrite)
 * boojum's AlgebraicSpongeBasedTranscript<F, 8, 12, 4, R>, pinned on the golden proofs (tools/golden_transcript.py): witnessed
 * elements are buffered; a challenge request absorbs the buffer FOLLOWED BY A ONE (then zero fill, one permutation per block of
 * 8, state never reset) and the TR_CHALLENGES = 8 rate lanes become the list of available challenges; when it runs out the
 * state is permuted again.  Query indexes (`BoolsBuffer`): every challenge contributes its 64 - log2(LDE domain) low bits,
 * LSB first. */
#define TR_CHALLENGES 8
typedef struct {
    uint64_t st[12];
    uint64_t *buf;
    size_t len, capacity;
    int pos; /* next lane to hand out; TR_CHALLENGES = exhausted */
    uint64_t bitbuf;
    int nbits;
} tr_t;
static void tr_init(tr_t *t) { memset(t, 0, sizeof(*t)); t->pos = TR_CHALLENGES; }
static void tr_free(tr_t *t) { free(t->buf); }
static void tr_absorb(tr_t *t, const uint64_t *v, size_t n) {
    if (t->len + n > t->capacity) {
        t->capacity = 2 * (t->len + n);
        t->buf = (uint64_t *)realloc(t->buf, t->capacity * 8);
    }
    memcpy(t->buf + t->len, v, n * 8);
    t->len += n;
}
static uint64_t tr_challenge(tr_t *t) {
    if (t->len) {
        const uint64_t one = 1;
        tr_absorb(t, &one, 1);
        for (size_t i = 0; i < t->len; i += 8) {
            for (size_t k = 0; k < 8; k++) t->st[k] = i + k < t->len ? t->buf[i + k] : 0;
            orc_poseidon2_permute(t->st);
        }
        t->len = 0;
        t->pos = 0;
    } else if (t->pos == TR_CHALLENGES) {
        orc_poseidon2_permute(t->st);
        t->pos = 0;
    }
    return t->st[t->pos++];
}
static size_t tr_query_index(tr_t *t, uint32_t bits) {
    const int take = 64 - (int)bits;
    while (t->nbits < (int)bits) {
        const uint64_t c = tr_challenge(t);
        t->bitbuf |= (take == 64 ? c : (c & (((uint64_t)1 << take) - 1))) << t->nbits;
        t->nbits += take;
    }
    const size_t idx = (size_t)(t->bitbuf & (((uint64_t)1 << bits) - 1));
    t->bitbuf >>= bits;
    t->nbits -= (int)bits;
    return idx;
}
static gl2 tr_challenge_ext(tr_t *t) { uint64_t a = tr_challenge(t); uint64_t b = tr_challenge(t); return gl2_make(a, b); }

/* ------------------------------------------------------------------ small helpers */
static uint64_t *alloc_u64(size_t n) {
    uint64_t *p = (uint64_t *)malloc((n ? n : 1) * 8);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
/* values on H (natural) -> monomials; then LDE (coset-major, bit-reversed) */
static void commit_columns(const uint64_t *vals, size_t n_cols, int log_n, int log_lde, size_t cap, uint64_t *mono, uint64_t *lde, uint64_t *tree) {
    size_t N = (size_t)1 << log_n, LN = N << log_lde;
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < n_cols; c++) {
        memcpy(mono + c * N, vals + c * N, N * 8);
        orc_ntt(mono + c * N, log_n, 1);
        for (uint32_t k = 0; k < (1u << log_lde); k++)
            orc_coset_evals_bitrev(mono + c * N, log_n, orc_lde_coset_shift(log_n, log_lde, k), lde + c * LN + (size_t)k * N);
    }
    orc_merkle_build(lde, LN, n_cols, LN, 1, cap, tree);
}
/* evaluate a base-coefficient polynomial at an Ext2 point */
static gl2 eval_at_ext(const uint64_t *mono, size_t n, gl2 z) {
    gl2 r = gl2_make(0, 0);
    for (size_t i = n; i-- > 0;) { r = gl2_mul(r, z); r.c0 = gl_add(r.c0, mono[i]); }
    return r;
}
static gl2 mul_by_u(gl2 a) { return gl2_make(gl_mul(7, a.c1), a.c0); } /* (a0 + a1 u) * u */

/* ------------------------------------------------------------------ proof buffer layout (DESIGN.md "Proof buffer") */
#define PROOF_MAGIC 0x5A4B50524F4F4631ULL
typedef struct {
    size_t N, LN, depth;
    uint32_t W, S, S2, Q, n_at_z, n_at_zw, n_at_0, n_final;
    size_t fri_dom_log[ZKGPU_MAX_FRI_ORACLES + 1], fri_leaves[ZKGPU_MAX_FRI_ORACLES], fri_cap[ZKGPU_MAX_FRI_ORACLES], fri_depth[ZKGPU_MAX_FRI_ORACLES];
} shape_t;
static void make_shape(const zkgpu_geometry *g, const zkgpu_proof_config *cfg, shape_t *s) {
    s->N = (size_t)1 << g->log_n;
    s->LN = s->N << cfg->log_lde;
    s->depth = ilog2(s->LN / cfg->cap_size);
    s->W = n_wit(g); s->S = n_setup(g); s->S2 = 2 * n_s2_ext(g); s->Q = 2 * g->quotient_degree;
    s->n_at_z = s->W + s->S + n_s2_ext(g) + g->quotient_degree;
    s->n_at_zw = 1;
    s->n_at_0 = g->lookup_reps ? g->lookup_reps + 1 : 0;
    size_t ld = g->log_n + cfg->log_lde;
    for (uint32_t k = 0; k < cfg->n_fri_oracles; k++) {
        s->fri_dom_log[k] = ld;
        s->fri_leaves[k] = ((size_t)1 << ld) >> cfg->fri_schedule[k];
        s->fri_cap[k] = cfg->cap_size < s->fri_leaves[k] ? cfg->cap_size : s->fri_leaves[k];
        s->fri_depth[k] = ilog2(s->fri_leaves[k] / s->fri_cap[k]);
        ld -= cfg->fri_schedule[k];
    }
    s->fri_dom_log[cfg->n_fri_oracles] = ld;
    s->n_final = (uint32_t)(((size_t)1 << ld) >> cfg->log_lde);
}
EXPORT size_t orc_proof_size_u64(const zkgpu_geometry *g, const zkgpu_proof_config *cfg) {
    shape_t s; make_shape(g, cfg, &s);
    size_t n = 32 + g->n_public_inputs + 3 * cfg->cap_size * 4 + 2 * s.n_final + 2 * (s.n_at_z + s.n_at_zw + s.n_at_0);
    for (uint32_t k = 0; k < cfg->n_fri_oracles; k++) n += s.fri_cap[k] * 4;
    size_t per_q = s.W + s.S2 + s.Q + s.S + 4 * s.depth * 4;
    for (uint32_t k = 0; k < cfg->n_fri_oracles; k++) per_q += 2 * ((size_t)1 << cfg->fri_schedule[k]) + s.fri_depth[k] * 4;
    return n + per_q * cfg->n_queries + 1;
}

/* ------------------------------------------------------------------ setup commitment (get_full_setup) */
EXPORT void orc_setup_cap(const zkgpu_geometry *g, const zkgpu_proof_config *cfg, const uint64_t *setup_cols, uint64_t *cap_out) {
    shape_t sh; make_shape(g, cfg, &sh);
    uint64_t *mono = alloc_u64(sh.S * sh.N), *lde = alloc_u64(sh.S * sh.LN), *tree = alloc_u64((2 * sh.LN - cfg->cap_size) * 4);
    commit_columns(setup_cols, sh.S, g->log_n, cfg->log_lde, cfg->cap_size, mono, lde, tree);
    memcpy(cap_out, tree + 4 * (2 * sh.LN - 2 * cfg->cap_size), cfg->cap_size * 32);
    free(mono); free(lde); free(tree);
}

/* ------------------------------------------------------------------ the quotient numerator at one point */
typedef struct {
    gl2 beta, gamma, lbeta, lgamma, alpha;
    gl2 *alpha_pow;     /* alpha^k for every term */
    uint32_t n_terms;
    uint64_t *pi_values; /* public input values */
    uint64_t knr[1024];  /* copy-permutation non-residues k_i (gl64.h gl_copy_permutation_non_residues) */
} chal_t;

static uint32_t count_terms(const zkgpu_geometry *g) {
    uint32_t t = og_total_terms(g) + (g->has_boolean_col ? 1 : 0); /* public inputs: DEEP openings, not quotient terms */
    if (g->lookup_reps) t += g->lookup_reps + 1;
    t += 1 + n_chunks(g);
    return t;
}

/* w: W witness values, s: S setup values, e2: stage-2 ext values, zs: z(omega x), x: the point, all at one point of a
 * coset.  Returns the combined numerator sum_k alpha^k term_k (before division by Z_H). */
static gl2 quotient_numerator(const zkgpu_geometry *g, const chal_t *ch, const uint64_t *w, const uint64_t *s, const gl2 *e2, gl2 zs, uint64_t x,
                              uint64_t xn_minus_1, uint64_t *scratch) {
    const uint32_t NP = n_perm(g), C = n_chunks(g), QD = g->quotient_degree;
    const uint64_t *sigma = s, *consts = s + NP, *tables = s + NP + g->n_const_cols;
    const uint64_t N = (uint64_t)1 << g->log_n;
    gl2 acc = gl2_make(0, 0);
    uint32_t k = 0;
    /* 1. gates */
    for (uint32_t gi = 0; gi < g->n_gates; gi++) {
        const zkgpu_gate *gt = &g->gates[gi];
        const uint64_t *cells = w;
        uint64_t cellbuf[520];
        if (g->n_witness_plain && NP != g->n_copy) { /* plain witness columns sit after ALL copy-permuted columns */
            memcpy(cellbuf, w, g->n_copy * 8);
            memcpy(cellbuf + g->n_copy, w + NP, g->n_witness_plain * 8);
            cells = cellbuf;
        }
        uint32_t nrel = og_eval_gate(gt, g, cells, consts + gt->path_len, ORC_P2_RC, scratch);
        if (!nrel) continue;
        uint64_t sel = 1;
        for (uint32_t b = 0; b < gt->path_len; b++) sel = gl_mul(sel, ((gt->path_bits >> b) & 1) ? consts[b] : gl_sub(1, consts[b]));
        gl2 ga = gl2_make(0, 0);
        for (uint32_t r = 0; r < nrel; r++) ga = gl2_add(ga, gl2_mul_base(ch->alpha_pow[k + r], scratch[r]));
        acc = gl2_add(acc, gl2_mul_base(ga, sel));
        k += nrel;
    }
    /* 2. boolean column */
    if (g->has_boolean_col) {
        uint64_t b = w[g->n_copy];
        acc = gl2_add(acc, gl2_mul_base(ch->alpha_pow[k++], gl_sub(gl_sqr(b), b)));
    }
    /* 3. (public inputs are opened through the DEEP polynomial -- see the DEEP section -- and are not quotient terms) */
    /* 4. lookup (log-derivative): A_i * den_i - 1 ; B * den_table - m */
    if (g->lookup_reps) {
        const uint32_t LW = g->lookup_width;
        const uint64_t *lw = w + g->n_copy + (g->has_boolean_col ? 1 : 0);
        gl2 gp[16];
        gp[0] = gl2_make(1, 0);
        for (uint32_t j = 1; j <= LW; j++) gp[j] = gl2_mul(gp[j - 1], ch->lgamma);
        gl2 tid = gl2_mul_base(gp[LW], consts[g->table_id_col]);
        for (uint32_t i = 0; i < g->lookup_reps; i++) {
            gl2 den = gl2_add(ch->lbeta, tid);
            for (uint32_t j = 0; j < LW; j++) den = gl2_add(den, gl2_mul_base(gp[j], lw[i * LW + j]));
            gl2 t = gl2_sub(gl2_mul(e2[C + i], den), gl2_make(1, 0));
            acc = gl2_add(acc, gl2_mul(ch->alpha_pow[k++], t));
        }
        gl2 den = ch->lbeta;
        for (uint32_t j = 0; j <= LW; j++) den = gl2_add(den, gl2_mul_base(gp[j], tables[j]));
        gl2 t = gl2_mul(e2[C + g->lookup_reps], den);
        t.c0 = gl_sub(t.c0, w[n_wit(g) - 1]);
        acc = gl2_add(acc, gl2_mul(ch->alpha_pow[k++], t));
    }
    /* 5. copy permutation */
    {
        uint64_t l0 = gl_mul(xn_minus_1, gl_inv(gl_mul(N % GL_P, gl_sub(x, 1))));
        gl2 t = gl2_mul_base(gl2_sub(e2[0], gl2_make(1, 0)), l0);
        acc = gl2_add(acc, gl2_mul(ch->alpha_pow[k++], t));
        for (uint32_t j = 0; j < C; j++) {
            gl2 num = gl2_make(1, 0), den = gl2_make(1, 0);
            for (uint32_t i = j * QD; i < (j + 1) * QD && i < NP; i++) {
                const uint64_t kx = gl_mul(ch->knr[i], x); /* k_i * x */
                gl2 a = gl2_add(gl2_mul_base(ch->beta, kx), ch->gamma);
                a.c0 = gl_add(a.c0, w[i]);
                gl2 b = gl2_add(gl2_mul_base(ch->beta, sigma[i]), ch->gamma);
                b.c0 = gl_add(b.c0, w[i]);
                num = gl2_mul(num, a);
                den = gl2_mul(den, b);
            }
            gl2 prev = e2[j]; /* j = 0: z, else p_{j-1} */
            gl2 cur = (j + 1 < C) ? e2[j + 1] : zs;
            gl2 t2 = gl2_sub(gl2_mul(cur, den), gl2_mul(prev, num));
            acc = gl2_add(acc, gl2_mul(ch->alpha_pow[k++], t2));
        }
    }
    return acc;
}

/* ------------------------------------------------------------------ prove */
EXPORT long orc_prove(const zkgpu_geometry *g, const zkgpu_proof_config *cfg, const uint64_t *wit_cols, const uint64_t *setup_cols,
                      uint64_t *proof, size_t proof_capacity) {
    shape_t sh; make_shape(g, cfg, &sh);
    const size_t N = sh.N, LN = sh.LN;
    const int log_n = g->log_n, log_lde = cfg->log_lde;
    const uint32_t W = sh.W, S = sh.S, S2 = sh.S2, Q = sh.Q, NP = n_perm(g), C = n_chunks(g), E2 = n_s2_ext(g), QD = g->quotient_degree;
    const size_t cap = cfg->cap_size, tree_len = (2 * LN - cap) * 4, cap_off = 4 * (2 * LN - 2 * cap);
    if (orc_proof_size_u64(g, cfg) > proof_capacity) return -1;
    if (QD & (QD - 1)) return -2;
    const uint64_t omega = gl_omega(log_n);

    /* ---- setup + witness commitments */
    uint64_t *mono_s = alloc_u64(S * N), *lde_s = alloc_u64(S * LN), *tree_s = alloc_u64(tree_len);
    commit_columns(setup_cols, S, log_n, log_lde, cap, mono_s, lde_s, tree_s);
    uint64_t *mono_w = alloc_u64(W * N), *lde_w = alloc_u64(W * LN), *tree_w = alloc_u64(tree_len);
    commit_columns(wit_cols, W, log_n, log_lde, cap, mono_w, lde_w, tree_w);

    uint64_t pi_values[ZKGPU_MAX_PUBLIC_INPUTS];
    for (uint32_t i = 0; i < g->n_public_inputs; i++) pi_values[i] = wit_cols[(size_t)g->pi_col[i] * N + g->pi_row[i]];

    tr_t tr; tr_init(&tr);
    tr_absorb(&tr, tree_s + cap_off, cap * 4);
    tr_absorb(&tr, pi_values, g->n_public_inputs);
    tr_absorb(&tr, tree_w + cap_off, cap * 4);
    chal_t ch; memset(&ch, 0, sizeof(ch));
    gl_copy_permutation_non_residues(ch.knr, n_perm(g), (int)g->log_n);
    ch.beta = tr_challenge_ext(&tr);
    ch.gamma = tr_challenge_ext(&tr);
    if (g->lookup_reps) { ch.lbeta = tr_challenge_ext(&tr); ch.lgamma = tr_challenge_ext(&tr); }
    ch.pi_values = pi_values;

    /* ---- stage 2 on H (natural order) */
    uint64_t *s2 = alloc_u64((size_t)S2 * N);
    {
        const uint64_t *sigma = setup_cols, *consts = setup_cols + (size_t)NP * N, *tables = setup_cols + (size_t)(NP + g->n_const_cols) * N;
        gl2 z = gl2_make(1, 0);
        uint64_t x = 1;
        gl2 *nums = (gl2 *)malloc(sizeof(gl2) * C), *dens = (gl2 *)malloc(sizeof(gl2) * C);
        for (size_t r = 0; r < N; r++) {
            s2[0 * N + r] = z.c0; s2[1 * N + r] = z.c1;
            for (uint32_t j = 0; j < C; j++) {
                gl2 num = gl2_make(1, 0), den = gl2_make(1, 0);
                for (uint32_t i = j * QD; i < (j + 1) * QD && i < NP; i++) {
                    uint64_t wv = wit_cols[(size_t)i * N + r];
                    const uint64_t kx = gl_mul(ch.knr[i], x);
                    gl2 a = gl2_add(gl2_mul_base(ch.beta, kx), ch.gamma); a.c0 = gl_add(a.c0, wv);
                    gl2 b = gl2_add(gl2_mul_base(ch.beta, sigma[(size_t)i * N + r]), ch.gamma); b.c0 = gl_add(b.c0, wv);
                    num = gl2_mul(num, a); den = gl2_mul(den, b);
                }
                nums[j] = num; dens[j] = den;
            }
            gl2 cur = z;
            for (uint32_t j = 0; j < C; j++) {
                cur = gl2_mul(gl2_mul(cur, nums[j]), gl2_inv(dens[j]));
                if (j + 1 < C) { s2[(size_t)(2 * (j + 1)) * N + r] = cur.c0; s2[(size_t)(2 * (j + 1) + 1) * N + r] = cur.c1; }
            }
            z = cur;
            x = gl_mul(x, omega);
        }
        free(nums); free(dens);
        if (!(z.c0 == 1 && z.c1 == 0)) { fprintf(stderr, "oracle: copy-permutation grand product does not close (witness does not satisfy sigma)\n"); }
        if (g->lookup_reps) {
            const uint32_t LW = g->lookup_width;
            gl2 gp[16]; gp[0] = gl2_make(1, 0);
            for (uint32_t j = 1; j <= LW; j++) gp[j] = gl2_mul(gp[j - 1], ch.lgamma);
            size_t lw0 = g->n_copy + (g->has_boolean_col ? 1 : 0);
#pragma omp parallel for schedule(static)
            for (size_t r = 0; r < N; r++) {
                gl2 tid = gl2_mul_base(gp[LW], consts[(size_t)g->table_id_col * N + r]);
                for (uint32_t i = 0; i < g->lookup_reps; i++) {
                    gl2 den = gl2_add(ch.lbeta, tid);
                    for (uint32_t j = 0; j < LW; j++) den = gl2_add(den, gl2_mul_base(gp[j], wit_cols[(lw0 + i * LW + j) * N + r]));
                    gl2 a = gl2_inv(den);
                    s2[(size_t)(2 * (C + i)) * N + r] = a.c0; s2[(size_t)(2 * (C + i) + 1) * N + r] = a.c1;
                }
                gl2 den = ch.lbeta;
                for (uint32_t j = 0; j <= LW; j++) den = gl2_add(den, gl2_mul_base(gp[j], tables[(size_t)j * N + r]));
                gl2 b = gl2_mul_base(gl2_inv(den), wit_cols[(size_t)(W - 1) * N + r]);
                s2[(size_t)(2 * (C + g->lookup_reps)) * N + r] = b.c0; s2[(size_t)(2 * (C + g->lookup_reps) + 1) * N + r] = b.c1;
            }
        }
    }
    uint64_t *mono_2 = alloc_u64((size_t)S2 * N), *lde_2 = alloc_u64((size_t)S2 * LN), *tree_2 = alloc_u64(tree_len);
    commit_columns(s2, S2, log_n, log_lde, cap, mono_2, lde_2, tree_2);
    free(s2);
    tr_absorb(&tr, tree_2 + cap_off, cap * 4);
    ch.alpha = tr_challenge_ext(&tr);
    ch.n_terms = count_terms(g);
    ch.alpha_pow = (gl2 *)malloc(sizeof(gl2) * ch.n_terms);
    ch.alpha_pow[0] = gl2_make(1, 0);
    for (uint32_t i = 1; i < ch.n_terms; i++) ch.alpha_pow[i] = gl2_mul(ch.alpha_pow[i - 1], ch.alpha);

    /* ---- quotient on QD cosets of size N (the 8N domain 7*<omega_8N>, coset-major bit-reversed) */
    const int log_qd = ilog2(QD);
    const size_t QN = N * QD;
    uint64_t *t0 = alloc_u64(QN), *t1 = alloc_u64(QN);
    {
        uint64_t *cw = alloc_u64((size_t)W * N), *cs = alloc_u64((size_t)S * N), *c2 = alloc_u64((size_t)S2 * N);
        for (uint32_t c = 0; c < QD; c++) {
            uint64_t shift = orc_lde_coset_shift(log_n, log_qd, c);
#pragma omp parallel for schedule(dynamic)
            for (size_t i = 0; i < (size_t)(W + S + S2); i++) {
                if (i < W) orc_coset_evals_bitrev(mono_w + i * N, log_n, shift, cw + i * N);
                else if (i < W + S) orc_coset_evals_bitrev(mono_s + (i - W) * N, log_n, shift, cs + (i - W) * N);
                else orc_coset_evals_bitrev(mono_2 + (i - W - S) * N, log_n, shift, c2 + (i - W - S) * N);
            }
            uint64_t xn_minus_1 = gl_sub(gl_pow(shift, N), 1);
            uint64_t zh_inv = gl_inv(xn_minus_1);
#pragma omp parallel
            {
                uint64_t *wv = alloc_u64(W), *sv = alloc_u64(S), *scratch = alloc_u64(1024);
                gl2 *ev = (gl2 *)malloc(sizeof(gl2) * E2);
#pragma omp for schedule(static)
                for (size_t j = 0; j < N; j++) {
                    uint32_t nat = bitrev32((uint32_t)j, log_n);
                    uint64_t x = gl_mul(shift, gl_pow(omega, nat));
                    size_t jn = bitrev32((nat + 1) & (uint32_t)(N - 1), log_n); /* position of omega*x */
                    for (uint32_t i = 0; i < W; i++) wv[i] = cw[(size_t)i * N + j];
                    for (uint32_t i = 0; i < S; i++) sv[i] = cs[(size_t)i * N + j];
                    for (uint32_t i = 0; i < E2; i++) ev[i] = gl2_make(c2[(size_t)(2 * i) * N + j], c2[(size_t)(2 * i + 1) * N + j]);
                    gl2 zs = gl2_make(c2[jn], c2[N + jn]);
                    gl2 num = quotient_numerator(g, &ch, wv, sv, ev, zs, x, xn_minus_1, scratch);
                    num = gl2_mul_base(num, zh_inv);
                    t0[(size_t)c * N + j] = num.c0; t1[(size_t)c * N + j] = num.c1;
                }
                free(wv); free(sv); free(scratch); free(ev);
            }
        }
        free(cw); free(cs); free(c2);
    }
    /* interpolate over the big coset: bit-reversed -> natural, inverse NTT, undo the shift 7^i, split in QD chunks */
    uint64_t *qvals = alloc_u64((size_t)Q * N); /* chunk monomials: col 2c = c0 of chunk c, 2c+1 = c1 */
    {
        orc_bitrev(t0, log_n + log_qd); orc_bitrev(t1, log_n + log_qd);
        orc_ntt(t0, log_n + log_qd, 1); orc_ntt(t1, log_n + log_qd, 1);
        uint64_t ginv = gl_inv(GL_GEN), s = 1;
        for (size_t i = 0; i < QN; i++) {
            size_t c = i / N, r = i % N;
            qvals[(2 * c) * N + r] = gl_mul(t0[i], s);
            qvals[(2 * c + 1) * N + r] = gl_mul(t1[i], s);
            s = gl_mul(s, ginv);
        }
    }
    free(t0); free(t1);
    uint64_t *mono_q = qvals, *lde_q = alloc_u64((size_t)Q * LN), *tree_q = alloc_u64(tree_len);
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < Q; c++)
        for (uint32_t k = 0; k < (1u << log_lde); k++)
            orc_coset_evals_bitrev(mono_q + c * N, log_n, orc_lde_coset_shift(log_n, log_lde, k), lde_q + c * LN + (size_t)k * N);
    orc_merkle_build(lde_q, LN, Q, LN, 1, cap, tree_q);
    tr_absorb(&tr, tree_q + cap_off, cap * 4);
    gl2 z = tr_challenge_ext(&tr);

    /* ---- openings, in the order of the reference's `values_at_z`: variables + plain witness columns, constants, sigmas,
     * z + partial products, lookup multiplicities, lookup A polys + B, lookup table columns, quotient chunks.  The lookup-free
     * part (witness leaf, constants, sigmas, stage 2, quotient; setup leaf stored sigmas-then-constants) is pinned hash-free on
     * golden proofs (tools/golden_deep.py, tests/golden/deep_*.json); the position of the lookup blocks was confirmed there too. */
    const uint32_t n_at_z = sh.n_at_z, n_at_0 = sh.n_at_0;
    open_src *src = (open_src *)malloc(sizeof(open_src) * n_at_z);
    if (opening_sources(g, src) != n_at_z) { fprintf(stderr, "oracle: opening count mismatch\n"); return -1; }
    gl2 *at_z = (gl2 *)malloc(sizeof(gl2) * n_at_z), *at_0 = (gl2 *)malloc(sizeof(gl2) * (n_at_0 ? n_at_0 : 1));
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < (size_t)n_at_z; i++) {
        const uint32_t e = src[i].idx;
        switch (src[i].kind) {
        case 0: at_z[i] = eval_at_ext(mono_w + (size_t)e * N, N, z); break;
        case 1: at_z[i] = eval_at_ext(mono_s + (size_t)e * N, N, z); break;
        case 2: at_z[i] = gl2_add(eval_at_ext(mono_2 + (size_t)(2 * e) * N, N, z), mul_by_u(eval_at_ext(mono_2 + (size_t)(2 * e + 1) * N, N, z))); break;
        default: at_z[i] = gl2_add(eval_at_ext(mono_q + (size_t)(2 * e) * N, N, z), mul_by_u(eval_at_ext(mono_q + (size_t)(2 * e + 1) * N, N, z))); break;
        }
    }
    gl2 zw = gl2_mul_base(z, omega);
    gl2 at_zw = gl2_add(eval_at_ext(mono_2, N, zw), mul_by_u(eval_at_ext(mono_2 + N, N, zw)));
    for (uint32_t i = 0; i < n_at_0; i++) at_0[i] = gl2_make(mono_2[(size_t)(2 * (C + i)) * N], mono_2[(size_t)(2 * (C + i) + 1) * N]);
    tr_absorb(&tr, (const uint64_t *)at_z, 2 * n_at_z);
    tr_absorb(&tr, (const uint64_t *)&at_zw, 2);
    tr_absorb(&tr, (const uint64_t *)at_0, 2 * n_at_0);
    gl2 phi = tr_challenge_ext(&tr);

    /* ---- DEEP polynomial on the LDE domain */
    const uint32_t n_deep = n_at_z + 1 + n_at_0 + g->n_public_inputs;
    gl2 *phip = (gl2 *)malloc(sizeof(gl2) * n_deep);
    phip[0] = gl2_make(1, 0);
    for (uint32_t i = 1; i < n_deep; i++) phip[i] = gl2_mul(phip[i - 1], phi);
    gl2 sum_at_z = gl2_make(0, 0);
    for (uint32_t i = 0; i < n_at_z; i++) sum_at_z = gl2_add(sum_at_z, gl2_mul(phip[i], at_z[i]));
    uint64_t *f0 = alloc_u64(LN), *f1 = alloc_u64(LN);
    const int log_ln = log_n + log_lde;
    const uint64_t omega_ln = gl_omega(log_ln);
    uint64_t pi_root[ZKGPU_MAX_PUBLIC_INPUTS];
    for (uint32_t i = 0; i < g->n_public_inputs; i++) pi_root[i] = gl_pow(omega, g->pi_row[i]);
#pragma omp parallel
    {
        uint64_t *wl = alloc_u64(W), *sl = alloc_u64(S), *l2 = alloc_u64(S2), *lq = alloc_u64(Q);
#pragma omp for schedule(static)
        for (size_t idx = 0; idx < LN; idx++) {
            uint64_t x = gl_mul(GL_GEN, gl_pow(omega_ln, bitrev32((uint32_t)idx, log_ln)));
            for (uint32_t i = 0; i < W; i++) wl[i] = lde_w[(size_t)i * LN + idx];
            for (uint32_t i = 0; i < S; i++) sl[i] = lde_s[(size_t)i * LN + idx];
            for (uint32_t i = 0; i < S2; i++) l2[i] = lde_2[(size_t)i * LN + idx];
            for (uint32_t i = 0; i < Q; i++) lq[i] = lde_q[(size_t)i * LN + idx];
            gl2 h = deep_point(g, src, n_at_z, n_at_0, wl, sl, l2, lq, phip, sum_at_z, at_zw, at_0, pi_values, pi_root, x, z, zw);
            f0[idx] = h.c0; f1[idx] = h.c1;
        }
        free(wl); free(sl); free(l2); free(lq);
    }

    /* ---- FRI */
    const uint32_t NF = cfg->n_fri_oracles;
    uint64_t *fri_c0[ZKGPU_MAX_FRI_ORACLES], *fri_c1[ZKGPU_MAX_FRI_ORACLES], *fri_tree[ZKGPU_MAX_FRI_ORACLES];
    uint64_t *cur0 = f0, *cur1 = f1;
    uint64_t shift = GL_GEN;
    for (uint32_t k = 0; k < NF; k++) {
        int ld = (int)sh.fri_dom_log[k];
        size_t D = (size_t)1 << ld;
        fri_c0[k] = cur0; fri_c1[k] = cur1;
        uint64_t *pair = alloc_u64(2 * D);
        memcpy(pair, cur0, D * 8); memcpy(pair + D, cur1, D * 8);
        fri_tree[k] = alloc_u64((2 * sh.fri_leaves[k] - sh.fri_cap[k]) * 4);
        orc_merkle_build(pair, D, 2, sh.fri_leaves[k], (size_t)1 << cfg->fri_schedule[k], sh.fri_cap[k], fri_tree[k]);
        free(pair);
        tr_absorb(&tr, fri_tree[k] + 4 * (2 * sh.fri_leaves[k] - 2 * sh.fri_cap[k]), sh.fri_cap[k] * 4);
        gl2 c = tr_challenge_ext(&tr);
        for (uint32_t st = 0; st < cfg->fri_schedule[k]; st++) {
            uint64_t *n0 = alloc_u64(D >> 1), *n1 = alloc_u64(D >> 1);
            uint64_t cc[2] = {c.c0, c.c1};
            orc_fri_fold(cur0, cur1, ld, shift, cc, n0, n1);
            if (st > 0) { free(cur0); free(cur1); }
            cur0 = n0; cur1 = n1;
            c = gl2_sqr(c); shift = gl_sqr(shift); ld--; D >>= 1;
        }
    }
    /* final polynomial: values on shift*<omega_D> (bit-reversed) -> monomials */
    const int ldf = (int)sh.fri_dom_log[NF];
    const size_t DF = (size_t)1 << ldf;
    uint64_t *fin0 = alloc_u64(DF), *fin1 = alloc_u64(DF);
    memcpy(fin0, cur0, DF * 8); memcpy(fin1, cur1, DF * 8);
    orc_bitrev(fin0, ldf); orc_bitrev(fin1, ldf);
    orc_ntt(fin0, ldf, 1); orc_ntt(fin1, ldf, 1);
    {
        uint64_t si = gl_inv(shift), s = 1;
        for (size_t i = 0; i < DF; i++) { fin0[i] = gl_mul(fin0[i], s); fin1[i] = gl_mul(fin1[i], s); s = gl_mul(s, si); }
        for (size_t i = sh.n_final; i < DF; i++)
            if (fin0[i] || fin1[i]) { fprintf(stderr, "oracle: final FRI polynomial has degree >= %u (constraints not satisfied?)\n", sh.n_final); break; }
    }
    tr_absorb(&tr, fin0, sh.n_final);
    tr_absorb(&tr, fin1, sh.n_final);

    /* ---- write the proof */
    uint64_t *p = proof;
    memset(p, 0, 32 * 8);
    p[0] = PROOF_MAGIC; p[1] = log_n; p[2] = log_lde; p[3] = cap; p[4] = cfg->n_queries; p[5] = NF; p[6] = W; p[7] = S2; p[8] = Q; p[9] = S;
    p[10] = n_at_z; p[11] = sh.n_at_zw; p[12] = n_at_0; p[13] = g->n_public_inputs; p[14] = sh.n_final; p[15] = cfg->pow_bits;
    for (uint32_t k = 0; k < NF; k++) p[16 + k] = cfg->fri_schedule[k];
    p += 32;
    memcpy(p, pi_values, g->n_public_inputs * 8); p += g->n_public_inputs;
    memcpy(p, tree_w + cap_off, cap * 32); p += cap * 4;
    memcpy(p, tree_2 + cap_off, cap * 32); p += cap * 4;
    memcpy(p, tree_q + cap_off, cap * 32); p += cap * 4;
    memcpy(p, fin0, sh.n_final * 8); p += sh.n_final;
    memcpy(p, fin1, sh.n_final * 8); p += sh.n_final;
    memcpy(p, at_z, n_at_z * 16); p += 2 * n_at_z;
    memcpy(p, &at_zw, 16); p += 2;
    memcpy(p, at_0, n_at_0 * 16); p += 2 * n_at_0;
    for (uint32_t k = 0; k < NF; k++) {
        memcpy(p, fri_tree[k] + 4 * (2 * sh.fri_leaves[k] - 2 * sh.fri_cap[k]), sh.fri_cap[k] * 32);
        p += sh.fri_cap[k] * 4;
    }
    for (uint32_t q = 0; q < cfg->n_queries; q++) {
        size_t idx = tr_query_index(&tr, (uint32_t)(log_n + log_lde));
        const uint64_t *ldes[4] = {lde_w, lde_2, lde_q, lde_s};
        const uint64_t *trees[4] = {tree_w, tree_2, tree_q, tree_s};
        const uint32_t widths[4] = {W, S2, Q, S};
        for (int o = 0; o < 4; o++) {
            for (uint32_t i = 0; i < widths[o]; i++) *p++ = ldes[o][(size_t)i * LN + idx];
            orc_merkle_path(trees[o], LN, cap, idx, p); p += sh.depth * 4;
        }
        size_t di = idx;
        for (uint32_t k = 0; k < NF; k++) {
            size_t epl = (size_t)1 << cfg->fri_schedule[k];
            size_t leaf = di >> cfg->fri_schedule[k];
            for (size_t e = 0; e < epl; e++) *p++ = fri_c0[k][leaf * epl + e];
            for (size_t e = 0; e < epl; e++) *p++ = fri_c1[k][leaf * epl + e];
            orc_merkle_path(fri_tree[k], sh.fri_leaves[k], sh.fri_cap[k], leaf, p); p += sh.fri_depth[k] * 4;
            di = leaf;
        }
    }
    *p++ = 0; /* pow_challenge (NoPow) */
    long written = (long)(p - proof);

    if (NF == 0 || cur0 != fri_c0[NF - 1]) { free(cur0); free(cur1); }
    for (uint32_t k = 0; k < NF; k++) { free(fri_c0[k]); free(fri_c1[k]); free(fri_tree[k]); }
    free(fin0); free(fin1); free(phip); free(at_z); free(at_0); free(ch.alpha_pow);
    free(mono_s); free(lde_s); free(tree_s); free(mono_w); free(lde_w); free(tree_w);
    free(mono_2); free(lde_2); free(tree_2); free(mono_q); free(lde_q); free(tree_q);
    tr_free(&tr);
    return written;
}

/* ------------------------------------------------------------------ verifier (the oracle's own acceptance check)
 * Mirrors `verifier.verify::<H, TR, POW>((), vk, proof)` (/root/reference/src/prover_utils.rs:351-372) for the flat proof buffer.
 * Written independently of the product's CPU verifier (csrc/host.cu): the DEEP check is deep_point() above -- the function the
 * golden fixtures pin --, the FRI check is orc_fri_fold_leaf (pinned by tests/golden/fri_chain_*.json), and the quotient identity
 * at z evaluates every gate over Ext2 WITHOUT an Ext2 gate library: a gate relation is a polynomial of degree <= 7 in its cells
 * and constants, so R(a + u b) is recovered from the eight base-field evaluations R(a + t b), t = 0..7, of the base-field gate
 * library (gates.h) by interpolation in t followed by t^2 -> 7.
 * Returns 0 when the proof is accepted, otherwise a positive code; `msg` gets a one-line reason. */
#define V_FAIL(code, ...) do { snprintf(msg, msg_len, __VA_ARGS__); rc = (code); goto done; } while (0)

static void interp8_init(uint64_t minv[8][8]) { /* inverse of the Vandermonde matrix of the points 0..7 */
    uint64_t a[8][16];
    for (int i = 0; i < 8; i++) {
        uint64_t pw = 1;
        for (int j = 0; j < 8; j++) { a[i][j] = pw; pw = gl_mul(pw, (uint64_t)i); a[i][8 + j] = (i == j); }
    }
    for (int c = 0; c < 8; c++) {
        int piv = c;
        while (a[piv][c] == 0) piv++;
        if (piv != c) for (int j = 0; j < 16; j++) { uint64_t t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        uint64_t inv = gl_inv(a[c][c]);
        for (int j = 0; j < 16; j++) a[c][j] = gl_mul(a[c][j], inv);
        for (int r = 0; r < 8; r++) {
            if (r == c || a[r][c] == 0) continue;
            uint64_t f = a[r][c];
            for (int j = 0; j < 16; j++) a[r][j] = gl_sub(a[r][j], gl_mul(f, a[c][j]));
        }
    }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) minv[i][j] = a[i][8 + j];
}

/* skip: bit 0 -- do not check the quotient identity at z (diagnostic: every OTHER check -- Fiat-Shamir replay, Merkle paths of all
 * queries against the proof's caps and the verification key, lookup sum, DEEP combination, FRI folds, final polynomial -- runs
 * unchanged; tools/golden_verify.py uses it on the reference's golden proofs, whose gate polynomials are not pinned yet). */
EXPORT int orc_verify_ex(const zkgpu_geometry *g, const zkgpu_proof_config *cfg, const uint64_t *vk_cap, const uint64_t *proof, size_t len,
                         unsigned skip, char *msg, size_t msg_len) {
    int rc = 0;
    shape_t sh; make_shape(g, cfg, &sh);
    const uint32_t W = sh.W, S = sh.S, S2 = sh.S2, Q = sh.Q, NP = n_perm(g), C = n_chunks(g), E2 = n_s2_ext(g), QD = g->quotient_degree;
    const uint32_t n_at_z = sh.n_at_z, n_at_0 = sh.n_at_0, NF = cfg->n_fri_oracles;
    const size_t N = sh.N, cap = cfg->cap_size;
    const int log_n = g->log_n, log_ln = log_n + cfg->log_lde;
    const uint64_t omega = gl_omega(log_n);
    open_src *src = NULL; gl2 *wz = NULL, *sz = NULL, *ez = NULL, *qz = NULL, *phip = NULL; uint64_t *scratch = NULL, *cells = NULL;
    chal_t ch; memset(&ch, 0, sizeof(ch));
    gl_copy_permutation_non_residues(ch.knr, n_perm(g), (int)g->log_n);
    tr_t tr; tr_init(&tr);
    uint64_t *canon_copy = NULL;
    if (msg_len) msg[0] = 0;
    if (len != orc_proof_size_u64(g, cfg)) V_FAIL(1, "proof length does not match geometry and config");
    if (proof[0] != PROOF_MAGIC || proof[1] != (uint64_t)log_n || proof[2] != cfg->log_lde || proof[3] != cap || proof[4] != cfg->n_queries ||
        proof[5] != NF || proof[10] != n_at_z || proof[12] != n_at_0 || proof[13] != g->n_public_inputs || proof[14] != sh.n_final)
        V_FAIL(2, "proof header does not match geometry and config");
    for (uint32_t k = 0; k < NF; k++) if (proof[16 + k] != cfg->fri_schedule[k]) V_FAIL(2, "folding schedule in the header differs");
    /* boojum serialises Goldilocks elements as raw u64 and accepts representatives >= p (golden base-layer proofs 4 and 8
     * contain some); they denote the same field elements, so they are reduced here instead of being rejected */
    {
        int noncanon = 0;
        for (size_t i = 32; i < len; i++) noncanon |= proof[i] >= GL_P;
        if (noncanon) {
            canon_copy = alloc_u64(len);
            memcpy(canon_copy, proof, 32 * 8);
            for (size_t i = 32; i < len; i++) canon_copy[i] = proof[i] >= GL_P ? proof[i] - GL_P : proof[i];
            proof = canon_copy;
        }
    }

    const uint64_t *p = proof + 32;
    const uint64_t *pi = p; p += g->n_public_inputs;
    const uint64_t *cap_w = p; p += cap * 4;
    const uint64_t *cap_2 = p; p += cap * 4;
    const uint64_t *cap_q = p; p += cap * 4;
    const uint64_t *fin0 = p; p += sh.n_final;
    const uint64_t *fin1 = p; p += sh.n_final;
    const gl2 *at_z = (const gl2 *)p; p += 2 * n_at_z;
    const gl2 at_zw = *(const gl2 *)p; p += 2;
    const gl2 *at_0 = (const gl2 *)p; p += 2 * n_at_0;
    const uint64_t *fri_cap[ZKGPU_MAX_FRI_ORACLES];
    for (uint32_t k = 0; k < NF; k++) { fri_cap[k] = p; p += sh.fri_cap[k] * 4; }
    const uint64_t *queries = p;

    /* ---- Fiat-Shamir replay, in the prover's order */
    tr_absorb(&tr, vk_cap, cap * 4);
    tr_absorb(&tr, pi, g->n_public_inputs);
    tr_absorb(&tr, cap_w, cap * 4);
    ch.beta = tr_challenge_ext(&tr); ch.gamma = tr_challenge_ext(&tr);
    if (g->lookup_reps) { ch.lbeta = tr_challenge_ext(&tr); ch.lgamma = tr_challenge_ext(&tr); }
    tr_absorb(&tr, cap_2, cap * 4);
    ch.alpha = tr_challenge_ext(&tr);
    tr_absorb(&tr, cap_q, cap * 4);
    const gl2 z = tr_challenge_ext(&tr);
    tr_absorb(&tr, (const uint64_t *)at_z, 2 * n_at_z);
    tr_absorb(&tr, (const uint64_t *)&at_zw, 2);
    tr_absorb(&tr, (const uint64_t *)at_0, 2 * n_at_0);
    const gl2 phi = tr_challenge_ext(&tr);
    gl2 fri_ch[ZKGPU_MAX_FRI_ORACLES];
    for (uint32_t k = 0; k < NF; k++) { tr_absorb(&tr, fri_cap[k], sh.fri_cap[k] * 4); fri_ch[k] = tr_challenge_ext(&tr); }
    tr_absorb(&tr, fin0, sh.n_final);
    tr_absorb(&tr, fin1, sh.n_final);

    /* ---- openings back in oracle order */
    src = (open_src *)malloc(sizeof(open_src) * n_at_z);
    if (opening_sources(g, src) != n_at_z) V_FAIL(4, "opening count mismatch");
    wz = (gl2 *)calloc(W, sizeof(gl2)); sz = (gl2 *)calloc(S, sizeof(gl2)); ez = (gl2 *)calloc(E2, sizeof(gl2)); qz = (gl2 *)calloc(QD, sizeof(gl2));
    for (uint32_t i = 0; i < n_at_z; i++) {
        gl2 *dst = src[i].kind == 0 ? wz : src[i].kind == 1 ? sz : src[i].kind == 2 ? ez : qz;
        dst[src[i].idx] = at_z[i];
    }
    const gl2 *sigma = sz, *consts = sz + NP, *tables = sz + NP + g->n_const_cols;

    /* ---- quotient identity at z */
    {
        const uint32_t n_terms = count_terms(g);
        gl2 *ap = (gl2 *)malloc(sizeof(gl2) * n_terms);
        ap[0] = gl2_make(1, 0);
        for (uint32_t i = 1; i < n_terms; i++) ap[i] = gl2_mul(ap[i - 1], ch.alpha);
        ch.alpha_pow = ap;
        uint64_t minv[8][8]; interp8_init(minv);
        const uint32_t n_cells = g->n_copy + g->n_witness_plain;
        scratch = alloc_u64(8 * 1024); cells = alloc_u64((size_t)8 * (n_cells + 64));
        gl2 acc = gl2_make(0, 0);
        uint32_t k = 0;
        for (uint32_t gi = 0; gi < g->n_gates; gi++) {
            const zkgpu_gate *gt = &g->gates[gi];
            uint32_t nrel = 0;
            for (uint64_t t = 0; t < 8; t++) { /* R(a + t b): cells and gate constants moved along the same line */
                uint64_t *cv = cells + t * (n_cells + 64), *kv = cv + n_cells;
                for (uint32_t c = 0; c < n_cells; c++) {
                    const gl2 v = wz[c < g->n_copy ? c : NP + (c - g->n_copy)];
                    cv[c] = gl_add(v.c0, gl_mul(t, v.c1));
                }
                for (uint32_t c = 0; c + gt->path_len < g->n_const_cols && c < 64; c++) {
                    const gl2 v = consts[gt->path_len + c];
                    kv[c] = gl_add(v.c0, gl_mul(t, v.c1));
                }
                nrel = og_eval_gate(gt, g, cv, kv, ORC_P2_RC, scratch + t * 1024);
            }
            if (!nrel) continue;
            gl2 sel = gl2_make(1, 0);
            for (uint32_t b = 0; b < gt->path_len; b++)
                sel = gl2_mul(sel, ((gt->path_bits >> b) & 1) ? consts[b] : gl2_sub(gl2_make(1, 0), consts[b]));
            gl2 ga = gl2_make(0, 0);
            for (uint32_t r = 0; r < nrel; r++) {
                gl2 val = gl2_make(0, 0);
                uint64_t p7 = 1; /* 7^(d/2) */
                for (int d = 0; d < 8; d++) {
                    uint64_t coef = 0;
                    for (int t = 0; t < 8; t++) coef = gl_add(coef, gl_mul(minv[d][t], scratch[t * 1024 + r]));
                    if (d & 1) { val.c1 = gl_add(val.c1, gl_mul(coef, p7)); p7 = gl_mul(p7, 7); }
                    else val.c0 = gl_add(val.c0, gl_mul(coef, p7));
                }
                ga = gl2_add(ga, gl2_mul(ap[k + r], val));
            }
            acc = gl2_add(acc, gl2_mul(ga, sel));
            k += nrel;
        }
        if (g->has_boolean_col) {
            const gl2 b = wz[g->n_copy];
            acc = gl2_add(acc, gl2_mul(ap[k++], gl2_sub(gl2_sqr(b), b)));
        }
        if (g->lookup_reps) {
            const uint32_t LW = g->lookup_width;
            const gl2 *lw = wz + g->n_copy + (g->has_boolean_col ? 1 : 0);
            gl2 gp[16]; gp[0] = gl2_make(1, 0);
            for (uint32_t j = 1; j <= LW; j++) gp[j] = gl2_mul(gp[j - 1], ch.lgamma);
            const gl2 tid = gl2_mul(gp[LW], consts[g->table_id_col]);
            for (uint32_t i = 0; i < g->lookup_reps; i++) {
                gl2 den = gl2_add(ch.lbeta, tid);
                for (uint32_t j = 0; j < LW; j++) den = gl2_add(den, gl2_mul(gp[j], lw[i * LW + j]));
                acc = gl2_add(acc, gl2_mul(ap[k++], gl2_sub(gl2_mul(ez[C + i], den), gl2_make(1, 0))));
            }
            gl2 den = ch.lbeta;
            for (uint32_t j = 0; j <= LW; j++) den = gl2_add(den, gl2_mul(gp[j], tables[j]));
            acc = gl2_add(acc, gl2_mul(ap[k++], gl2_sub(gl2_mul(ez[C + g->lookup_reps], den), wz[W - 1])));
        }
        const gl2 zn_minus_1 = gl2_sub(gl2_pow(z, N), gl2_make(1, 0));
        {
            const gl2 l0 = gl2_mul(zn_minus_1, gl2_inv(gl2_mul_base(gl2_sub(z, gl2_make(1, 0)), (uint64_t)N % GL_P)));
            acc = gl2_add(acc, gl2_mul(ap[k++], gl2_mul(gl2_sub(ez[0], gl2_make(1, 0)), l0)));
            for (uint32_t j = 0; j < C; j++) {
                gl2 num = gl2_make(1, 0), den = gl2_make(1, 0);
                for (uint32_t i = j * QD; i < (j + 1) * QD && i < NP; i++) {
                    const gl2 kx = gl2_mul_base(z, ch.knr[i]);
                    num = gl2_mul(num, gl2_add(gl2_add(gl2_mul(ch.beta, kx), ch.gamma), wz[i]));
                    den = gl2_mul(den, gl2_add(gl2_add(gl2_mul(ch.beta, sigma[i]), ch.gamma), wz[i]));
                }
                const gl2 cur = (j + 1 < C) ? ez[j + 1] : at_zw;
                acc = gl2_add(acc, gl2_mul(ap[k++], gl2_sub(gl2_mul(cur, den), gl2_mul(ez[j], num))));
            }
        }
        if (k != n_terms) V_FAIL(5, "term count mismatch");
        gl2 quot = gl2_make(0, 0);
        const gl2 zn = gl2_pow(z, N);
        for (uint32_t c = QD; c-- > 0;) quot = gl2_add(gl2_mul(quot, zn), qz[c]);
        if (!(skip & 1) && !gl2_eq(acc, gl2_mul(quot, zn_minus_1))) V_FAIL(6, "quotient identity fails at z");
    }
    /* ---- lookup: sum_i A_i(0) = B(0) */
    if (n_at_0) {
        gl2 s = gl2_make(0, 0);
        for (uint32_t i = 0; i + 1 < n_at_0; i++) s = gl2_add(s, at_0[i]);
        if (!gl2_eq(s, at_0[n_at_0 - 1])) V_FAIL(7, "lookup sum check fails at 0");
    }
    /* ---- queries */
    {
        const uint32_t n_deep = n_at_z + 1 + n_at_0 + g->n_public_inputs;
        phip = (gl2 *)malloc(sizeof(gl2) * n_deep);
        phip[0] = gl2_make(1, 0);
        for (uint32_t i = 1; i < n_deep; i++) phip[i] = gl2_mul(phip[i - 1], phi);
        gl2 sum_at_z = gl2_make(0, 0);
        for (uint32_t i = 0; i < n_at_z; i++) sum_at_z = gl2_add(sum_at_z, gl2_mul(phip[i], at_z[i]));
        uint64_t pi_root[ZKGPU_MAX_PUBLIC_INPUTS];
        for (uint32_t i = 0; i < g->n_public_inputs; i++) pi_root[i] = gl_pow(omega, g->pi_row[i]);
        const gl2 zw = gl2_mul_base(z, omega);
        const uint64_t omega_ln = gl_omega(log_ln);
        const uint64_t *qp = queries;
        for (uint32_t q = 0; q < cfg->n_queries; q++) {
            const size_t idx = tr_query_index(&tr, (uint32_t)log_ln);
            const uint64_t *leaf[4], *caps[4] = {cap_w, cap_2, cap_q, vk_cap};
            const uint32_t widths[4] = {W, S2, Q, S};
            static const char *names[4] = {"witness", "stage 2", "quotient", "setup"};
            for (int o = 0; o < 4; o++) {
                leaf[o] = qp; qp += widths[o];
                if (!orc_merkle_verify(leaf[o], widths[o], qp, sh.depth, caps[o], idx)) V_FAIL(8, "query %u: %s oracle Merkle path fails", q, names[o]);
                qp += sh.depth * 4;
            }
            const uint64_t x = gl_mul(GL_GEN, gl_pow(omega_ln, bitrev32((uint32_t)idx, log_ln)));
            gl2 expect = deep_point(g, src, n_at_z, n_at_0, leaf[0], leaf[3], leaf[1], leaf[2], phip, sum_at_z, at_zw, at_0, pi, pi_root, x, z, zw);
            size_t di = idx;
            uint64_t shift = GL_GEN;
            for (uint32_t k = 0; k < NF; k++) {
                const uint32_t s = cfg->fri_schedule[k];
                const size_t epl = (size_t)1 << s, lf = di >> s, pos = di & (epl - 1);
                const uint64_t *c0 = qp, *c1 = qp + epl;
                if (!orc_merkle_verify(qp, 2 * epl, qp + 2 * epl, sh.fri_depth[k], fri_cap[k], lf)) V_FAIL(9, "query %u: FRI oracle %u Merkle path fails", q, k);
                if (c0[pos] != expect.c0 || c1[pos] != expect.c1)
                    V_FAIL(10, k == 0 ? "query %u: DEEP value differs from FRI base oracle (oracle %u)" : "query %u: fold does not land in FRI oracle %u", q, k);
                uint64_t cc[2] = {fri_ch[k].c0, fri_ch[k].c1}, out[2];
                orc_fri_fold_leaf(c0, c1, epl, (int)sh.fri_dom_log[k], shift, lf << s, cc, out);
                expect = gl2_make(out[0], out[1]);
                for (uint32_t i = 0; i < s; i++) shift = gl_sqr(shift);
                qp += 2 * epl + sh.fri_depth[k] * 4;
                di = lf;
            }
            const int ldf = (int)sh.fri_dom_log[NF];
            uint64_t out[2];
            orc_eval_ext_poly_at_base(fin0, fin1, sh.n_final, gl_mul(shift, gl_pow(gl_omega(ldf), bitrev32((uint32_t)di, ldf))), out);
            if (out[0] != expect.c0 || out[1] != expect.c1) V_FAIL(11, "query %u: last fold differs from the final polynomial", q);
        }
        if (*qp != 0) V_FAIL(12, "non-zero proof-of-work nonce (NoPow)");
    }
done:
    free(src); free(wz); free(sz); free(ez); free(qz); free(phip); free(scratch); free(cells); free(ch.alpha_pow); free(canon_copy);
    tr_free(&tr);
    return rc;
}
EXPORT int orc_verify(const zkgpu_geometry *g, const zkgpu_proof_config *cfg, const uint64_t *vk_cap, const uint64_t *proof, size_t len,
                      char *msg, size_t msg_len) {
    return orc_verify_ex(g, cfg, vk_cap, proof, len, 0, msg, msg_len);
}
