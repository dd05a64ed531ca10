/* oracle/synth.c -- TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline / reference arm).
 *
 * Synthetic satisfying trace of a circuit geometry, in plain C: the same generator as the product's host-side
 * zkgpu_synth_trace (csrc/host.cu), restated here so that the CPU reference arm of bench.py can build its input without loading
 * libzkgpu.so, and so that the two generators can be checked against each other (tests/test_prover_cpu.py).
 * It stands in for the reference's Rust circuit synthesis (`synthesis_inner`,
 * /root/reference/circuit_definitions/src/circuit_definitions/base_layer/mod.rs:286-313), which cannot run in this image.
 *
 * Layout (column-major u64, natural row order): witness = copy columns, boolean column, lookup columns, plain witness columns,
 * multiplicities; setup = sigmas, constant columns, lookup table columns (width + 1).  Row r carries gate r mod n_gates with
 * that gate's selector path and constants in the constant columns.
 */
#include <stdlib.h>
#include <string.h>
#include "gates.h"
#include "poseidon2_consts.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct { uint64_t s; } splitmix;
static uint64_t sm_next(splitmix *m) {
    uint64_t z = (m->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static uint64_t sm_field(splitmix *m) { return sm_next(m) % GL_P; }

static uint64_t table_entry(uint32_t t, uint32_t j, uint32_t width) {
    if (j == width) return 1; /* table id */
    if (j == 0) return t;
    splitmix m = {0x7AB1E000ULL + (uint64_t)t * 16 + j};
    return sm_next(&m) & 0xFFFFFFFFULL;
}

/* split = 0: one stream (seed) drives witness values and gate constants; split = 1: setup_seed fixes the gate constants of the
 * circuit TYPE, seed the free witness values of one INSTANCE. */
EXPORT int orc_synth_trace(const zkgpu_geometry *g, uint64_t setup_seed, uint64_t seed, int split, uint64_t *wit, uint64_t *setup) {
    const size_t N = (size_t)1 << g->log_n;
    const uint32_t NP = g->n_copy + (g->has_boolean_col ? 1 : 0) + g->lookup_width * g->lookup_reps;
    const uint32_t lookup_col0 = g->n_copy + (g->has_boolean_col ? 1 : 0), plain_col0 = NP;
    const uint32_t W = NP + g->n_witness_plain + (g->lookup_reps ? 1 : 0);
    uint64_t *sigma = setup, *consts = setup + (size_t)NP * N, *tables = consts + (size_t)g->n_const_cols * N;
    memset(consts, 0, (size_t)g->n_const_cols * N * 8);
    const uint64_t omega = gl_omega((int)g->log_n);

    uint64_t knr[1024]; /* copy-permutation non-residues k_i (gl64.h) */
    gl_copy_permutation_non_residues(knr, NP, (int)g->log_n);
    { /* identity permutation: sigma_i(w^r) = k_i * w^r */
        uint64_t *wp = (uint64_t *)malloc(N * 8), x = 1;
        for (size_t r = 0; r < N; r++) { wp[r] = x; x = gl_mul(x, omega); }
        for (uint32_t i = 0; i < NP; i++)
            for (size_t r = 0; r < N; r++) sigma[(size_t)i * N + r] = gl_mul(knr[i], wp[r]);
        free(wp);
    }
    uint64_t *mult = NULL;
    if (g->lookup_reps) {
        mult = (uint64_t *)calloc(N, 8);
        for (uint32_t j = 0; j <= g->lookup_width; j++)
            for (size_t r = 0; r < N; r++) tables[(size_t)j * N + r] = r < g->table_len ? table_entry((uint32_t)r, j, g->lookup_width) : 0;
    }
    long last_fma_row = -1;
    splitmix rng = {seed ^ 0xB2000000ULL}, rng_setup = {setup_seed ^ 0x5E7000000ULL};
    splitmix *rk = split ? &rng_setup : &rng;
    const uint32_t n_cells = g->n_copy + g->n_witness_plain;
    uint64_t *v = (uint64_t *)malloc((n_cells + 1) * 8), kc[64];
    for (size_t r = 0; r < N; r++) {
        for (uint32_t c = 0; c < n_cells; c++) v[c] = sm_field(&rng);
        memset(kc, 0, sizeof kc);
        const zkgpu_gate *gt = g->n_gates ? &g->gates[r % g->n_gates] : NULL;
        if (gt) {
            for (uint32_t b = 0; b < gt->path_len; b++) kc[b] = (gt->path_bits >> b) & 1;
            uint64_t *k = kc + gt->path_len;
            const uint32_t inst = og_instances(gt, g);
            switch (gt->kind) {
            case ZKGPU_GATE_CONSTANTS_ALLOCATOR:
                for (uint32_t t = 0; t < inst; t++) { k[t] = sm_field(rk); v[t] = k[t]; }
                break;
            case ZKGPU_GATE_FMA: {
                k[0] = sm_field(rk); k[1] = sm_field(rk);
                const long prev = last_fma_row;
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 4 * t;
                    if (prev >= 0) { /* wire: input a of this row = output d of the previous FMA row (a 2-cycle in sigma) */
                        x[0] = wit[(size_t)(4 * t + 3) * N + prev];
                        const uint64_t ka = knr[4 * t], kd = knr[4 * t + 3];
                        sigma[(size_t)(4 * t) * N + r] = gl_mul(kd, gl_pow(omega, (uint64_t)prev));
                        sigma[(size_t)(4 * t + 3) * N + prev] = gl_mul(ka, gl_pow(omega, r));
                    }
                    x[3] = gl_add(gl_mul(k[0], gl_mul(x[0], x[1])), gl_mul(k[1], x[2]));
                }
                last_fma_row = (long)r;
            } break;
            case ZKGPU_GATE_REDUCTION4:
                for (int i = 0; i < 4; i++) k[i] = sm_field(rk);
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 5 * t, s = 0;
                    for (int i = 0; i < 4; i++) s = gl_add(s, gl_mul(k[i], x[i]));
                    x[4] = s;
                }
                break;
            case ZKGPU_GATE_SELECTION:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 4 * t;
                    x[0] = sm_next(&rng) & 1;
                    x[3] = x[0] ? x[1] : x[2];
                }
                break;
            case ZKGPU_GATE_PARALLEL_SELECTION4:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 13 * t;
                    x[0] = sm_next(&rng) & 1;
                    for (int i = 0; i < 4; i++) x[3 + 3 * i] = x[0] ? x[1 + 3 * i] : x[2 + 3 * i];
                }
                break;
            case ZKGPU_GATE_ZERO_CHECK:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 3 * t;
                    if (sm_next(&rng) & 3) { if (!x[0]) x[0] = 5; x[1] = gl_inv(x[0]); x[2] = 0; }
                    else { x[0] = 0; x[2] = 1; }
                }
                break;
            case ZKGPU_GATE_UINTX_ADD:
                k[0] = (uint64_t)1 << 32;
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 5 * t;
                    x[0] = sm_next(&rng) & 0xFFFFFFFFULL; x[1] = sm_next(&rng) & 0xFFFFFFFFULL; x[2] = sm_next(&rng) & 1;
                    const uint64_t s = x[0] + x[1] + x[2];
                    x[3] = s & 0xFFFFFFFFULL; x[4] = s >> 32;
                }
                break;
            case ZKGPU_GATE_U32_TRI_ADD_CARRY:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 5 * t;
                    for (int i = 0; i < 3; i++) x[i] = sm_next(&rng) & 0xFFFFFFFFULL;
                    const uint64_t s = x[0] + x[1] + x[2];
                    x[3] = s & 0xFFFFFFFFULL; x[4] = s >> 32;
                }
                break;
            case ZKGPU_GATE_BOUNDED_BOOLEAN:
            case ZKGPU_GATE_BOOLEAN_ALL:
                for (uint32_t t = 0; t < inst; t++) v[t] = sm_next(&rng) & 1;
                break;
            case ZKGPU_GATE_MATMUL12_EXTERNAL:
            case ZKGPU_GATE_MATMUL12_INNER:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 24 * t, s[12];
                    for (int i = 0; i < 12; i++) s[i] = x[i];
                    if (gt->kind == ZKGPU_GATE_MATMUL12_EXTERNAL) og_p2_external(s);
                    else og_p2_internal(s);
                    for (int i = 0; i < 12; i++) x[12 + i] = s[i];
                }
                break;
            case ZKGPU_GATE_NONLINEARITY7:
                k[0] = sm_field(rk);
                for (uint32_t t = 0; t < inst; t++) v[2 * t + 1] = og_pow7(gl_add(v[2 * t], k[0]));
                break;
            case ZKGPU_GATE_CONDITIONAL_SWAP4:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 17 * t;
                    x[8] = sm_next(&rng) & 1;
                    for (int i = 0; i < 4; i++) { x[9 + i] = x[8] ? x[4 + i] : x[i]; x[13 + i] = x[8] ? x[i] : x[4 + i]; }
                }
                break;
            case ZKGPU_GATE_ZERO_CHECK_WITNESS:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 2 * t, *inv = &v[g->n_copy + t];
                    if (sm_next(&rng) & 3) { if (!x[0]) x[0] = 5; *inv = gl_inv(x[0]); x[1] = 0; }
                    else { x[0] = 0; x[1] = 1; }
                }
                break;
            case ZKGPU_GATE_DOT_PRODUCT4:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 9 * t, s = 0;
                    for (int i = 0; i < 4; i++) s = gl_add(s, gl_mul(x[2 * i], x[2 * i + 1]));
                    x[8] = s;
                }
                break;
            case ZKGPU_GATE_U8X4_FMA:
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 26 * t;
                    const uint64_t a = sm_next(&rng) & 0x7FFFFFFFULL, b = sm_next(&rng) & 0x7FFFFFFFULL, c = sm_next(&rng) & 0xFFFFFFFFULL,
                                   ci = sm_next(&rng) & 0xFFFFFFFFULL;
                    const uint64_t tot = a * b + c + ci;
                    for (int i = 0; i < 4; i++) {
                        x[i] = (a >> (8 * i)) & 0xFF; x[4 + i] = (b >> (8 * i)) & 0xFF; x[8 + i] = (c >> (8 * i)) & 0xFF;
                        x[12 + i] = (ci >> (8 * i)) & 0xFF; x[16 + i] = (tot >> (8 * i)) & 0xFF; x[20 + i] = (tot >> (32 + 8 * i)) & 0xFF;
                    }
                }
                break;
            case ZKGPU_GATE_FMA_EXT:
                for (int i = 0; i < 4; i++) k[i] = sm_field(rk);
                for (uint32_t t = 0; t < inst; t++) {
                    uint64_t *x = v + 8 * t;
                    const gl2 d = gl2_add(gl2_mul(gl2_make(k[0], k[1]), gl2_mul(gl2_make(x[0], x[1]), gl2_make(x[2], x[3]))),
                                          gl2_mul(gl2_make(k[2], k[3]), gl2_make(x[4], x[5])));
                    x[6] = d.c0; x[7] = d.c1;
                }
                break;
            case ZKGPU_GATE_POSEIDON2_FLATTENED:
                if (inst) {
                    uint64_t s[12];
                    for (int i = 0; i < 12; i++) s[i] = v[i];
                    og_p2_external(s);
                    uint32_t col = 12;
                    int rr = 0;
                    for (int q = 0; q < 4; q++, rr++) {
                        for (int i = 0; i < 12; i++) { s[i] = og_pow7(gl_add(s[i], ORC_P2_RC[12 * rr + i])); v[col + i] = s[i]; }
                        col += 12;
                        og_p2_external(s);
                    }
                    for (int q = 0; q < 22; q++, rr++) {
                        s[0] = og_pow7(gl_add(s[0], ORC_P2_RC[12 * rr]));
                        v[col++] = s[0];
                        og_p2_internal(s);
                    }
                    for (int q = 0; q < 4; q++, rr++) {
                        for (int i = 0; i < 12; i++) { s[i] = og_pow7(gl_add(s[i], ORC_P2_RC[12 * rr + i])); v[col + i] = s[i]; }
                        col += 12;
                        og_p2_external(s);
                    }
                }
                break;
            default: break;
            }
        }
        for (uint32_t c = 0; c < g->n_copy; c++) wit[(size_t)c * N + r] = v[c];
        for (uint32_t c = 0; c < g->n_witness_plain; c++) wit[(size_t)(plain_col0 + c) * N + r] = v[g->n_copy + c];
        if (g->has_boolean_col) wit[(size_t)g->n_copy * N + r] = sm_next(&rng) & 1;
        if (g->lookup_reps) {
            kc[g->table_id_col] = 1;
            for (uint32_t i = 0; i < g->lookup_reps; i++) {
                const uint32_t t = (uint32_t)(sm_next(&rng) % (g->table_len ? g->table_len : 1));
                mult[t]++;
                for (uint32_t j = 0; j < g->lookup_width; j++)
                    wit[(size_t)(lookup_col0 + i * g->lookup_width + j) * N + r] = table_entry(t, j, g->lookup_width);
            }
        }
        for (uint32_t c = 0; c < g->n_const_cols; c++) consts[(size_t)c * N + r] = kc[c];
    }
    if (g->lookup_reps) {
        for (size_t r = 0; r < N; r++) wit[(size_t)(W - 1) * N + r] = mult[r];
        free(mult);
    }
    free(v);
    return 0;
}
