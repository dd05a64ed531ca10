/* oracle/gates.h -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the gate constraint polynomials (base field),
 * independent of the product's templated csrc/gates.cuh.  Gate kinds and variable layouts: include/zkgpu.h.
 * Reference: gate sets per circuit /root/reference/circuit_definitions/src/circuit_definitions/base_layer/vm_main.rs:55-117
 * (and siblings); the polynomials are boojum's evaluators (un-vendored) -- restated from their published definitions.
 */
#ifndef ORACLE_GATES_H
#define ORACLE_GATES_H
#include "../include/zkgpu.h"
#include "gl64.h"

static inline uint32_t og_width(uint32_t kind) {
    static const uint32_t W[ZKGPU_GATE_KINDS] = {0, 1, 4, 5, 4, 13, 3, 5, 9, 26, 130, 8, 5, 1, 24, 24, 2, 17, 2, 1};
    return kind < ZKGPU_GATE_KINDS ? W[kind] : 0;
}
static inline uint32_t og_relations(uint32_t kind) {
    static const uint32_t R[ZKGPU_GATE_KINDS] = {0, 1, 1, 1, 1, 4, 2, 2, 1, 1, 118, 2, 1, 1, 12, 12, 1, 8, 2, 1};
    return kind < ZKGPU_GATE_KINDS ? R[kind] : 0;
}
/* gate cells: the copy columns followed by the plain witness columns (include/zkgpu.h: n_witness_plain) */
static inline uint32_t og_instances(const zkgpu_gate *g, const zkgpu_geometry *geo) {
    uint32_t w = og_width(g->kind), n_copy = geo->n_copy, n_plain = geo->n_witness_plain;
    if (!w) return 0;
    switch (g->kind) {
    case ZKGPU_GATE_CONSTANTS_ALLOCATOR: return g->n_consts;
    case ZKGPU_GATE_POSEIDON2_FLATTENED: return (n_copy + n_plain) / 130;
    case ZKGPU_GATE_BOUNDED_BOOLEAN: return n_copy < 10 ? n_copy : 10;
    case ZKGPU_GATE_ZERO_CHECK_WITNESS: return n_copy / 2 < n_plain ? n_copy / 2 : n_plain;
    default: return n_copy / w;
    }
}
static inline uint32_t og_total_terms(const zkgpu_geometry *geo) {
    uint32_t t = 0;
    for (uint32_t i = 0; i < geo->n_gates; i++) t += og_instances(&geo->gates[i], geo) * og_relations(geo->gates[i].kind);
    return t;
}

/* Poseidon2 linear layers as explicit matrices (deliberately NOT the addition chains of the product code) */
static void og_p2_external(uint64_t *s) {
    static const uint64_t M4[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
    uint64_t o[12];
    for (int i = 0; i < 12; i++) {
        uint64_t a = 0;
        for (int j = 0; j < 12; j++) {
            uint64_t c = M4[i % 4][j % 4] * ((i / 4 == j / 4) ? 2 : 1);
            a = gl_add(a, gl_mul(c, s[j]));
        }
        o[i] = a;
    }
    for (int i = 0; i < 12; i++) s[i] = o[i];
}
static void og_p2_internal(uint64_t *s) {
    static const int SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    uint64_t sum = 0;
    for (int i = 0; i < 12; i++) sum = gl_add(sum, s[i]);
    for (int i = 0; i < 12; i++) s[i] = gl_add(gl_mul(s[i], (uint64_t)1 << SH[i]), sum);
}
static inline uint64_t og_pow7(uint64_t x) { return gl_mul(gl_mul(gl_sqr(gl_sqr(x)), gl_sqr(x)), x); }

/* writes the relation values of gate g at one point into out[], returns how many.  v = gate CELL values (copy columns,
 * then plain witness columns), k = the gate's constants (constant columns starting at path_len), rc = Poseidon2 round constants. */
static __attribute__((unused)) uint32_t og_eval_gate(const zkgpu_gate *g, const zkgpu_geometry *geo, const uint64_t *v, const uint64_t *k, const uint64_t *rc, uint64_t *out) {
    uint32_t inst = og_instances(g, geo), n = 0;
    const uint32_t n_copy = geo->n_copy;
    switch (g->kind) {
    case ZKGPU_GATE_CONSTANTS_ALLOCATOR:
        for (uint32_t t = 0; t < inst; t++) out[n++] = gl_sub(v[t], k[t]);
        break;
    case ZKGPU_GATE_FMA:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 4 * t;
            out[n++] = gl_sub(gl_add(gl_mul(gl_mul(k[0], x[0]), x[1]), gl_mul(k[1], x[2])), x[3]);
        }
        break;
    case ZKGPU_GATE_REDUCTION4:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 5 * t;
            uint64_t s = 0;
            for (int i = 0; i < 4; i++) s = gl_add(s, gl_mul(k[i], x[i]));
            out[n++] = gl_sub(s, x[4]);
        }
        break;
    case ZKGPU_GATE_SELECTION:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 4 * t;
            uint64_t one_minus_s = gl_sub(1, x[0]);
            out[n++] = gl_sub(gl_add(gl_mul(x[0], x[1]), gl_mul(one_minus_s, x[2])), x[3]);
        }
        break;
    case ZKGPU_GATE_PARALLEL_SELECTION4:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 13 * t;
            uint64_t one_minus_s = gl_sub(1, x[0]);
            for (int i = 0; i < 4; i++)
                out[n++] = gl_sub(gl_add(gl_mul(x[0], x[1 + 3 * i]), gl_mul(one_minus_s, x[2 + 3 * i])), x[3 + 3 * i]);
        }
        break;
    case ZKGPU_GATE_ZERO_CHECK:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 3 * t;
            out[n++] = gl_sub(gl_mul(x[0], x[1]), gl_sub(1, x[2]));
            out[n++] = gl_mul(x[0], x[2]);
        }
        break;
    case ZKGPU_GATE_UINTX_ADD:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 5 * t;
            uint64_t lhs = gl_add(gl_add(x[0], x[1]), x[2]);
            uint64_t rhs = gl_add(x[3], gl_mul(k[0], x[4]));
            out[n++] = gl_sub(lhs, rhs);
            out[n++] = gl_sub(gl_sqr(x[4]), x[4]);
        }
        break;
    case ZKGPU_GATE_U32_TRI_ADD_CARRY: /* a + b + c = out + 2^32 * carry */
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 5 * t;
            uint64_t rhs = gl_add(x[3], gl_mul(x[4], (uint64_t)1 << 32));
            out[n++] = gl_sub(gl_add(gl_add(x[0], x[1]), x[2]), rhs);
        }
        break;
    case ZKGPU_GATE_BOUNDED_BOOLEAN:
    case ZKGPU_GATE_BOOLEAN_ALL:
        for (uint32_t t = 0; t < inst; t++) out[n++] = gl_sub(gl_sqr(v[t]), v[t]);
        break;
    case ZKGPU_GATE_MATMUL12_EXTERNAL:
    case ZKGPU_GATE_MATMUL12_INNER:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 24 * t;
            uint64_t s[12];
            for (int i = 0; i < 12; i++) s[i] = x[i];
            if (g->kind == ZKGPU_GATE_MATMUL12_EXTERNAL) og_p2_external(s);
            else og_p2_internal(s);
            for (int i = 0; i < 12; i++) out[n++] = gl_sub(x[12 + i], s[i]);
        }
        break;
    case ZKGPU_GATE_NONLINEARITY7:
        for (uint32_t t = 0; t < inst; t++) out[n++] = gl_sub(v[2 * t + 1], og_pow7(gl_add(v[2 * t], k[0])));
        break;
    case ZKGPU_GATE_CONDITIONAL_SWAP4: /* a[4], b[4], swap, result_a[4], result_b[4] */
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 17 * t;
            for (int i = 0; i < 4; i++) {
                /* swap ? b : a  and  swap ? a : b */
                out[n++] = gl_sub(gl_add(gl_mul(x[8], x[4 + i]), gl_mul(gl_sub(1, x[8]), x[i])), x[9 + i]);
                out[n++] = gl_sub(gl_add(gl_mul(x[8], x[i]), gl_mul(gl_sub(1, x[8]), x[4 + i])), x[13 + i]);
            }
        }
        break;
    case ZKGPU_GATE_ZERO_CHECK_WITNESS:
        for (uint32_t t = 0; t < inst; t++) {
            out[n++] = gl_sub(gl_mul(v[2 * t], v[n_copy + t]), gl_sub(1, v[2 * t + 1]));
            out[n++] = gl_mul(v[2 * t], v[2 * t + 1]);
        }
        break;
    case ZKGPU_GATE_DOT_PRODUCT4:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 9 * t;
            uint64_t s = 0;
            for (int i = 0; i < 4; i++) s = gl_add(s, gl_mul(x[2 * i], x[2 * i + 1]));
            out[n++] = gl_sub(s, x[8]);
        }
        break;
    case ZKGPU_GATE_U8X4_FMA:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 26 * t;
            uint64_t r = 0;
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) r = gl_add(r, gl_mul(gl_mul(x[i], x[4 + j]), (uint64_t)1 << (8 * (i + j))));
            for (int i = 0; i < 4; i++) {
                uint64_t sh = (uint64_t)1 << (8 * i);
                r = gl_add(r, gl_mul(gl_add(x[8 + i], x[12 + i]), sh));
                r = gl_sub(r, gl_mul(x[16 + i], sh));
                r = gl_sub(r, gl_mul(x[20 + i], gl_mul(sh, (uint64_t)1 << 32)));
            }
            out[n++] = r;
        }
        break;
    case ZKGPU_GATE_FMA_EXT:
        for (uint32_t t = 0; t < inst; t++) {
            const uint64_t *x = v + 8 * t;
            gl2 k0 = gl2_make(k[0], k[1]), k1 = gl2_make(k[2], k[3]);
            gl2 r = gl2_sub(gl2_add(gl2_mul(k0, gl2_mul(gl2_make(x[0], x[1]), gl2_make(x[2], x[3]))), gl2_mul(k1, gl2_make(x[4], x[5]))),
                            gl2_make(x[6], x[7]));
            out[n++] = r.c0;
            out[n++] = r.c1;
        }
        break;
    case ZKGPU_GATE_POSEIDON2_FLATTENED:
        if (inst) {
            uint64_t s[12];
            for (int i = 0; i < 12; i++) s[i] = v[i];
            og_p2_external(s);
            uint32_t col = 12;
            int r = 0;
            for (int q = 0; q < 4; q++, r++) {
                for (int i = 0; i < 12; i++) {
                    out[n++] = gl_sub(v[col + i], og_pow7(gl_add(s[i], rc[12 * r + i])));
                    s[i] = v[col + i];
                }
                col += 12;
                og_p2_external(s);
            }
            for (int q = 0; q < 22; q++, r++) {
                out[n++] = gl_sub(v[col], og_pow7(gl_add(s[0], rc[12 * r])));
                s[0] = v[col];
                col++;
                og_p2_internal(s);
            }
            for (int q = 0; q < 4; q++, r++) {
                for (int i = 0; i < 12; i++) {
                    out[n++] = gl_sub(v[col + i], og_pow7(gl_add(s[i], rc[12 * r + i])));
                    s[i] = v[col + i];
                }
                col += 12;
                og_p2_external(s);
            }
        }
        break;
    default:
        break;
    }
    return n;
}
#endif
