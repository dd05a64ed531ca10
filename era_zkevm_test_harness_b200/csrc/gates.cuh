// gates.cuh -- constraint polynomials of the gate set, written once over a generic field element type so the same
// code evaluates (a) base-field LDE values in the CUDA quotient kernel and (b) Ext2 openings at z in the CPU verifier.
//
// Reference: the gate set is chosen per circuit by `configure_builder` (e.g.
// /root/reference/circuit_definitions/src/circuit_definitions/base_layer/vm_main.rs:55-117); the polynomials themselves
// are boojum's `GateConstraintEvaluator`s (un-vendored).  Relations and their order are this framework's restatement
// (include/zkgpu.h gate kinds; DESIGN.md "Gate library") -- the CPU oracle (oracle/gates.h) restates them
// independently and tests require bit-identical quotients.
#pragma once
#include "../../include/zkgpu.h"
#include "gl.cuh"
#include "poseidon2_consts.cuh"

namespace zk {

// field-generic helpers: F is uint64_t (base) or gl::e2
GL_HD uint64_t f_add(uint64_t a, uint64_t b) { return gl::add(a, b); }
GL_HD uint64_t f_sub(uint64_t a, uint64_t b) { return gl::sub(a, b); }
GL_HD uint64_t f_mul(uint64_t a, uint64_t b) { return gl::mul(a, b); }
GL_HD uint64_t f_mulc(uint64_t a, uint64_t c) { return gl::mul(a, c); }  // by a base-field constant
GL_HD uint64_t f_addc(uint64_t a, uint64_t c) { return gl::add(a, c); }
GL_HD uint64_t f_shl(uint64_t a, unsigned k) { return gl::mul_pow2(a, k); }
GL_HD gl::e2 f_add(gl::e2 a, gl::e2 b) { return gl::add(a, b); }
GL_HD gl::e2 f_sub(gl::e2 a, gl::e2 b) { return gl::sub(a, b); }
GL_HD gl::e2 f_mul(gl::e2 a, gl::e2 b) { return gl::mul(a, b); }
GL_HD gl::e2 f_mulc(gl::e2 a, uint64_t c) { return gl::mul_base(a, c); }
GL_HD gl::e2 f_addc(gl::e2 a, uint64_t c) { return gl::make2(gl::add(a.c0, c), a.c1); }
GL_HD gl::e2 f_shl(gl::e2 a, unsigned k) { return gl::make2(gl::mul_pow2(a.c0, k), gl::mul_pow2(a.c1, k)); }
template <typename F> GL_HD F f_zero();
template <> GL_HD uint64_t f_zero<uint64_t>() { return 0; }
template <> GL_HD gl::e2 f_zero<gl::e2>() { return gl::make2(0, 0); }
template <typename F> GL_HD F f_one() { return f_addc(f_zero<F>(), 1); }
template <typename F> GL_HD F f_dbl(F a) { return f_add(a, a); }
template <typename F> GL_HD F f_pow7(F x) {
    F x2 = f_mul(x, x), x4 = f_mul(x2, x2);
    return f_mul(f_mul(x4, x2), x);
}

GL_HD uint32_t gate_width(uint32_t kind) {
    switch (kind) {
        case ZKGPU_GATE_CONSTANTS_ALLOCATOR: return 1;
        case ZKGPU_GATE_FMA: return 4;
        case ZKGPU_GATE_REDUCTION4: return 5;
        case ZKGPU_GATE_SELECTION: return 4;
        case ZKGPU_GATE_PARALLEL_SELECTION4: return 13;
        case ZKGPU_GATE_ZERO_CHECK: return 3;
        case ZKGPU_GATE_UINTX_ADD: return 5;
        case ZKGPU_GATE_DOT_PRODUCT4: return 9;
        case ZKGPU_GATE_U8X4_FMA: return 26;
        case ZKGPU_GATE_POSEIDON2_FLATTENED: return 130;
        case ZKGPU_GATE_FMA_EXT: return 8;
        case ZKGPU_GATE_U32_TRI_ADD_CARRY: return 5;
        case ZKGPU_GATE_BOUNDED_BOOLEAN: return 1;
        case ZKGPU_GATE_BOOLEAN_ALL: return 1;
        case ZKGPU_GATE_MATMUL12_EXTERNAL: return 24;
        case ZKGPU_GATE_MATMUL12_INNER: return 24;
        case ZKGPU_GATE_NONLINEARITY7: return 2;
        case ZKGPU_GATE_CONDITIONAL_SWAP4: return 17;
        case ZKGPU_GATE_ZERO_CHECK_WITNESS: return 2;   // x, flag under copy permutation; the inverse is a plain witness cell
        default: return 0;
    }
}
GL_HD uint32_t gate_relations(uint32_t kind) {
    switch (kind) {
        case ZKGPU_GATE_CONSTANTS_ALLOCATOR: return 1;
        case ZKGPU_GATE_FMA: return 1;
        case ZKGPU_GATE_REDUCTION4: return 1;
        case ZKGPU_GATE_SELECTION: return 1;
        case ZKGPU_GATE_PARALLEL_SELECTION4: return 4;
        case ZKGPU_GATE_ZERO_CHECK: return 2;
        case ZKGPU_GATE_UINTX_ADD: return 2;
        case ZKGPU_GATE_DOT_PRODUCT4: return 1;
        case ZKGPU_GATE_U8X4_FMA: return 1;
        case ZKGPU_GATE_POSEIDON2_FLATTENED: return 118;
        case ZKGPU_GATE_FMA_EXT: return 2;
        case ZKGPU_GATE_U32_TRI_ADD_CARRY: return 1;
        case ZKGPU_GATE_BOUNDED_BOOLEAN: return 1;
        case ZKGPU_GATE_BOOLEAN_ALL: return 1;
        case ZKGPU_GATE_MATMUL12_EXTERNAL: return 12;
        case ZKGPU_GATE_MATMUL12_INNER: return 12;
        case ZKGPU_GATE_NONLINEARITY7: return 1;
        case ZKGPU_GATE_CONDITIONAL_SWAP4: return 8;
        case ZKGPU_GATE_ZERO_CHECK_WITNESS: return 2;
        default: return 0;
    }
}
// Gate CELLS: cell c < n_copy is copy column c; cell c >= n_copy is plain witness column c - n_copy (compression modes 1-3).
// Only the flattened Poseidon2 gate spans both kinds of column (its 130 cells are the copy columns followed by the plain
// witness columns: 52 + 78, 56 + 74, 68 + 62 in the reference's compression circuits); ZeroCheck-with-witness keeps x and the
// flag under the copy permutation and puts instance t's inverse in plain witness cell t.
constexpr uint32_t BOUNDED_BOOLEAN_MAX_ON_ROW = 10;  // mode_{2,3,4}.rs: BoundedBooleanConstraintGate::configure_builder(.., 10)
GL_HD uint32_t gate_instances(const zkgpu_gate& g, const zkgpu_geometry& geo) {
    uint32_t w = gate_width(g.kind);
    if (!w) return 0;
    if (g.kind == ZKGPU_GATE_CONSTANTS_ALLOCATOR) return g.n_consts;
    if (g.kind == ZKGPU_GATE_POSEIDON2_FLATTENED) return (geo.n_copy + geo.n_witness_plain) / w;
    if (g.kind == ZKGPU_GATE_BOUNDED_BOOLEAN) return geo.n_copy < BOUNDED_BOOLEAN_MAX_ON_ROW ? geo.n_copy : BOUNDED_BOOLEAN_MAX_ON_ROW;
    if (g.kind == ZKGPU_GATE_ZERO_CHECK_WITNESS) return geo.n_copy / 2 < geo.n_witness_plain ? geo.n_copy / 2 : geo.n_witness_plain;
    return geo.n_copy / w;
}
GL_HD uint32_t total_gate_terms(const zkgpu_geometry& geo) {
    uint32_t t = 0;
    for (uint32_t i = 0; i < geo.n_gates; i++) t += gate_instances(geo.gates[i], geo) * gate_relations(geo.gates[i].kind);
    return t;
}

// ---- Poseidon2 linear layers over F (same matrices as poseidon2.cu) ----
template <typename F>
GL_HD void p2g_external(F (&s)[12]) {
#pragma unroll
    for (int b = 0; b < 3; b++) {
        F x0 = s[4 * b], x1 = s[4 * b + 1], x2 = s[4 * b + 2], x3 = s[4 * b + 3];
        F t0 = f_add(x0, x1), t1 = f_add(x2, x3);
        F t2 = f_add(f_dbl(x1), t1), t3 = f_add(f_dbl(x3), t0);
        F t4 = f_add(f_dbl(f_dbl(t1)), t3), t5 = f_add(f_dbl(f_dbl(t0)), t2);
        s[4 * b] = f_add(t3, t5);
        s[4 * b + 1] = t5;
        s[4 * b + 2] = f_add(t2, t4);
        s[4 * b + 3] = t4;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        F t = f_add(f_add(s[i], s[4 + i]), s[8 + i]);
        s[i] = f_add(s[i], t);
        s[4 + i] = f_add(s[4 + i], t);
        s[8 + i] = f_add(s[8 + i], t);
    }
}
template <typename F>
GL_HD void p2g_internal(F (&s)[12]) {
    constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    F sum = s[0];
#pragma unroll
    for (int i = 1; i < 12; i++) sum = f_add(sum, s[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = f_add(f_shl(s[i], SH[i]), sum);
}

// Evaluate every relation of one gate over its tiled instances.
//   acc(c) : returns the value of gate CELL c at the current point (type F); the caller maps cells to columns
//   kc(i)  : returns gate constant i (constant column path_len + i) at the current point
//   sink(r): receives each relation value in canonical order
//   rc     : Poseidon2 round constant table (360 u64)
template <typename F, typename Acc, typename Kc, typename Sink>
GL_HD void eval_gate(const zkgpu_gate& g, const zkgpu_geometry& geo, const uint64_t* __restrict__ rc, Acc&& acc, Kc&& kc, Sink&& sink) {
    const uint32_t inst = gate_instances(g, geo);
    const uint32_t n_copy = geo.n_copy;
    switch (g.kind) {
        case ZKGPU_GATE_CONSTANTS_ALLOCATOR:
            for (uint32_t t = 0; t < inst; t++) sink(f_sub(acc(t), kc(t)));
            break;
        case ZKGPU_GATE_FMA: {
            F k0 = kc(0), k1 = kc(1);
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 4 * t;
                sink(f_sub(f_add(f_mul(k0, f_mul(acc(b), acc(b + 1))), f_mul(k1, acc(b + 2))), acc(b + 3)));
            }
        } break;
        case ZKGPU_GATE_REDUCTION4: {
            F k0 = kc(0), k1 = kc(1), k2 = kc(2), k3 = kc(3);
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 5 * t;
                F s = f_add(f_add(f_mul(k0, acc(b)), f_mul(k1, acc(b + 1))), f_add(f_mul(k2, acc(b + 2)), f_mul(k3, acc(b + 3))));
                sink(f_sub(s, acc(b + 4)));
            }
        } break;
        case ZKGPU_GATE_SELECTION:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 4 * t;
                F bb = acc(b + 2);
                sink(f_sub(f_add(f_mul(acc(b), f_sub(acc(b + 1), bb)), bb), acc(b + 3)));
            }
            break;
        case ZKGPU_GATE_PARALLEL_SELECTION4:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 13 * t;
                F s = acc(b);
                for (uint32_t i = 0; i < 4; i++) {
                    F bb = acc(b + 2 + 3 * i);
                    sink(f_sub(f_add(f_mul(s, f_sub(acc(b + 1 + 3 * i), bb)), bb), acc(b + 3 + 3 * i)));
                }
            }
            break;
        case ZKGPU_GATE_ZERO_CHECK:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 3 * t;
                F x = acc(b), zf = acc(b + 2);
                sink(f_sub(f_add(f_mul(x, acc(b + 1)), zf), f_one<F>()));
                sink(f_mul(x, zf));
            }
            break;
        case ZKGPU_GATE_UINTX_ADD: {
            F k0 = kc(0);
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 5 * t;
                F co = acc(b + 4);
                sink(f_sub(f_sub(f_add(f_add(acc(b), acc(b + 1)), acc(b + 2)), acc(b + 3)), f_mul(k0, co)));
                sink(f_sub(f_mul(co, co), co));
            }
        } break;
        case ZKGPU_GATE_U32_TRI_ADD_CARRY:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 5 * t;
                sink(f_sub(f_add(f_add(acc(b), acc(b + 1)), acc(b + 2)), f_add(acc(b + 3), f_shl(acc(b + 4), 32))));
            }
            break;
        case ZKGPU_GATE_BOUNDED_BOOLEAN:
        case ZKGPU_GATE_BOOLEAN_ALL:
            for (uint32_t t = 0; t < inst; t++) {
                F x = acc(t);
                sink(f_sub(f_mul(x, x), x));
            }
            break;
        case ZKGPU_GATE_MATMUL12_EXTERNAL:
        case ZKGPU_GATE_MATMUL12_INNER:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 24 * t;
                F s[12];
                for (int i = 0; i < 12; i++) s[i] = acc(b + i);
                if (g.kind == ZKGPU_GATE_MATMUL12_EXTERNAL) p2g_external(s);
                else p2g_internal(s);
                for (int i = 0; i < 12; i++) sink(f_sub(acc(b + 12 + i), s[i]));
            }
            break;
        case ZKGPU_GATE_NONLINEARITY7: {
            F k0 = kc(0);
            for (uint32_t t = 0; t < inst; t++) sink(f_sub(acc(2 * t + 1), f_pow7(f_add(acc(2 * t), k0))));
        } break;
        case ZKGPU_GATE_CONDITIONAL_SWAP4:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 17 * t;   // a[4], b[4], should_swap, result_a[4], result_b[4]
                F sw = acc(b + 8);
                for (uint32_t i = 0; i < 4; i++) {
                    F a = acc(b + i), bb = acc(b + 4 + i);
                    F d = f_mul(sw, f_sub(bb, a));
                    sink(f_sub(f_add(d, a), acc(b + 9 + i)));
                    sink(f_sub(f_sub(bb, d), acc(b + 13 + i)));
                }
            }
            break;
        case ZKGPU_GATE_ZERO_CHECK_WITNESS:
            for (uint32_t t = 0; t < inst; t++) {
                F x = acc(2 * t), zf = acc(2 * t + 1), inv = acc(n_copy + t);
                sink(f_sub(f_add(f_mul(x, inv), zf), f_one<F>()));
                sink(f_mul(x, zf));
            }
            break;
        case ZKGPU_GATE_DOT_PRODUCT4:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 9 * t;
                F s = f_add(f_add(f_mul(acc(b), acc(b + 1)), f_mul(acc(b + 2), acc(b + 3))),
                            f_add(f_mul(acc(b + 4), acc(b + 5)), f_mul(acc(b + 6), acc(b + 7))));
                sink(f_sub(s, acc(b + 8)));
            }
            break;
        case ZKGPU_GATE_U8X4_FMA:
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 26 * t;
                // sum_{i,j} a_i b_j 2^(8(i+j)) + sum (c_i + cin_i) 2^(8i) - sum lo_i 2^(8i) - 2^32 sum hi_i 2^(8i)
                F bv[4];
                for (uint32_t j = 0; j < 4; j++) bv[j] = acc(b + 4 + j);
                F r = f_zero<F>();
                for (uint32_t i = 0; i < 4; i++) {
                    F ai = acc(b + i);
                    F row = f_zero<F>();
                    for (uint32_t j = 0; j < 4; j++) row = f_add(row, f_shl(f_mul(ai, bv[j]), 8 * j));
                    r = f_add(r, f_shl(row, 8 * i));
                }
                for (uint32_t i = 0; i < 4; i++) {
                    F lin = f_sub(f_add(acc(b + 8 + i), acc(b + 12 + i)), f_add(acc(b + 16 + i), f_shl(acc(b + 20 + i), 32)));
                    r = f_add(r, f_shl(lin, 8 * i));
                }
                sink(r);
            }
            break;
        case ZKGPU_GATE_FMA_EXT: {
            // (k0 a b + k1 c - d) over Ext2 with coordinates in separate variables; u^2 = 7
            F k00 = kc(0), k01 = kc(1), k10 = kc(2), k11 = kc(3);
            for (uint32_t t = 0; t < inst; t++) {
                uint32_t b = 8 * t;
                F a0 = acc(b), a1 = acc(b + 1), b0 = acc(b + 2), b1 = acc(b + 3), c0 = acc(b + 4), c1 = acc(b + 5);
                F ab0 = f_add(f_mul(a0, b0), f_mulc(f_mul(a1, b1), 7)), ab1 = f_add(f_mul(a0, b1), f_mul(a1, b0));
                F p0 = f_add(f_mul(k00, ab0), f_mulc(f_mul(k01, ab1), 7)), p1 = f_add(f_mul(k00, ab1), f_mul(k01, ab0));
                F q0 = f_add(f_mul(k10, c0), f_mulc(f_mul(k11, c1), 7)), q1 = f_add(f_mul(k10, c1), f_mul(k11, c0));
                sink(f_sub(f_add(p0, q0), acc(b + 6)));
                sink(f_sub(f_add(p1, q1), acc(b + 7)));
            }
        } break;
        case ZKGPU_GATE_POSEIDON2_FLATTENED: {
            if (inst == 0) break;
            // columns: [0,12) input state; then one variable per S-box output in round order (48 + 22 + 48)
            F s[12];
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = acc(i);
            p2g_external(s);
            uint32_t col = 12;
            int r = 0;
#pragma unroll 1
            for (int k = 0; k < 4; k++, r++) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    F v = acc(col + i);
                    sink(f_sub(v, f_pow7(f_addc(s[i], rc[12 * r + i]))));
                    s[i] = v;
                }
                col += 12;
                p2g_external(s);
            }
#pragma unroll 1
            for (int k = 0; k < 22; k++, r++) {
                F v = acc(col);
                sink(f_sub(v, f_pow7(f_addc(s[0], rc[12 * r]))));
                s[0] = v;
                col += 1;
                p2g_internal(s);
            }
#pragma unroll 1
            for (int k = 0; k < 4; k++, r++) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    F v = acc(col + i);
                    sink(f_sub(v, f_pow7(f_addc(s[i], rc[12 * r + i]))));
                    s[i] = v;
                }
                col += 12;
                p2g_external(s);
            }
        } break;
        default: break;
    }
}

}  // namespace zk
