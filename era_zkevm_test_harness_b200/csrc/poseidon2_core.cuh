// poseidon2_core.cuh -- the Poseidon2-Goldilocks permutation (width 12, x^7, 4 + 22 + 4 rounds) on lazy residues.
//
// Same function as oracle/primitives.c:orc_poseidon2_permute (the tree/transcript hash `H` of the reference,
// /root/reference/src/prover_utils.rs:43), restructured for the GPU integer pipes:
//   * state lanes are arbitrary u64 representatives (glx.cuh); nothing is canonicalised between rounds,
//   * both linear layers are computed as exact 96-bit integer combinations (coefficients are tiny) followed by one
//     reduce96 per lane instead of ~60 modular additions per layer,
//   * the caller canonicalises only the lanes that leave the kernel.
// RC is any indexable holder of the 360 round constants (a __constant__ array on the device, a plain array on the host,
// which is how tests/test_lazy_poseidon_cpu.py checks this header against the oracle without a GPU).
#pragma once
#include "glx.cuh"

namespace zk {

// y = circ(2*M4, M4, M4) * x,  M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]; integer result < 64 * 2^64
GL_HD void p2x_external(uint64_t (&s)[12]) {
    glx::w96 y[12];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        glx::w96 x0 = glx::widen(s[4 * b]), x1 = glx::widen(s[4 * b + 1]), x2 = glx::widen(s[4 * b + 2]), x3 = glx::widen(s[4 * b + 3]);
        glx::w96 t0 = glx::add(x0, x1), t1 = glx::add(x2, x3);
        glx::w96 t2 = glx::add(glx::shl(x1, 1), t1), t3 = glx::add(glx::shl(x3, 1), t0);
        glx::w96 t4 = glx::add(glx::shl(t1, 2), t3), t5 = glx::add(glx::shl(t0, 2), t2);
        y[4 * b] = glx::add(t3, t5);
        y[4 * b + 1] = t5;
        y[4 * b + 2] = glx::add(t2, t4);
        y[4 * b + 3] = t4;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        glx::w96 t = glx::add(glx::add(y[i], y[4 + i]), y[8 + i]);
        s[i] = glx::reduce(glx::add(y[i], t));
        s[4 + i] = glx::reduce(glx::add(y[4 + i], t));
        s[8 + i] = glx::reduce(glx::add(y[8 + i], t));
    }
}

// y_i = 2^sh_i * x_i + sum(x),  sh = [4,14,11,8,0,5,2,9,13,6,3,12]; integer result < (2^14 + 12) * 2^64
GL_HD void p2x_internal(uint64_t (&s)[12]) {
    constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    glx::w96 sum = glx::widen(s[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) sum = glx::add(sum, s[i]);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        glx::w96 x = glx::widen(s[i]);
        if (SH[i]) x = glx::shl(x, SH[i]);
        s[i] = glx::reduce(glx::add(x, sum));
    }
}

// Two partial rounds with ONE reduction of lanes 1..11: after the first round the lanes stay exact 96-bit integers
// (2^sh * x + sum < 2^79), the second round shifts and sums those (< 2^94) and reduces.  Saves 11 reduce96 per pair of rounds.
template <typename RC>
GL_HD void p2x_partial_pair(uint64_t (&s)[12], const RC& rc, int r) {
    constexpr unsigned SH[12] = {4, 14, 11, 8, 0, 5, 2, 9, 13, 6, 3, 12};
    s[0] = glx::pow7(glx::add_canon(s[0], rc[12 * r]));
    glx::w96 sum = glx::widen(s[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) sum = glx::add(sum, s[i]);
    glx::w96 y[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        glx::w96 x = glx::widen(s[i]);
        if (SH[i]) x = glx::shl(x, SH[i]);
        y[i] = glx::add(x, sum);
    }
    const uint64_t y0 = glx::pow7(glx::add_canon(glx::reduce(y[0]), rc[12 * (r + 1)]));
    glx::w96 sum2 = glx::widen(y0);
#pragma unroll
    for (int i = 1; i < 12; i++) sum2 = glx::add(sum2, y[i]);
    s[0] = glx::reduce(glx::add(glx::shl(glx::widen(y0), SH[0]), sum2));
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = glx::reduce(glx::add(SH[i] ? glx::shl(y[i], SH[i]) : y[i], sum2));
}

// One full round on 4 lanes at a time with the state rotated by 4 between iterations, so the S-box code (4 x 4
// multiplies) exists ONCE in the instruction stream and is executed 3 times per round: the whole permutation is ~12 KB of
// SASS instead of ~45 KB fully unrolled.  The hash kernels are instruction-fetch-bound otherwise (ncu: stall_no_instruction
// dominated at 96 KB of code, profiles/r01_b_*).
template <typename RC>
GL_HD void p2x_full_round(uint64_t (&s)[12], const RC& rc, int r) {
#ifdef ZK_P2_ROTATE_STATE
#pragma unroll 1
    for (int it = 0; it < 3; it++) {
        uint64_t n0 = glx::pow7(glx::add_canon(s[0], rc[12 * r + 4 * it]));
        uint64_t n1 = glx::pow7(glx::add_canon(s[1], rc[12 * r + 4 * it + 1]));
        uint64_t n2 = glx::pow7(glx::add_canon(s[2], rc[12 * r + 4 * it + 2]));
        uint64_t n3 = glx::pow7(glx::add_canon(s[3], rc[12 * r + 4 * it + 3]));
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = s[i + 4];
        s[8] = n0; s[9] = n1; s[10] = n2; s[11] = n3;
    }
#else
    // S-box layer in place, 12 lanes straight-line: 3x the S-box code of the rotating form, but no 16 register moves per 4 lanes
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = glx::pow7(glx::add_canon(s[i], rc[12 * r + i]));
#endif
    p2x_external(s);
}

// in: any representatives; out: any representatives (canonicalise what you export)
template <typename RC>
GL_HD void p2x_permute(uint64_t (&s)[12], const RC& rc) {
    p2x_external(s);
    int r = 0;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
        for (int k = 0; k < 4; k++, r++) p2x_full_round(s, rc, r);
        if (half == 0) {
#ifndef ZK_P2_SINGLE_PARTIAL   // paired partial rounds: 1.575 -> 1.619 G permutations/s (tools/microbench/p2bench.cu)
#pragma unroll 1
            for (int k = 0; k < 11; k++, r += 2) p2x_partial_pair(s, rc, r);
#else
#pragma unroll 1
            for (int k = 0; k < 22; k++, r++) {
                s[0] = glx::pow7(glx::add_canon(s[0], rc[12 * r]));
                p2x_internal(s);
            }
#endif
        }
    }
}

}  // namespace zk
