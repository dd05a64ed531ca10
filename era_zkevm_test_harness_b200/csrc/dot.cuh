// dot.cuh -- unreduced dot products  sum_k r_k * a_k  of Goldilocks values (device only).
//
// A 64x64 product is four IMAD.WIDE.U32 accumulated into three 96-bit column accumulators (1 fma-pipe + 1 alu-pipe
// instruction per partial product) and the whole sum is reduced mod p ONCE, instead of a modular multiply (28 instructions)
// and a modular add per term.  Used wherever the prover combines many base-field values with fixed Ext2 weights: the
// alpha-powers of the quotient (quotient.cu), the DEEP challenge powers and the powers of z of the openings (prover.cu).
#pragma once
#include "gl.cuh"

namespace zk {

// sum_k r_k * a_k over the integers: A0 + 2^32*A1 + 2^64*A2, each A_i a 96-bit (l, h, c) accumulator.
struct Dot {
    uint32_t l0, h0, c0, l1, h1, c1, l2, h2, c2;
};
__device__ __forceinline__ void dot_zero(Dot& d) { d.l0 = d.h0 = d.c0 = d.l1 = d.h1 = d.c1 = d.l2 = d.h2 = d.c2 = 0; }
__device__ __forceinline__ void dot_add(Dot& d, uint64_t r, uint64_t a) {
    uint32_t r0 = (uint32_t)r, r1 = (uint32_t)(r >> 32), a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32);
    asm("mad.lo.cc.u32 %0, %9, %11, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %11, %1;\n\t"
        "addc.u32 %2, %2, 0;\n\t"
        "mad.lo.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.hi.cc.u32 %4, %9, %12, %4;\n\t"
        "addc.u32 %5, %5, 0;\n\t"
        "mad.lo.cc.u32 %3, %10, %11, %3;\n\t"
        "madc.hi.cc.u32 %4, %10, %11, %4;\n\t"
        "addc.u32 %5, %5, 0;\n\t"
        "mad.lo.cc.u32 %6, %10, %12, %6;\n\t"
        "madc.hi.cc.u32 %7, %10, %12, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(d.l0), "+r"(d.h0), "+r"(d.c0), "+r"(d.l1), "+r"(d.h1), "+r"(d.c1), "+r"(d.l2), "+r"(d.h2), "+r"(d.c2)
        : "r"(r0), "r"(r1), "r"(a0), "r"(a1));
}
// value mod p, canonical.  limbs j0..j4 of the 160-bit total (before carry normalisation):
//   j0 = l0, j1 = h0 + l1, j2 = c0 + h1 + l2, j3 = c1 + h2, j4 = c2
// and 2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32 (mod p)  =>  value = (j0 - j2 - j3) + 2^32 * (j1 + j2 - j4)
__device__ __forceinline__ uint64_t dot_reduce(const Dot& d) {
    int64_t j0 = d.l0, j1 = (int64_t)d.h0 + d.l1, j2 = (int64_t)d.c0 + d.h1 + d.l2, j3 = (int64_t)d.c1 + d.h2, j4 = d.c2;
    int64_t lo = j0 - j2 - j3;   // |lo| < 2^35
    int64_t hi = j1 + j2 - j4;   // |hi| < 2^35
    // lo + 2^32*hi, made non-negative by adding 2^40 * p, then reduced as a 128-bit value
    __int128 t = (__int128)lo + ((__int128)hi << 32) + ((__int128)GL_P << 40);
    return gl::reduce128((uint64_t)t, (uint64_t)((unsigned __int128)t >> 64));
}

struct DotE2 {  // sum_k r_k * alpha_k for Ext2 weights
    Dot a, b;
};
__device__ __forceinline__ void dote_zero(DotE2& d) { dot_zero(d.a); dot_zero(d.b); }
__device__ __forceinline__ void dote_add(DotE2& d, uint64_t r, const ulonglong2 w) { dot_add(d.a, r, w.x); dot_add(d.b, r, w.y); }
__device__ __forceinline__ gl::e2 dote_reduce(const DotE2& d) { return gl::make2(dot_reduce(d.a), dot_reduce(d.b)); }

}  // namespace zk
