// gl.cuh -- Goldilocks field (p = 2^64 - 2^32 + 1) and its quadratic extension F_p[u]/(u^2 - 7) on the device.
//
// Replaces, for the GPU path, `boojum::field::goldilocks::{GoldilocksField, GoldilocksExt2}` as used by the
// reference's prover call (/root/reference/src/prover_utils.rs:12,16,36-44).  All values are kept CANONICAL
// (< p) at function boundaries so results are bit-identical to the CPU oracle (oracle/gl64.h).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL
#define GL_ROOT_2_32 0x185629dcda58878cULL
#define GL_GEN 7ULL

#ifdef __CUDACC__
// 2^32 - 1 as a value ptxas cannot see through: with the literal, `mad.lo.cc x, 0xffffffff, y` is strength-reduced to a subtract
// and its `madc.hi` partner is left alone as IMAD.HI.U32 -- quarter rate on sm_100a (tools/microbench/pipes.cu: 29 vs 61
// thread-ops/clk/SM); with a register operand the pair fuses into ONE IMAD.WIDE.U32 with carry-out.
static __constant__ uint32_t gl_eps_opaque_c = 0xFFFFFFFFu;
#define GL_EPS_OPAQUE gl_eps_opaque_c
#define GL_HD __host__ __device__ __forceinline__
#define GL_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define GL_HD inline
#define GL_HD_NOINLINE static inline
#endif

namespace gl {

// x mod p for any u64 x.  Device: x + (2^32 - 1) carries out exactly when x >= p and then IS x - p (mod 2^64); the carry
// predicate of the add drives the two selects directly -- IADD3, IADD3.X, SEL, SEL: 4 alu-pipe instructions against the 6 of the
// compare form (2 ISETP + 2 IADD3 + 2 SEL).  (The variant that applies the carry with an IMAD.WIDE -- 1 instruction fewer, but on
// the fma pipe -- measured 1 % SLOWER on the whole proof: the multiply pipe is the contended one.)
// Measured per kernel (ncu time of one MainVM proof, profiles/r02_*): quotient_perm -2.8 %, quotient_gates -2 %, NTT pass B -3 %,
// but NTT pass A +6 % (it sits on its 128-register limit and spills more), so ntt1024.cu keeps the compare form (ZK_CANON_SEL 0).
#ifndef ZK_CANON_SEL
#define ZK_CANON_SEL 1
#endif
GL_HD uint64_t canon(uint64_t x) {
#if defined(__CUDA_ARCH__) && ZK_CANON_SEL
    uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), r0, r1;
    asm("{\n\t.reg .u32 t0,t1,k;\n\t.reg .pred q;\n\t"
        "add.cc.u32 t0, %2, 0xffffffff;\n\t"
        "addc.cc.u32 t1, %3, 0;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "setp.ne.u32 q, k, 0;\n\t"
        "selp.u32 %0, t0, %2, q;\n\t"
        "selp.u32 %1, t1, %3, q;\n\t}"
        : "=r"(r0), "=r"(r1)
        : "r"(x0), "r"(x1));
    return ((uint64_t)r1 << 32) | r0;
#else
    return x >= GL_P ? x - GL_P : x;
#endif
}

// canonical add/sub.  Device: PTX borrow chains (sub = 4 alu + 1 fma SASS instructions, add = 5 alu + 2 fma) instead of the
// 64-bit compare + select sequences the C form compiles to (9 each).
GL_HD uint64_t sub(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), r0, r1;
    asm("{\n\t.reg .u32 k;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 k, 0, 0;\n\t"        // 0 or 0xffffffff
        "sub.cc.u32 %0, %0, k;\n\t"    // on borrow: d += p  <=>  d -= 2^32 - 1 (mod 2^64)
        "subc.u32 %1, %1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)r1 << 32) | r0;
#else
    uint64_t d = a - b;
    if (a < b) d += GL_P;
    return d;
#endif
}
GL_HD uint64_t add(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), r0, r1;
    // u = a + (2^32 - 1) cannot overflow (a < p); the carry of u + b says a + b >= p, and then u + b mod 2^64 IS a + b - p
    asm("{\n\t.reg .u32 k,u0,u1;\n\t"
        "add.cc.u32 u0, %2, 0xffffffff;\n\t"
        "addc.u32 u1, %3, 0;\n\t"
        "add.cc.u32 u0, u0, %4;\n\t"
        "addc.cc.u32 u1, u1, %5;\n\t"
        "addc.u32 k, 0xffffffff, 0;\n\t"   // carry - 1: 0 when reduced, 0xffffffff when 2^32 - 1 has to come off again
        "sub.cc.u32 %0, u0, k;\n\t"
        "subc.u32 %1, u1, 0;\n\t}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((uint64_t)r1 << 32) | r0;
#else
    uint64_t s = a + b;
    // a, b < p: either the 64-bit add wrapped (true sum >= 2^64 > p) or s may be in [p, 2^64)
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
#endif
}
GL_HD uint64_t neg(uint64_t a) { return a ? GL_P - a : 0; }
GL_HD uint64_t dbl(uint64_t a) { return add(a, a); }

// 128-bit -> canonical.  2^64 = 2^32 - 1 and 2^96 = -1 (mod p)
GL_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;       // + p (mod 2^64)
    uint64_t t1 = (hl << 32) - hl;   // hl * (2^32 - 1)
    uint64_t r = t0 + t1;
    if (r < t0) r += GL_EPS;         // wrapped: - p (mod 2^64) -- cannot wrap twice
    return canon(r);
}

GL_HD uint64_t mul(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    // 64x64->128 as four IMAD.WIDE.U32 with PTX carry chains, reduction with single carry/borrow fix-ups, then one
    // conditional subtract: 12 fma-pipe + 12 alu-pipe SASS instructions (the plain C form costs 12 + 21; see glx.cuh)
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 l0,l1,h0,h1,k,c;\n\t"
        ".reg .u64 w;\n\t"
        "mul.wide.u32 w, %2, %4;\n\t"
        "mov.b64 {l0,l1}, w;\n\t"
        "mad.lo.cc.u32 l1, %2, %5, l1;\n\t"
        "madc.hi.u32 h0, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 l1, %3, %4, l1;\n\t"
        "madc.hi.cc.u32 h0, %3, %4, h0;\n\t"
        "addc.u32 h1, 0, 0;\n\t"
        "mad.lo.cc.u32 h0, %3, %5, h0;\n\t"
        "madc.hi.u32 h1, %3, %5, h1;\n\t"
        "sub.cc.u32 l0, l0, h1;\n\t"
        "subc.cc.u32 l1, l1, 0;\n\t"
        "subc.u32 k, 0, 0;\n\t"
        "sub.cc.u32 l0, l0, k;\n\t"
        "subc.u32 l1, l1, 0;\n\t"
        "mad.lo.cc.u32 l0, h0, %6, l0;\n\t"
        "madc.hi.cc.u32 l1, h0, %6, l1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, c, %6, l0;\n\t"
        "madc.hi.u32 %1, c, %6, l1;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(GL_EPS_OPAQUE));
    return canon(((uint64_t)r1 << 32) | r0);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    return reduce128((uint64_t)x, (uint64_t)(x >> 64));
#endif
}
GL_HD uint64_t sqr(uint64_t a) { return mul(a, a); }

GL_HD uint64_t pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}
// a^(2^n) * b
GL_HD uint64_t sqr_n_mul(uint64_t a, int n, uint64_t b) {
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int i = 0; i < n; i++) a = sqr(a);
    return mul(a, b);
}
// a^(p-2) by an addition chain: p - 2 = 2^64 - 2^32 - 1 = 2*((2^31 - 1)*2^32 + (2^31 - 1)) + 1 -- 63 squarings + 9 multiplications
// instead of the 64 + 63 of square-and-multiply (the exponent has 63 one bits); the loops stay rolled (a fully unrolled inversion is
// 40 KB of SASS per call site, and stage 2 / the quotient / the DEEP kernel each invert per row).  inv(0) = 0.
GL_HD_NOINLINE uint64_t inv(uint64_t a) {
    const uint64_t t2 = sqr_n_mul(a, 1, a);        // 2^2 - 1
    const uint64_t t3 = sqr_n_mul(t2, 1, a);       // 2^3 - 1
    const uint64_t t6 = sqr_n_mul(t3, 3, t3);      // 2^6 - 1
    const uint64_t t12 = sqr_n_mul(t6, 6, t6);     // 2^12 - 1
    const uint64_t t24 = sqr_n_mul(t12, 12, t12);  // 2^24 - 1
    const uint64_t t30 = sqr_n_mul(t24, 6, t6);    // 2^30 - 1
    const uint64_t t31 = sqr_n_mul(t30, 1, a);     // 2^31 - 1
    const uint64_t t63 = sqr_n_mul(t31, 32, t31);  // (2^31 - 1) * 2^32 + 2^31 - 1
    return sqr_n_mul(t63, 1, a);
}
GL_HD uint64_t omega(int log_n) {
    uint64_t w = GL_ROOT_2_32;
    for (int i = log_n; i < 32; i++) w = sqr(w);
    return w;
}

// x * 2^k for 0 <= k < 64 without a full multiply
GL_HD uint64_t mul_pow2(uint64_t x, unsigned k) {
    if (k == 0) return x;
    return reduce128(x << k, x >> (64 - k));
}

// ---- Ext2 ----
struct e2 {
    uint64_t c0, c1;
};
GL_HD e2 make2(uint64_t a, uint64_t b) { e2 r; r.c0 = a; r.c1 = b; return r; }
GL_HD e2 add(e2 a, e2 b) { return make2(add(a.c0, b.c0), add(a.c1, b.c1)); }
GL_HD e2 sub(e2 a, e2 b) { return make2(sub(a.c0, b.c0), sub(a.c1, b.c1)); }
GL_HD e2 neg(e2 a) { return make2(neg(a.c0), neg(a.c1)); }
#ifdef __CUDA_ARCH__
// 64x64 -> 128-bit product as four 32-bit limbs (PTX carry chains; ptxas fuses each mad.lo/madc.hi pair into IMAD.WIDE.U32)
__device__ __forceinline__ void mul_wide(uint64_t a, uint64_t b, uint32_t& p0, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.u32 %2, %4, %7, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, 0, 0;\n\t"
        "mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;\n\t"
        "}"
        : "=&r"(p0), "=&r"(p1), "=&r"(p2), "=&r"(p3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// s0 + s1 2^32 + s2 2^64 + s3 2^96 + s4 2^128 mod p (s4 < 2^31), canonical:  2^128 = -2^32 (mod p)
__device__ __forceinline__ uint64_t reduce160(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t s4) {
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 k,c;\n\t"
        "sub.cc.u32 %2, %2, %5;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.u32 k, 0, 0;\n\t"
        "sub.cc.u32 %2, %2, k;\n\t"
        "subc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %2, %4, %6, %2;\n\t"
        "madc.hi.cc.u32 %3, %4, %6, %3;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, c, %6, %2;\n\t"
        "madc.hi.u32 %1, c, %6, %3;\n\t"
        "}"
        : "=r"(r0), "=r"(r1), "+r"(s0), "+r"(s1)
        : "r"(s2), "r"(s3), "r"(GL_EPS_OPAQUE));
    return sub(canon(((uint64_t)r1 << 32) | r0), (uint64_t)s4 << 32);
}
#endif
GL_HD e2 mul(e2 a, e2 b) {
#ifdef __CUDA_ARCH__
    // schoolbook with the two products of each coordinate summed as 160-bit integers and reduced ONCE:
    // 43 fma-pipe + 50 alu-pipe instructions (Karatsuba over canonical mul/add/sub: 47 + 75)
    uint32_t p0, p1, p2, p3, q0, q1, q2, q3, s4;
    mul_wide(a.c0, b.c1, p0, p1, p2, p3);
    mul_wide(a.c1, b.c0, q0, q1, q2, q3);
    asm("add.cc.u32 %0,%0,%5;\n\taddc.cc.u32 %1,%1,%6;\n\taddc.cc.u32 %2,%2,%7;\n\taddc.cc.u32 %3,%3,%8;\n\taddc.u32 %4,0,0;"
        : "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3), "=r"(s4)
        : "r"(q0), "r"(q1), "r"(q2), "r"(q3));
    const uint64_t c1 = reduce160(p0, p1, p2, p3, s4);
    mul_wide(a.c0, b.c0, p0, p1, p2, p3);
    mul_wide(a.c1, b.c1, q0, q1, q2, q3);
    // 7*q = (q << 3) - q over 160 bits, then + p
    uint32_t t0 = q0 << 3, t1 = (q1 << 3) | (q0 >> 29), t2 = (q2 << 3) | (q1 >> 29), t3 = (q3 << 3) | (q2 >> 29), t4 = q3 >> 29;
    asm("sub.cc.u32 %0,%0,%5;\n\tsubc.cc.u32 %1,%1,%6;\n\tsubc.cc.u32 %2,%2,%7;\n\tsubc.cc.u32 %3,%3,%8;\n\tsubc.u32 %4,%4,0;"
        : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4)
        : "r"(q0), "r"(q1), "r"(q2), "r"(q3));
    asm("add.cc.u32 %0,%0,%5;\n\taddc.cc.u32 %1,%1,%6;\n\taddc.cc.u32 %2,%2,%7;\n\taddc.cc.u32 %3,%3,%8;\n\taddc.u32 %4,%4,0;"
        : "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3), "+r"(t4)
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3));
    return make2(reduce160(p0, p1, p2, p3, t4), c1);
#else
    uint64_t v0 = mul(a.c0, b.c0), v1 = mul(a.c1, b.c1);
    // Karatsuba for the cross term: (a0+a1)(b0+b1) - v0 - v1
    uint64_t cross = sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), v0), v1);
    uint64_t v1_7 = sub(mul_pow2(v1, 3), v1);
    return make2(add(v0, v1_7), cross);
#endif
}
GL_HD e2 mul_base(e2 a, uint64_t b) { return make2(mul(a.c0, b), mul(a.c1, b)); }
GL_HD e2 sqr(e2 a) { return mul(a, a); }
GL_HD e2 inv(e2 a) {
    uint64_t c1s = sqr(a.c1);
    uint64_t n = sub(sqr(a.c0), sub(mul_pow2(c1s, 3), c1s));
    uint64_t ni = inv(n);
    return make2(mul(a.c0, ni), mul(neg(a.c1), ni));
}
GL_HD e2 pow(e2 b, uint64_t e) {
    e2 r = make2(1, 0);
    while (e) {
        if (e & 1) r = mul(r, b);
        b = sqr(b);
        e >>= 1;
    }
    return r;
}
GL_HD bool eq(e2 a, e2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }

GL_HD uint32_t bitrev(uint32_t x, int bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace gl
