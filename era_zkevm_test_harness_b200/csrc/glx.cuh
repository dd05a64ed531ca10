// glx.cuh -- "lazy" Goldilocks arithmetic for the ALU-bound kernels (Poseidon2 hashing, gate evaluation).
//
// gl.cuh keeps every value canonical (< p) after every operation, which costs a 64-bit compare + select per add and per
// multiply.  The functions here work on ANY 64-bit representative of a residue (x and x + p denote the same element when
// x + p < 2^64) and only promise congruence mod p; callers canonicalise once, with glx::canon(), where a value leaves the
// kernel.  Results after canon() are therefore bit-identical to gl.cuh / the CPU oracle (oracle/gl64.h).
//
// Blackwell has no 64-bit integer multiplier: a 64x64->128 product is four IMAD.WIDE.U32 (fma pipe), and the cost that
// matters is the number of alu-pipe instructions around them (both pipes issue one warp instruction every 2 cycles per SM
// sub-partition, see B300_MICROARCH.md "Pipe rates").  mul() below is 9 fma-pipe + 8 alu-pipe SASS instructions against
// 12 + 22 for the canonical gl::mul; the PTX carry chains are what lets ptxas fuse the adds into IMAD.WIDE with carry-out.
#pragma once
#include "gl.cuh"

namespace glx {

GL_HD uint64_t canon(uint64_t x) { return gl::canon(x); }

// a * b mod p for any u64 a, b; result in [0, 2^64), not canonical.
//   a*b = lo + 2^64*hl + 2^96*hh,  2^64 = 2^32 - 1,  2^96 = -1 (mod p)
//   t = lo - hh           (on borrow: t -= 2^32 - 1, i.e. + p mod 2^64; cannot borrow twice)
//   r = t + hl*(2^32 - 1) (on carry:  r += 2^32 - 1, i.e. - p mod 2^64; cannot carry twice)
GL_HD uint64_t mul(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
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
    return ((uint64_t)r1 << 32) | r0;
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t = lo - hh;
    if (lo < hh) t -= GL_EPS;
    uint64_t m = hl * GL_EPS;
    uint64_t r = t + m;
    if (r < t) r += GL_EPS;
    return r;
#endif
}
GL_HD uint64_t sqr(uint64_t a) { return mul(a, a); }

// a * b + c mod p for any u64 a, b, c; result in [0, 2^64), not canonical.  The addend rides in the accumulator of the first wide
// product and its carry joins the 2^64 column: three instructions more than mul().
GL_HD uint64_t fma(uint64_t a, uint64_t b, uint64_t c) {
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), c0 = (uint32_t)c, c1 = (uint32_t)(c >> 32);
    uint32_t r0, r1;
    asm("{\n\t"
        ".reg .u32 l0,l1,h0,h1,k,c,cy;\n\t"
        "mad.lo.cc.u32 l0, %2, %4, %7;\n\t"
        "madc.hi.cc.u32 l1, %2, %4, %8;\n\t"
        "addc.u32 cy, 0, 0;\n\t"
        "mad.lo.cc.u32 l1, %2, %5, l1;\n\t"
        "madc.hi.u32 h0, %2, %5, 0;\n\t"
        "mad.lo.cc.u32 l1, %3, %4, l1;\n\t"
        "madc.hi.cc.u32 h0, %3, %4, h0;\n\t"
        "addc.u32 h1, 0, 0;\n\t"
        "add.cc.u32 h0, h0, cy;\n\t"     // the addend's carry enters the 2^64 column here (hi(a0*b1) + 2 carries could wrap)
        "addc.u32 h1, h1, 0;\n\t"
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
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(GL_EPS_OPAQUE), "r"(c0), "r"(c1));
    return ((uint64_t)r1 << 32) | r0;
#else
    unsigned __int128 x = (unsigned __int128)a * b + c;
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t = lo - hh;
    if (lo < hh) t -= GL_EPS;
    uint64_t m = hl * GL_EPS;
    uint64_t r = t + m;
    if (r < t) r += GL_EPS;
    return r;
#endif
}

// lo + 2^64*hi mod p for a small hi (< 2^32): one multiply-add with carry fix-up.  Not canonical.
GL_HD uint64_t reduce96(uint64_t lo, uint32_t hi) {
#ifdef __CUDA_ARCH__
    uint32_t l0 = (uint32_t)lo, l1 = (uint32_t)(lo >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 t0,t1,c;\n\t"
        "mad.lo.cc.u32 t0, %4, %5, %2;\n\t"
        "madc.hi.cc.u32 t1, %4, %5, %3;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, c, %5, t0;\n\t"
        "madc.hi.u32 %1, c, %5, t1;\n\t"
        "}"
        : "=r"(r0), "=r"(r1)
        : "r"(l0), "r"(l1), "r"(hi), "r"(GL_EPS_OPAQUE));
    return ((uint64_t)r1 << 32) | r0;
#else
    uint64_t m = (uint64_t)hi * GL_EPS;
    uint64_t r = lo + m;
    if (r < lo) r += GL_EPS;  // r_wrapped < m <= (2^32-1)^2, so this cannot carry again
    return r;
#endif
}

// x + c for any u64 x and a CANONICAL c (< p): single carry fix-up is enough (x + c < 2^64 + p).
GL_HD uint64_t add_canon(uint64_t x, uint64_t c) {
#ifdef __CUDA_ARCH__
    // add with carry-out, then + carry * (2^32 - 1) as one IMAD.WIDE: 4 SASS instructions (the C form below compiles to 8:
    // two compares and two selects)
    uint32_t x0 = (uint32_t)x, x1 = (uint32_t)(x >> 32), c0 = (uint32_t)c, c1 = (uint32_t)(c >> 32), r0, r1;
    asm("{\n\t"
        ".reg .u32 k;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 k, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, k, %6, %0;\n\t"
        "madc.hi.u32 %1, k, %6, %1;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(x0), "r"(x1), "r"(c0), "r"(c1), "r"(GL_EPS_OPAQUE));
    return ((uint64_t)r1 << 32) | r0;
#else
    uint64_t s = x + c;
    return s < x ? s + GL_EPS : s;
#endif
}
// a - b for any u64 a, b
GL_HD uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    if (a < b) {            // true value a - b + 2^64 = d; want d - (2^32-1); may borrow again only if d < 2^32-1
        uint64_t e = d - GL_EPS;
        d = d < GL_EPS ? e - GL_EPS : e;
    }
    return d;
}

// 96-bit accumulator for the small-coefficient linear layers: exact integer sums, one reduce96 at the end.
struct w96 {
    uint64_t lo;
    uint32_t hi;
};
GL_HD w96 widen(uint64_t x) { w96 r; r.lo = x; r.hi = 0; return r; }
GL_HD w96 add(w96 a, w96 b) {
    w96 r;
#ifdef __CUDA_ARCH__
    uint32_t a0 = (uint32_t)a.lo, a1 = (uint32_t)(a.lo >> 32), b0 = (uint32_t)b.lo, b1 = (uint32_t)(b.lo >> 32), r0, r1;
    asm("add.cc.u32 %0, %3, %6;\n\taddc.cc.u32 %1, %4, %7;\n\taddc.u32 %2, %5, %8;"
        : "=r"(r0), "=r"(r1), "=r"(r.hi)
        : "r"(a0), "r"(a1), "r"(a.hi), "r"(b0), "r"(b1), "r"(b.hi));
    r.lo = ((uint64_t)r1 << 32) | r0;
#else
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1u : 0u);
#endif
    return r;
}
GL_HD w96 add(w96 a, uint64_t b) { return add(a, widen(b)); }
GL_HD w96 shl(w96 a, unsigned k) {  // 0 < k < 32, result must fit 96 bits
    w96 r;
    r.hi = (a.hi << k) | (uint32_t)(a.lo >> (64 - k));
    r.lo = a.lo << k;
    return r;
}
GL_HD uint64_t reduce(w96 a) { return reduce96(a.lo, a.hi); }

// x^7
GL_HD uint64_t pow7(uint64_t x) {
    uint64_t x2 = sqr(x), x3 = mul(x2, x), x4 = sqr(x2);
    return mul(x4, x3);
}

}  // namespace glx

namespace glx {

// lo + 2^64*hi mod p for ANY 64-bit hi (the tail of mul()); result in [0, 2^64), not canonical
GL_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
#ifdef __CUDA_ARCH__
    uint32_t l0 = (uint32_t)lo, l1 = (uint32_t)(lo >> 32), h0 = (uint32_t)hi, h1 = (uint32_t)(hi >> 32), r0, r1;
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
        : "=r"(r0), "=r"(r1), "+r"(l0), "+r"(l1)
        : "r"(h0), "r"(h1), "r"(GL_EPS_OPAQUE));
    return ((uint64_t)r1 << 32) | r0;
#else
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t = lo - hh;
    if (lo < hh) t -= GL_EPS;
    uint64_t m = hl * GL_EPS;
    uint64_t r = t + m;
    if (r < t) r += GL_EPS;
    return r;
#endif
}

// x * 2^S mod p for a compile-time 0 <= S < 96 and CANONICAL x; canonical result.  2 is a 192nd root of unity in
// Goldilocks (2^96 = -1), so every twiddle of a 32-point (even 64-point) NTT is +-2^S: shifts instead of multiplies.
template <int S>
GL_HD uint64_t mul_2exp(uint64_t x) {
    static_assert(S >= 0 && S < 96, "mul_2exp: exponent out of range");
    if constexpr (S == 0) {
        return x;
    } else if constexpr (S <= 32) {
        return canon(reduce96(x << S, (uint32_t)(x >> (64 - S))));
    } else if constexpr (S < 64) {
        return canon(reduce128(x << S, x >> (64 - S)));
    } else {
        // x*2^S = V*2^64 with V = x << (S-64) = vlo + 2^64*vhi:  vlo*2^64 + vhi*2^128 = hl*(2^32-1) - hh - vhi*2^32
        constexpr int U = S - 64;
        uint64_t vlo = x << U;
        uint64_t vhi = U ? (x >> (64 - (U ? U : 1))) : 0;
        uint64_t a = (uint64_t)(uint32_t)vlo * 0xFFFFFFFFull;   // <= (2^32-1)^2 < p
        uint64_t b = (vlo >> 32) + (vhi << 32);                 // < 2^32 + 2^63 < p
        return gl::sub(a, b);
    }
}

}  // namespace glx
