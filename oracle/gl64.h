/* oracle/gl64.h -- TEST INFRASTRUCTURE ONLY (CPU restatement, never shipped, never on the product path).
 *
 * Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[u]/(u^2 - 7).
 *
 * The reference (matter-labs/era-zkevm_test_harness) holds no field arithmetic of its own: it uses
 * `boojum::field::goldilocks::{GoldilocksField, GoldilocksExt2}` (un-vendored git dependency, branch `main`, no
 * lockfile; imported at /root/reference/src/prover_utils.rs:12,16 and circuit_definitions/src/lib.rs:70).
 * The conventions restated here are the ones CONFIRMED against the reference's golden proofs without going through
 * the hash (tests/test_golden_fri.py re-checks them on committed fixtures):
 *   - extension non-residue 7, element = c0 + c1*u
 *   - 2^32-th root of unity G = 0x185629dcda58878c = 7^((p-1)/2^32); omega_{2^k} = G^(2^(32-k))
 *   - multiplicative generator / LDE coset shift 7
 */
#ifndef ORACLE_GL64_H
#define ORACLE_GL64_H
#include <stdint.h>
#include <stddef.h>

typedef unsigned __int128 u128;
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_ROOT_2_32 0x185629dcda58878cULL
#define GL_GEN 7ULL

static inline uint64_t gl_canon(uint64_t x) { return x >= GL_P ? x - GL_P : x; }

static inline uint64_t gl_add(uint64_t a, uint64_t b) { /* canonical in, canonical out */
    uint64_t s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline uint64_t gl_neg(uint64_t a) { return a ? GL_P - a : 0; }

static inline uint64_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    /* 2^64 = 2^32 - 1, 2^96 = -1 (mod p) */
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS; /* borrow: add p == subtract eps mod 2^64 */
    uint64_t t1 = hi_lo * GL_EPS;
    uint64_t r = t0 + t1;
    if (r < t0) r += GL_EPS;
    return gl_canon(r);
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }
static inline uint64_t gl_sqr(uint64_t a) { return gl_mul(a, a); }
static inline uint64_t gl_pow(uint64_t b, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }
static inline uint64_t gl_omega(int log_n) { /* primitive 2^log_n-th root of unity */
    uint64_t w = GL_ROOT_2_32;
    for (int i = log_n; i < 32; i++) w = gl_sqr(w);
    return w;
}

/* ---- quadratic extension, u^2 = 7 ---- */
typedef struct { uint64_t c0, c1; } gl2;
static inline gl2 gl2_make(uint64_t a, uint64_t b) { gl2 r = {a, b}; return r; }
static inline gl2 gl2_add(gl2 a, gl2 b) { return gl2_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline gl2 gl2_sub(gl2 a, gl2 b) { return gl2_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline gl2 gl2_neg(gl2 a) { return gl2_make(gl_neg(a.c0), gl_neg(a.c1)); }
static inline gl2 gl2_mul(gl2 a, gl2 b) {
    uint64_t v0 = gl_mul(a.c0, b.c0), v1 = gl_mul(a.c1, b.c1);
    uint64_t c0 = gl_add(v0, gl_mul(7, v1));
    uint64_t c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return gl2_make(c0, c1);
}
static inline gl2 gl2_mul_base(gl2 a, uint64_t b) { return gl2_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
static inline gl2 gl2_sqr(gl2 a) { return gl2_mul(a, a); }
static inline gl2 gl2_inv(gl2 a) { /* 1/(c0 + c1 u) = (c0 - c1 u)/(c0^2 - 7 c1^2) */
    uint64_t n = gl_sub(gl_sqr(a.c0), gl_mul(7, gl_sqr(a.c1)));
    uint64_t ni = gl_inv(n);
    return gl2_make(gl_mul(a.c0, ni), gl_mul(gl_neg(a.c1), ni));
}
static inline gl2 gl2_pow(gl2 b, uint64_t e) {
    gl2 r = {1, 0};
    while (e) {
        if (e & 1) r = gl2_mul(r, b);
        b = gl2_sqr(b);
        e >>= 1;
    }
    return r;
}
static inline int gl2_eq(gl2 a, gl2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }

static inline uint32_t bitrev32(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

/* Copy-permutation non-residues k_0 .. k_{n-1}: column i of the permutation argument lives on the coset k_i * H.  boojum
 * (`non_residues_for_copy_permutation` -> `make_non_residues`, the bellman routine): k_0 = 1, then the successive smallest
 * quadratic non-residues whose cosets k * H are new -- 7, 11, 13, 14, 19, 21, 22, ... in Goldilocks for every domain size used
 * here [recalled; not observable on the golden proofs until the quotient identity is pinned, DESIGN.md section 5]. */
static inline void gl_copy_permutation_non_residues(uint64_t *k, uint32_t n, int log_n) {
    uint64_t cur = 1;
    uint32_t have = 0;
    uint64_t seen[1024];
    if (n == 0) return;
    k[have] = 1; seen[have++] = 1;
    while (have < n && have < 1024) {
        cur++;
        if (gl_pow(cur, (GL_P - 1) / 2) != GL_P - 1) continue;         /* a square */
        uint64_t t = cur;
        for (int i = 0; i < log_n; i++) t = gl_sqr(t);                  /* cur^(domain size) decides the coset */
        int dup = 0;
        for (uint32_t j = 0; j < have; j++) dup |= seen[j] == t;
        if (dup) continue;
        k[have] = cur; seen[have++] = t;
    }
}
#endif
