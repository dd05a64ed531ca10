/* tools/deep/deep_solve_pi.c -- helper of tools/golden_deep_pi.py (test tooling, not product code).
 *
 * The DEEP relation of a proof whose public-input ROW is unknown (golden proofs made with an older circuit layout than the VK in
 * the repository: the public inputs sit on another row).  With E0 = (h - everything but the public-input terms) * Q, cleared of
 * its denominators Q = (x - z)(x - z w), the relation of one query point reads
 *        (x - r) * E0(phi, z) + phi^n0 * N(phi) * Q(z) = 0,      r = w^row  (base field, unknown),  N = sum_t phi^t (w_t(x) - pi_t)
 * and is LINEAR in r: two queries eliminate it,
 *        G_1q = (x_1 - x_q) E0_1 E0_q + phi^n0 (N_1 Q_1 E0_q - N_q Q_q E0_1) = 0          (quartic in z),
 * three queries eliminate z (8x8 Sylvester resultant of G_12 and G_13, a polynomial of degree <= 8 (2 n0 + 2) in phi, computed by
 * evaluation at 8192 roots of unity and an inverse NTT), a fourth query singles phi out as the root of a gcd; a resultant with a
 * different pivot query removes the roots that belong to the pivot (E0_1 = N_1 = 0).  The position j of each query point inside its
 * FRI leaf is unknown, so all 8^4 combinations are tried.
 *
 * Input (binary u64): nq (=4), nj, n0, omega, then per (q, j): x, d[4], and the polynomials a, b, c (n0 Ext2 coefficients each,
 * ascending) of E0 = a + b z + c z^2.  Prints "HIT j1 j2 j3 j4 phi.c0 phi.c1 z.c0 z.c1 r.c0 r.c1". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../oracle/gl64.h"

#define LOGK 13
#define K (1 << LOGK)
static uint64_t W[K]; /* powers of omega_K */

static void ntt(uint64_t *a, int inverse) { /* in place, natural in / natural out */
    for (uint32_t i = 0; i < K; i++) { uint32_t j = bitrev32(i, LOGK); if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; } }
    for (int len = 2; len <= K; len <<= 1) {
        int step = K / len;
        for (int i = 0; i < K; i += len)
            for (int k = 0; k < len / 2; k++) {
                uint64_t w = W[(inverse ? (K - k * step) % K : k * step)];
                uint64_t u = a[i + k], v = gl_mul(a[i + k + len / 2], w);
                a[i + k] = gl_add(u, v); a[i + k + len / 2] = gl_sub(u, v);
            }
    }
    if (inverse) { uint64_t ninv = gl_inv(K); for (int i = 0; i < K; i++) a[i] = gl_mul(a[i], ninv); }
}
typedef struct { gl2 *v; } evals; /* K values */
static evals eval_poly(const gl2 *coef, int n) {
    uint64_t *c0 = calloc(K, 8), *c1 = calloc(K, 8);
    for (int i = 0; i < n; i++) { c0[i] = coef[i].c0; c1[i] = coef[i].c1; }
    ntt(c0, 0); ntt(c1, 0);
    evals e; e.v = malloc(sizeof(gl2) * K);
    for (int i = 0; i < K; i++) e.v[i] = gl2_make(c0[i], c1[i]);
    free(c0); free(c1);
    return e;
}
typedef struct { gl2 *c; int n; } poly;
static void ptrim(poly *p) { while (p->n > 0 && p->c[p->n - 1].c0 == 0 && p->c[p->n - 1].c1 == 0) p->n--; }
static poly interpolate(const gl2 *vals) { /* inverse NTT, then strip the power of phi dividing it */
    uint64_t *c0 = malloc(K * 8), *c1 = malloc(K * 8);
    for (int i = 0; i < K; i++) { c0[i] = vals[i].c0; c1[i] = vals[i].c1; }
    ntt(c0, 1); ntt(c1, 1);
    int lo = 0;
    while (lo < K && c0[lo] == 0 && c1[lo] == 0) lo++;
    poly p; p.n = K - lo; p.c = malloc(sizeof(gl2) * (p.n > 0 ? p.n : 1));
    for (int i = lo; i < K; i++) p.c[i - lo] = gl2_make(c0[i], c1[i]);
    ptrim(&p);
    free(c0); free(c1);
    return p;
}
static poly pgcd(poly a0, poly b0) {
    poly a, b;
    a.n = a0.n; a.c = malloc(sizeof(gl2) * (a.n > 0 ? a.n : 1)); memcpy(a.c, a0.c, sizeof(gl2) * a.n);
    b.n = b0.n; b.c = malloc(sizeof(gl2) * (b.n > 0 ? b.n : 1)); memcpy(b.c, b0.c, sizeof(gl2) * b.n);
    while (b.n > 0) {
        gl2 inv = gl2_inv(b.c[b.n - 1]);
        while (a.n >= b.n) {
            gl2 f = gl2_mul(a.c[a.n - 1], inv);
            int d = a.n - b.n;
            for (int i = 0; i < b.n - 1; i++) a.c[d + i] = gl2_sub(a.c[d + i], gl2_mul(f, b.c[i]));
            a.n--;
            ptrim(&a);
        }
        poly t = a; a = b; b = t;
    }
    free(b.c);
    return a;
}
/* 8x8 determinant over Ext2 by elimination */
static gl2 det8(gl2 m[8][8]) {
    gl2 det = gl2_make(1, 0);
    for (int c = 0; c < 8; c++) {
        int piv = -1;
        for (int r = c; r < 8; r++) if (m[r][c].c0 || m[r][c].c1) { piv = r; break; }
        if (piv < 0) return gl2_make(0, 0);
        if (piv != c) { for (int j = 0; j < 8; j++) { gl2 t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; } det = gl2_neg(det); }
        det = gl2_mul(det, m[c][c]);
        gl2 inv = gl2_inv(m[c][c]);
        for (int r = c + 1; r < 8; r++) {
            if (!(m[r][c].c0 || m[r][c].c1)) continue;
            gl2 f = gl2_mul(m[r][c], inv);
            for (int j = c; j < 8; j++) m[r][j] = gl2_sub(m[r][j], gl2_mul(f, m[c][j]));
        }
    }
    return det;
}
static gl2 sylvester44(const gl2 *f, const gl2 *g) { /* f, g: 5 coefficients each, ascending */
    gl2 m[8][8];
    memset(m, 0, sizeof m);
    for (int r = 0; r < 4; r++) for (int i = 0; i <= 4; i++) { m[r][r + (4 - i)] = f[i]; m[4 + r][r + (4 - i)] = g[i]; }
    return det8(m);
}

typedef struct { uint64_t x; uint64_t d[4]; evals a, b, c; gl2 *acoef, *bcoef, *ccoef; } qj_t;
static uint64_t OMEGA; static int N0;
/* G_1q(z) coefficients at evaluation point k (phi = omega_K^k) */
static void gcoef(const qj_t *p1, const qj_t *pq, int k, gl2 *out) {
    const uint64_t phi = W[k];
    uint64_t phin0 = gl_pow(phi, (uint64_t)N0);
    uint64_t n1 = 0, nq = 0, pw = 1;
    for (int t = 0; t < 4; t++) { n1 = gl_add(n1, gl_mul(p1->d[t], pw)); nq = gl_add(nq, gl_mul(pq->d[t], pw)); pw = gl_mul(pw, phi); }
    n1 = gl_mul(n1, phin0); nq = gl_mul(nq, phin0);
    gl2 e1[3] = {p1->a.v[k], p1->b.v[k], p1->c.v[k]}, eq[3] = {pq->a.v[k], pq->b.v[k], pq->c.v[k]};
    uint64_t q1[3] = {gl_mul(p1->x, p1->x), gl_neg(gl_mul(p1->x, gl_add(1, OMEGA))), OMEGA};
    uint64_t qq[3] = {gl_mul(pq->x, pq->x), gl_neg(gl_mul(pq->x, gl_add(1, OMEGA))), OMEGA};
    const uint64_t dx = gl_sub(p1->x, pq->x);
    for (int i = 0; i < 5; i++) out[i] = gl2_make(0, 0);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            gl2 t = gl2_mul_base(gl2_mul(e1[i], eq[j]), dx);
            t = gl2_add(t, gl2_mul_base(eq[j], gl_mul(n1, q1[i])));   /* + phi^n0 N_1 Q_1 E0_q */
            t = gl2_sub(t, gl2_mul_base(e1[j], gl_mul(nq, qq[i])));   /* - phi^n0 N_q Q_q E0_1 */
            out[i + j] = gl2_add(out[i + j], t);
        }
}
static gl2 peval(const gl2 *c, int n, gl2 x) { gl2 r = gl2_make(0, 0); for (int i = n; i-- > 0;) r = gl2_add(gl2_mul(r, x), c[i]); return r; }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4) return 2;
    const int nq = (int)hdr[0], nj = (int)hdr[1];
    N0 = (int)hdr[2]; OMEGA = hdr[3];
    if (nq != 4 || 8 * (2 * N0 + 2) >= K) { fprintf(stderr, "unsupported sizes\n"); return 2; }
    W[0] = 1; { uint64_t w = gl_omega(LOGK); for (int i = 1; i < K; i++) W[i] = gl_mul(W[i - 1], w); }
    qj_t *P = calloc(nq * nj, sizeof(qj_t));
    for (int i = 0; i < nq * nj; i++) {
        if (fread(&P[i].x, 8, 1, f) != 1 || fread(P[i].d, 8, 4, f) != 4) return 2;
        P[i].acoef = malloc(sizeof(gl2) * N0); P[i].bcoef = malloc(sizeof(gl2) * N0); P[i].ccoef = malloc(sizeof(gl2) * N0);
        if (fread(P[i].acoef, sizeof(gl2), N0, f) != (size_t)N0 || fread(P[i].bcoef, sizeof(gl2), N0, f) != (size_t)N0 ||
            fread(P[i].ccoef, sizeof(gl2), N0, f) != (size_t)N0) return 2;
    }
    fclose(f);
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nq * nj; i++) { P[i].a = eval_poly(P[i].acoef, N0); P[i].b = eval_poly(P[i].bcoef, N0); P[i].c = eval_poly(P[i].ccoef, N0); }
    /* resultants: R3[j1][j2][j3] = Res_z(G_12, G_13), R4[j1][j2][j4] = Res_z(G_12, G_14), R5[j2][j3][j4] = Res_z(G_23, G_24) (pivot 2) */
    const int nt = nj * nj * nj;
    poly *R3 = malloc(sizeof(poly) * nt), *R4 = malloc(sizeof(poly) * nt), *R5 = malloc(sizeof(poly) * nt);
#pragma omp parallel for schedule(dynamic)
    for (int t = 0; t < 3 * nt; t++) {
        const int which = t / nt, u = t % nt, ja = u / (nj * nj), jb = (u / nj) % nj, jc = u % nj;
        const qj_t *pp, *p2, *p3;
        if (which == 0) { pp = &P[0 * nj + ja]; p2 = &P[1 * nj + jb]; p3 = &P[2 * nj + jc]; }
        else if (which == 1) { pp = &P[0 * nj + ja]; p2 = &P[1 * nj + jb]; p3 = &P[3 * nj + jc]; }
        else { pp = &P[1 * nj + ja]; p2 = &P[2 * nj + jb]; p3 = &P[3 * nj + jc]; }
        gl2 *vals = malloc(sizeof(gl2) * K);
        for (int k = 0; k < K; k++) { gl2 g1[5], g2[5]; gcoef(pp, p2, k, g1); gcoef(pp, p3, k, g2); vals[k] = sylvester44(g1, g2); }
        poly r = interpolate(vals);
        free(vals);
        (which == 0 ? R3 : which == 1 ? R4 : R5)[u] = r;
    }
    fprintf(stderr, "resultants done (degrees e.g. %d %d %d)\n", R3[0].n - 1, R4[0].n - 1, R5[0].n - 1);
    int hits = 0;
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int j1 = 0; j1 < nj; j1++)
        for (int j2 = 0; j2 < nj; j2++)
            for (int j3 = 0; j3 < nj; j3++)
                for (int j4 = 0; j4 < nj; j4++) {
                    poly g = pgcd(R3[(j1 * nj + j2) * nj + j3], R4[(j1 * nj + j2) * nj + j4]);
                    if (g.n >= 2) {
                        poly g2 = pgcd(g, R5[(j2 * nj + j3) * nj + j4]);
                        if (g2.n == 2) {
                            gl2 phi = gl2_mul(gl2_neg(g2.c[0]), gl2_inv(g2.c[1]));
                            /* z: common root of G_12(phi, .) and G_13(phi, .) */
                            const qj_t *p1 = &P[0 * nj + j1], *p2 = &P[1 * nj + j2], *p3 = &P[2 * nj + j3];
                            gl2 G[2][5];
                            for (int s = 0; s < 2; s++) {
                                const qj_t *pq = s ? p3 : p2;
                                gl2 e1[3] = {peval(p1->acoef, N0, phi), peval(p1->bcoef, N0, phi), peval(p1->ccoef, N0, phi)};
                                gl2 eq[3] = {peval(pq->acoef, N0, phi), peval(pq->bcoef, N0, phi), peval(pq->ccoef, N0, phi)};
                                gl2 phin0 = gl2_pow(phi, (uint64_t)N0), n1 = gl2_make(0, 0), nqv = gl2_make(0, 0), pw = gl2_make(1, 0);
                                for (int t = 0; t < 4; t++) { n1 = gl2_add(n1, gl2_mul_base(pw, p1->d[t])); nqv = gl2_add(nqv, gl2_mul_base(pw, pq->d[t])); pw = gl2_mul(pw, phi); }
                                n1 = gl2_mul(n1, phin0); nqv = gl2_mul(nqv, phin0);
                                uint64_t q1[3] = {gl_mul(p1->x, p1->x), gl_neg(gl_mul(p1->x, gl_add(1, OMEGA))), OMEGA};
                                uint64_t qq[3] = {gl_mul(pq->x, pq->x), gl_neg(gl_mul(pq->x, gl_add(1, OMEGA))), OMEGA};
                                const uint64_t dx = gl_sub(p1->x, pq->x);
                                for (int i = 0; i < 5; i++) G[s][i] = gl2_make(0, 0);
                                for (int i = 0; i < 3; i++)
                                    for (int j = 0; j < 3; j++) {
                                        gl2 t = gl2_mul_base(gl2_mul(e1[i], eq[j]), dx);
                                        t = gl2_add(t, gl2_mul(eq[j], gl2_mul_base(n1, q1[i])));
                                        t = gl2_sub(t, gl2_mul(e1[j], gl2_mul_base(nqv, qq[i])));
                                        G[s][i + j] = gl2_add(G[s][i + j], t);
                                    }
                            }
                            poly ga = {G[0], 5}, gb = {G[1], 5};
                            ptrim(&ga); ptrim(&gb);
                            poly gz = pgcd(ga, gb);
                            if (gz.n == 2) {
                                gl2 z = gl2_mul(gl2_neg(gz.c[0]), gl2_inv(gz.c[1]));
                                /* r = x_1 + phi^n0 N_1 Q_1 / E0_1 */
                                gl2 e0 = gl2_add(peval(p1->acoef, N0, phi), gl2_mul(z, gl2_add(peval(p1->bcoef, N0, phi), gl2_mul(z, peval(p1->ccoef, N0, phi)))));
                                gl2 n1 = gl2_make(0, 0), pw = gl2_make(1, 0);
                                for (int t = 0; t < 4; t++) { n1 = gl2_add(n1, gl2_mul_base(pw, p1->d[t])); pw = gl2_mul(pw, phi); }
                                n1 = gl2_mul(n1, gl2_pow(phi, (uint64_t)N0));
                                gl2 xe = gl2_make(p1->x, 0);
                                gl2 Q1 = gl2_mul(gl2_sub(xe, z), gl2_sub(xe, gl2_mul_base(z, OMEGA)));
                                gl2 r = gl2_add(xe, gl2_mul(gl2_mul(n1, Q1), gl2_inv(e0)));
#pragma omp critical
                                {
                                    hits++;
                                    printf("HIT %d %d %d %d %llu %llu %llu %llu %llu %llu\n", j1, j2, j3, j4, (unsigned long long)phi.c0,
                                           (unsigned long long)phi.c1, (unsigned long long)z.c0, (unsigned long long)z.c1, (unsigned long long)r.c0,
                                           (unsigned long long)r.c1);
                                    fflush(stdout);
                                }
                            }
                            free(gz.c);
                        }
                        free(g2.c);
                    }
                    free(g.c);
                }
    printf("done hits=%d\n", hits);
    return 0;
}
