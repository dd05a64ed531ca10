/* tools/deep/deep_solve.c -- helper of tools/golden_deep.py (test tooling, not product code).
 *
 * Input (binary, u64): nq, nj, n, then for q < nq, j < nj three polynomials a, b, c in phi (n Ext2 coefficients each, ascending),
 * the coefficients of  E_{q,j}(phi, z) = a + b z + c z^2  (see golden_deep.py).  For the true positions j_q the three queries share
 * a root (phi, z).  Eliminates z with the resultant of two quadratics and finds phi as the root of
 * gcd(Res_z(E_0, E_1), Res_z(E_0, E_2)) over all nj^3 position triples.  Prints "HIT j0 j1 j2 deg phi.c0 phi.c1". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../oracle/gl64.h"

typedef struct { gl2 *c; int n; } poly;   /* n = number of coefficients */
static poly pnew(int n) { poly p; p.n = n; p.c = (gl2 *)calloc(n > 0 ? n : 1, sizeof(gl2)); return p; }
static void ptrim(poly *p) { while (p->n > 0 && p->c[p->n - 1].c0 == 0 && p->c[p->n - 1].c1 == 0) p->n--; }
static poly pmul(poly a, poly b) {
    poly r = pnew(a.n + b.n > 0 ? a.n + b.n - 1 : 0);
    for (int i = 0; i < a.n; i++) {
        if (!a.c[i].c0 && !a.c[i].c1) continue;
        for (int j = 0; j < b.n; j++) r.c[i + j] = gl2_add(r.c[i + j], gl2_mul(a.c[i], b.c[j]));
    }
    ptrim(&r);
    return r;
}
static poly psub(poly a, poly b) {
    int n = a.n > b.n ? a.n : b.n;
    poly r = pnew(n);
    for (int i = 0; i < n; i++) {
        gl2 x = i < a.n ? a.c[i] : gl2_make(0, 0), y = i < b.n ? b.c[i] : gl2_make(0, 0);
        r.c[i] = gl2_sub(x, y);
    }
    ptrim(&r);
    return r;
}
static poly pcopy(poly a) { poly r = pnew(a.n); memcpy(r.c, a.c, sizeof(gl2) * a.n); return r; }
/* gcd by Euclid; consumes copies */
static poly pgcd(poly a0, poly b0) {
    poly a = pcopy(a0), b = pcopy(b0);
    ptrim(&a); ptrim(&b);
    while (b.n > 0) {
        /* a = a mod b */
        gl2 inv = gl2_inv(b.c[b.n - 1]);
        while (a.n >= b.n) {
            gl2 f = gl2_mul(a.c[a.n - 1], inv);
            int d = a.n - b.n;
            for (int i = 0; i < b.n; i++) a.c[d + i] = gl2_sub(a.c[d + i], gl2_mul(f, b.c[i]));
            a.c[a.n - 1] = gl2_make(0, 0);
            ptrim(&a);
        }
        poly t = a; a = b; b = t;
    }
    free(b.c);
    return a;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) return 2;
    int nq = (int)hdr[0], nj = (int)hdr[1], n = (int)hdr[2];
    if (nq != 3) return 2;
    poly *A = malloc(sizeof(poly) * nq * nj), *B = malloc(sizeof(poly) * nq * nj), *C = malloc(sizeof(poly) * nq * nj);
    for (int i = 0; i < nq * nj; i++) {
        poly *dst[3] = {&A[i], &B[i], &C[i]};
        for (int k = 0; k < 3; k++) {
            *dst[k] = pnew(n);
            if (fread(dst[k]->c, sizeof(gl2), n, f) != (size_t)n) return 2;
            ptrim(dst[k]);
        }
    }
    fclose(f);
    /* resultants of E_{0,j1} with E_{q,j2}, q = 1, 2 */
    poly *R = malloc(sizeof(poly) * 2 * nj * nj);
#pragma omp parallel for collapse(3) schedule(dynamic)
    for (int q = 1; q <= 2; q++)
        for (int j1 = 0; j1 < nj; j1++)
            for (int j2 = 0; j2 < nj; j2++) {
                poly a1 = A[j1], b1 = B[j1], c1 = C[j1], a2 = A[q * nj + j2], b2 = B[q * nj + j2], c2 = C[q * nj + j2];
                poly ac = psub(pmul(a1, c2), pmul(a2, c1));
                poly ab = psub(pmul(a1, b2), pmul(a2, b1));
                poly bc = psub(pmul(b1, c2), pmul(b2, c1));
                R[((q - 1) * nj + j1) * nj + j2] = psub(pmul(ac, ac), pmul(ab, bc));
            }
    int hits = 0;
#pragma omp parallel for collapse(3) schedule(dynamic)
    for (int j1 = 0; j1 < nj; j1++)
        for (int j2 = 0; j2 < nj; j2++)
            for (int j3 = 0; j3 < nj; j3++) {
                poly g = pgcd(R[(0 * nj + j1) * nj + j2], R[(1 * nj + j1) * nj + j3]);
                if (g.n >= 2) {
#pragma omp critical
                    {
                        hits++;
                        if (g.n == 2) {
                            gl2 root = gl2_mul(gl2_neg(g.c[0]), gl2_inv(g.c[1]));
                            printf("HIT %d %d %d deg 1 %llu %llu\n", j1, j2, j3, (unsigned long long)root.c0, (unsigned long long)root.c1);
                        } else
                            printf("HIT %d %d %d deg %d\n", j1, j2, j3, g.n - 1);
                    }
                }
                free(g.c);
            }
    printf("done hits=%d\n", hits);
    return 0;
}
