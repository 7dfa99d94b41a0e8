/* lml_ld.c -- TEST INFRASTRUCTURE ONLY (see oracle/gp_oracle.py).
 *
 * Extended-precision (x87 80-bit long double, eps = 1.08e-19) restatement of the reference's
 * log-marginal likelihood and gradient, used as the "truth" that decides whether the CUDA path is as
 * close to the exact value as the reference's own FP64 LAPACK path is.  It restates
 *   sklearn/gaussian_process/_gpr.py:583-651  (K, Cholesky, alpha, LML, K^-1, gradient traces) with
 *   sklearn/gaussian_process/kernels.py:1559-1578, 1279-1292, 1407-1414 (RBF on X/ell, constant, white)
 * -- the code GP-BayesOpInf's codebase/gpkernels.py:330-348 delegates to -- with every operation
 * (including exp) carried out in long double on the FP64 inputs t, y, theta.
 *
 * Usage: lml_ld <in.bin> <out.bin>
 *   in : int64 m; double t[m]; double y[m]; double theta[3]
 *   out: double lml; double grad[3]; double alpha[m]; double lml_hi_lo[2]; double grad_hi_lo[6]
 *        (x_hi_lo: the long double value split as hi + lo doubles, for consumers that want > 53 bits)
 * OpenMP-parallel; memory 2 * m^2 * 16 bytes.
 */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#else /* built without OpenMP (a compiler lacking libgomp): serial, same results */
static inline int omp_get_num_threads(void) { return 1; }
static inline int omp_get_thread_num(void) { return 0; }
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef long double ld;

static inline ld dotl(const ld* a, const ld* b, long n) {
    ld s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    long k = 0;
    for (; k + 3 < n; k += 4) {
        s0 += a[k] * b[k];
        s1 += a[k + 1] * b[k + 1];
        s2 += a[k + 2] * b[k + 2];
        s3 += a[k + 3] * b[k + 3];
    }
    for (; k < n; ++k) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open input"); return 2; }
    int64_t m64;
    if (fread(&m64, 8, 1, f) != 1) return 2;
    const long m = (long)m64;
    double* t = malloc(8 * m); double* y = malloc(8 * m); double th[3];
    if (fread(t, 8, m, f) != (size_t)m || fread(y, 8, m, f) != (size_t)m || fread(th, 8, 3, f) != 3) return 2;
    fclose(f);

    const ld sig2 = expl((ld)th[0]), ell = expl((ld)th[1]), chi = expl((ld)th[2]);
    ld* x = malloc(sizeof(ld) * m);
    for (long i = 0; i < m; ++i) x[i] = (ld)t[i] / ell;                      /* kernels.py:1559 */
    ld* L = malloc(sizeof(ld) * m * m);   /* lower: K then L, row-major */
    ld* U = malloc(sizeof(ld) * m * m);   /* upper: U[j][k] = W[k][j], W = L^-1 */
    if (!L || !U) { fprintf(stderr, "out of memory\n"); return 3; }

    /* K = sigma^2 R + chi I, R = exp(-d/2) with unit diagonal  (kernels.py:1559-1565, 1279-1292, 1407-1414) */
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < m; ++i) {
        for (long j = 0; j < i; ++j) {
            const ld d = x[i] - x[j];
            L[i * m + j] = sig2 * expl(-0.5L * d * d);
        }
        L[i * m + i] = sig2 + chi;
    }

    /* Cholesky K = L L^T (_gpr.py:589-593), left-looking by row blocks; not PD -> lml = -inf, grad = 0 */
    const long NB = 64;
    int bad = 0;
    for (long i0 = 0; i0 < m && !bad; i0 += NB) {
        const long i1 = i0 + NB < m ? i0 + NB : m;
        /* columns j < i0 of the rows of this block: rows are independent; each thread sweeps j once for all of
         * its rows so a factor row L[j][:j] is read once per thread, not once per row */
#pragma omp parallel
        {
            const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
            for (long j = 0; j < i0; ++j) {
                const ld* Lj = L + j * m;
                const ld inv = 1.0L / Lj[j];
                for (long i = i0 + tid; i < i1; i += nt) {
                    ld* Li = L + i * m;
                    Li[j] = (Li[j] - dotl(Li, Lj, j)) * inv;
                }
            }
        }
        /* the diagonal block */
        for (long i = i0; i < i1; ++i) {
            ld* Li = L + i * m;
            for (long j = i0; j < i; ++j) Li[j] = (Li[j] - dotl(Li, L + j * m, j)) / L[j * m + j];
            const ld d = Li[i] - dotl(Li, Li, i);
            if (!(d > 0)) { bad = 1; break; }
            Li[i] = sqrtl(d);
        }
    }

    double out_lml = -INFINITY, out_grad[3] = {0, 0, 0};
    double* alpha_d = calloc(m, 8);
    double hl[8] = {0};
    if (!bad) {
        /* W = L^-1, stored transposed: U[j][i] = W[i][j] = -(sum_{k=j}^{i-1} L[i][k] W[k][j]) / L[i][i] */
        for (long i = 0; i < m; ++i) {
            const ld inv = 1.0L / L[i * m + i];
            U[i * m + i] = inv;
#pragma omp parallel for schedule(static) if (i > 256)
            for (long j = 0; j < i; ++j) U[j * m + i] = -dotl(L + i * m + j, U + j * m + j, i - j) * inv;
        }
        /* z = W y, alpha = W^T z  (_gpr.py:601) */
        ld* yl = malloc(sizeof(ld) * m); ld* z = malloc(sizeof(ld) * m); ld* al = malloc(sizeof(ld) * m);
        for (long i = 0; i < m; ++i) yl[i] = y[i];
#pragma omp parallel for schedule(dynamic, 64)
        for (long i = 0; i < m; ++i) {      /* z_i = sum_{j<=i} W[i][j] y_j = sum_j U[j][i] y_j */
            ld s = 0;
            for (long j = 0; j <= i; ++j) s += U[j * m + i] * yl[j];
            z[i] = s;
        }
#pragma omp parallel for schedule(dynamic, 64)
        for (long j = 0; j < m; ++j) al[j] = dotl(U + j * m + j, z + j, m - j);   /* alpha_j = sum_{k>=j} W[k][j] z_k */
        /* LML (_gpr.py:613-617) */
        ld quad = dotl(yl, al, m), logdet = 0;
        for (long i = 0; i < m; ++i) logdet += logl(L[i * m + i]);
        const ld lml = -0.5L * quad - logdet - 0.5L * (ld)m * logl(2.0L * acosl(-1.0L));
        /* gradient (_gpr.py:629-651): 1/2 sum_ij (a_i a_j - Kinv_ij) dK_ij/dtheta, Kinv_ij = sum_{k>=max} U[i][k] U[j][k] */
        ld s0 = 0, s1 = 0, s2 = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : s0, s1, s2)
        for (long i = 0; i < m; ++i) {
            ld r0 = 0, r1 = 0;
            for (long j = 0; j < i; ++j) {
                const ld kinv = dotl(U + i * m + i, U + j * m + i, m - i);
                const ld w = al[i] * al[j] - kinv;
                const ld d = x[i] - x[j];
                const ld d2 = d * d;
                const ld kr = sig2 * expl(-0.5L * d2);
                r0 += w * kr;
                r1 += w * kr * d2;                                            /* kernels.py:1575-1577, 966-969 */
            }
            const ld kii = dotl(U + i * m + i, U + i * m + i, m - i);
            const ld wd = al[i] * al[i] - kii;
            s0 += 2 * r0 + wd * sig2;
            s1 += 2 * r1;
            s2 += wd;
        }
        const ld g0 = 0.5L * s0, g1 = 0.5L * s1, g2 = 0.5L * chi * s2;        /* kernels.py:1407-1414 */
        out_lml = (double)lml;
        out_grad[0] = (double)g0; out_grad[1] = (double)g1; out_grad[2] = (double)g2;
        for (long i = 0; i < m; ++i) alpha_d[i] = (double)al[i];
        hl[0] = (double)lml; hl[1] = (double)(lml - (ld)hl[0]);
        const ld gv[3] = {g0, g1, g2};
        for (int k = 0; k < 3; ++k) { hl[2 + 2 * k] = (double)gv[k]; hl[3 + 2 * k] = (double)(gv[k] - (ld)hl[2 + 2 * k]); }
    }
    f = fopen(argv[2], "wb");
    if (!f) { perror("open output"); return 2; }
    fwrite(&out_lml, 8, 1, f);
    fwrite(out_grad, 8, 3, f);
    fwrite(alpha_d, 8, m, f);
    fwrite(hl, 8, 8, f);
    fclose(f);
    return 0;
}
