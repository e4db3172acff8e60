/* friedrich_oracle.c — CPU restatement of friedrich's as-coded f64 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * *** PARITY UNPINNED ***  The reference (Rust crate friedrich v0.5.1 @ 6a3bb02 + nalgebra 0.31.4) cannot be
 * compiled or run in this image (no cargo/rustc, nalgebra source not on disk) and its own tests assert no numeric
 * value.  This file therefore restates the reference's algorithm from its source, and is pinned
 * against (a) independent known-answer anchors (numpy / 50-digit mpmath, tests/golden/) and (b) LAPACK via scipy.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file.
 * The product library (friedrich_b200/csrc) never links, calls or falls back to it.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC   (no FMA contraction: Rust never fuses a*b+c)
 *
 * Conventions: all matrices are f64 column-major with an explicit leading dimension (nalgebra DMatrix /
 * EMatrix::as_matrix() slices, src/algebra/extendable_matrix.rs:52-55). A "row" of an n x d input matrix X with
 * leading dimension ld is x[r + k*ld], k = 0..d-1.
 *
 * Third-party arithmetic (nalgebra 0.31.4, Cargo.toml:23; not vendored) restated from its published source:
 *   linalg/cholesky.rs  Cholesky::new_internal / new_with_substitute / solve_mut / inverse / insert_column
 *   linalg/solve.rs     solve_lower_triangular_vector_mut (axpy form), ad_solve (dot form)
 *   base/blas.rs        dotx (8 accumulators per column), axcpy, gemv, gemv_tr/gemm_tr, tr_dot
 *   base/norm.rs        norm_squared (sum over columns of dotc), norm = sqrt(norm_squared)
 *   base/statistics.rs  mean, variance (population, two-pass), row_variance
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/fgp_kernel_desc.h"

#define FO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------
 * nalgebra base/blas.rs `dotx` on two column vectors of length n with unit stride:
 * eight running accumulators over chunks of 8, folded as (0+4)+(1+5)+(2+6)+(3+7), then the tail sequentially. */
static double na_dot(const double *a, const double *b, int64_t n) {
    double res = 0.0;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, acc6 = 0, acc7 = 0;
    int64_t i = 0;
    while (n - i >= 8) {
        acc0 += a[i] * b[i];
        acc1 += a[i + 1] * b[i + 1];
        acc2 += a[i + 2] * b[i + 2];
        acc3 += a[i + 3] * b[i + 3];
        acc4 += a[i + 4] * b[i + 4];
        acc5 += a[i + 5] * b[i + 5];
        acc6 += a[i + 6] * b[i + 6];
        acc7 += a[i + 7] * b[i + 7];
        i += 8;
    }
    res += acc0 + acc4;
    res += acc1 + acc5;
    res += acc2 + acc6;
    res += acc3 + acc7;
    for (; i < n; ++i) res += a[i] * b[i];
    return res;
}

/* y <- a*x + y elementwise, product and sum rounded separately (blas.rs axcpy with c = b = 1). */
static void na_axpy(double a, const double *x, double *y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = a * x[i] + y[i];
}

/* (x1 - x2).norm_squared() for two ROW views of length d (strides inc1/inc2): the difference vector is a
 * 1 x d matrix, norm_squared sums dotc over its d one-element columns => plain left-to-right sum. */
static double row_dist2(const double *x1, int64_t inc1, const double *x2, int64_t inc2, int64_t d) {
    double res = 0.0;
    for (int64_t k = 0; k < d; ++k) {
        double diff = x1[k * inc1] - x2[k * inc2];
        res += diff * diff;
    }
    return res;
}

/* x1.dot(x2) for two row views: same one-element-column structure => left-to-right sum. */
static double row_dot(const double *x1, int64_t inc1, const double *x2, int64_t inc2, int64_t d) {
    double res = 0.0;
    for (int64_t k = 0; k < d; ++k) res += x1[k * inc1] * x2[k * inc2];
    return res;
}

static double signum(double v) { /* f64::signum: +-1 (also for +-0), NaN for NaN */
    if (v != v) return v;
    return signbit(v) ? -1.0 : 1.0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Leaf kernels, exactly as coded in src/parameters/kernel.rs (quirks kept). */
static double leaf_kernel(int tag, const double *p, const double *x1, int64_t i1, const double *x2, int64_t i2,
                          int64_t d) {
    switch (tag) {
        case FGP_K_LINEAR: /* kernel.rs:376-382 */
            return row_dot(x1, i1, x2, i2, d) + p[0];
        case FGP_K_POLYNOMIAL: /* kernel.rs:451-457 */
            return pow(p[0] * row_dot(x1, i1, x2, i2, d) + p[1], p[2]);
        case FGP_K_SQUARED_EXP: { /* kernel.rs:550-561 */
            double ampl = fabs(p[1]);
            double d2 = row_dist2(x1, i1, x2, i2, d);
            double x = -d2 / (2.0 * p[0] * p[0]);
            return ampl * exp(x);
        }
        case FGP_K_EXPONENTIAL: { /* kernel.rs:655-666 */
            double ampl = fabs(p[1]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double x = -dist / (2.0 * p[0] * p[0]);
            return ampl * exp(x);
        }
        case FGP_K_MATERN1: { /* kernel.rs:760-772 */
            double ampl = fabs(p[1]), l = fabs(p[0]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double x = sqrt(3.0) * dist / l;
            return ampl * (1.0 + x) * exp(-x);
        }
        case FGP_K_MATERN2: { /* kernel.rs:867-879 */
            double ampl = fabs(p[1]), l = fabs(p[0]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double x = sqrt(5.0) * dist / l;
            return ampl * (1.0 + x + (5.0 * dist * dist) / (3.0 * l * l)) * exp(-x);
        }
        case FGP_K_HYPERTAN: /* kernel.rs:971-977 */
            return tanh(p[0] * row_dot(x1, i1, x2, i2, d) + p[1]);
        case FGP_K_MULTIQUADRIC: /* kernel.rs:1044-1050: hypot(r^2, c), not hypot(r, c) */
            return hypot(row_dist2(x1, i1, x2, i2, d), p[0]);
        case FGP_K_RATIONAL_QUADRATIC: { /* kernel.rs:1116-1123 */
            double d2 = row_dist2(x1, i1, x2, i2, d);
            return pow(1.0 + d2 / (2.0 * p[0] * p[1] * p[1]), -p[0]);
        }
        default: return NAN;
    }
}

static double powi(double x, int n) { /* f64::powi for the small positive exponents used */
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}

/* gradient of a leaf, in get_parameters order; returns the number of values written */
static int leaf_gradient(int tag, const double *p, const double *x1, int64_t i1, const double *x2, int64_t i2,
                         int64_t d, double *g) {
    switch (tag) {
        case FGP_K_LINEAR: /* kernel.rs:384-391 */
            g[0] = 1.0;
            return 1;
        case FGP_K_POLYNOMIAL: { /* kernel.rs:459-472 */
            double x = row_dot(x1, i1, x2, i2, d);
            double inner = p[0] * x + p[1];
            double grad_c = p[2] * pow(inner, p[2] - 1.0);
            g[0] = x * grad_c;
            g[1] = grad_c;
            g[2] = log(inner) * pow(inner, p[2]);
            return 3;
        }
        case FGP_K_SQUARED_EXP: { /* kernel.rs:563-576 */
            double ampl = fabs(p[1]);
            double d2 = row_dist2(x1, i1, x2, i2, d);
            double e = exp(-d2 / (2.0 * p[0] * p[0]));
            g[0] = (d2 * ampl * e) / powi(p[0], 3);
            g[1] = signum(p[1]) * e;
            return 2;
        }
        case FGP_K_EXPONENTIAL: { /* kernel.rs:668-681 */
            double ampl = fabs(p[1]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double e = exp(-dist / (2.0 * p[0] * p[0]));
            g[0] = (dist * ampl * e) / powi(p[0], 3);
            g[1] = signum(p[1]) * e;
            return 2;
        }
        case FGP_K_MATERN1: { /* kernel.rs:774-788 */
            double ampl = fabs(p[1]), l = fabs(p[0]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double x = sqrt(3.0) * dist / l;
            g[0] = (3.0 * ampl * powi(dist, 2) * exp(-x)) / powi(p[0], 3);
            g[1] = signum(p[1]) * (1.0 + x) * exp(-x);
            return 2;
        }
        case FGP_K_MATERN2: { /* kernel.rs:881-900 (x uses the SIGNED ls; formula kept as coded) */
            double ampl = fabs(p[1]), l = fabs(p[0]);
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            double x = sqrt(5.0) * dist / p[0];
            g[0] = signum(p[0]) * ampl *
                   ((2.0 * l / 3.0 + 1.0) + dist * sqrt(5.0) * ((powi(l, 2) / 3.0 + l + 1.0) / powi(l, 2))) * exp(-x);
            g[1] = signum(p[1]) * (1.0 + x + (5.0 * dist * dist) / (3.0 * l * l)) * exp(-x);
            return 2;
        }
        case FGP_K_HYPERTAN: { /* kernel.rs:979-989 */
            double x = row_dot(x1, i1, x2, i2, d);
            double grad_c = 1.0 / powi(cosh(p[0] * x + p[1]), 2);
            g[0] = x * grad_c;
            g[1] = grad_c;
            return 2;
        }
        case FGP_K_MULTIQUADRIC: { /* kernel.rs:1052-1059 */
            double dist = sqrt(row_dist2(x1, i1, x2, i2, d));
            g[0] = p[0] / hypot(dist, p[0]);
            return 1;
        }
        case FGP_K_RATIONAL_QUADRATIC: { /* kernel.rs:1125-1145 */
            double alpha = p[0], l = fabs(p[1]);
            double d2 = row_dist2(x1, i1, x2, i2, d);
            double l2 = powi(l, 2);
            g[0] = pow((d2 + 2.0 * l2 * alpha) / (l2 * alpha), -alpha) *
                   (pow(2.0, alpha) * (1.0 - log((d2 + 2.0 * l2 * alpha) / (2.0 * l2 * alpha))) -
                    (l2 * pow(2.0, alpha + 1.0) * alpha) / (d2 + 2.0 * l2 * alpha));
            g[1] = d2 * pow(d2 / (2.0 * alpha * l * l) + 1.0, -alpha - 1.0) / powi(p[1], 3);
            return 2;
        }
        default: return 0;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Postfix evaluation of Sum/Prod trees (kernel.rs:155-161, :244-250, gradients :163-172, :252-262). */
FO_API int fo_desc_nb_parameters(const fgp_kernel_desc *k) {
    int n = 0;
    for (int i = 0; i < k->n_ops; ++i) {
        int c = fgp_leaf_nparams(k->op[i]);
        if (c < 0) return -1;
        n += c;
    }
    return n;
}

FO_API double fo_kernel(const fgp_kernel_desc *k, const double *x1, int64_t inc1, const double *x2, int64_t inc2,
                        int64_t d) {
    double st[FGP_MAX_OPS];
    int sp = 0, po = 0;
    for (int i = 0; i < k->n_ops; ++i) {
        int tag = k->op[i];
        if (tag == FGP_K_SUM) { st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; }
        else if (tag == FGP_K_PROD) { st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; }
        else { st[sp++] = leaf_kernel(tag, k->param + po, x1, inc1, x2, inc2, d); po += fgp_leaf_nparams(tag); }
    }
    return st[0];
}

/* gradient wrt all parameters, in parameter order; g must hold fo_desc_nb_parameters values */
FO_API int fo_kernel_gradient(const fgp_kernel_desc *k, const double *x1, int64_t inc1, const double *x2,
                              int64_t inc2, int64_t d, double *g) {
    /* stack entries: value, and the [start,count) range of g owned by the subtree */
    double val[FGP_MAX_OPS];
    int gs[FGP_MAX_OPS], gc[FGP_MAX_OPS];
    int sp = 0, po = 0;
    for (int i = 0; i < k->n_ops; ++i) {
        int tag = k->op[i];
        if (tag == FGP_K_SUM) {
            val[sp - 2] = val[sp - 2] + val[sp - 1];
            gc[sp - 2] += gc[sp - 1];
            --sp;
        } else if (tag == FGP_K_PROD) {
            double k1 = val[sp - 2], k2 = val[sp - 1];
            for (int t = 0; t < gc[sp - 2]; ++t) g[gs[sp - 2] + t] = g[gs[sp - 2] + t] * k2;
            for (int t = 0; t < gc[sp - 1]; ++t) g[gs[sp - 1] + t] = g[gs[sp - 1] + t] * k1;
            val[sp - 2] = k1 * k2;
            gc[sp - 2] += gc[sp - 1];
            --sp;
        } else {
            val[sp] = leaf_kernel(tag, k->param + po, x1, inc1, x2, inc2, d);
            gs[sp] = po;
            gc[sp] = leaf_gradient(tag, k->param + po, x1, inc1, x2, inc2, d, g + po);
            po += fgp_leaf_nparams(tag);
            ++sp;
        }
    }
    return po;
}

static int leaf_scalable(int tag) { /* kernel.rs:544, :649, :754, :861 */
    return tag == FGP_K_SQUARED_EXP || tag == FGP_K_EXPONENTIAL || tag == FGP_K_MATERN1 || tag == FGP_K_MATERN2;
}

/* subtree bookkeeping for is_scalable / rescale */
typedef struct { int first_op, last_op, first_param, scalable; } fo_node;

static int build_nodes(const fgp_kernel_desc *k, fo_node *node /* one per op */, int *lhs /* per op */) {
    int st[FGP_MAX_OPS], sp = 0, po = 0;
    for (int i = 0; i < k->n_ops; ++i) {
        int tag = k->op[i];
        if (tag == FGP_K_SUM || tag == FGP_K_PROD) {
            if (sp < 2) return -1;
            int b = st[sp - 1], a = st[sp - 2];
            node[i].first_op = node[a].first_op;
            node[i].last_op = i;
            node[i].first_param = node[a].first_param;
            node[i].scalable = (tag == FGP_K_SUM) ? (node[a].scalable && node[b].scalable)  /* kernel.rs:150-153 */
                                                  : (node[a].scalable || node[b].scalable); /* kernel.rs:239-242 */
            lhs[i] = a;
            sp -= 2;
            st[sp++] = i;
        } else {
            if (fgp_leaf_nparams(tag) < 0) return -1;
            node[i].first_op = node[i].last_op = i;
            node[i].first_param = po;
            node[i].scalable = leaf_scalable(tag);
            lhs[i] = -1;
            po += fgp_leaf_nparams(tag);
            st[sp++] = i;
        }
    }
    return sp == 1 ? st[0] : -1;
}

FO_API int fo_desc_is_scalable(const fgp_kernel_desc *k) {
    fo_node node[FGP_MAX_OPS];
    int lhs[FGP_MAX_OPS];
    int root = build_nodes(k, node, lhs);
    return root < 0 ? -1 : node[root].scalable;
}

static void rescale_node(fgp_kernel_desc *k, const fo_node *node, const int *lhs, int i, double scale) {
    int tag = k->op[i];
    if (tag == FGP_K_SUM) { /* kernel.rs:174-178 */
        rescale_node(k, node, lhs, lhs[i], scale);
        rescale_node(k, node, lhs, i - 1, scale);
    } else if (tag == FGP_K_PROD) { /* kernel.rs:264-274 */
        if (node[lhs[i]].scalable) rescale_node(k, node, lhs, lhs[i], scale);
        else rescale_node(k, node, lhs, i - 1, scale);
    } else if (leaf_scalable(tag)) {
        k->param[node[i].first_param + 1] *= scale; /* ampl *= scale (kernel.rs:578-581 etc.) */
    }
    /* a non-scalable leaf panics in the reference (kernel.rs:43-54); callers check is_scalable first */
}

FO_API int fo_desc_rescale(fgp_kernel_desc *k, double scale) {
    fo_node node[FGP_MAX_OPS];
    int lhs[FGP_MAX_OPS];
    int root = build_nodes(k, node, lhs);
    if (root < 0) return -1;
    rescale_node(k, node, lhs, root, scale);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * src/algebra/mod.rs */

/* make_covariance_matrix (algebra/mod.rs:41-54): out[r,c] = kernel(m1.row(r), m2.row(c)); out is n1 x n2. */
FO_API void fo_make_covariance_matrix(const fgp_kernel_desc *k, const double *m1, int64_t ld1, int64_t n1,
                                      const double *m2, int64_t ld2, int64_t n2, int64_t d, double *out,
                                      int64_t ldo) {
    for (int64_t c = 0; c < n2; ++c)
        for (int64_t r = 0; r < n1; ++r) out[r + c * ldo] = fo_kernel(k, m1 + r, ld1, m2 + c, ld2, d);
}

/* Gram matrix part of make_cholesky_cov_matrix (algebra/mod.rs:67-79): lower triangle, NaN above, noise^2 on the
 * diagonal added AFTER the kernel value; kernel(x_col, x_row) argument order. */
FO_API void fo_gram_lower(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n, int64_t d,
                          double noise, double *K, int64_t ldk) {
    for (int64_t c = 0; c < n; ++c) {
        for (int64_t r = 0; r < c; ++r) K[r + c * ldk] = NAN;
        for (int64_t r = c; r < n; ++r) K[r + c * ldk] = fo_kernel(k, X + c, ldx, X + r, ldx, d);
        K[c + c * ldk] += noise * noise;
    }
}

/* nalgebra Cholesky::new_internal (linalg/cholesky.rs): unblocked left-looking column algorithm on the lower
 * triangle. Returns 0 on success, j+1 when column j's pivot is zero/negative/NaN (and the substitute, if any,
 * is too). The upper triangle is never touched. */
FO_API int64_t fo_cholesky_inplace(double *A, int64_t lda, int64_t n, int has_substitute, double substitute) {
    for (int64_t j = 0; j < n; ++j) {
        for (int64_t k = 0; k < j; ++k) {
            double factor = -A[j + k * lda];
            na_axpy(factor, A + j + k * lda, A + j + j * lda, n - j);
        }
        double diag = A[j + j * lda];
        double denom;
        int ok = 0;
        if (diag != 0.0 && diag >= 0.0) { denom = sqrt(diag); ok = 1; } /* is_zero / try_sqrt (NaN fails >=) */
        else if (has_substitute && substitute != 0.0 && substitute >= 0.0) { denom = sqrt(substitute); ok = 1; }
        if (!ok) return j + 1;
        A[j + j * lda] = denom;
        for (int64_t i = j + 1; i < n; ++i) A[i + j * lda] /= denom;
    }
    return 0;
}

/* make_cholesky_cov_matrix (algebra/mod.rs:59-92) */
FO_API int64_t fo_make_cholesky_cov_matrix(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n,
                                           int64_t d, double noise, int has_eps, double eps, double *L,
                                           int64_t ldl) {
    fo_gram_lower(k, X, ldx, n, d, noise, L, ldl);
    return fo_cholesky_inplace(L, ldl, n, has_eps, eps);
}

/* solve.rs solve_lower_triangular_mut: per RHS column, column-oriented forward substitution with axpy.
 * Returns 0 when a diagonal entry is exactly zero (the reference then panics via expect). */
FO_API int fo_solve_lower(const double *L, int64_t ldl, int64_t n, double *B, int64_t ldb, int64_t q) {
    for (int64_t c = 0; c < q; ++c) {
        double *b = B + c * ldb;
        for (int64_t i = 0; i < n; ++i) {
            double diag = L[i + i * ldl];
            if (diag == 0.0) return 0;
            double coeff = b[i] / diag;
            b[i] = coeff;
            na_axpy(-coeff, L + (i + 1) + i * ldl, b + i + 1, n - i - 1);
        }
    }
    return 1;
}

/* solve.rs ad_solve_lower_triangular_unchecked_mut: L^T x = b, backward, dot-product form. */
FO_API void fo_ad_solve_lower(const double *L, int64_t ldl, int64_t n, double *B, int64_t ldb, int64_t q) {
    for (int64_t c = 0; c < q; ++c) {
        double *b = B + c * ldb;
        for (int64_t i = n - 1; i >= 0; --i) {
            double dot = na_dot(L + (i + 1) + i * ldl, b + i + 1, n - i - 1);
            b[i] = (b[i] - dot) / L[i + i * ldl];
        }
    }
}

/* Cholesky::solve_mut = forward (unchecked) then adjoint solve */
FO_API void fo_chol_solve(const double *L, int64_t ldl, int64_t n, double *B, int64_t ldb, int64_t q) {
    fo_solve_lower(L, ldl, n, B, ldb, q);
    fo_ad_solve_lower(L, ldl, n, B, ldb, q);
}

/* Cholesky::inverse = solve_mut on a dense identity; out is n x n, ld = n */
FO_API void fo_chol_inverse(const double *L, int64_t ldl, int64_t n, double *out) {
    memset(out, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (int64_t i = 0; i < n; ++i) out[i + i * n] = 1.0;
    fo_chol_solve(L, ldl, n, out, n, n);
}

/* add_rows_cholesky_cov_matrix (algebra/mod.rs:97-126) with nalgebra insert_column(j = end):
 * L has room for (n_old + n_new)^2 with leading dimension ldl; X holds all n_old + n_new rows. */
FO_API int fo_add_rows_cholesky(const fgp_kernel_desc *k, double *L, int64_t ldl, const double *X, int64_t ldx,
                                int64_t n_old, int64_t n_new, int64_t d, double noise) {
    double *col = (double *)malloc(sizeof(double) * (size_t)(n_old + n_new));
    for (int64_t i = 0; i < n_new; ++i) {
        int64_t j = n_old + i; /* col_index */
        for (int64_t t = 0; t <= j; ++t) col[t] = fo_kernel(k, X + t, ldx, X + j, ldx, d);
        col[j] += noise * noise;
        double cjj = col[j];
        if (!fo_solve_lower(L, ldl, j, col, j, 1)) { free(col); return 0; } /* assert!(solve ok) */
        for (int64_t t = 0; t < j; ++t) L[j + t * ldl] = col[t];           /* adjoint_to row j */
        L[j + j * ldl] = sqrt(cjj - na_dot(col, col, j));                  /* unchecked sqrt: NaN if negative */
        for (int64_t t = 0; t < j; ++t) L[t + j * ldl] = 0.0;              /* new matrix starts zeroed */
    }
    free(col);
    return 1;
}

/* make_gradient_covariance_matrices (algebra/mod.rs:129-155): P full symmetric n x n matrices, G[p] at
 * out + p*n*n (ld = n). */
FO_API void fo_make_gradient_covariance_matrices(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n,
                                                 int64_t d, double *out) {
    int P = fo_desc_nb_parameters(k);
    double g[FGP_MAX_PARAMS];
    for (int64_t c = 0; c < n; ++c)
        for (int64_t r = c; r < n; ++r) {
            fo_kernel_gradient(k, X + c, ldx, X + r, ldx, d, g);
            for (int p = 0; p < P; ++p) {
                out[(size_t)p * n * n + r + c * n] = g[p];
                out[(size_t)p * n * n + c + r * n] = g[p];
            }
        }
}

/* ------------------------------------------------------------------------------------------------------------
 * src/parameters/kernel.rs heuristics and nalgebra statistics */

/* fit_bandwidth_mean (kernel.rs:94-113) */
FO_API double fo_fit_bandwidth_mean(const double *X, int64_t ldx, int64_t n, int64_t d) {
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) sum += sqrt(row_dist2(X + i, ldx, X + j, ldx, d));
    double nb = (double)((n * n - n) / 2);
    return sum / nb;
}

FO_API double fo_mean(const double *y, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += y[i];
    return s / (double)n;
}

/* statistics.rs variance: population, two-pass (fit_amplitude_var kernel.rs:116-119; builder.rs:73) */
FO_API double fo_variance(const double *y, int64_t n) {
    if (n == 0) return 0.0;
    double mean = fo_mean(y, n), acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc = acc + (y[i] - mean) * (y[i] - mean);
    return acc / (double)n;
}

/* ------------------------------------------------------------------------------------------------------------
 * src/gaussian_process/mod.rs — formulas as coded. `y` is training_outputs = y_raw - prior(X). */

/* likelihood (mod.rs:196-220); returns NaN-safe value, *ok = 0 if the solve hit a zero diagonal */
FO_API double fo_likelihood(const fgp_kernel_desc *k, double noise, const double *X, int64_t ldx, int64_t n,
                            int64_t d, const double *y, const double *L, int64_t ldl, int *ok) {
    double *ol = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(ol, y, sizeof(double) * (size_t)n);
    *ok = fo_solve_lower(L, ldl, n, ol, n, 1);
    double data_fit = na_dot(ol, ol, n);
    free(ol);
    double penalty = 0.0;
    for (int64_t r = 0; r < n; ++r) penalty += log(fabs(fo_kernel(k, X + r, ldx, X + r, ldx, d) + noise * noise));
    double norm = (double)n * log(2.0 * M_PI);
    return -(data_fit + penalty + norm) / 2.0;
}

/* predict (mod.rs:226-244): mean[i] = 1*dot(weights.col(i), y) + 1*prior[i]; `mean` holds prior(Xq) on entry. */
FO_API void fo_predict(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n, int64_t d,
                       const double *y, const double *L, int64_t ldl, const double *Xq, int64_t ldq, int64_t q,
                       double *mean) {
    double *w = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    fo_make_covariance_matrix(k, X, ldx, n, Xq, ldq, q, d, w, n);
    fo_chol_solve(L, ldl, n, w, n, q);
    for (int64_t i = 0; i < q; ++i) mean[i] = 1.0 * na_dot(w + i * n, y, n) + 1.0 * mean[i];
    free(w);
}

/* predict_variance (mod.rs:248-273) */
FO_API int fo_predict_variance(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n, int64_t d,
                               const double *L, int64_t ldl, const double *Xq, int64_t ldq, int64_t q,
                               double *var) {
    double *kl = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    fo_make_covariance_matrix(k, X, ldx, n, Xq, ldq, q, d, kl, n);
    int ok = fo_solve_lower(L, ldl, n, kl, n, q);
    for (int64_t i = 0; i < q; ++i)
        var[i] = fo_kernel(k, Xq + i, ldq, Xq + i, ldq, d) - na_dot(kl + i * n, kl + i * n, n);
    free(kl);
    return ok;
}

/* predict_mean_variance (mod.rs:290-326) */
FO_API void fo_predict_mean_variance(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n, int64_t d,
                                     const double *y, const double *L, int64_t ldl, const double *Xq, int64_t ldq,
                                     int64_t q, double *mean, double *var) {
    double *c = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    double *w = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    fo_make_covariance_matrix(k, X, ldx, n, Xq, ldq, q, d, c, n);
    memcpy(w, c, sizeof(double) * (size_t)n * (size_t)q);
    fo_chol_solve(L, ldl, n, w, n, q);
    for (int64_t i = 0; i < q; ++i) {
        mean[i] = 1.0 * na_dot(w + i * n, y, n) + 1.0 * mean[i];
        var[i] = fo_kernel(k, Xq + i, ldq, Xq + i, ldq, d) - na_dot(c + i * n, w + i * n, n);
    }
    free(c);
    free(w);
}

/* predict_covariance (mod.rs:329-350) mode 0: Kqq - kl^T kl ;  sample_at (mod.rs:371-392) mode 1: Kqq - Knq^T W,
 * also fills mean (holding prior on entry) in mode 1. cov is q x q, ld = q. */
FO_API int fo_predict_covariance(const fgp_kernel_desc *k, const double *X, int64_t ldx, int64_t n, int64_t d,
                                 const double *y, const double *L, int64_t ldl, const double *Xq, int64_t ldq,
                                 int64_t q, int mode, double *cov, double *mean) {
    double *c = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    double *w = (double *)malloc(sizeof(double) * (size_t)n * (size_t)q);
    int ok = 1;
    fo_make_covariance_matrix(k, X, ldx, n, Xq, ldq, q, d, c, n);
    fo_make_covariance_matrix(k, Xq, ldq, q, Xq, ldq, q, d, cov, q);
    memcpy(w, c, sizeof(double) * (size_t)n * (size_t)q);
    if (mode == 0) {
        ok = fo_solve_lower(L, ldl, n, w, n, q);
        for (int64_t j = 0; j < q; ++j)     /* gemm_tr(-1, kl, kl, 1): per output column, gemv_tr */
            for (int64_t i = 0; i < q; ++i)
                cov[i + j * q] = -1.0 * na_dot(w + i * n, w + j * n, n) + 1.0 * cov[i + j * q];
    } else {
        fo_chol_solve(L, ldl, n, w, n, q);
        for (int64_t j = 0; j < q; ++j)     /* gemm_tr(-1, cov_train_inputs, weights, 1) */
            for (int64_t i = 0; i < q; ++i)
                cov[i + j * q] = -1.0 * na_dot(c + i * n, w + j * n, n) + 1.0 * cov[i + j * q];
        if (mean)
            for (int64_t i = 0; i < q; ++i) mean[i] = 1.0 * na_dot(w + i * n, y, n) + 1.0 * mean[i];
    }
    free(c);
    free(w);
    return ok;
}

/* ------------------------------------------------------------------------------------------------------------
 * src/gaussian_process/optimizer.rs */

/* alpha = &cov_inv * y : nalgebra gemv, column-oriented axcpy accumulation */
static void na_gemv(const double *A, int64_t n, const double *x, double *yout) {
    for (int64_t i = 0; i < n; ++i) yout[i] = 1.0 * A[i] * x[0];
    for (int64_t j = 1; j < n; ++j) {
        const double *col = A + j * n;
        double xj = x[j];
        for (int64_t i = 0; i < n; ++i) yout[i] = 1.0 * col[i] * xj + yout[i];
    }
}

/* gradient_marginal_likelihood (optimizer.rs:24-60) when scaled == 0: returns P kernel gradients + noise gradient.
 * scaled_gradient_marginal_likelihood (optimizer.rs:159-203) when scaled != 0: returns scale and P gradients. */
FO_API void fo_gradient_marginal_likelihood(const fgp_kernel_desc *k, double noise, const double *X, int64_t ldx,
                                            int64_t n, int64_t d, const double *y, const double *L, int64_t ldl,
                                            int scaled, double *scale_out, double *grads) {
    int P = fo_desc_nb_parameters(k);
    double *inv = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)n);
    double *G = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n * (size_t)(P > 0 ? P : 1));
    fo_chol_inverse(L, ldl, n, inv);
    na_gemv(inv, n, y, alpha);
    double scale = 1.0;
    if (scaled) scale = na_dot(y, alpha, n) / (double)n;
    fo_make_gradient_covariance_matrices(k, X, ldx, n, d, G);
    for (int p = 0; p < P; ++p) {
        const double *Gp = G + (size_t)p * n * n;
        double data_fit = 0.0;
        for (int64_t c = 0; c < n; ++c) data_fit += na_dot(alpha, Gp + c * n, n) * alpha[c];
        if (scaled) data_fit = data_fit / scale;
        double penalty = 0.0;
        for (int64_t i = 0; i < n; ++i) { /* cov_inv.row(i).tr_dot(G.column(i)): sequential */
            double res = 0.0;
            for (int64_t t = 0; t < n; ++t) res += inv[i + t * n] * Gp[t + i * n];
            penalty += res;
        }
        grads[p] = (data_fit - penalty) / 2.0;
    }
    if (!scaled) {
        double data_fit = na_dot(alpha, alpha, n);
        double trace = 0.0;
        for (int64_t i = 0; i < n; ++i) trace += inv[i + i * n];
        grads[P] = noise * (data_fit - trace);
    }
    if (scale_out) *scale_out = scale;
    free(inv);
    free(alpha);
    free(G);
}

/* One record per optimiser iteration for trajectory comparison. */
typedef struct {
    double scale;                       /* 1.0 in the unscaled optimiser */
    double grads[FGP_MAX_PARAMS + 1];   /* as returned by the gradient function (noise gradient corrected to log space) */
    double params[FGP_MAX_PARAMS];      /* kernel parameters after the update (+rescale) */
    double noise;                       /* noise after the update */
} fo_opt_record;

/* scaled_optimize_parameters (optimizer.rs:211-283) / optimize_parameters (optimizer.rs:69-149).
 * max_time is treated as infinite (wall-clock stop is non-deterministic). L (n x n, ld = ldl) is refit in place.
 * Returns the number of iterations executed, or -(failing column + 1) when a refit Cholesky fails. */
FO_API int64_t fo_optimize_parameters(fgp_kernel_desc *k, double *noise_io, const double *X, int64_t ldx, int64_t n,
                                      int64_t d, const double *y, double *L, int64_t ldl, int has_eps, double eps,
                                      int scaled, int64_t max_iter, double convergence_fraction,
                                      fo_opt_record *trace /* may be NULL, max_iter entries */) {
    const double beta1 = 0.9, beta2 = 0.999, epsilon = 1e-8, learning_rate = 0.1;
    int P = fo_desc_nb_parameters(k);
    int np = scaled ? P : P + 1;
    double parameters[FGP_MAX_PARAMS + 1], mean_grad[FGP_MAX_PARAMS + 1], var_grad[FGP_MAX_PARAMS + 1];
    double gradients[FGP_MAX_PARAMS + 1];
    for (int p = 0; p < P; ++p) parameters[p] = (k->param[p] == 0.0) ? epsilon : k->param[p];
    if (!scaled) parameters[P] = log(*noise_io);
    for (int p = 0; p < np; ++p) mean_grad[p] = var_grad[p] = 0.0;
    int64_t it = 0;
    for (int64_t i = 1; i <= max_iter; ++i) {
        double scale = 1.0;
        fo_gradient_marginal_likelihood(k, *noise_io, X, ldx, n, d, y, L, ldl, scaled, &scale, gradients);
        if (!scaled) gradients[P] *= *noise_io; /* optimizer.rs:106-110 */
        int progress = 0;
        double b1 = 1.0, b2 = 1.0;
        for (int64_t t = 0; t < i; ++t) { b1 *= beta1; b2 *= beta2; } /* powi(i) */
        for (int p = 0; p < np; ++p) {
            mean_grad[p] = beta1 * mean_grad[p] + (1.0 - beta1) * gradients[p];
            var_grad[p] = beta2 * var_grad[p] + (1.0 - beta2) * (gradients[p] * gradients[p]);
            double bcm = mean_grad[p] / (1.0 - b1);
            double bcv = var_grad[p] / (1.0 - b2);
            double delta = learning_rate * bcm / (sqrt(bcv) + epsilon);
            progress |= fabs(delta) > convergence_fraction;
            parameters[p] *= 1.0 + delta;
        }
        for (int p = 0; p < P; ++p) k->param[p] = parameters[p];
        if (scaled) {
            fo_desc_rescale(k, scale);
            *noise_io *= scale;
            for (int p = 0; p < P; ++p) parameters[p] = k->param[p];
        } else {
            *noise_io = exp(parameters[P]);
        }
        int64_t fail = fo_make_cholesky_cov_matrix(k, X, ldx, n, d, *noise_io, has_eps, eps, L, ldl);
        it = i;
        if (trace) {
            trace[i - 1].scale = scale;
            for (int p = 0; p < np; ++p) trace[i - 1].grads[p] = gradients[p];
            for (int p = 0; p < P; ++p) trace[i - 1].params[p] = k->param[p];
            trace[i - 1].noise = *noise_io;
        }
        if (fail) return -fail;
        if (!progress) break;
    }
    return it;
}
