// kernel_eval.cuh — device evaluation of a friedrich `Kernel` value (and its gradient) from the pair statistics
// (dot = x.y, d2 = ||x-y||^2).  Formulas follow src/parameters/kernel.rs AS CODED, quirks included:
// signed ls in SquaredExp/Exponential exponents squared, |ampl|, signum(ampl) in gradients, Exponential's
// -r/(2 ls^2), Matern2's non-analytic ls gradient with signed x, Multiquadric = hypot(r^2, c).
#pragma once

#include <math.h>
#include <string.h>

#include "../../include/fgp_kernel_desc.h"
#include "common.cuh"
#include "exp_table.cuh"

namespace fgp {

// by-value copy of fgp_kernel_desc passed as a kernel parameter (256 bytes)
struct DevKernel {
    int32_t n_ops;
    int32_t op[FGP_MAX_OPS];
    double param[FGP_MAX_PARAMS];
};

enum KernelKind { KIND_GENERIC = 0, KIND_SQEXP = 3, KIND_MATERN2 = 6 };

// device twin of fgp_leaf_nparams (include/fgp_kernel_desc.h)
__host__ __device__ inline int leaf_nparams(int tag) {
    switch (tag) {
        case FGP_K_LINEAR: case FGP_K_MULTIQUADRIC: return 1;
        case FGP_K_POLYNOMIAL: return 3;
        case FGP_K_SQUARED_EXP: case FGP_K_EXPONENTIAL: case FGP_K_MATERN1: case FGP_K_MATERN2:
        case FGP_K_HYPERTAN: case FGP_K_RATIONAL_QUADRATIC: return 2;
        case FGP_K_SUM: case FGP_K_PROD: return 0;
        default: return -1;
    }
}

// exp(x) for x <= 0 without branches (the library routine's special-case branch would keep the 32 independent kernel
// evaluations of a pair-tile thread from being interleaved): round-to-nearest range reduction x = n ln2 + r, |r| <= 0.347,
// degree-13 Taylor polynomial (truncation 4e-18), exponent insertion.  <= 1.5 ulp on [-708, 0] (tests/test_host_logic.py
// through fgp_dbg_exp); below -708 (results under 3.3e-308, where exponent insertion would leave the normal range) it
// returns 0, NaN stays NaN.
__host__ __device__ __forceinline__ double exp_nonpos(double x) {
    const double xc = (x < -708.0) ? -708.0 : x;
    const double t = fma(xc, 1.4426950408889634074, 6755399441055744.0);  // 1.5 * 2^52: the low bits of t hold n
    const double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, xc);
    r = fma(n, -1.90821492927058770002e-10, r);
    double q = 1.6059043836821613e-10;            // 1/13!
    q = fma(q, r, 2.08767569878681e-09);          // 1/12!
    q = fma(q, r, 2.505210838544172e-08);         // 1/11!
    q = fma(q, r, 2.755731922398589e-07);         // 1/10!
    q = fma(q, r, 2.7557319223985893e-06);        // 1/9!
    q = fma(q, r, 2.48015873015873e-05);          // 1/8!
    q = fma(q, r, 1.984126984126984e-04);         // 1/7!
    q = fma(q, r, 1.388888888888889e-03);         // 1/6!
    q = fma(q, r, 8.333333333333333e-03);         // 1/5!
    q = fma(q, r, 4.1666666666666664e-02);        // 1/4!
    q = fma(q, r, 1.6666666666666666e-01);        // 1/3!
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    const long long shift = (long long)(int)n << 52;
#ifdef __CUDA_ARCH__
    const double v = __longlong_as_double(__double_as_longlong(q) + shift);
#else
    long long bits;
    memcpy(&bits, &q, 8);
    bits += shift;
    double v;
    memcpy(&v, &bits, 8);
#endif
    return (x < -708.0) ? 0.0 : v;
}

// Table-assisted exp for the interior-tile fast path of pair_tile_kernel, where the fp64 pipe is the limit (the 13-term
// polynomial above is half of a pair's fp64 instructions): x = (256 e + j) ln2/256 + r with |r| <= ln2/512, so
// exp(x) = T + T p(r), T = 2^e 2^(j/256), p = exp(r) - 1 by a degree-4 Taylor polynomial (truncation 3.8e-17) and a 256-entry
// table (tools/gen_exp_table.py, correctly rounded).  `tab` holds s * 2^(j/256) for a caller-chosen scale s (the kernel
// amplitude, folded in when the CTA copies the table into shared memory; 2^-150 <= s <= 2^150 so that the exponent
// insertion stays inside the normal range): returns s * exp(x), <= 1.5 ulp of exp on [-600, 0] (tests/test_abi.py through
// fgp_dbg_exp_tab).  The clamp and the exponent insertion are integer instructions; 9 fp64 instructions in all.
// x <= -600 gives 0 (true value < 1e-260), NaN stays NaN.
__host__ __device__ __forceinline__ void dbl_split(double v, int& hi, int& lo) {
#ifdef __CUDA_ARCH__
    hi = __double2hiint(v);
    lo = __double2loint(v);
#else
    long long b;
    memcpy(&b, &v, 8);
    hi = (int)(b >> 32);
    lo = (int)(b & 0xffffffffll);
#endif
}
__host__ __device__ __forceinline__ double dbl_join(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    const long long b = ((long long)hi << 32) | (long long)(unsigned)lo;
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}
__host__ __device__ __forceinline__ double exp_nonpos_tab(double x, const double* tab) {
    const double t = fma(x, 369.3299304675746, 6755399441055744.0);   // 256/ln2; the low word of t holds n = 256 e + j
    const double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01 * 0.00390625, x);   // ln2_hi / 256: 32 significant bits, n * it is exact
    r = fma(n, -1.90821492927058770002e-10 * 0.00390625, r);
    const double r2 = r * r;
    double q = fma(4.1666666666666664e-02, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    const double p = fma(q, r2, r);                                   // exp(r) - 1
    int thi, ni, Thi, Tlo, xhi, xlo;
    dbl_split(t, thi, ni);
    dbl_split(x, xhi, xlo);
    dbl_split(tab[ni & 255], Thi, Tlo);
    const double T = dbl_join(Thi + (int)((unsigned)(ni >> 8) << 20), Tlo);   // * 2^e, e >= -866
    const double v = fma(T, p, T);
    return ((unsigned)xhi >= 0xC082C000u) ? 0.0 : v;                  // x <= -600 (negative doubles order like their bits)
}

__device__ __forceinline__ double dsignum(double v) { return (v != v) ? v : (signbit(v) ? -1.0 : 1.0); }

__device__ __forceinline__ double leaf_value(int tag, const double* p, double dot, double d2) {
    switch (tag) {
        case FGP_K_LINEAR: return dot + p[0];                                    // kernel.rs:381
        case FGP_K_POLYNOMIAL: return pow(p[0] * dot + p[1], p[2]);              // kernel.rs:456
        case FGP_K_SQUARED_EXP: return fabs(p[1]) * exp_nonpos(-d2 / (2.0 * p[0] * p[0]));  // kernel.rs:556-560
        case FGP_K_EXPONENTIAL: return fabs(p[1]) * exp_nonpos(-sqrt(d2) / (2.0 * p[0] * p[0]));  // kernel.rs:661-665
        case FGP_K_MATERN1: {                                                    // kernel.rs:766-771
            double l = fabs(p[0]), r = sqrt(d2);
            double x = sqrt(3.0) * r / l;
            return fabs(p[1]) * (1.0 + x) * exp_nonpos(-x);
        }
        case FGP_K_MATERN2: {                                                    // kernel.rs:873-878
            double l = fabs(p[0]), r = sqrt(d2);
            double x = sqrt(5.0) * r / l;
            return fabs(p[1]) * (1.0 + x + (5.0 * r * r) / (3.0 * l * l)) * exp_nonpos(-x);
        }
        case FGP_K_HYPERTAN: return tanh(p[0] * dot + p[1]);                     // kernel.rs:976
        case FGP_K_MULTIQUADRIC: return hypot(d2, p[0]);                         // kernel.rs:1049
        case FGP_K_RATIONAL_QUADRATIC: return pow(1.0 + d2 / (2.0 * p[0] * p[1] * p[1]), -p[0]);  // kernel.rs:1121-1122
        default: return nan("");
    }
}

// gradient of a leaf in get_parameters order; returns the number of values written
__device__ __forceinline__ int leaf_grad(int tag, const double* p, double dot, double d2, double* g) {
    switch (tag) {
        case FGP_K_LINEAR: g[0] = 1.0; return 1;                                 // kernel.rs:389-390
        case FGP_K_POLYNOMIAL: {                                                 // kernel.rs:464-471
            double inner = p[0] * dot + p[1];
            double gc = p[2] * pow(inner, p[2] - 1.0);
            g[0] = dot * gc; g[1] = gc; g[2] = log(inner) * pow(inner, p[2]);
            return 3;
        }
        case FGP_K_SQUARED_EXP: {                                                // kernel.rs:569-575
            double e = exp_nonpos(-d2 / (2.0 * p[0] * p[0]));
            g[0] = (d2 * fabs(p[1]) * e) / (p[0] * p[0] * p[0]);
            g[1] = dsignum(p[1]) * e;
            return 2;
        }
        case FGP_K_EXPONENTIAL: {                                                // kernel.rs:674-680
            double r = sqrt(d2);
            double e = exp_nonpos(-r / (2.0 * p[0] * p[0]));
            g[0] = (r * fabs(p[1]) * e) / (p[0] * p[0] * p[0]);
            g[1] = dsignum(p[1]) * e;
            return 2;
        }
        case FGP_K_MATERN1: {                                                    // kernel.rs:780-787
            double l = fabs(p[0]), r = sqrt(d2);
            double x = sqrt(3.0) * r / l;
            double e = exp(-x);
            g[0] = (3.0 * fabs(p[1]) * (r * r) * e) / (p[0] * p[0] * p[0]);
            g[1] = dsignum(p[1]) * (1.0 + x) * e;
            return 2;
        }
        case FGP_K_MATERN2: {                                                    // kernel.rs:887-899 (signed x)
            double l = fabs(p[0]), r = sqrt(d2);
            double x = sqrt(5.0) * r / p[0];
            double e = exp(-x);
            g[0] = dsignum(p[0]) * fabs(p[1]) *
                   ((2.0 * l / 3.0 + 1.0) + r * sqrt(5.0) * (((l * l) / 3.0 + l + 1.0) / (l * l))) * e;
            g[1] = dsignum(p[1]) * (1.0 + x + (5.0 * r * r) / (3.0 * l * l)) * e;
            return 2;
        }
        case FGP_K_HYPERTAN: {                                                   // kernel.rs:984-988
            double c = cosh(p[0] * dot + p[1]);
            double gc = 1.0 / (c * c);
            g[0] = dot * gc; g[1] = gc;
            return 2;
        }
        case FGP_K_MULTIQUADRIC: g[0] = p[0] / hypot(sqrt(d2), p[0]); return 1;  // kernel.rs:1057
        case FGP_K_RATIONAL_QUADRATIC: {                                         // kernel.rs:1131-1144
            double alpha = p[0], l = fabs(p[1]);
            double l2 = l * l;
            double num = d2 + 2.0 * l2 * alpha;
            g[0] = pow(num / (l2 * alpha), -alpha) *
                   (pow(2.0, alpha) * (1.0 - log(num / (2.0 * l2 * alpha))) - (l2 * pow(2.0, alpha + 1.0) * alpha) / num);
            g[1] = d2 * pow(d2 / (2.0 * alpha * l * l) + 1.0, -alpha - 1.0) / (p[1] * p[1] * p[1]);
            return 2;
        }
        default: return 0;
    }
}

// value of the whole postfix program
template <int KIND>
__device__ __forceinline__ double kernel_value(const DevKernel& k, double dot, double d2) {
    if (KIND == KIND_SQEXP) return leaf_value(FGP_K_SQUARED_EXP, k.param, dot, d2);
    if (KIND == KIND_MATERN2) return leaf_value(FGP_K_MATERN2, k.param, dot, d2);
    double st[FGP_MAX_OPS];
    int sp = 0, po = 0;
    for (int i = 0; i < k.n_ops; ++i) {
        int tag = k.op[i];
        if (tag == FGP_K_SUM) { st[sp - 2] = st[sp - 2] + st[sp - 1]; --sp; }       // kernel.rs:160
        else if (tag == FGP_K_PROD) { st[sp - 2] = st[sp - 2] * st[sp - 1]; --sp; }  // kernel.rs:249
        else { st[sp++] = leaf_value(tag, k.param + po, dot, d2); po += leaf_nparams(tag); }
    }
    return st[0];
}

// value + gradient wrt all parameters (parameter order), g must hold >= nb_parameters values; returns P
__device__ __forceinline__ int kernel_value_grad(const DevKernel& k, double dot, double d2, double* value, double* g) {
    double val[FGP_MAX_OPS];
    int gs[FGP_MAX_OPS], gc[FGP_MAX_OPS];
    int sp = 0, po = 0;
    for (int i = 0; i < k.n_ops; ++i) {
        int tag = k.op[i];
        if (tag == FGP_K_SUM) {                       // kernel.rs:163-172
            val[sp - 2] = val[sp - 2] + val[sp - 1];
            gc[sp - 2] += gc[sp - 1];
            --sp;
        } else if (tag == FGP_K_PROD) {               // kernel.rs:252-262
            double k1 = val[sp - 2], k2 = val[sp - 1];
            for (int t = 0; t < gc[sp - 2]; ++t) g[gs[sp - 2] + t] *= k2;
            for (int t = 0; t < gc[sp - 1]; ++t) g[gs[sp - 1] + t] *= k1;
            val[sp - 2] = k1 * k2;
            gc[sp - 2] += gc[sp - 1];
            --sp;
        } else {
            val[sp] = leaf_value(tag, k.param + po, dot, d2);
            gs[sp] = po;
            gc[sp] = leaf_grad(tag, k.param + po, dot, d2, g + po);
            po += leaf_nparams(tag);
            ++sp;
        }
    }
    *value = val[0];
    return po;
}

// host-side classification of a descriptor
struct KernelTraits {
    int kind;       // KernelKind fast path
    bool need_d2;   // some leaf is a function of ||x-y||
    bool need_dot;  // some leaf is a function of x.y
    int nparams;
    bool valid;
};

inline KernelTraits classify(const fgp_kernel_desc* d) {
    KernelTraits t{KIND_GENERIC, false, false, 0, true};
    if (!d || d->n_ops < 1 || d->n_ops > FGP_MAX_OPS) { t.valid = false; return t; }
    int depth = 0;
    for (int i = 0; i < d->n_ops; ++i) {
        int tag = d->op[i];
        int np = fgp_leaf_nparams(tag);
        if (np < 0) { t.valid = false; return t; }
        if (tag == FGP_K_SUM || tag == FGP_K_PROD) {
            if (depth < 2) { t.valid = false; return t; }
            --depth;
        } else {
            ++depth;
            t.nparams += np;
            if (tag == FGP_K_LINEAR || tag == FGP_K_POLYNOMIAL || tag == FGP_K_HYPERTAN) t.need_dot = true;
            else t.need_d2 = true;
        }
    }
    if (depth != 1 || t.nparams > FGP_MAX_PARAMS) { t.valid = false; return t; }
    if (d->n_ops == 1 && d->op[0] == FGP_K_SQUARED_EXP) t.kind = KIND_SQEXP;
    if (d->n_ops == 1 && d->op[0] == FGP_K_MATERN2) t.kind = KIND_MATERN2;
    return t;
}

inline DevKernel to_dev(const fgp_kernel_desc* d) {
    DevKernel k{};
    k.n_ops = d->n_ops;
    for (int i = 0; i < FGP_MAX_OPS; ++i) k.op[i] = d->op[i];
    for (int i = 0; i < FGP_MAX_PARAMS; ++i) k.param[i] = d->param[i];
    return k;
}

}  // namespace fgp
