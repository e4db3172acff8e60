/* fgp_kernel_desc.h — plain-C description of a friedrich `Kernel` value.
 *
 * The reference's `Kernel` trait (src/parameters/kernel.rs:22-86) is generic Rust code that is called once
 * per matrix element. A device cannot call back into Rust, so a kernel value crosses the C-ABI as a small
 * postfix program: leaves are the nine built-in kernels (kernel.rs:342-1157), operators are `KernelSum`
 * (kernel.rs:132-211) and `KernelProd` (kernel.rs:221-307). Leaf parameters are stored in `param[]` in the
 * order the leaves appear in the program, each leaf in the order of its own `get_parameters()` — which is
 * exactly the concatenation order k1-then-k2 of KernelSum/KernelProd::get_parameters (kernel.rs:180-186,
 * :276-282).
 *
 * This header only defines the wire format. It is shared by the product library (include/fgp.h) and by
 * the test oracle (oracle/), so both sides are fed the very same bytes.
 */
#ifndef FGP_KERNEL_DESC_H
#define FGP_KERNEL_DESC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGP_MAX_OPS 15
#define FGP_MAX_PARAMS 24

/* leaf tags (value = position in kernel.rs) and their parameter lists */
enum fgp_kernel_tag {
    FGP_K_LINEAR = 1,             /* kernel.rs:342  params: c                  k = x.y + c                       */
    FGP_K_POLYNOMIAL = 2,         /* kernel.rs:411  params: alpha, c, d        k = (alpha x.y + c)^d             */
    FGP_K_SQUARED_EXP = 3,        /* kernel.rs:507  params: ls, ampl           k = |ampl| exp(-r2/(2 ls^2))      */
    FGP_K_EXPONENTIAL = 4,        /* kernel.rs:612  params: ls, ampl           k = |ampl| exp(-r/(2 ls^2))       */
    FGP_K_MATERN1 = 5,            /* kernel.rs:717  params: ls, ampl           nu = 3/2                          */
    FGP_K_MATERN2 = 6,            /* kernel.rs:824  params: ls, ampl           nu = 5/2                          */
    FGP_K_HYPERTAN = 7,           /* kernel.rs:934  params: alpha, c           k = tanh(alpha x.y + c)           */
    FGP_K_MULTIQUADRIC = 8,       /* kernel.rs:1010 params: c                  k = hypot(r2, c)  (as coded)      */
    FGP_K_RATIONAL_QUADRATIC = 9, /* kernel.rs:1079 params: alpha, ls          k = (1 + r2/(2 alpha ls^2))^-alpha*/
    FGP_K_SUM = 100,              /* kernel.rs:132  pops two, pushes k1 + k2                                      */
    FGP_K_PROD = 101              /* kernel.rs:221  pops two, pushes k1 * k2                                      */
};

typedef struct fgp_kernel_desc {
    int32_t n_ops;               /* 1 .. FGP_MAX_OPS */
    int32_t op[FGP_MAX_OPS];     /* postfix program of fgp_kernel_tag */
    double param[FGP_MAX_PARAMS];
} fgp_kernel_desc;

/* number of parameters a leaf consumes (0 for SUM/PROD, -1 for an unknown tag) */
static inline int fgp_leaf_nparams(int tag) {
    switch (tag) {
        case FGP_K_LINEAR: return 1;
        case FGP_K_POLYNOMIAL: return 3;
        case FGP_K_SQUARED_EXP: case FGP_K_EXPONENTIAL: case FGP_K_MATERN1: case FGP_K_MATERN2:
        case FGP_K_HYPERTAN: case FGP_K_RATIONAL_QUADRATIC: return 2;
        case FGP_K_MULTIQUADRIC: return 1; /* get_parameters returns 1 value (kernel.rs:1061-1064) */
        case FGP_K_SUM: case FGP_K_PROD: return 0;
        default: return -1;
    }
}

#ifdef __cplusplus
}
#endif
#endif
