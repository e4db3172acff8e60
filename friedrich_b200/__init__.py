"""friedrich_b200 — B200-native (sm_100a) Gaussian-process fit/predict engine behind the API of the Rust crate
`friedrich` (GaussianProcess / GaussianProcessBuilder / Kernel / Prior).  See DESIGN.md and INTEGRATION.md."""
from .gp import (ConstantPrior, GaussianProcess, GaussianProcessBuilder, LinearPrior, MultivariateNormal,  # noqa: F401
                 ZeroPrior)
from .kernels import (Exponential, Gaussian, HyperTan, Kernel, KernelProd, KernelSum, Linear, Matern1,  # noqa: F401
                      Matern2, Multiquadric, Polynomial, RationalQuadratic, SquaredExp)
