"""Host-side mirror of the reference's `Kernel` trait and its implementations (src/parameters/kernel.rs).

The reference calls `Kernel::kernel` / `Kernel::gradient` once per matrix element on the CPU.  On the B200 path the
element loop runs inside CUDA kernels, so a kernel VALUE crosses the C-ABI as a postfix program
(`device_desc()` -> `fgp_kernel_desc`, include/fgp_kernel_desc.h); parameters, rescaling rules and the heuristic fit
stay on the host exactly as in the reference.  No per-element arithmetic happens in this file.
"""
from __future__ import annotations

from ._native import KernelDesc

K_LINEAR, K_POLYNOMIAL, K_SQUARED_EXP, K_EXPONENTIAL, K_MATERN1, K_MATERN2 = 1, 2, 3, 4, 5, 6
K_HYPERTAN, K_MULTIQUADRIC, K_RATIONAL_QUADRATIC, K_SUM, K_PROD = 7, 8, 9, 100, 101


class Kernel:
    """kernel.rs:22-86.  Subclasses define `_tag`, `_names` (parameter order of get_parameters)."""
    _tag = 0
    _names: tuple = ()

    # -- trait methods ------------------------------------------------------------------------------------------
    def nb_parameters(self):
        return len(self._names)

    def is_scalable(self):  # kernel.rs:33-36
        return False

    def rescale(self, scale):  # kernel.rs:43-46
        raise NotImplementedError("rescale is only defined for scalable kernels")

    def get_parameters(self):
        return [getattr(self, n) for n in self._names]

    def set_parameters(self, parameters):
        for n, v in zip(self._names, parameters):
            setattr(self, n, float(v))

    def heuristic_fit(self, bandwidth_mean, amplitude_var):  # kernel.rs:81-85 (default: nothing)
        """`bandwidth_mean()` / `amplitude_var()` are thunks: the O(n^2 d) mean pair distance runs on the device
        (fgp_mean_pair_distance) only when a kernel asks for it."""

    # -- device description -------------------------------------------------------------------------
    def _program(self):
        return [self._tag], self.get_parameters()

    def device_desc(self) -> KernelDesc:
        ops, params = self._program()
        return KernelDesc.make(ops, params)

    # -- KernelArith (kernel.rs:312-332) ------------------------------------------------------------------------
    def __add__(self, other):
        return KernelSum(self, other)

    def __mul__(self, other):
        return KernelProd(self, other)

    def __repr__(self):
        args = ", ".join(f"{n}={getattr(self, n)!r}" for n in self._names)
        return f"{type(self).__name__}({args})"


class _Combinator(Kernel):
    def __init__(self, k1: Kernel, k2: Kernel):
        self.k1, self.k2 = k1, k2

    def nb_parameters(self):
        return self.k1.nb_parameters() + self.k2.nb_parameters()

    def get_parameters(self):  # k1 then k2 (kernel.rs:180-186, :276-282)
        return self.k1.get_parameters() + self.k2.get_parameters()

    def set_parameters(self, parameters):
        n1 = self.k1.nb_parameters()
        self.k1.set_parameters(parameters[:n1])
        self.k2.set_parameters(parameters[n1:])

    def heuristic_fit(self, bandwidth_mean, amplitude_var):
        self.k1.heuristic_fit(bandwidth_mean, amplitude_var)
        self.k2.heuristic_fit(bandwidth_mean, amplitude_var)

    def _program(self):
        o1, p1 = self.k1._program()
        o2, p2 = self.k2._program()
        return o1 + o2 + [self._tag], p1 + p2

    def __repr__(self):
        return f"{type(self).__name__}({self.k1!r}, {self.k2!r})"


class KernelSum(_Combinator):
    """kernel.rs:132-211"""
    _tag = K_SUM

    def is_scalable(self):
        return self.k1.is_scalable() and self.k2.is_scalable()

    def rescale(self, scale):
        self.k1.rescale(scale)
        self.k2.rescale(scale)


class KernelProd(_Combinator):
    """kernel.rs:221-307"""
    _tag = K_PROD

    def is_scalable(self):
        return self.k1.is_scalable() or self.k2.is_scalable()

    def rescale(self, scale):
        if self.k1.is_scalable():
            self.k1.rescale(scale)
        else:
            self.k2.rescale(scale)


class Linear(Kernel):
    """k = x.y + c  (kernel.rs:342-402)"""
    _tag, _names = K_LINEAR, ("c",)

    def __init__(self, c=0.0):
        self.c = float(c)


class Polynomial(Kernel):
    """k = (alpha x.y + c)^d  (kernel.rs:411-485)"""
    _tag, _names = K_POLYNOMIAL, ("alpha", "c", "d")

    def __init__(self, alpha=1.0, c=0.0, d=1.0):
        self.alpha, self.c, self.d = float(alpha), float(c), float(d)


class _LsAmpl(Kernel):
    _names = ("ls", "ampl")

    def __init__(self, ls=1.0, ampl=1.0):
        self.ls, self.ampl = float(ls), float(ampl)

    def is_scalable(self):
        return True

    def rescale(self, scale):
        self.ampl *= scale

    def heuristic_fit(self, bandwidth_mean, amplitude_var):
        self.ls = bandwidth_mean()
        self.ampl = amplitude_var()


class SquaredExp(_LsAmpl):
    """k = |ampl| exp(-|x-y|^2 / (2 ls^2))  (kernel.rs:496-601)"""
    _tag = K_SQUARED_EXP


Gaussian = SquaredExp  # kernel.rs:496


class Exponential(_LsAmpl):
    """k = |ampl| exp(-|x-y| / (2 ls^2))  (kernel.rs:612-706)"""
    _tag = K_EXPONENTIAL


class Matern1(_LsAmpl):
    """nu = 3/2  (kernel.rs:717-813)"""
    _tag = K_MATERN1


class Matern2(_LsAmpl):
    """nu = 5/2  (kernel.rs:824-925)"""
    _tag = K_MATERN2


class HyperTan(Kernel):
    """k = tanh(alpha x.y + c)  (kernel.rs:934-1001)"""
    _tag, _names = K_HYPERTAN, ("alpha", "c")

    def __init__(self, alpha=1.0, c=0.0):
        self.alpha, self.c = float(alpha), float(c)


class Multiquadric(Kernel):
    """k = hypot(|x-y|^2, c) as coded (kernel.rs:1010-1070).  nb_parameters() is 2 in the reference while
    get_parameters() returns one value (kernel.rs:1039-1042, :1061-1064), so the reference cannot optimise it;
    here get/set are consistent with the single parameter."""
    _tag, _names = K_MULTIQUADRIC, ("c",)

    def __init__(self, c=0.0):
        self.c = float(c)


class RationalQuadratic(Kernel):
    """k = (1 + |x-y|^2 / (2 alpha ls^2))^-alpha  (kernel.rs:1079-1157)"""
    _tag, _names = K_RATIONAL_QUADRATIC, ("alpha", "ls")

    def __init__(self, alpha=1.0, ls=1.0):
        self.alpha, self.ls = float(alpha), float(ls)
