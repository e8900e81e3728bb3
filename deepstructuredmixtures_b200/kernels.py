"""Kernel-function parameter holders (host side) and the stand-alone Gram operator.

Mirrors the structs of /root/reference/src/kernels.jl: IsoSE :59-66, ArdSE :109-116, IsoLinear :174-179,
ArdLinear :209-214.  Hyper-parameters are stored in LOG scale exactly like the reference (`logl`, `logs`);
all arithmetic on them happens inside libdsmgp.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from . import _native as nat


class KernelFunction:
    type: int = -1

    def __init__(self, logl, logs: float = 0.0):
        self.logl = np.atleast_1d(np.asarray(logl, dtype=np.float64)).copy()
        self.logs = float(logs)

    # gaussianprocess.jl:139-145 -> (lengthscale, variance, noise): len(logl) + 2 hyper-parameters
    @property
    def nparams(self) -> int:
        return self.logl.size + 2

    def copy(self):
        k = self.__class__.__new__(self.__class__)
        k.logl = self.logl.copy()
        k.logs = self.logs
        return k

    def getvariance(self, logscale: bool = False) -> float:
        return self.logs if logscale else math.exp(2.0 * self.logs)

    def getlengthscales(self, logscale: bool = False):
        v = self.logl if logscale else np.exp(self.logl)
        return float(v[0]) if v.size == 1 and self.type in (nat.ISO_SE, nat.ISO_LINEAR) else v.copy()

    def theta(self) -> np.ndarray:
        """[logl..., log sigma] as dsmgp_kernelmatrix expects."""
        return np.concatenate([self.logl, [self.logs]])

    def __repr__(self):
        return f"{self.__class__.__name__}(logl={self.logl.tolist()}, logs={self.logs})"


class IsoSE(KernelFunction):          # kernels.jl:59-66
    type = nat.ISO_SE

    def __init__(self, logl: float, logs: float):
        super().__init__([float(logl)], logs)


class ArdSE(KernelFunction):          # kernels.jl:109-116
    type = nat.ARD_SE

    def __init__(self, logl: Sequence[float], logs: float):
        super().__init__(logl, logs)


class _Linear(KernelFunction):
    def getvariance(self, logscale: bool = False) -> float:      # kernels.jl:181,216
        return 0.0 if logscale else 1.0


class IsoLinear(_Linear):             # kernels.jl:174-179
    type = nat.ISO_LINEAR

    def __init__(self, logl: float):
        super().__init__([float(logl)], 0.0)


class ArdLinear(_Linear):             # kernels.jl:209-214
    type = nat.ARD_LINEAR

    def __init__(self, logl: Sequence[float]):
        super().__init__(logl, 0.0)


def kernelmatrix(kernel: KernelFunction, x1: np.ndarray, x2: Optional[np.ndarray] = None) -> np.ndarray:
    """kernelmatrix(kernel, x1, x2)  kernels.jl:15-18, computed on the GPU (gram_rect_kernel)."""
    x1 = nat.colmajor(np.asarray(x1, dtype=np.float64).reshape(len(x1), -1))
    x2 = x1 if x2 is None else nat.colmajor(np.asarray(x2, dtype=np.float64).reshape(len(x2), -1))
    n1, D = x1.shape
    n2 = x2.shape[0]
    if kernel.type in (nat.ARD_SE, nat.ARD_LINEAR) and kernel.logl.size != D:
        raise ValueError("ARD kernel needs one length scale per input dimension")
    K = np.zeros((n1, n2), order="F")
    th = nat.f64(kernel.theta())
    nat.check(nat.lib().dsmgp_kernelmatrix(kernel.type, nat.p_d(th), D, nat.p_d(x1), n1, nat.p_d(x2), n2, nat.p_d(K)))
    return K
