"""`ConjugateGradientSolver_rowMajor`: drop-in for the reference's Jacobi-PCG class
(`/root/reference/conjugateGradientSolver.py:8-127`) running on the B200 library.

Constructor contract kept: `spm` is the ELL value array [N, W] (f64), `sparseIJ` the ELL index
array [N, W+1] (i32, count in column 0, -1 padding), `b` the right-hand side [N]; all three are
*aliased*, not copied (the reference keeps references to the caller's fields, :17-19), so
`re_init()` / `solve()` see whatever the caller has written into them since.  The result is in
`.x` (a field with `to_numpy()`).

The ELL arrays are converted on the device to a scalar SELL-32 matrix (`femcy_cg_from_ell`); the
iteration itself is `femcy_cg_solve` (3 kernels per iteration, no host round trips, same
stopping rule max|r| < eps*max|r0| evaluated every iteration).
"""
import ctypes as C

import numpy as np

from ._lib import VEC, Context, as_d, as_i32
from .fields import DeviceVector


def _np(a):
    return a.to_numpy() if hasattr(a, "to_numpy") else np.asarray(a)


class ConjugateGradientSolver_rowMajor:
    def __init__(self, spm, sparseIJ, b, eps=1.0e-3, device: int = 0, check_every: int = 32, quiet: bool = True):
        self.A = spm
        self.ij = sparseIJ
        self.b = b
        self.eps = eps
        self.check_every = check_every
        self.quiet = quiet
        self.ctx = Context(device)
        self._N = int(_np(b).shape[0])
        self._upload_matrix()
        self.x = DeviceVector(self.ctx, "x", self._N)
        self.r = DeviceVector(self.ctx, "r", self._N)
        self.d = DeviceVector(self.ctx, "d", self._N)
        self.M = DeviceVector(self.ctx, "M", self._N)
        self.Ad = DeviceVector(self.ctx, "Ad", self._N)
        self.iterations = 0

    def _upload_matrix(self):
        spm = np.ascontiguousarray(_np(self.A), dtype=np.float64)
        ij = np.ascontiguousarray(_np(self.ij), dtype=np.int32)
        N, W = spm.shape
        if ij.shape != (N, W + 1):
            raise ValueError(f"sparseIJ must have shape {(N, W + 1)}, got {ij.shape}")
        self.ctx.call("femcy_cg_from_ell", N, W, as_d(spm), as_i32(ij))

    def re_init(self):
        """x, r, d, Ad <- 0 and M <- 1/diag(A) from the (possibly updated) aliased matrix
        (conjugateGradientSolver.py:32-38); the vectors are reset inside `solve`."""
        self._upload_matrix()

    def compute_Ad(self):
        self.ctx.call("femcy_spmv", VEC["d"], VEC["Ad"])

    def solve(self, max_iter=None, fixed_iters=False):
        b = np.ascontiguousarray(_np(self.b), dtype=np.float64)
        self.ctx.vec_set("rhs", b)
        it, r0, r1 = C.c_int64(0), C.c_double(0.), C.c_double(0.)
        if max_iter is None:
            max_iter = self._N   # "CG will converge within at most b.shape[0] loops" (:109)
        self.ctx.call("femcy_cg_solve", VEC["rhs"], float(self.eps), int(max_iter), int(self.check_every),
                      1 if fixed_iters else 0, C.byref(it), C.byref(r0), C.byref(r1))
        self.iterations = int(it.value)
        self.r0, self.rmax_final = r0.value, r1.value
        if not self.quiet:
            print(f"\033[32;1m the initial residual scale is {r0.value} \033[0m")
            print(f"\033[35;1m the {self.iterations - 1}-th loop, norm of residual is {r1.value} \033[0m")
            print(f"\033[32;1m CG solver's computation time is {self.ctx.time_ms(1) / 1e3} sec \033[0m")
        return self.x

    def rmax(self):
        return float(self.ctx.norms("r")[1])

    def close(self):
        self.ctx.close()
