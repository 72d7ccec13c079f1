"""Sequential pure-Python stand-in for the `taichi` package (TEST INFRASTRUCTURE ONLY).

Purpose: let the *unmodified* reference sources under /root/reference execute in a
container that has no Taichi.  Every `@ti.kernel` body
then runs as ordinary Python, one loop iteration per Taichi thread, so the
reference's own code becomes the tier-1 oracle on small decks.

Not a product path: nothing under `femcy_b200/` may import this module.  It is used
only by `oracle/run_reference.py` (golden-vector generation, run in the build
container where /root/reference exists).

Semantics covered = exactly the API surface the reference touches
(grep over /root/reference: ti.init, data_oriented, kernel, func, template, static,
grouped, f64/f32/i32, field, Vector(.field), Matrix(.field/.zero), abs/sin/cos/log,
GUI/ui dummies).  `ti.atomic_max/min` on kernel-local scalars cannot be emulated
by-reference; the runner replaces the five reductions that use them
(conjugateGradientSolver.py:67-72, tiGadgets.py:20-64) with NumPy one-liners.
"""
import math

import numpy as np

f64 = np.float64
f32 = np.float32
i32 = np.int32
float64 = np.float64
cuda = "cuda"
cpu = "cpu"


def init(**kwargs):
    return None


def data_oriented(cls):
    return cls


def kernel(fn):
    return fn


def func(fn):
    return fn


def template():
    return None


def static(*args):
    return args if len(args) != 1 else args[0]


class _Mat(np.ndarray):
    """ndarray view carrying the handful of ti.Matrix attributes the reference uses."""

    @property
    def n(self):
        return self.shape[0]

    @property
    def m(self):
        return self.shape[1]

    def inverse(self):
        return np.linalg.inv(np.asarray(self)).view(_Mat)

    def determinant(self):
        return np.linalg.det(np.asarray(self))

    def transpose(self):
        return np.asarray(self).T.view(_Mat)


def _as_mat(data, dt=None):
    arr = np.array(data) if dt is None else np.array(data, dtype=dt)
    return arr.view(_Mat)


class _Field:
    """Dense field: numpy storage of shape `shape + element_shape`; iterating yields indices."""

    def __init__(self, eshape, dtype, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.eshape = tuple(eshape)
        self.a = np.zeros(self.shape + self.eshape, dtype if dtype is not None else np.float64)

    def __getitem__(self, idx):
        if idx is None:
            idx = ()
        out = self.a[idx]
        if isinstance(out, np.ndarray) and out.ndim > 0:
            return out.view(_Mat)  # a view: `f[i][j] += v` writes through
        return out

    def __setitem__(self, idx, val):
        if idx is None:
            idx = ()
        self.a[idx] = val

    def __iter__(self):
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return iter(np.ndindex(*self.shape))

    def from_numpy(self, x):
        self.a[...] = np.asarray(x).reshape(self.a.shape)

    def to_numpy(self):
        return self.a.copy()

    def fill(self, v):
        self.a[...] = v

    def copy_from(self, other):
        self.a[...] = other.a


def field(dtype, shape=(), **kwargs):
    return _Field((), dtype, shape)


class Vector:
    def __new__(cls, data, dt=None):
        return _as_mat(data, dt)

    @staticmethod
    def field(n, dtype=None, shape=(), **kwargs):
        return _Field((n,), dtype, shape)


class Matrix:
    def __new__(cls, data, dt=None):
        return _as_mat(data, dt)

    @staticmethod
    def field(n, m, dtype=None, shape=(), **kwargs):
        return _Field((n, m), dtype, shape)

    @staticmethod
    def zero(dt, n, m=None):
        return np.zeros((n, m) if m is not None else (n,), dt).view(_Mat)


def grouped(f):
    return iter(np.ndindex(*f.shape))


abs = np.abs  # noqa: A001  (mirrors ti.abs)
sin = math.sin
cos = math.cos
log = math.log


def atomic_max(a, b):
    raise NotImplementedError("by-reference scalar reduction: patched by oracle/run_reference.py")


atomic_min = atomic_max


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


class ui:  # noqa: N801
    Window = _Dummy
    Camera = _Dummy
    Scene = _Dummy
    LMB = 0


GUI = _Dummy


def rgb_to_hex(c):
    return 0
