"""Vector helpers with the reference's names (`/root/reference/tiGadgets.py`), running on the
device vectors of a femcy context (or on NumPy arrays when given arrays).

  c_equals_a_minus_b       tiGadgets.py:6-10
  a_equals_b_plus_c_mul_d  tiGadgets.py:13-17
  field_abs_max            tiGadgets.py:20-26
  field_norm               tiGadgets.py:29-37   (an RMS: sqrt(sum f^2 / N))
  field_multiply           tiGadgets.py:68-70
"""
import numpy as np

from ._lib import VEC
from .fields import DeviceVector


def _dev(*fs):
    return all(isinstance(f, DeviceVector) for f in fs)


def c_equals_a_minus_b(c, a, b):
    if _dev(c, a, b):
        c.ctx.call("femcy_vec_lincomb", VEC[c.name], VEC[a.name], -1.0, VEC[b.name])
    else:
        c[...] = np.asarray(a) - np.asarray(b)


def a_equals_b_plus_c_mul_d(a, b, c: float, d):
    if _dev(a, b, d):
        a.ctx.call("femcy_vec_lincomb", VEC[a.name], VEC[b.name], float(c), VEC[d.name])
    else:
        a[...] = np.asarray(b) + c * np.asarray(d)


def field_abs_max(f) -> float:
    if _dev(f):
        return float(f.ctx.norms(f.name)[1])
    return float(np.abs(np.asarray(f)).max())


def field_norm(f) -> float:
    if _dev(f):
        return float(f.ctx.norms(f.name)[0])
    a = np.asarray(f)
    return float((np.sum(a ** 2) / a.size) ** 0.5)


def field_multiply(field, num: float):
    if _dev(field):
        field.ctx.call("femcy_vec_scale", VEC[field.name], float(num))
    else:
        field *= num


def get_index_ti(arr, val) -> int:
    idx = np.nonzero(np.asarray(arr) == val)[0]
    return int(idx[-1]) if len(idx) else -1


def relative_error(a, b):
    m = max(abs(a), abs(b))
    return abs(a - b) / m if m > 1.e-9 else abs(a - b)
