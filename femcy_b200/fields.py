"""Small field objects that keep the reference's `ti.field` call surface on top of our storage.

The reference reads and writes simulation state through Taichi fields
(`field.to_numpy()`, `field.from_numpy(a)`, `field.fill(v)`, `field.copy_from(other)`,
`field.shape`, `field[i]`).  Callers written against that surface keep working:

* HostField   -- plugin tables (Gauss points, weights ...) that live on the host; an ndarray
                 subclass with the extra methods.
* DeviceVector / DeviceGPArray -- views of the named device arrays owned by a femcy_ctx
                 (include/femcy_b200.h: enum femcy_vec / femcy_gp_array); every access is an
                 explicit host<->device copy.
"""
import numpy as np

from ._lib import VEC


class HostField(np.ndarray):
    def __new__(cls, data, dtype=np.float64):
        return np.asarray(data, dtype=dtype).view(cls)

    def to_numpy(self):
        return np.array(self)

    def from_numpy(self, a):
        self[...] = np.asarray(a).reshape(self.shape)

    def copy_from(self, other):
        self[...] = np.asarray(other)


class DeviceVector:
    """One of the ctx's named length-N vectors (dof, rhs, residual, x, ...)."""

    def __init__(self, ctx, name, n):
        self.ctx, self.name, self.n = ctx, name, int(n)
        self.shape = (self.n,)

    def to_numpy(self):
        return self.ctx.vec_get(self.name, self.n)

    def from_numpy(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size != self.n:
            raise ValueError(f"{self.name}: expected {self.n} entries, got {a.size}")
        self.ctx.vec_set(self.name, a)

    def fill(self, v):
        self.ctx.call("femcy_vec_fill", VEC[self.name], float(v))

    def copy_from(self, other):
        self.ctx.call("femcy_vec_copy", VEC[self.name], VEC[other.name])

    def __getitem__(self, i):
        return self.to_numpy()[i]

    def __len__(self):
        return self.n


class DeviceGPArray:
    """Per-Gauss-point array (vol, dsdx, F, cauchy_stress, mises_stress, strain, energy density)."""

    def __init__(self, ctx, name, shape, perm=None, section=None):
        self.ctx, self.name, self.shape = ctx, name, tuple(int(s) for s in shape)
        self.perm = perm      # device element k is the caller's element perm[k] (locality reordering)
        self.section = section   # row f4: index of the mesh section this array belongs to (None: single-section mesh)

    def _select(self):
        if self.section is not None:
            self.ctx.call("femcy_select_section", int(self.section))

    def to_numpy(self):
        self._select()
        a = self.ctx.gp_get(self.name, self.shape)
        if self.perm is None:
            return a
        out = np.empty_like(a)
        out[self.perm] = a
        return out

    def from_numpy(self, a):
        a = np.asarray(a, dtype=np.float64).reshape(self.shape)
        self._select()
        self.ctx.gp_set(self.name, a if self.perm is None else a[self.perm])

    def extrapolate_on_device(self, E, comp=0, n_en=None, nn=None):
        """(per-element nodal values [ne, n_en], nodal means [nn] or None) of component `comp` of this field:
        nodal = E . Gauss-point values on the device (femcy_extrapolate), E = ELE.extrapolation_matrix()."""
        from ._lib import GP, as_d
        E = np.ascontiguousarray(E, dtype=np.float64)
        ne = self.shape[0]
        en = np.empty((ne, E.shape[0]))
        mean = np.empty(int(nn)) if nn else None
        self._select()
        self.ctx.call("femcy_extrapolate", GP[self.name], int(comp), as_d(E), as_d(en), as_d(mean) if mean is not None else None)
        if self.perm is not None:
            out = np.empty_like(en)
            out[self.perm] = en
            en = out
        return en, mean

    def __getitem__(self, i):
        return self.to_numpy()[i]


class SectionedGPField:
    """Row f4: a per-Gauss-point field of a mesh of several sections = one DeviceGPArray per section (the sections may
    differ in nodes and Gauss points per element, so there is no single [ne, n_gp, ...] array).  `to_numpy()` returns the
    list of the sections' arrays; `parts[i]` is section i's field."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.name = self.parts[0].name
        self.ctx = self.parts[0].ctx

    def to_numpy(self):
        return [p.to_numpy() for p in self.parts]

    def from_numpy(self, arrays):
        for p, a in zip(self.parts, arrays):
            p.from_numpy(a)

    def max(self):
        return max(float(a.max()) for a in self.to_numpy() if a.size)

    def __getitem__(self, i):
        return self.parts[i]

    def __len__(self):
        return len(self.parts)
