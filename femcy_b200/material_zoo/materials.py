"""Material plugins of the B200 path.

Surface kept from the reference (`/root/reference/material_zoo/mater_base.py:9-27`): attributes
`type` ("3d" | "planeStrain" | "planeStress"), `dm`, `C` (constant tangent used as ddsdde,
stiffnessMtrx.py:124-129), optional `C_6x6`, and the methods
`constitutiveOfSmallDeform / constitutiveOfLargeDeform(deformationGradient, cauchy_stress, ddsdde)`
and `elasticEnergyDensity(F)`.

In the reference those methods are Taichi kernels over fields.  Here a material is a *kind id* +
parameter blob + C matrix handed to the CUDA library (`femcy_set_material`); the constitutive
methods dispatch to the device kernel `femcy_constitutive` when given device fields of a
System_of_equations, and are plain NumPy when given arrays (host-side use, unit tests).

  kind 0 LinearIsotropic            /root/reference/material_zoo/linear_isotropic.py:10-99
  kind 1 LinearIsotropicPlaneStrain /root/reference/material_zoo/linear_isotropic_plane_strain.py:10-100
  kind 2 LinearIsotropicPlaneStress /root/reference/material_zoo/linear_isotropic_plane_stress.py:10-114
  kind 3 NeoHookean                 /root/reference/material_zoo/neo_hookean.py:9-89
"""
import abc

import numpy as np

from ..fields import HostField, DeviceGPArray, SectionedGPField


class MaterBase(abc.ABC):
    kind: int       # id understood by femcy_set_material
    type: str
    dm: int

    def device_params(self):
        """float64 parameter blob for femcy_set_material (see include/femcy_b200.h)."""
        raise NotImplementedError

    # device dispatch shared by all materials -------------------------------------------------
    def _constitutive(self, F, cauchy, large):
        if isinstance(F, (DeviceGPArray, SectionedGPField)):     # (several sections: the library call covers them all)
            if not (isinstance(cauchy, (DeviceGPArray, SectionedGPField)) and cauchy.ctx is F.ctx):
                raise TypeError("deformationGradient and cauchy_stress must be fields of the same system")
            F.ctx.call("femcy_constitutive", 1 if large else 0)
            return None
        Fa = np.asarray(F, dtype=np.float64)
        out = self.cauchy_from_F(Fa, large)
        if cauchy is not None:
            cauchy[...] = out
        return out

    def constitutiveOfSmallDeform(self, deformationGradient, cauchy_stress=None, ddsdde=None):
        return self._constitutive(deformationGradient, cauchy_stress, False)

    def constitutiveOfLargeDeform(self, deformationGradient, cauchy_stress=None, ddsdde=None):
        return self._constitutive(deformationGradient, cauchy_stress, True)

    @abc.abstractmethod
    def cauchy_from_F(self, F, large):
        """NumPy statement of the constitutive law on arrays F[..., dm, dm] (host-side use)."""

    @abc.abstractmethod
    def elasticEnergyDensity(self, F):
        pass


def _voigt3(E):
    return np.stack([E[..., 0, 0], E[..., 1, 1], E[..., 2, 2], 2. * E[..., 0, 1], 2. * E[..., 2, 0], 2. * E[..., 1, 2]], axis=-1)


def _unvoigt3(s):
    return np.stack([np.stack([s[..., 0], s[..., 3], s[..., 4]], -1),
                     np.stack([s[..., 3], s[..., 1], s[..., 5]], -1),
                     np.stack([s[..., 4], s[..., 5], s[..., 2]], -1)], -2)


def _T(A):
    return np.swapaxes(A, -1, -2)


class LinearIsotropic(MaterBase):
    kind = 0

    def __init__(self, modulus: float, poisson_ratio: float):
        self.type, self.dm = "3d", 3
        self.modulus, self.poisson_ratio = modulus, poisson_ratio
        nu = poisson_ratio
        self.G = G = modulus / 2. / (1. + nu)
        c00 = modulus * (1. - nu) / (1. + nu) / (1. - 2. * nu)
        c01 = modulus * nu / (1. + nu) / (1. - 2. * nu)
        C = np.zeros((6, 6))
        C[:3, :3] = c01
        C[np.arange(3), np.arange(3)] = c00
        C[np.arange(3, 6), np.arange(3, 6)] = G
        self.C = HostField(C)

    def device_params(self):
        return np.array([self.modulus, self.poisson_ratio])

    def cauchy_from_F(self, F, large):
        I = np.eye(3)
        C = np.asarray(self.C)
        if not large:
            E = (F + _T(F)) / 2. - I
            return _unvoigt3(_voigt3(E) @ C.T)
        E = (_T(F) @ F - I) / 2.
        pk2 = _unvoigt3(_voigt3(E) @ C.T)
        return F @ pk2 @ _T(F) / np.linalg.det(F)[..., None, None]

    def elasticEnergyDensity(self, F):
        F = np.asarray(F, dtype=np.float64)
        ev = _voigt3((_T(F) @ F - np.eye(3)) / 2.)
        return np.einsum("...p,pq,...q->...", ev, np.asarray(self.C), ev) / 2.


class LinearIsotropicPlaneStrain(MaterBase):
    kind = 1

    def __init__(self, modulus: float, poisson_ratio: float):
        self.type, self.dm = "planeStrain", 2
        self.modulus, self.poisson_ratio = modulus, poisson_ratio
        self.G = G = modulus / 2. / (1. + poisson_ratio)
        t1 = modulus / (1. + poisson_ratio)
        t2 = poisson_ratio / (abs(1. - 2. * poisson_ratio) + 1.e-30)
        c00, c01 = t1 * (1. + t2), t1 * t2
        self.C = HostField([[c00, c01, 0.], [c01, c00, 0.], [0., 0., G]])
        C6 = np.zeros((6, 6))
        C6[:3, :3] = c01
        C6[0, 0] = C6[1, 1] = c00
        C6[2, 2] = 0.
        C6[3, 3] = G
        self.C_6x6 = HostField(C6)

    def device_params(self):
        return np.array([self.modulus, self.poisson_ratio])

    def cauchy_from_F(self, F, large):
        I = np.eye(2)
        C = np.asarray(self.C)
        E = ((F + _T(F)) / 2. - I) if not large else (_T(F) @ F - I) / 2.
        ev = np.stack([E[..., 0, 0], E[..., 1, 1], E[..., 0, 1] + E[..., 1, 0]], -1)
        s = ev @ C.T
        S = np.stack([np.stack([s[..., 0], s[..., 2]], -1), np.stack([s[..., 2], s[..., 1]], -1)], -2)
        if not large:
            return S
        return F @ S @ _T(F) / np.linalg.det(F)[..., None, None]

    def elasticEnergyDensity(self, F):
        F = np.asarray(F, dtype=np.float64)
        F3 = np.zeros(F.shape[:-2] + (3, 3))
        F3[..., :2, :2] = F
        F3[..., 2, 2] = 1.
        ev = _voigt3((_T(F3) @ F3 - np.eye(3)) / 2.)
        return np.einsum("...p,pq,...q->...", ev, np.asarray(self.C_6x6), ev) / 2.


class LinearIsotropicPlaneStress(MaterBase):
    kind = 2

    def __init__(self, modulus: float, poisson_ratio: float):
        self.type, self.dm = "planeStress", 2
        self.modulus, self.poisson_ratio = modulus, poisson_ratio
        self.G = G = modulus / 2. / (1. + poisson_ratio)
        c00 = modulus / (1. - poisson_ratio ** 2)
        c01 = c00 * poisson_ratio
        self.C = HostField([[c00, c01, 0.], [c01, c00, 0.], [0., 0., G]])
        C6 = np.zeros((6, 6))
        C6[0, 0] = C6[1, 1] = c00
        C6[0, 1] = C6[1, 0] = c01
        C6[3, 3] = G
        self.C_6x6 = HostField(C6)

    def device_params(self):
        return np.array([self.modulus, self.poisson_ratio])

    def _embed(self, F):
        nu = self.poisson_ratio
        F3 = np.zeros(F.shape[:-2] + (3, 3))
        F3[..., :2, :2] = F
        F3[..., 2, 2] = -nu / (1. - nu) * (F[..., 0, 0] + F[..., 1, 1] - 2.) + 1.
        return F3

    def cauchy_from_F(self, F, large):
        I = np.eye(3)
        F3 = self._embed(F)
        C6 = np.asarray(self.C_6x6)
        if not large:
            E = (F3 + _T(F3)) / 2. - I
            return _unvoigt3(_voigt3(E) @ C6.T)[..., :2, :2]
        E = (_T(F3) @ F3 - I) / 2.
        pk2 = _unvoigt3(_voigt3(E) @ C6.T)
        return (F3 @ pk2 @ _T(F3) / np.linalg.det(F3)[..., None, None])[..., :2, :2]

    def elasticEnergyDensity(self, F):
        F3 = self._embed(np.asarray(F, dtype=np.float64))
        ev = _voigt3((_T(F3) @ F3 - np.eye(3)) / 2.)
        return np.einsum("...p,pq,...q->...", ev, np.asarray(self.C_6x6), ev) / 2.


class NeoHookean(MaterBase):
    """psi = C1 (I1 - 3 - 2 ln J) + D1 (J-1)^2 ;  sigma = 2 C1/J (B - I) + 2 D1 (J-1) I."""
    kind = 3

    def __init__(self, C1: float = 0.4, D1: float = 0.00025):
        self.type, self.dm = "3d", 3
        self.C1, self.D1 = C1, D1
        self.C = HostField(self.get_C())

    def get_C(self):
        vol = np.zeros((6, 6))
        vol[:3, :3] = 1.
        return 4. * self.C1 * np.eye(6) + 2. * self.D1 * vol

    def device_params(self):
        return np.array([self.C1, self.D1])

    def cauchy_from_F(self, F, large):
        I = np.eye(3)
        J = np.linalg.det(F)[..., None, None]
        return 2. * self.C1 / J * (F @ _T(F) - I) + 2. * self.D1 * (J - 1.) * I

    def elasticEnergyDensity(self, F):
        F = np.asarray(F, dtype=np.float64)
        J = np.linalg.det(F)
        trB = np.einsum("...ij,...ij->...", F, F)
        return self.C1 * (trB - 3. - 2. * np.log(J)) + self.D1 * (J - 1.) ** 2
