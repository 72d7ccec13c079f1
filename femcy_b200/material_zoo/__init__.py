"""Material plugins (same class names as /root/reference/material_zoo/__init__.py:5-8)."""
from .materials import (MaterBase, LinearIsotropic, LinearIsotropicPlaneStrain, LinearIsotropicPlaneStress,
                        NeoHookean)

__all__ = ["MaterBase", "LinearIsotropic", "LinearIsotropicPlaneStrain", "LinearIsotropicPlaneStress", "NeoHookean"]
