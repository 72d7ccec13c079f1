"""femcy_b200 -- B200-native (sm_100a CUDA) implementation of FEMcy's data-parallel hot path:
K assembly (Gauss-point loop, Bt.C.B, scatter into the global sparse matrix) and the
Jacobi-preconditioned CG loop, behind the reference's own class / plugin surface.

    from femcy_b200 import InpInfo, Body, System_of_equations
    inp = InpInfo("deck.inp")
    body = Body(inp.nodes, list(inp.eSets.values())[0], inp.ELE)
    system = System_of_equations(body, list(inp.materials.values())[0], inp.geometric_nonlinear)
    system.solve(inp)
    u = system.dof.to_numpy()

The CUDA library is built in-tree by `python -m femcy_b200.build`; nothing here falls back to a
CPU implementation.
"""
from .body import Body
from .conjugateGradientSolver import ConjugateGradientSolver_rowMajor
from .reader import InpInfo
from .stiffnessMtrx import System_of_equations

__all__ = ["Body", "ConjugateGradientSolver_rowMajor", "InpInfo", "System_of_equations"]
__version__ = "0.1.0"
