"""User-defined boundary conditions (plugin point kept from `/root/reference/user_defined/user_api.py`).

`user_dirichletBC` prescribes, on the nodes of a set, the displacement component of a rigid
rotation by `time*pi` about the z axis through (40, 5, 0) -- the BC that drives the twist decks
(user_api.py:11-29).  It returns the prescribed values; the caller writes them into `dof` on the
device (`femcy_dirichlet_val`).
"""
import numpy as np


def user_dirichletBC(dof, nodeSet, dm: int, dm_specified: int, nodes, time: float):
    nodeSet = np.asarray(nodeSet, dtype=np.int64)
    X = np.asarray(nodes, dtype=np.float64)[nodeSet]
    center = np.array([40., 5., 0.])[:X.shape[1]]
    angle = time * 3.141592653589793
    c, s = np.cos(angle), np.sin(angle)
    rota = np.array([[c, s, 0.], [-s, c, 0.], [0., 0., 1.]])[:X.shape[1], :X.shape[1]]
    new_x = (X - center) @ rota.T + center
    vals = (new_x - X)[:, dm_specified]
    if dof is not None and hasattr(dof, "ctx"):
        from .._lib import as_d, as_i32
        n32 = np.ascontiguousarray(nodeSet, dtype=np.int32)
        comps = np.full(len(n32), dm_specified, dtype=np.int32)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        dof.ctx.call("femcy_dirichlet_val", as_i32(n32), as_i32(comps), as_d(v), len(n32))
    return vals
