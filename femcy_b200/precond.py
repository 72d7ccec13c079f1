"""Row f2: opt-in two-level preconditioner for `solve_by_CG` (library side: csrc/precond.cu).

The reference solves with plain Jacobi-PCG (conjugateGradientSolver.py:48-51); that remains the default.  With
`System_of_equations.set_preconditioner("two_level")` the same `femcy_cg_solve` call runs PCG preconditioned by
2 Chebyshev-Jacobi steps + a coarse correction over node aggregates x rigid-body modes.  This module only forms the
aggregates (host, once per mesh): geometric bins of the bounding box, sized so that the dense coarse level stays small."""
import numpy as np


def geometric_aggregates(nodes, max_coarse_unknowns=6000, min_nodes_per_aggregate=256):
    """aggregate id (0..nagg-1, all used) of every node: uniform bins of the bounding box.

    max_coarse_unknowns bounds (rigid-body modes per aggregate) x nagg -- the coarse matrix is inverted densely at every
    solve (cost ~ nc^3: 6000 unknowns = ~25 ms on a B200), so small meshes get few aggregates (>= 256 nodes each)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    nn, dm = nodes.shape
    nr = 3 if dm == 2 else 6
    target = max(1, min(int(max_coarse_unknowns) // nr, nn // int(min_nodes_per_aggregate)))
    lo, hi = nodes.min(axis=0), nodes.max(axis=0)
    ext = np.maximum(hi - lo, 1e-300)
    h0 = (np.prod(ext) / target) ** (1.0 / dm)
    nb = np.maximum(1, np.rint(ext / h0)).astype(np.int64)
    while int(np.prod(nb)) > target and nb.max() > 1:          # rounding may overshoot the budget
        nb[int(np.argmax(nb))] -= 1
    ijk = np.minimum((((nodes - lo) / ext) * nb).astype(np.int64), nb - 1)
    lin = ijk[:, 0]
    for c in range(1, dm):
        lin = lin * nb[c] + ijk[:, c]
    _, agg = np.unique(lin, return_inverse=True)
    return agg.astype(np.int32), int(agg.max()) + 1
