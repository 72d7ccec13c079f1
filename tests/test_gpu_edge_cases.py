"""GPU tests of the C-ABI's error behaviour and of edge cases (empty sets, repeated nodes, zero loads,
1x1-block matrices, reuse of a context)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import GoldenDeck, load_golden, rel_err, system_from_deck

pytestmark = pytest.mark.gpu


def test_errors_are_reported_not_swallowed():
    from femcy_b200._lib import Context, FemcyError, as_d, as_i32
    ctx = Context(0)
    with pytest.raises(FemcyError, match="set_mesh first"):
        ctx.call("femcy_build_pattern", C.byref(C.c_int64()))
    nodes = np.zeros((4, 3))
    conn = np.array([[0, 1, 2, 3]], dtype=np.int32)
    with pytest.raises(FemcyError, match="unsupported"):
        ctx.call("femcy_set_mesh", 3, 4, 4, as_d(nodes), 1, 5, as_i32(conn))       # 5-node 3-D element
    ctx.call("femcy_set_mesh", 3, 4, 4, as_d(nodes), 1, 4, as_i32(conn))
    with pytest.raises(FemcyError, match="no matrix|build_pattern"):
        ctx.call("femcy_cg_solve", 1, 1e-3, 10, 1, 0, None, None, None)
    C3 = np.eye(3)
    p = np.array([1.0, 0.3])
    with pytest.raises(FemcyError, match="Voigt"):
        ctx.call("femcy_set_material", 0, as_d(p), 2, as_d(C3), 3)                   # 3x3 tangent on a 3-D mesh
    with pytest.raises(FemcyError, match="2-D material"):
        ctx.call("femcy_set_material", 2, as_d(p), 2, as_d(np.eye(6)), 6)            # plane stress on tets
    with pytest.raises(FemcyError, match="n_gp"):
        ctx.call("femcy_set_element", 9, as_d(np.zeros(200)), as_d(np.zeros(9)))
    ctx.close()


def test_empty_and_repeated_dirichlet_sets():
    g = load_golden("c3d4_ellip")
    s = system_from_deck(GoldenDeck(g))
    s.geometric_nonlinear = False
    s.assemble_stiffnessMtrx()
    K0 = s.csr().copy()
    s.rhs.from_numpy(g["rhs_neumann"])
    s.dirichletBC_linearEquations(np.zeros(0, dtype=np.int64), 0, 1.0)             # empty set: nothing changes
    assert abs(s.csr() - K0).max() == 0.0
    assert np.array_equal(s.rhs.to_numpy(), g["rhs_neumann"])
    ns = g["bc_nodes"][g["bc_ptr"][0]:g["bc_ptr"][1]]
    s.dirichletBC_linearEquations(np.concatenate([ns, ns[:3]]), int(g["bc_dof"][0]), 0.25)   # repeats in one set
    ref = system_from_deck(GoldenDeck(g))
    ref.assemble_stiffnessMtrx()
    ref.rhs.from_numpy(g["rhs_neumann"])
    ref.dirichletBC_linearEquations(ns, int(g["bc_dof"][0]), 0.25)
    assert abs(s.csr() - ref.csr()).max() < 1e-13 * abs(K0).max()      # two atomic assemblies: order noise only
    assert rel_err(s.rhs.to_numpy(), ref.rhs.to_numpy()) < 1e-12
    with pytest.raises(Exception, match="out of range"):
        s.dirichletBC_linearEquations(np.array([10 ** 6]), 0, 0.0)
    s.close()
    ref.close()


def test_nonzero_dirichlet_matches_sequential_oracle():
    """Prescribed non-zero values on neighbouring dofs: the reference kernel is racy there; the
    device kernel must give the sequential result (oracle)."""
    from oracle import femcy_oracle as O
    g = load_golden("cps3_ellip")
    s = system_from_deck(GoldenDeck(g))
    s.assemble_stiffnessMtrx()
    K0 = O.assemble_K(g["nodes"], g["elements"].astype(np.int64), np.zeros(g["nodes"].size), "CPS3", g["C"])
    rng = np.random.default_rng(2)
    rhs = rng.standard_normal(g["nodes"].size)
    s.rhs.from_numpy(rhs)
    nodes = np.arange(0, 40)                      # a connected patch: many constrained neighbours
    s.dirichletBC_linearEquations(nodes, 0, 0.37)
    s.dirichletBC_linearEquations(nodes[::2], 1, -1.5)
    dofs = np.concatenate([nodes * 2, nodes[::2] * 2 + 1])
    vals = np.concatenate([np.full(40, 0.37), np.full(20, -1.5)])
    Kref, rref = O.dirichlet_linear(K0, rhs, dofs, vals)
    assert abs(s.csr() - Kref).max() < 1e-12 * abs(K0).max()
    assert rel_err(s.rhs.to_numpy(), rref) < 1e-13
    s.close()


def test_zero_load_newton_increment_and_context_reuse():
    g = load_golden("c3d4_neohookean_newton")
    deck = GoldenDeck(g)
    deck.neumann_bc_info = []                      # no load: residual is zero from the start
    s = system_from_deck(deck)
    s.solve(deck)
    assert s.inc_trace and all(c for _, c, _ in s.inc_trace)
    assert np.abs(s.dof.to_numpy()).max() == 0.0
    # the same system object can be solved again after a reset of the clock (ini_residual is captured
    # once per object, like in the reference -- quirk B4 -- so it is cleared too)
    s.time0 = s.time1 = 0.0
    del s.ini_residual
    deck2 = GoldenDeck(g)
    s.solve(deck2)
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-6
    s.close()


def test_scalar_matrix_cg_with_ragged_rows():
    """ConjugateGradientSolver_rowMajor on a scalar (1x1-block) SPD matrix with very uneven row lengths
    (SELL padding exercised), incl. a size that is not a multiple of the slice height."""
    from femcy_b200 import ConjugateGradientSolver_rowMajor as CG
    rng = np.random.default_rng(4)
    N = 1003
    A = sp.random(N, N, density=0.01, random_state=5, format="lil")
    A[0, :200] = rng.standard_normal(200)           # one long row
    A = (A + A.T).tocsr()
    A = A + sp.diags(np.abs(A).sum(axis=1).A1 + 1.0)
    A = A.tocsr()
    A.sort_indices()
    cnt = np.diff(A.indptr)
    W = int(cnt.max())
    ij = -np.ones((N, W + 1), dtype=np.int32)
    ij[:, 0] = cnt
    spm = np.zeros((N, W))
    pos = np.arange(A.nnz) - np.repeat(A.indptr[:-1], cnt)
    r = np.repeat(np.arange(N), cnt)
    # shuffle the column order inside each row: the reference's rows are unsorted (Python sets)
    perm = np.concatenate([rng.permutation(c) for c in cnt])
    ij[r, perm + 1] = A.indices
    spm[r, perm] = A.data
    b = rng.standard_normal(N)
    cg = CG(spm, ij, b, eps=1e-12)
    cg.solve(max_iter=5000)
    import scipy.sparse.linalg as sl
    assert rel_err(cg.x.to_numpy(), sl.spsolve(A.tocsc(), b)) < 1e-9
    # aliasing semantics: the caller edits b in place, solve() sees it
    b *= 2.0
    cg.re_init()
    cg.solve(max_iter=5000)
    assert rel_err(cg.x.to_numpy(), sl.spsolve(A.tocsc(), b)) < 1e-9
    cg.close()
