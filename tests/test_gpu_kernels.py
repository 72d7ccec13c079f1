"""GPU parity of every kernel choice the library still offers (round 2: 2 assembly formulations, 3 PCG kernels, the
upper-half SpMV, sigma-sorted rows), on the reference goldens and on seeded synthetic meshes.  Nothing here is gated: the
driver's `pytest -m gpu` covers all of it."""
import ctypes as C

import numpy as np
import pytest

from helpers import load_golden, rel_err
from test_gpu_parity import K_on_golden_pattern, build_system

pytestmark = pytest.mark.gpu

DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip", "c3d4_cook", "c3d10_cook"]
PCG_KERNELS = {"three_kernel_graph": {"cg_kernel": 1}, "persistent_plain_loads": {"cg_kernel": 2},
               "persistent_streaming": {"cg_kernel": 3}}


def _linear_system(deck, kind, **kw):
    from femcy_b200 import Body, System_of_equations
    conn, mat = deck.eSets[kind], list(deck.materials.values())[0]
    s = System_of_equations(Body(deck.nodes, conn, deck.ELE), mat, False, quiet=True, **kw)
    s.assemble_stiffnessMtrx()
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    return s


@pytest.mark.parametrize("kind,n", [("C3D4", 24), ("C3D10", 9)])
def test_assembly_formulations_agree_on_synthetic_mesh(kind, n):
    """thousands of slices / blocks: gather (default and explicit) against the atomic scatter; the gather is
    bit-reproducible and must not accumulate over repeated calls."""
    from femcy_b200 import Body, System_of_equations, meshgen
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    conn, mat = deck.eSets[kind], list(deck.materials.values())[0]
    u = 1e-3 * np.random.default_rng(0).standard_normal(deck.nodes.size)
    ref = None
    for variant in (1, 0, 2):
        s = System_of_equations(Body(deck.nodes, conn, deck.ELE), mat, False, quiet=True, assembly_variant=variant)
        s.dof.from_numpy(u)
        s.assemble_stiffnessMtrx()
        K = s.csr()
        if ref is None:
            ref = K
        else:
            assert abs(K - ref).max() <= 1e-12 * abs(ref).max(), (kind, variant)
            s.assemble_stiffnessMtrx()
            assert (s.csr() != K).nnz == 0, (kind, variant, "not bit-reproducible")
        s.close()


@pytest.mark.parametrize("general_tangent", [False, True])
def test_gather_with_a_general_tangent(general_tangent):
    """the gradient-product epilogue K = L(P) for a tangent that is NOT of cubic form (full 6x6 path) against the scatter,
    which multiplies B^T C B out per Gauss point."""
    from femcy_b200 import Body, System_of_equations, meshgen
    deck = meshgen.SyntheticDeck("C3D10", n=4)
    mat = list(deck.materials.values())[0]
    if general_tangent:
        rng = np.random.default_rng(5)
        A = rng.standard_normal((6, 6))
        mat.C = np.asarray(mat.C) + 0.05 * np.abs(np.asarray(mat.C)).max() * (A + A.T)     # symmetric, fully populated
    u = 1e-3 * np.random.default_rng(1).standard_normal(deck.nodes.size)
    Ks = []
    for variant in (1, 2):
        s = System_of_equations(Body(deck.nodes, deck.eSets["C3D10"], deck.ELE), mat, False, quiet=True, assembly_variant=variant)
        s.dof.from_numpy(u)
        s.assemble_stiffnessMtrx()
        Ks.append(s.csr())
        s.close()
    assert abs(Ks[0] - Ks[1]).max() <= 1e-12 * abs(Ks[0]).max()


@pytest.mark.parametrize("kind,n,eps", [("C3D4", 12, 1e-3), ("C3D4", 30, 1e-8), ("C3D10", 6, 1e-8)])
def test_pcg_kernels_agree(kind, n, eps):
    """the three PCG kernels run the same recurrence: same stop within an iteration or two, same solution to the stop
    rule's accuracy, first iterates equal to rounding; the streaming kernel is re-entered several times per solve
    (check_every) and solves twice on the same context."""
    from femcy_b200 import meshgen
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    s = _linear_system(deck, kind)
    out, fixed = {}, {}
    for name, opts in PCG_KERNELS.items():
        for k, v in opts.items():
            s.ctx.set_option(k, v)
        for rep in range(2):
            s.solve_by_CG(eps=eps, max_iter=20000, check_every=8)
            out[(name, rep)] = (s._x.to_numpy(), s.last_cg_iters, s.last_cg_residuals)
        for k in (1, 5, 17):
            s.solve_by_CG(eps=1e-30, max_iter=k, check_every=4, fixed_iters=True)
            assert s.last_cg_iters == k
            fixed[(name, k)] = s._x.to_numpy()
    s.close()
    xa, ia, _ = out[("three_kernel_graph", 0)]
    for name in PCG_KERNELS:
        for rep in range(2):
            xb, ib, (r0, r1) = out[(name, rep)]
            assert abs(ia - ib) <= max(2, ia // 50), (name, ia, ib)
            assert r1 < eps * r0
            assert np.abs(xa - xb).max() <= max(10 * eps, 1e-9) * np.abs(xa).max()
        for k in (1, 5, 17):
            a, b = fixed[("three_kernel_graph", k)], fixed[(name, k)]
            assert np.abs(a - b).max() <= 1e-10 * np.abs(a).max(), (name, k)


@pytest.mark.parametrize("kind,n,eps", [("C3D4", 12, 1e-3), ("C3D4", 30, 1e-8), ("C3D10", 6, 1e-8)])
@pytest.mark.parametrize("kernel", [2, 3])
def test_symmetric_half_storage_pcg_matches_default(kind, n, eps, kernel):
    """option cg_sym: the SpMV over the upper half of the matrix (transposed products scattered with fp64 atomics)."""
    from femcy_b200 import meshgen
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    s = _linear_system(deck, kind)
    s.ctx.set_option("cg_kernel", kernel)
    out, fixed = {}, {}
    for sym in (0, 1):
        s.ctx.set_option("cg_sym", sym)
        for rep in range(2):
            s.solve_by_CG(eps=eps, max_iter=20000, check_every=8)
            out[(sym, rep)] = (s._x.to_numpy(), s.last_cg_iters, s.last_cg_residuals)
        for k in (1, 5, 17):
            s.solve_by_CG(eps=1e-30, max_iter=k, check_every=4, fixed_iters=True)
            fixed[(sym, k)] = s._x.to_numpy()
    s.close()
    xa, ia, _ = out[(0, 0)]
    for rep in range(2):
        xb, ib, (r0, r1) = out[(1, rep)]
        assert abs(ia - ib) <= max(2, ia // 50), (ia, ib)          # the stopping point is summation-order sensitive (+-2 %)
        assert r1 < eps * r0
        assert np.abs(xa - xb).max() <= max(10 * eps, 1e-9) * np.abs(xa).max()
    for k in (1, 5, 17):
        assert np.abs(fixed[(0, k)] - fixed[(1, k)]).max() <= 1e-10 * np.abs(fixed[(0, k)]).max(), k


def test_zero_right_hand_side_returns_zero():
    """b = 0: x = 0 after 0 iterations (the unguarded recurrence divides 0 by 0)."""
    from femcy_b200 import meshgen
    deck = meshgen.SyntheticDeck("C3D4", n=6, jitter=0.1)
    s = _linear_system(deck, "C3D4")
    s.rhs.fill(0.)
    for opts in PCG_KERNELS.values():
        for k, v in opts.items():
            s.ctx.set_option(k, v)
        s.solve_by_CG(eps=1e-8, max_iter=100, check_every=8)
        assert s.last_cg_iters == 0 and not s.last_cg_breakdown
        assert np.array_equal(s._x.to_numpy(), np.zeros(s.N))
    s.close()


def test_breakdown_is_reported():
    """an indefinite system (K -> -K on part of the diagonal is not needed: a NaN in b does it) sets the breakdown flag."""
    from femcy_b200 import meshgen
    deck = meshgen.SyntheticDeck("C3D4", n=5, jitter=0.1)
    s = _linear_system(deck, "C3D4")
    b = s.rhs.to_numpy()
    b[7] = np.nan
    s.rhs.from_numpy(b)
    s.solve_by_CG(eps=1e-8, max_iter=50, check_every=4)
    assert s.last_cg_breakdown
    s.close()


# ---- SELL-32-sigma row order (option sell_sigma) -----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3d10_ellip", "cps6_ellip", "cps8_ellip", "c3d4_cook", "c3d4_neohookean_newton"])
def test_sigma_sorted_pattern_assembly(name, monkeypatch):
    g = load_golden(name)
    monkeypatch.setenv("FEMCY_OPT_SELL_SIGMA", "64")       # applied by Context() before the pattern is built
    s = build_system(g)
    K = s.csr().tocoo()                                   # exported in natural row order
    order = np.lexsort((K.col, K.row))
    assert np.array_equal(K.row[order], g["K_rows"]) and np.array_equal(K.col[order], g["K_cols"])
    for variant in (1, 2):
        s.assembly_variant = variant
        s.dof.from_numpy(g["u1"])
        s.assemble_stiffnessMtrx()
        v1, _ = K_on_golden_pattern(s, g)
        assert rel_err(v1, g["K1_vals"]) < 1e-12, (name, variant)
    s.close()


def test_sigma_sorted_solve_matches_natural_order(monkeypatch):
    from femcy_b200 import Body, System_of_equations, meshgen
    from femcy_b200.material_zoo import LinearIsotropic
    deck = meshgen.SyntheticDeck("C3D10", n=6)
    deck.geometric_nonlinear = False
    mat = LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3)
    out, slots = {}, {}
    for sigma in ("0", "256"):
        monkeypatch.setenv("FEMCY_OPT_SELL_SIGMA", sigma)
        s = System_of_equations(Body(deck.nodes, deck.eSets["C3D10"], deck.ELE), mat, False, quiet=True, cg_eps=1e-10)
        st = (C.c_int64 * 4)()
        s.ctx.call("femcy_pattern_stats", st)
        slots[sigma] = (int(st[0]), int(st[1]))
        s.solve(deck)
        out[sigma] = (s.dof.to_numpy(), s.last_cg_iters)
        s.close()
    assert slots["0"][0] == slots["256"][0]
    assert slots["256"][1] < 0.8 * slots["0"][1]           # the padding is gone
    assert abs(out["0"][1] - out["256"][1]) <= 2
    assert np.abs(out["0"][0] - out["256"][0]).max() <= 1e-8 * np.abs(out["0"][0]).max()


# ---- row f2: opt-in two-level preconditioner ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3d4_ellip", "c3d10_ellip", "cps6_ellip", "c3d4_cook"])
def test_two_level_preconditioner_matches_direct_solve(name):
    """converged solution <= 1e-8 of scipy's direct solve of the SAME eliminated system (2-D and 3-D, linear and quadratic
    elements), in far fewer iterations than Jacobi-PCG; the default path is untouched (same iterate bit for bit before and
    after the option was used)."""
    import scipy.sparse.linalg as sl
    g = load_golden(name)
    s = build_system(g, nlgeom=False)
    s.dof.fill(0.)
    s.assemble_stiffnessMtrx()
    s.rhs.from_numpy(g["rhs_neumann"])
    for k in range(len(g["bc_ptr"]) - 1):
        sl_ = slice(g["bc_ptr"][k], g["bc_ptr"][k + 1])
        s.dirichletBC_linearEquations(g["bc_nodes"][sl_], int(g["bc_dof"][k]), float(g["bc_val"][k]))
    K = s.csr().tocsc()
    b = s.rhs.to_numpy()
    x_ref = sl.spsolve(K, b)
    s.solve_by_CG(eps=1e-12, max_iter=100000, check_every=8)
    x_jac, it_jac = s._x.to_numpy(), s.last_cg_iters
    s.set_preconditioner("two_level", max_coarse_unknowns=600)
    s.solve_by_CG(eps=1e-12, max_iter=100000, check_every=4)
    x_two, it_two = s._x.to_numpy(), s.last_cg_iters
    assert np.abs(x_two - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    assert it_two < it_jac, (it_two, it_jac)          # (a few hundred dofs, 1-2 aggregates: the large gains are tested below)
    s.set_preconditioner("jacobi")
    s.solve_by_CG(eps=1e-12, max_iter=100000, check_every=8)
    assert np.array_equal(s._x.to_numpy(), x_jac) and s.last_cg_iters == it_jac
    s.close()


def test_two_level_preconditioner_reports_breakdown_on_nan():
    """a NaN in K (a diverged Newton step): no exception -- x = NaN and the breakdown flag, like the Jacobi recurrence."""
    from femcy_b200 import meshgen
    deck = meshgen.SyntheticDeck("C3D4", n=6, jitter=0.1)
    s = _linear_system(deck, "C3D4")
    u = np.zeros(s.N)
    u[10] = np.nan
    s.dof.from_numpy(u)
    s.assemble_stiffnessMtrx()
    s.set_preconditioner("two_level")
    s.solve_by_CG(eps=1e-8, max_iter=100, check_every=4)
    assert s.last_cg_breakdown and np.isnan(s._x.to_numpy()).any()
    s.close()


def test_two_level_preconditioner_on_synthetic_meshes():
    """cube and thin plate (the shape of BASELINE.json's twist deck): >= 4x fewer iterations at eps = 1e-10 even with a dozen aggregates, same solution."""
    from femcy_b200 import meshgen
    for cells, lengths in (((16, 16, 16), (1., 1., 1.)), ((22, 3, 33), (80., 10., 120.))):
        deck = meshgen.SyntheticDeck("C3D4", cells=cells, lengths=lengths, jitter=0.0)
        s = _linear_system(deck, "C3D4")
        s.solve_by_CG(eps=1e-10, max_iter=100000, check_every=8)
        x0, it0 = s._x.to_numpy(), s.last_cg_iters
        s.set_preconditioner("two_level")
        s.solve_by_CG(eps=1e-10, max_iter=100000, check_every=4)
        x1, it1 = s._x.to_numpy(), s.last_cg_iters
        assert it1 * 4 <= it0, (cells, it1, it0)
        assert np.abs(x1 - x0).max() <= 1e-8 * np.abs(x0).max()
        s.close()


# ---- row f3: Gauss-point field -> nodes on the device ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cps3_ellip", "cps6_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip"])
def test_device_extrapolation_matches_host(name):
    """femcy_extrapolate (ELE.extrapolate + nodal averaging on the device) against the host NumPy statement
    vals @ E^T and vtk.nodal_average, for a scalar field (Mises) and a tensor component (Cauchy_yy)."""
    from femcy_b200.vtk import nodal_average
    g = load_golden(name)
    s = build_system(g)
    s.dof.from_numpy(g["u1"])
    s.compute_strain_stress()
    E = s.ELE.extrapolation_matrix()
    nn = s.body.np_nodes.shape[0]
    mises = s.mises_stress.to_numpy()
    en, mean = s.mises_stress.extrapolate_on_device(E, nn=nn)
    ref = mises @ E.T
    assert np.abs(en - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1e-300)
    assert np.abs(mean - nodal_average(s.body, ref)).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300)
    assert np.array_equal(s.ELE.extrapolate(s.mises_stress), en)            # the plugin method routes device fields to the library
    dm = s.dm
    cauchy = s.cauchy_stress.to_numpy()
    en2, mean2 = s.cauchy_stress.extrapolate_on_device(E, comp=1 * dm + 1, nn=nn)
    ref2 = cauchy[:, :, 1, 1] @ E.T
    assert np.abs(en2 - ref2).max() <= 1e-13 * max(np.abs(ref2).max(), 1e-300)
    assert np.abs(mean2 - nodal_average(s.body, ref2)).max() <= 1e-12 * max(np.abs(ref2).max(), 1e-300)
    s.close()


# ---- row f2: the opt-in consistent tangent (k_assemble_scatter_ct) ------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3d4_neohookean_newton", "c3d10_ellip", "cpe6_cook", "c3d4_cook", "cps4_ellip", "cps8_ellip", "cps3_ellip"])
def test_consistent_tangent_matches_the_oracle(name):
    """K of option consistent_tangent against the oracle's exact linearisation of the internal force (which tests/test_tangent.py
    pins on finite differences of the oracle's f_int); both sides difference the constitutive law with h = 1e-6, so they agree to
    the round-off of that quotient.  The default assembly of the same system is untouched by the option being set and reset."""
    from helpers import material_params
    from oracle import femcy_oracle as O
    g = load_golden(name)
    s = build_system(g, nlgeom=True)
    nodes, el = g["nodes"], g["elements"].astype(np.int64)
    span = float((nodes.max(axis=0) - nodes.min(axis=0)).max())
    u = 0.02 * span * np.random.default_rng(0).standard_normal(nodes.size)
    s.dof.from_numpy(u)
    s.assemble_stiffnessMtrx()
    K_default = s.csr()
    s.set_tangent("consistent")
    s.assemble_stiffnessMtrx()
    s.assemble_stiffnessMtrx()                  # zero-fill + scatter: must not accumulate
    K = s.csr()
    mc, p = material_params(g)
    Kref = O.assemble_K_consistent(nodes, el, u, str(g["elem_type"]), mc, p, g["C"])
    assert abs(K - Kref).max() <= 1e-4 * abs(Kref).max()
    assert abs(K - K.T).max() <= 1e-9 * abs(K).max()
    s.set_tangent("reference")
    s.assemble_stiffnessMtrx()
    assert abs(s.csr() - K_default).max() == 0.0
    s.geometric_nonlinear = False
    with pytest.raises(ValueError, match="non-linear"):
        s.set_tangent("consistent")                # a linear analysis is its stiffness matrix: refused
    s.close()


def test_consistent_tangent_newton_needs_fewer_loops():
    """neo-Hookean C3D4 Newton deck through the C-ABI: same increments, fewer Newton loops and PCG iterations than the reference's
    modified Newton (whose loop counts still equal the reference's own trace), same state within the driver's stopping tolerance"""
    from helpers import GoldenDeck, system_from_deck
    g = load_golden("c3d4_neohookean_newton")
    deck = GoldenDeck(g)
    out = {}
    for kind in ("reference", "consistent"):
        s = system_from_deck(deck, cg_eps=1e-10)
        s.set_tangent(kind)
        s.solve(deck)
        out[kind] = (s.dof.to_numpy(), list(s.inc_trace), s.cg_iters_total, s.tangent_fallbacks)
        s.close()
    ref, ct = out["reference"], out["consistent"]
    assert [int(l) for _, _, l in ref[1]] == [int(v) for v in g["inc_trace"][:, 2]]
    assert [(t, c) for t, c, _ in ref[1]] == [(t, c) for t, c, _ in ct[1]] and all(c for _, c, _ in ct[1])
    assert sum(l for _, _, l in ct[1]) < sum(l for _, _, l in ref[1]) and ct[2] < ref[2] and ct[3] == 0
    assert rel_err(ct[0], ref[0]) < 5e-3


def test_consistent_tangent_on_a_mesh_of_several_sections():
    """two neo-Hookean materials (row f4 + row f2): every section differentiates its own law"""
    from femcy_b200 import System_of_equations, meshgen
    from helpers import material_oracle_args
    from oracle import femcy_oracle as O
    deck = meshgen.SectionedDeck("bar_bimaterial", n=4, nlgeom=True)
    s = System_of_equations(deck.body(), None, True, quiet=True)
    s.set_tangent("consistent")
    nn, dm = deck.nodes.shape
    u = 0.02 * np.random.default_rng(2).standard_normal(nn * dm)
    s.dof.from_numpy(u)
    s.assemble_stiffnessMtrx()
    K = s.csr()
    Kref = None
    for sec in deck.sections:
        name, params, Cm = material_oracle_args(sec["material"])
        Kp = O.assemble_K_consistent(deck.nodes, sec["elements"], u, sec["etype"], name, params, Cm)
        Kref = Kp if Kref is None else Kref + Kp
    assert abs(K - Kref).max() <= 1e-4 * abs(Kref).max()
    s.close()
