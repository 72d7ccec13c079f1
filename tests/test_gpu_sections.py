"""GPU parity tests of row f4: meshes of several sections -- element kinds and / or materials
(`*Solid Section`) -- through the C-ABI (femcy_add_section / femcy_select_section) against the NumPy oracle.

The reference rejects these decks (`/root/reference/reader/inp_info.py:125-128` raises on several element types;
`main.py:24` uses the first material only), so the oracle is the reference's single-kind arithmetic (oracle/femcy_oracle.py,
pinned on the reference's goldens by tests/test_oracle.py) summed over the sections.  Tolerances as in test_gpu_parity.py:
pattern exact, K 1e-12 of max|K|, vectors 1e-11, converged solution 1e-8."""
import ctypes as C

import numpy as np
import pytest

from helpers import material_oracle_args, rel_err, sectioned_K, sectioned_direct_solution

from femcy_b200 import meshgen
from oracle import femcy_oracle as O

pytestmark = pytest.mark.gpu

KINDS = ["plate_linear", "plate_quadratic", "bar_bimaterial", "bar_mixed"]


def build(deck, **kw):
    from femcy_b200 import System_of_equations
    kw.setdefault("quiet", True)
    return System_of_equations(deck.body(), None, deck.geometric_nonlinear, **kw)


@pytest.mark.parametrize("kind", KINDS)
def test_sectioned_pattern_and_assembly_match_the_oracle(kind):
    deck = meshgen.SectionedDeck(kind, n=6)
    s = build(deck)
    nn, dm = deck.nodes.shape
    s.ctx.call("femcy_select_section", len(deck.sections) - 1)     # the all-section calls do not depend on the selection
    rng = np.random.default_rng(5)
    for u in (np.zeros(nn * dm), 0.01 * rng.standard_normal(nn * dm)):
        s.dof.from_numpy(u)
        s.assemble_stiffnessMtrx()
        s.assemble_stiffnessMtrx()            # twice: the zero-fill must precede the sections' scatter passes every time
        K = s.csr()
        Kref = sectioned_K(deck, u)
        assert K.nnz == Kref.nnz and np.array_equal(K.indptr, Kref.indptr) and np.array_equal(K.indices, Kref.indices)
        assert abs(K - Kref).max() <= 1e-12 * abs(Kref).max()
    s.close()


@pytest.mark.parametrize("kind", KINDS)
def test_sectioned_linear_solve_and_stress_recovery(kind):
    deck = meshgen.SectionedDeck(kind, n=6)
    s = build(deck)
    s.solve(deck)
    dm = deck.nodes.shape[1]
    rhs = s.neumann_vector(deck.neumann_bc_info[0]["face_set"], 1.0, deck.neumann_bc_info[0]["direction"])
    u_ref, K0 = sectioned_direct_solution(deck, rhs)
    u = s.dof.to_numpy()
    assert rel_err(u, u_ref) < 1e-8
    s.compute_strain_stress()
    sig, mis, F = s.cauchy_stress.to_numpy(), s.mises_stress.to_numpy(), s.F.to_numpy()
    vol = s.vol.to_numpy()
    for k, sec in enumerate(deck.sections):
        name, params, Cm = material_oracle_args(sec["material"])
        Fr = O.deformation_gradient(deck.nodes, sec["elements"], u, sec["etype"])
        assert rel_err(F[k], Fr) < 1e-12
        ref = O.cauchy_stress(Fr, name, params, Cm, False)
        assert rel_err(sig[k], ref) < 1e-11
        mt = {"LinearIsotropicPlaneStrain": "planeStrain", "LinearIsotropicPlaneStress": "planeStress"}.get(name, "3d")
        assert rel_err(mis[k], O.mises(ref, mt, params[1])) < 1e-11
        assert abs(vol[k].sum() - 1.0) < 1e-3                      # each section covers half of the 2 x 1 (x 1) body
    # elastic energy of the converged linear solution = u.K.u / 2 (to the order of the displacement gradient)
    e = s.get_elasEng()
    assert abs(e - 0.5 * u_ref @ (K0 @ u_ref)) < 1e-3 * e
    # nodal extrapolation per section on the device (row f3) against E . Gauss-point values
    for k, sec in enumerate(deck.sections):
        E = sec["ELE"].extrapolation_matrix()
        nodal, _ = s.mises_stress[k].extrapolate_on_device(E)
        assert rel_err(nodal, mis[k] @ np.asarray(E).T) < 1e-12
    s.close()


@pytest.mark.parametrize("kind", ["bar_bimaterial", "bar_mixed"])
def test_sectioned_internal_force_matches_the_oracle(kind):
    """f_int over two neo-Hookean materials (and two element kinds): the sum of the sections' oracle vectors"""
    deck = meshgen.SectionedDeck(kind, n=4, nlgeom=True)
    s = build(deck)
    nn, dm = deck.nodes.shape
    u = 0.02 * np.random.default_rng(11).standard_normal(nn * dm)
    s.dof.from_numpy(u)
    s.assemble_nodal_force_GN()
    f = s.nodal_force.to_numpy()
    f_ref = np.zeros(nn * dm)
    sig = s.cauchy_stress.to_numpy()
    for k, sec in enumerate(deck.sections):
        name, params, Cm = material_oracle_args(sec["material"])
        fk, sk, _ = O.internal_force(deck.nodes, sec["elements"], u, sec["etype"], name, params, Cm)
        f_ref += fk
        assert rel_err(sig[k], sk) < 1e-11
    assert rel_err(f, f_ref) < 1e-11
    s.close()


def test_sectioned_newton_solve_follows_the_oracle_backed_driver():
    """two neo-Hookean materials, nlgeom: the same host driver over the CUDA library and over the oracle-backed context
    (tests/fake_ctx.py) takes the same increments / Newton loops and ends at the same displacement"""
    import femcy_b200.stiffnessMtrx as sm
    from fake_ctx import SectionedFakeContext
    deck = meshgen.SectionedDeck("bar_bimaterial", n=4, nlgeom=True, traction=0.02)
    s = build(deck, cg_eps=1e-10)
    s.solve(deck)
    u, trace = s.dof.to_numpy(), list(s.inc_trace)
    s.close()
    real = sm.Context
    sm.Context = SectionedFakeContext
    try:
        r = sm.System_of_equations(deck.body(), None, True, quiet=True)
        r.solve(deck)
        u_ref, trace_ref = r.dof.to_numpy(), list(r.inc_trace)
    finally:
        sm.Context = real
    assert trace == trace_ref and all(c for _, c, _ in trace)
    assert rel_err(u, u_ref) < 1e-6


def test_section_calls_report_errors():
    from femcy_b200._lib import Context, FemcyError, as_d, as_i32
    ctx = Context(0)
    conn = np.array([[0, 1, 2, 3]], dtype=np.int32)
    with pytest.raises(FemcyError, match="set_mesh first"):
        ctx.call("femcy_add_section", 1, 4, as_i32(conn), None)
    nodes = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.], [1., 1., 1.]])
    ctx.call("femcy_set_mesh", 3, 5, 5, as_d(nodes), 1, 4, as_i32(conn))
    assert ctx.lib.femcy_section_count(ctx.h) == 1
    with pytest.raises(FemcyError, match="no such section"):
        ctx.call("femcy_select_section", 1)
    with pytest.raises(FemcyError, match="unsupported"):
        ctx.call("femcy_add_section", 1, 3, as_i32(conn), None)                    # a triangle on a 3-D mesh
    sec = C.c_int(-1)
    conn2 = np.array([[1, 2, 3, 4]], dtype=np.int32)
    ctx.call("femcy_add_section", 1, 4, as_i32(conn2), C.byref(sec))
    assert sec.value == 1 and ctx.lib.femcy_section_count(ctx.h) == 2
    ELE = meshgen.Element_linear_tetrahedral()
    dN, w = ELE.device_tables()
    mat = meshgen.LinearIsotropic(modulus=1.0, poisson_ratio=0.3)
    Cm, p = np.ascontiguousarray(mat.C, dtype=np.float64), np.ascontiguousarray(mat.device_params(), dtype=np.float64)
    for k in (0, 1):
        ctx.call("femcy_select_section", k)
        ctx.call("femcy_set_element", 1, as_d(dN), as_d(w))
        if k == 0:
            ctx.call("femcy_set_material", int(mat.kind), as_d(p), len(p), as_d(Cm), 6)
    ctx.call("femcy_build_pattern", C.byref(C.c_int64()))
    with pytest.raises(FemcyError, match="every section"):
        ctx.call("femcy_assemble_K", 0)                                            # section 1 has no material yet
    ctx.call("femcy_select_section", 1)
    ctx.call("femcy_set_material", int(mat.kind), as_d(p), len(p), as_d(Cm), 6)
    with pytest.raises(FemcyError, match="scatter-add"):
        ctx.call("femcy_assemble_K", 2)                                            # the gather needs one record format
    ctx.call("femcy_assemble_K", 0)
    # a new mesh drops the sections
    ctx.call("femcy_set_mesh", 3, 5, 5, as_d(nodes), 1, 4, as_i32(conn))
    assert ctx.lib.femcy_section_count(ctx.h) == 1
    ctx.close()
