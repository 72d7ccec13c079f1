"""Row f1: topology builders and the Neumann vector as device code (csrc/topology.cu).

CPU side: the KERNEL SOURCE on the SIMT emulation against the host NumPy versions (`Body.boundary_arrays`,
`Body.node_element_csr`, `neumann.neumann_vector`) and against the rhs the reference's own `neumannBC` produced
(`/root/reference/stiffnessMtrx.py:369-411`, golden `rhs_neumann`); the reader's (element, face) pairs; the `Body` queries
answered through a context.  The hardware tests are in test_gpu_topology.py."""
import os

import numpy as np
import pytest

from helpers import GoldenDeck, golden_names, load_golden, make_element, rel_err

DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "cpe6_cook", "c3d4_ellip", "c3d10_ellip", "c3d4_cook", "c3d10_cook"]


class _Pairs:
    def __init__(self, ele, kid):
        self.ele, self.kid = ele, kid


@pytest.mark.parametrize("name", DECKS)
def test_emulated_boundary_facets_and_node_elements_match_the_host_versions(name):
    import simt
    from femcy_b200 import Body
    g = load_golden(name)
    ELE = make_element(g)
    body = Body(g["nodes"], g["elements"], ELE)
    T = simt.Topology(ELE, g["nodes"], g["elements"])
    be, bk = T.boundary_facets()
    _, ele, kid = body.boundary_arrays()
    assert np.array_equal(be, ele) and np.array_equal(bk, kid)          # same facets, same order
    ptr, lst = T.node_elements()
    hp, hl = body.node_element_csr()
    assert np.array_equal(ptr, hp) and np.array_equal(lst, hl)


@pytest.mark.parametrize("name", DECKS)
def test_emulated_neumann_kernel_matches_the_host_integration(name):
    """pressure (outward normal from the pushed-forward natural normal) and TRVEC loads on boundary facets"""
    import simt
    from femcy_b200 import Body
    from femcy_b200.neumann import neumann_vector
    g = load_golden(name)
    ELE = make_element(g)
    body = Body(g["nodes"], g["elements"], ELE)
    _, ele, kid = body.boundary_arrays()
    T = simt.Topology(ELE, g["nodes"], g["elements"])
    assert rel_err(T.neumann(ele, kid, 2.5), neumann_vector(body, _Pairs(ele, kid), 2.5)) < 1e-13
    d = np.array([0.3, -1.0, 0.5])[: body.dm]
    assert rel_err(T.neumann(ele[::2], kid[::2], -1.5, d), neumann_vector(body, _Pairs(ele[::2], kid[::2]), -1.5, d)) < 1e-13
    assert not T.neumann(ele[:0], kid[:0], 1.0).any()                     # no facet: rhs = 0 (rhs.fill(0), :384)


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in ("cps3_dense_cg",)])
def test_emulated_neumann_kernel_reproduces_the_reference_rhs(name):
    """the rhs the reference's own neumannBC produced (golden rhs_neumann)"""
    import simt
    from femcy_b200 import Body
    g = load_golden(name)
    if "rhs_neumann" not in g.files or int(g["n_neumann"]) == 0:
        pytest.skip("no load in this deck")
    deck = GoldenDeck(g)
    body = Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE)
    nbc = deck.neumann_bc_info[-1]
    facets = np.array(sorted(nbc["face_set"]), dtype=np.int64)
    ele, kid = body.locate_boundary_facets(facets)
    T = simt.Topology(deck.ELE, deck.nodes, list(deck.eSets.values())[0])
    rhs = T.neumann(ele, kid, nbc["traction"], nbc.get("direction"))
    assert rel_err(rhs, g["rhs_neumann"]) < 1e-13


def test_body_queries_go_through_the_context(monkeypatch):
    """a System_of_equations built on a Body answers get_boundary / get_nodeEles from the library (here: the kernel source
    on the emulation); the answers equal the NumPy ones, and a closed system hands the queries back to NumPy"""
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    from femcy_b200 import Body
    monkeypatch.setattr(sm, "Context", EmuContext)
    g = load_golden("c3d4_ellip")
    deck = GoldenDeck(g)
    host = Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE)
    body = Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE)
    calls = []
    real = EmuContext.call

    def spy(self, name, *a):
        calls.append(name)
        return real(self, name, *a)
    monkeypatch.setattr(EmuContext, "call", spy)
    s = sm.System_of_equations(body, list(deck.materials.values())[0], False, quiet=True)
    assert "femcy_set_facet_tables" in calls
    for a, b in zip(body.boundary_arrays(), host.boundary_arrays()):
        assert np.array_equal(a, b)
    assert "femcy_boundary_facets" in calls
    assert body.get_boundary() == host.get_boundary() and body.boundaryNodes == host.boundaryNodes
    assert body.get_nodeEles() == host.get_nodeEles() and "femcy_node_elements" in calls
    # the loaded surface of the deck, looked up among the device's boundary facets, then integrated by the device kernel
    nbc = deck.neumann_bc_info[-1]
    s.neumannBC(nbc["face_set"], nbc["traction"], nbc.get("direction", np.array([])))
    assert "femcy_neumann" in calls and rel_err(s.rhs.to_numpy(), g["rhs_neumann"]) < 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference decks not present")
@pytest.mark.parametrize("deck", ["elliptic_membrane/element_quadratic/ellip_membrane_quadritic_trig_neumann.inp",
                                  "cook_membrane/3D/smallDef_qualEl_coarse/cook_3d_quadEl_smallDef.inp",
                                  "elliptic_membrane/element_quadrilateral/ellip_CPS8.inp"])
def test_reader_face_sets_carry_element_face_pairs(deck):
    """`elset, Sx` -> (element, facet key index): the pairs name exactly the facets of the reference-style tuple set"""
    from femcy_b200 import Body, InpInfo
    inp = InpInfo(os.path.join("/root/reference/tests", deck))
    body = Body(inp.nodes, list(inp.eSets.values())[0], inp.ELE)
    keys = np.asarray(inp.ELE.element_facets(), dtype=np.int64)
    assert inp.face_sets
    for fs in inp.face_sets.values():
        tuples = np.sort(np.take_along_axis(body.np_elements[fs.ele], keys[fs.kid], axis=1), axis=1)
        assert set(map(tuple, tuples.tolist())) == set(fs) and len(fs.ele) == len(fs)
        ele, kid = body.locate_boundary_facets(np.array(sorted(fs), dtype=np.int64))
        assert sorted(zip(ele.tolist(), kid.tolist())) == sorted(zip(fs.ele.tolist(), fs.kid.tolist()))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference decks not present")
@pytest.mark.parametrize("name", ["cps6_ellip", "c3d4_ellip", "cps8_ellip"])
def test_reader_deck_solves_through_the_device_neumann_path_without_a_facet_search(name, monkeypatch):
    """InpInfo -> FaceSet (survives the driver's deepcopy) -> femcy_neumann with the deck's own (element, face) pairs: no boundary
    search at all, and the solution is the reference's (golden dof_final), kernels on the emulation"""
    import copy
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    from femcy_b200 import Body, InpInfo
    monkeypatch.setattr(sm, "Context", EmuContext)
    g = load_golden(name)
    inp = InpInfo(os.path.join("/root/reference", str(g["deck"])))
    assert copy.deepcopy(inp.neumann_bc_info)[0]["face_set"].kid is not None
    calls = []
    real = EmuContext.call

    def spy(self, n, *a):
        calls.append(n)
        return real(self, n, *a)
    monkeypatch.setattr(EmuContext, "call", spy)
    s = sm.System_of_equations(Body(inp.nodes, list(inp.eSets.values())[0], inp.ELE), list(inp.materials.values())[0],
                               inp.geometric_nonlinear, quiet=True)
    s.solve(inp)
    assert "femcy_neumann" in calls and "femcy_boundary_facets" not in calls
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-8


@pytest.mark.parametrize("dim,seed", [(2, 0), (2, 1), (3, 2), (3, 3)])
def test_emulated_topology_kernels_on_unstructured_delaunay_meshes(dim, seed):
    """random point clouds, Delaunay triangles / tetrahedra (irregular valence, long runs of facets sharing their two smallest nodes):
    boundary facets and node -> elements equal the NumPy versions; a unit pressure on the closed boundary sums to zero"""
    import simt
    from scipy.spatial import Delaunay
    from femcy_b200 import Body
    from femcy_b200.element_zoo import ELEMENT_TYPES
    rng = np.random.default_rng(seed)
    pts = rng.random((120 if dim == 2 else 80, dim))
    simp = Delaunay(pts).simplices.astype(np.int64)
    if dim == 3:                                   # the C3D4 convention: det[x1-x2, x3-x2, x0-x2] > 0
        x = pts[simp]
        det = np.linalg.det(np.stack([x[:, 1] - x[:, 2], x[:, 3] - x[:, 2], x[:, 0] - x[:, 2]], axis=1))
        simp[det < 0] = simp[det < 0][:, [1, 0, 2, 3]]
    else:                                          # counter-clockwise triangles
        x = pts[simp]
        a, b = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        area = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        simp[area < 0] = simp[area < 0][:, [1, 0, 2]]
    ELE = ELEMENT_TYPES["CPS3" if dim == 2 else "C3D4"]()
    body = Body(pts, simp, ELE)
    T = simt.Topology(ELE, pts, simp)
    be, bk = T.boundary_facets()
    _, ele, kid = body.boundary_arrays()
    assert np.array_equal(be, ele) and np.array_equal(bk, kid)
    ptr, lst = T.node_elements()
    hp, hl = body.node_element_csr()
    assert np.array_equal(ptr, hp) and np.array_equal(lst, hl)
    rhs = T.neumann(ele, kid, 1.0).reshape(-1, dim)
    assert np.abs(rhs.sum(axis=0)).max() < 1e-12          # closed surface: the pressure resultant vanishes
    from femcy_b200.neumann import neumann_vector
    assert rel_err(rhs.reshape(-1), neumann_vector(body, _Pairs(ele, kid), 1.0)) < 1e-12


# ---- pinned on the reference's own topology code (tests/golden/topology_reference.npz, made by make_topology_golden.py) ---------
TOPO_DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip", "c3d4_cook"]


def reference_topology(name):
    t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "topology_reference.npz"))
    return {k: t[f"{name}_{k}"] for k in ("facets", "owner", "ne_ptr", "ne_list", "boundary_nodes")}


def as_reference_boundary(facs, ele):
    """(facets, owner) in the golden's order: lexicographic by the sorted facet node tuple"""
    order = np.lexsort(tuple(facs[:, c] for c in range(facs.shape[1] - 1, -1, -1)))
    return facs[order], ele[order]


@pytest.mark.parametrize("name", TOPO_DECKS)
def test_topology_matches_the_reference_body(name):
    """`Body.get_boundary` / `get_nodeEles` of the REFERENCE (body.py:165-234, run unmodified under the shim) against (a) the NumPy
    versions and (b) the kernel source of femcy_boundary_facets / femcy_node_elements on the emulation"""
    import simt
    from femcy_b200 import Body
    g = load_golden(name)
    ref = reference_topology(name)
    ELE = make_element(g)
    body = Body(g["nodes"], g["elements"], ELE)
    facs, ele, _ = body.boundary_arrays()
    f, o = as_reference_boundary(facs, ele)
    assert np.array_equal(f, ref["facets"]) and np.array_equal(o, ref["owner"])
    ptr, lst = body.node_element_csr()
    assert np.array_equal(ptr, ref["ne_ptr"]) and np.array_equal(lst, ref["ne_list"])
    body.get_boundary()
    assert sorted(body.boundaryNodes) == ref["boundary_nodes"].tolist()
    T = simt.Topology(ELE, g["nodes"], g["elements"])
    be, bk = T.boundary_facets()
    keys = np.asarray(ELE.element_facets(), dtype=np.int64)
    dfacs = np.sort(np.take_along_axis(g["elements"].astype(np.int64)[be], keys[bk], axis=1), axis=1)
    f, o = as_reference_boundary(dfacs, be.astype(np.int64))
    assert np.array_equal(f, ref["facets"]) and np.array_equal(o, ref["owner"])
    dptr, dlst = T.node_elements()
    assert np.array_equal(dptr, ref["ne_ptr"]) and np.array_equal(dlst, ref["ne_list"])
