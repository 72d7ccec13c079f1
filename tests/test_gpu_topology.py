"""GPU tests of row f1 (csrc/topology.cu) through the C-ABI: femcy_boundary_facets / femcy_node_elements against the host
NumPy versions (exact), femcy_neumann against the rhs the reference's own neumannBC produced (golden `rhs_neumann`,
/root/reference/stiffnessMtrx.py:369-411; 1e-13) and against the host integration for pressure and TRVEC loads."""
import ctypes as C

import numpy as np
import pytest

from helpers import GoldenDeck, golden_names, load_golden, rel_err, system_from_deck

pytestmark = pytest.mark.gpu

DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "cpe6_cook", "c3d4_ellip", "c3d10_ellip", "c3d4_cook", "c3d10_cook"]


class _Pairs:
    def __init__(self, ele, kid):
        self.ele, self.kid = ele, kid


@pytest.mark.parametrize("name", DECKS)
def test_device_topology_matches_the_host_versions(name):
    from femcy_b200 import Body
    g = load_golden(name)
    deck = GoldenDeck(g)
    s = system_from_deck(deck)
    host = Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE)
    for a, b in zip(s.body.boundary_arrays(), host.boundary_arrays()):          # answered by femcy_boundary_facets
        assert np.array_equal(a, b)
    hp, hl = host.node_element_csr()
    dp, dl = s.body.node_element_csr()                                          # answered by femcy_node_elements
    assert np.array_equal(dp, hp) and np.array_equal(dl, hl)
    s.close()


@pytest.mark.parametrize("name", ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip", "c3d4_cook"])
def test_device_topology_matches_the_reference_body(name):
    """femcy_boundary_facets / femcy_node_elements against what the REFERENCE's own `Body.get_boundary` / `get_nodeEles`
    (body.py:165-234, run unmodified under the shim) produced: tests/golden/topology_reference.npz"""
    from test_topology import as_reference_boundary, reference_topology
    g = load_golden(name)
    ref = reference_topology(name)
    s = system_from_deck(GoldenDeck(g))
    n = C.c_int64(0)
    s.ctx.call("femcy_boundary_facets", C.byref(n))
    assert n.value == len(ref["owner"])
    facs, ele, _ = s.body.boundary_arrays()
    f, o = as_reference_boundary(facs, ele)
    assert np.array_equal(f, ref["facets"]) and np.array_equal(o, ref["owner"])
    ptr, lst = s.body.node_element_csr()
    assert np.array_equal(ptr, ref["ne_ptr"]) and np.array_equal(lst, ref["ne_list"])
    s.body.get_boundary()
    assert sorted(s.body.boundaryNodes) == ref["boundary_nodes"].tolist()
    s.close()


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in ("cps3_dense_cg",)])
def test_device_neumann_reproduces_the_reference_rhs(name):
    g = load_golden(name)
    if "rhs_neumann" not in g.files or int(g["n_neumann"]) == 0:
        pytest.skip("no load in this deck")
    deck = GoldenDeck(g)
    s = system_from_deck(deck)
    nbc = deck.neumann_bc_info[-1]
    s.rhs.fill(7.0)                                                             # must be overwritten, not accumulated
    s.neumannBC(nbc["face_set"], nbc["traction"], nbc.get("direction", np.array([])))
    assert rel_err(s.rhs.to_numpy(), g["rhs_neumann"]) < 1e-13
    s.close()


@pytest.mark.parametrize("name", DECKS)
def test_device_neumann_matches_the_host_integration(name):
    from femcy_b200.neumann import neumann_vector
    g = load_golden(name)
    deck = GoldenDeck(g)
    s = system_from_deck(deck)
    _, ele, kid = s.body.boundary_arrays()
    s.neumannBC(_Pairs(ele, kid), 2.5)                                          # pressure on the whole boundary
    assert rel_err(s.rhs.to_numpy(), neumann_vector(s.body, _Pairs(ele, kid), 2.5)) < 1e-13
    d = np.array([0.3, -1.0, 0.5])[: s.dm]
    s.neumannBC(_Pairs(ele[::2], kid[::2]), -1.5, d)
    assert rel_err(s.rhs.to_numpy(), neumann_vector(s.body, _Pairs(ele[::2], kid[::2]), -1.5, d)) < 1e-13
    s.neumannBC(_Pairs(ele[:0], kid[:0]), 1.0)
    assert not s.rhs.to_numpy().any()
    s.close()


def test_large_mesh_topology_properties():
    """1.3 M C3D4 elements (n = 60): 6 n^2 boundary faces per side of the cube, every node's element list complete, total
    load = traction x area"""
    from femcy_b200 import Body, System_of_equations, meshgen
    n = 60
    deck = meshgen.SyntheticDeck("C3D4", n=n)
    conn = deck.eSets["C3D4"]
    s = System_of_equations(Body(deck.nodes, conn, deck.ELE), deck.materials["Elastic"], False, quiet=True)
    cnt = C.c_int64(0)
    s.ctx.call("femcy_boundary_facets", C.byref(cnt))
    assert cnt.value == 6 * 2 * n * n
    ptr, lst = s.body.node_element_csr()
    assert ptr[-1] == conn.size and np.array_equal(np.diff(ptr), np.bincount(conn.reshape(-1), minlength=deck.nodes.shape[0]))
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], 2.0, nb["direction"])
    load = s.rhs.to_numpy().reshape(-1, 3).sum(axis=0)
    assert np.allclose(load, [0.0, 2.0, 0.0], atol=1e-10)
    s.close()


def test_topology_calls_report_errors():
    from femcy_b200._lib import Context, FemcyError, as_d, as_i32
    from femcy_b200 import meshgen
    ctx = Context(0)
    nodes = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    conn = np.array([[0, 1, 2, 3]], dtype=np.int32)
    ctx.call("femcy_set_mesh", 3, 4, 4, as_d(nodes), 1, 4, as_i32(conn))
    with pytest.raises(FemcyError, match="femcy_set_facet_tables first"):
        ctx.call("femcy_boundary_facets", C.byref(C.c_int64()))
    ELE = meshgen.Element_linear_tetrahedral()
    dN, w = ELE.device_tables()
    ctx.call("femcy_set_element", 1, as_d(dN), as_d(w))
    kn, fw, fn, fN, fdN = ELE.device_facet_tables()
    bad = kn.copy()
    bad[0, 0] = 9
    with pytest.raises(FemcyError, match="local node out of range"):
        ctx.call("femcy_set_facet_tables", 4, 3, 1, as_i32(bad), as_d(fw), as_d(fn), as_d(fN), as_d(fdN))
    ctx.call("femcy_set_facet_tables", 4, 3, 1, as_i32(kn), as_d(fw), as_d(fn), as_d(fN), as_d(fdN))
    cnt = C.c_int64(0)
    ctx.call("femcy_boundary_facets", C.byref(cnt))
    assert cnt.value == 4                                                       # a single tetrahedron: all four faces
    e, k = np.array([0], dtype=np.int32), np.array([4], dtype=np.int32)
    with pytest.raises(FemcyError, match="out of range"):
        ctx.call("femcy_neumann", 1, as_i32(e), as_i32(k), 1.0, None)
    ctx.close()


@pytest.mark.parametrize("kind,n,nranks", [("C3D4", 12, 2), ("C3D4", 12, 8), ("C3D10", 6, 4), ("C3D4", 40, 8)])
def test_device_partitioner_equals_the_numpy_statement(kind, n, nranks):
    """femcy_partition on one GPU, for every rank of an nranks-way split: owners, local elements, numbering, coordinates and
    halo plan equal the host NumPy statement of the scheme array for array (partition.py)"""
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition
    from test_partition import assert_same_partition
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1) if kind == "C3D4" else meshgen.SyntheticDeck(kind, n=n)
    nodes, conn = deck.nodes, deck.eSets[kind]
    for rank in sorted({0, nranks // 2, nranks - 1}):
        dev = Partition(nodes, conn, rank, nranks, device=0)
        assert dev.built_on == "device"
        assert_same_partition(Partition(nodes, conn, rank, nranks), dev)
    w = np.linspace(1.0, 1.3, nranks)
    assert_same_partition(Partition(nodes, conn, 1, nranks, weights=w), Partition(nodes, conn, 1, nranks, weights=w, device=0))
