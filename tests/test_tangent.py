"""Row f2, second half: the opt-in CONSISTENT tangent -- the exact linearisation of the reference's
internal force (`/root/reference/stiffnessMtrx.py:609-644`) in place of its constant-C stiffness (ddsdde is never updated:
`material_zoo/neo_hookean.py:62-64` is commented out).  CPU side: the oracle statement against finite differences of the
oracle's internal force, the KERNEL SOURCE (k_assemble_scatter_ct) on the SIMT emulation against the oracle statement, and the
Newton driver on the emulated kernels.  Hardware tests: test_gpu_kernels.py."""
import numpy as np
import pytest

from helpers import GoldenDeck, load_golden, make_element, make_material, material_params, rel_err

from oracle import femcy_oracle as O

DECKS = ["c3d4_neohookean_newton", "c3d10_ellip", "cpe6_cook", "c3d4_cook", "cps4_ellip", "cps8_ellip", "cps3_ellip"]


def _state(g, amp=0.02, seed=0):
    nodes = g["nodes"]
    span = float((nodes.max(axis=0) - nodes.min(axis=0)).max())
    return amp * span * np.random.default_rng(seed).standard_normal(nodes.size), span


@pytest.mark.parametrize("name", ["c3d4_neohookean_newton", "c3d10_ellip", "cpe6_cook", "c3d4_cook"])
def test_oracle_consistent_tangent_is_the_derivative_of_the_internal_force(name):
    """columns of K against central differences of f_int (hyperelastic laws: the symmetrised tensor is the exact one)"""
    g = load_golden(name)
    nodes, el = g["nodes"], g["elements"].astype(np.int64)
    mc, p = material_params(g)
    et = str(g["elem_type"])
    u, span = _state(g)
    K = O.assemble_K_consistent(nodes, el, u, et, mc, p, g["C"]).toarray()
    assert abs(K - K.T).max() <= 1e-10 * abs(K).max()
    h = 1e-6 * span
    for j in np.random.default_rng(1).choice(nodes.size, 8, replace=False):
        up, um = u.copy(), u.copy()
        up[j] += h
        um[j] -= h
        col = (O.internal_force(nodes, el, up, et, mc, p, g["C"])[0] - O.internal_force(nodes, el, um, et, mc, p, g["C"])[0]) / (2 * h)
        assert abs(col - K[:, j]).max() <= 2e-5 * abs(K).max()


def test_oracle_consistent_tangent_at_rest_is_the_small_strain_stiffness():
    """at u = 0 the exact tangent of a linear-elastic law is B^T C B (no stress, no geometric part)"""
    g = load_golden("c3d4_cook")
    nodes, el = g["nodes"], g["elements"].astype(np.int64)
    mc, p = material_params(g)
    u0 = np.zeros(nodes.size)
    Kc = O.assemble_K_consistent(nodes, el, u0, "C3D4", mc, p, g["C"])
    K0 = O.assemble_K(nodes, el, u0, "C3D4", g["C"])
    assert abs(Kc - K0).max() <= 1e-7 * abs(K0).max()


@pytest.mark.parametrize("name", DECKS)
def test_emulated_consistent_tangent_kernel_matches_the_oracle(name):
    import simt
    g = load_golden(name)
    nodes, el = g["nodes"], g["elements"].astype(np.int64)
    mc, p = material_params(g)
    ELE, mat = make_element(g), make_material(g)
    u, _ = _state(g)
    Kref = O.assemble_K_consistent(nodes, el, u, str(g["elem_type"]), mc, p, g["C"])
    pat = simt.SellPattern(el, nodes.shape[0], dm=nodes.shape[1])
    dN, _ = ELE.device_tables()
    val, _, _ = simt.assemble_raw(simt.make_tables(ELE, mat), dN.shape, nodes, el, u, pat, variant=4, knob=int(mat.kind))
    K = pat.to_csr(val).tocsr()
    # both sides difference the constitutive law numerically (h = 1e-6): agreement to the round-off of that quotient
    assert abs(K - Kref).max() <= 1e-4 * abs(Kref).max()
    assert abs(K - K.T).max() <= 1e-9 * abs(K).max()


def test_newton_driver_with_the_consistent_tangent_on_emulated_kernels(monkeypatch):
    """neo-Hookean C3D4 deck: full Newton needs fewer loops than the reference's modified Newton and ends within the
    driver's own stopping tolerance of the same state"""
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    monkeypatch.setattr(sm, "Context", EmuContext)
    g = load_golden("c3d4_neohookean_newton")
    deck = GoldenDeck(g)
    out = {}
    for kind in ("reference", "consistent"):
        s = sm.System_of_equations(sm.Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE), list(deck.materials.values())[0], True,
                                   quiet=True, cg_eps=1e-10)
        s.set_tangent(kind)
        s.solve(deck)
        out[kind] = (s.dof.to_numpy(), list(s.inc_trace), s.cg_iters_total, s.tangent_fallbacks)
    ref, ct = out["reference"], out["consistent"]
    assert [(t, c) for t, c, _ in ref[1]] == [(t, c) for t, c, _ in ct[1]] and all(c for _, c, _ in ct[1])
    assert sum(l for _, _, l in ct[1]) < sum(l for _, _, l in ref[1]) and ct[2] < ref[2] and ct[3] == 0
    assert [int(l) for _, _, l in ref[1]] == [int(v) for v in g["inc_trace"][:, 2]]          # the default path is untouched
    assert rel_err(ct[0], ref[0]) < 5e-3
    with pytest.raises(ValueError):
        s.set_tangent("secant")
