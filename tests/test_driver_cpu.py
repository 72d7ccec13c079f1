"""CPU tests of the increment / Newton DRIVER (`System_of_equations.solve / advance_inc`, the host control flow
transcribed from /root/reference/stiffnessMtrx.py:647-822) against the traces of the reference's own run.

Two stand-ins for the CUDA context, both test infrastructure that nothing under femcy_b200/ knows about:
  * `fake_ctx.FakeContext` answers the C-ABI calls with the NumPy oracle + a direct sparse solve: isolates the HOST
    logic (boundary-condition plumbing, load stepping, the residual / relaxation / step-cutting decisions, the
    multi-increment quirks B14/B15);
  * `emu_ctx.EmuContext` answers them by running the product's CUDA KERNEL SOURCE on the SIMT emulation (tests/simt):
    real host code + real kernel code end to end, incl. the library's default choices (assembly variant, persistent PCG).
The GPU parity tests (tests/test_gpu_parity.py) remain the parity gate for the shipped binary."""
import numpy as np
import pytest

import femcy_b200.stiffnessMtrx as sm
from helpers import GoldenDeck, load_golden, rel_err, system_from_deck


def _solve(monkeypatch, ctx_cls, name, **kw):
    monkeypatch.setattr(sm, "Context", ctx_cls)
    g = load_golden(name)
    deck = GoldenDeck(g)
    if g["inc_trace"].shape[0] and float(g["inc_trace"][-1, 0]) < deck.time_incs["max_time"]:
        deck.time_incs["max_time"] = float(g["inc_trace"][-1, 0])      # the golden run was stopped after these increments
    s = system_from_deck(deck, **kw)
    s.solve(deck)
    return g, s


def _same_trace(s, g):
    got = [(round(t, 12), bool(c), int(n)) for t, c, n in s.inc_trace]
    want = [(round(float(t), 12), bool(c), int(n)) for t, c, n, _ in g["inc_trace"]]
    assert got == want, (got, want)


LINEAR = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "cpe3_cook", "cpe6_cook", "c3d4_ellip", "c3d10_ellip",
          "c3d4_cook", "c3d10_cook", "cps3_bydisp_4inc", "cps3_dirforce_4inc"]
NEWTON = ["c3d4_neohookean_newton", "cps6_beam_largedef_newton", "c3d4_twist_2inc"]


@pytest.mark.parametrize("name", LINEAR)
def test_host_driver_linear_decks(monkeypatch, name):
    from fake_ctx import FakeContext
    g, s = _solve(monkeypatch, FakeContext, name)
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-9
    _same_trace(s, g)


@pytest.mark.parametrize("name", NEWTON)
def test_host_driver_newton_decks(monkeypatch, name):
    """identical increment / Newton-loop trace and the reference's final displacement."""
    from fake_ctx import FakeContext
    g, s = _solve(monkeypatch, FakeContext, name)
    _same_trace(s, g)
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-6


# ---- real host code + real kernel source, end to end on the CPU ----------------------------------------------------
EMU_DECKS = ["cps3_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip", "cps3_bydisp_4inc", "c3d4_neohookean_newton"]


@pytest.mark.parametrize("name", EMU_DECKS)
def test_driver_over_emulated_kernels_matches_reference(monkeypatch, name):
    """`System_of_equations.solve` with every C-ABI call answered by the product's kernel source on the SIMT emulation
    (default assembly variant, persistent PCG to the direct-solve tolerance): the reference's increment / Newton trace
    and its final displacement; the Mises field of the final state."""
    from emu_ctx import EmuContext
    log = []
    monkeypatch.setattr(EmuContext, "assembly_log", log)
    g, s = _solve(monkeypatch, EmuContext, name)
    _same_trace(s, g)
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-6
    assert set(log) == {2}       # the library default (gather) was exercised
    s.compute_strain_stress()
    scale = np.abs(g["cauchy_final"]).max()
    assert np.abs(s.mises_stress.to_numpy() - g["mises_final"]).max() < 1e-5 * scale


@pytest.mark.parametrize("options", [{"cg_kernel": 2}, {"cg_sym": 1}, {"cg_kernel": 1}])
def test_driver_over_emulated_kernels_with_pcg_options(monkeypatch, options):
    """the same end-to-end path with the other PCG kernels: persistent with plain loads, upper-half SpMV, three-kernel."""
    from emu_ctx import EmuContext
    monkeypatch.setattr(EmuContext, "options", dict(options))
    g, s = _solve(monkeypatch, EmuContext, "c3d4_ellip")
    _same_trace(s, g)
    assert rel_err(s.dof.to_numpy(), g["dof_final"]) < 1e-6


def test_headless_driver_on_a_reference_deck_with_vtk(monkeypatch, tmp_path):
    """`python -m femcy_b200.main deck.inp --vtk out.vtk` (the reference's main.py without prompts / GUI): .inp reader ->
    Body -> System_of_equations.solve over the emulated kernels -> stress recovery -> extrapolate -> VTK file; checked
    against the golden of the same deck (BASELINE.json configs[0])."""
    import os
    from emu_ctx import EmuContext
    from femcy_b200 import main as headless
    from femcy_b200.vtk import read_vtk
    g = load_golden("cps3_ellip")
    deck = os.path.join(os.environ.get("FEMCY_REFERENCE", "/root/reference"), str(g["deck"]))
    if not os.path.exists(deck):
        pytest.skip("reference decks not present")
    monkeypatch.setattr(sm, "Context", EmuContext)
    monkeypatch.setattr(headless, "System_of_equations", sm.System_of_equations)
    out = headless.run(deck, quiet=True, vtk=str(tmp_path / "res.vtk"), stress_index=1)
    assert rel_err(out["dof"], g["dof_final"]) < 1e-6
    assert np.abs(out["mises"] - g["mises_final"]).max() < 1e-5 * np.abs(g["cauchy_final"]).max()
    r = read_vtk(str(tmp_path / "res.vtk"))
    assert np.array_equal(r["cells"], g["elements"]) and r["point_data"]["U"].shape == (g["nodes"].shape[0], 3)
    assert np.allclose(r["point_data"]["U"][:, :2].reshape(-1), out["dof"], rtol=0, atol=0)
