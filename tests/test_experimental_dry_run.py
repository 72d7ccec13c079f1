"""Dry run, on the CPU, of the gated hardware tests (tests/test_gpu_experimental.py): the same test functions, with the
CUDA context replaced by emu_ctx.EmuContext (kernel source on the SIMT emulation).  Purpose: the first GPU minutes of the
next round must not be spent on a Python error or a wrong tolerance in a test that could never run here."""
import pytest

import femcy_b200.stiffnessMtrx as sm
import test_gpu_experimental as X


@pytest.fixture
def emu(monkeypatch):
    from emu_ctx import EmuContext
    monkeypatch.setattr(sm, "Context", EmuContext)
    return EmuContext


@pytest.mark.parametrize("name", ["cps3_ellip", "cps6_ellip", "c3d4_ellip", "c3d10_cook"])
def test_dry_experimental_assembly_matches_reference(emu, name):
    X.test_experimental_assembly_matches_reference(name)


@pytest.mark.parametrize("kind,n", [("C3D4", 4), ("C3D10", 2)])
def test_dry_experimental_assembly_on_synthetic_mesh(emu, kind, n):
    X.test_experimental_assembly_on_synthetic_mesh(kind, n)


def test_dry_single_reduction_pcg(emu, monkeypatch):
    # the env switch is read by the C library; the emulated context takes it from a class attribute
    calls = []
    orig = emu._femcy_cg_solve

    def spy(self, *a):
        import os
        self.cg_variant = 1 if os.environ.get("FEMCY_CG_VARIANT") == "sr" else 0
        calls.append(self.cg_variant)
        return orig(self, *a)
    monkeypatch.setattr(emu, "_femcy_cg_solve", spy)
    X.test_single_reduction_pcg_matches_default(5, 1e-8, monkeypatch)
    assert calls == [0, 1]
    X.test_single_reduction_pcg_fixed_iterations(monkeypatch)


@pytest.mark.parametrize("name", ["c3d10_ellip", "cps6_ellip", "c3d4_cook"])
def test_dry_sigma_sorted_pattern_assembly_and_solve(emu, name, monkeypatch):
    X.test_sigma_sorted_pattern_assembly_and_solve(name, monkeypatch)


def test_dry_sigma_sorted_solve(emu, monkeypatch):
    # the hardware test uses a C3D10 cube of 6 cells per edge; 2 per edge here
    import femcy_b200.meshgen as mg
    real = mg.SyntheticDeck
    monkeypatch.setattr(mg, "SyntheticDeck", lambda kind, n=6, **kw: real(kind, n=2, **kw))
    X.test_sigma_sorted_solve_matches_natural_order(monkeypatch)


def test_dry_symmetric_half_storage_pcg(emu, monkeypatch):
    """Python-level rehearsal of the FEMCY_CG_SYM hardware test on a 3-cell cube; the emulated context maps the switch to
    the emulated kernel's `sym` flag like the C library maps the environment variable."""
    import femcy_b200.meshgen as mg
    real = mg.SyntheticDeck
    monkeypatch.setattr(mg, "SyntheticDeck", lambda kind, n=6, **kw: real(kind, n=3 if kind == "C3D4" else 2, **kw))
    calls = []
    orig = emu._femcy_cg_solve

    def spy(self, *a):
        import os
        self.cg_sym = 1 if os.environ.get("FEMCY_CG_SYM") == "1" else 0
        self.cg_variant = 1 if os.environ.get("FEMCY_CG_VARIANT") == "sr" else 0
        calls.append((self.cg_sym, self.cg_variant))
        return orig(self, *a)
    monkeypatch.setattr(emu, "_femcy_cg_solve", spy)
    X.test_symmetric_half_storage_pcg_matches_default("C3D4", 12, 1e-8, monkeypatch)
    assert calls == [(0, 0)] * 5 + [(1, 0)] * 5 + [(1, 1)] * 5


def test_rehearsal_of_the_gpu_parity_suite_fast_subset():
    """the hardware parity tests themselves (tests/test_gpu_parity.py, test_gpu_edge_cases.py), unchanged, against the
    emulated context (-p emu_plugin): pattern, assembly, geometry / stress, Dirichlet, PCG iterates, the ELL drop-in and
    the ragged scalar matrix -- 56 tests in ~25 s.  The full rehearsal (all 76, ~6 min) is the command in README.md."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sel = ("pattern_matches or assembly_matches or geometry_and_stress or dirichlet or cg_iterates or synthetic "
           "or scalar_matrix or cg_dropin")
    r = subprocess.run([sys.executable, "-m", "pytest", "test_gpu_parity.py", "test_gpu_edge_cases.py", "-m", "gpu", "-q", "-x",
                        "-p", "emu_plugin", "-p", "no:cacheprovider", "-k", sel], cwd=here, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
