"""Dry run of bench.py's own-arm on the CPU: torch.cuda is replaced by inert stand-ins (streams, events that report
1 ms) and the CUDA context by emu_ctx.EmuContext, on a 3-cells-per-edge cube.  The NUMBERS are meaningless; the point is
that every line of `run_ours` -- the step loop, the in-loop profiling call, the end-to-end leg through the public API
(H2D u, get_dsdx_and_vol, assemble, the mesh-volume read-back, H2D rhs, Dirichlet, PCG, D2H x), the roofline / e2e /
config fields of the JSON line -- executes and keeps the contract's keys, so a Python slip cannot eat the GPU run."""
import argparse
import json
import types

import numpy as np
import pytest


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 1.0


class _Stream:
    cuda_stream = 0


class _StreamCtx:
    def __init__(self, s):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


@pytest.mark.parametrize("gp_sum_fails", [False, True])
def test_bench_own_arm_dry_run(monkeypatch, gp_sum_fails):
    import torch
    import bench
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext
    from fake_ctx import _arr, _set

    class BenchCtx(EmuContext):
        def time_ms(self, kind):
            return 1.0

        def launches(self):
            return self.n_launch

        def _femcy_set_stream(self, s):
            pass

        def _femcy_vec_set(self, which, ptr, n):
            self.vec_set(which, _arr(ptr, n))

        def _femcy_vec_get(self, which, ptr, n):
            _arr(ptr, n)[:] = self.vec_get(which, n)

        def _femcy_gp_sum(self, which, out):
            if gp_sum_fails:
                raise RuntimeError("simulated failure of the new entry point")
            _set(out, float(self.gp["vol"].sum()))

        def _femcy_gp_get(self, which, ptr, n):
            _arr(ptr, n)[:] = self.gp["vol"].reshape(-1)[:n]

    monkeypatch.setattr(sm, "Context", BenchCtx)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "stream", _StreamCtx)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["dry run"]})
    monkeypatch.setattr(bench, "cpu_baseline", lambda **kw: {"value": 1.0, "unit": "elem/s", "cores": 1, "kind": "port", "sample": "dry run"})
    from femcy_b200 import meshgen
    real_deck = meshgen.SyntheticDeck
    monkeypatch.setattr(meshgen, "SyntheticDeck", lambda kind="C3D4", n=8, **kw: real_deck(kind, n=min(n, 3 if kind == "C3D4" else 2), **kw))
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", n=3, cg_iters=4, cpu_sample_n=4, no_cpu_baseline=False,
                              ref_n=4, ref_cg_iters=2, balance="equal", no_parity=False, write_parity_golden=False,
                              no_builder_timings=False)
    out = bench.run_ours(args)
    line = json.loads(json.dumps(out))                      # it must serialise
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["dtype"] == "f64" and line["vs_baseline"] is None
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert "workload" in line["config"] and line["config"]["assembly_variant"] == 0
    assert abs(line["e2e"]["mesh_volume"] - 1.0) < 1e-12      # the unit cube, read back through femcy_gp_sum's handler
    assert ("vol array" in line["e2e"]["what"]) == gp_sum_fails  # the fallback read-back keeps the run alive
    assert line["roofline_assembly"]["kernel"].startswith("k_elem_geometry4t (TMA tensor store) + k_assemble_gather_h")
    assert line["roofline"]["kernel"].startswith("k_cg_stream") and line["roofline"]["spmv_phase"]["ms"] > 0
    for key in ("frac_dram", "phase_us_per_iteration", "ms_per_launch", "algorithmic_bytes_per_launch"):
        assert key in line["roofline"], key
    assert line["gpu_launches"] > 0
    # post-run correctness leg: eps = 1e-8 solve, true residual recomputed with one more SpMV (no committed fixture at n = 3)
    assert line["parity_ok"] is True and line["parity"]["golden"] is None
    assert line["parity"]["residual_inf_rel"] < 1e-6 and line["parity"]["iters"] > 0
    # post-run timings of the device builders / opt-in kernels (rows f1, f2): every entry ran (a failure would be a string)
    b = line["builders_ms"]
    for key in ("pattern_build", "boundary_facets", "node_elements_incl_d2h", "neumann_device", "neumann_host_numpy", "assembly_scatter",
                "assembly_consistent_tangent"):
        assert isinstance(b[key], float), (key, b[key])
    assert b["boundary_facets_found"] == 6 * 2 * 3 * 3 and b["neumann_facets"] == 2 * 3 * 3
    assert isinstance(b["partition_device_rank3_of_8"], str)         # no GPU here: reported as text, the run goes on
    assert isinstance(b["assembly_two_sections_scatter"], float) and b["two_sections_nnz"] == line["config"].get("nnz", b["two_sections_nnz"])
    nw = b["newton_c3d10_n16"]                                        # (n = 2 here, see the patched SyntheticDeck below)
    assert all(nw["consistent"]["converged"]) and sum(nw["consistent"]["newton_loops"]) <= sum(nw["reference"]["newton_loops"])


def test_smoke_entry_dry_run(monkeypatch, capsys):
    """__graft_entry__.smoke() -- the driver's first GPU step -- against the emulated context."""
    import __graft_entry__ as entry
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext

    class SmokeCtx(EmuContext):
        def launches(self):
            return self.n_launch

    monkeypatch.setattr(sm, "Context", SmokeCtx)
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
