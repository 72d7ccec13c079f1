"""Dry run of the round-opener measurement tools (tools/quick_ab.py, tools/ab_variants.py, tools/pick_defaults.py) on the
CPU: the CUDA context is replaced by emu_ctx.EmuContext and the meshes shrink to a few cells per edge.  The timings
are meaningless; the point is that every variant list, every C-ABI call and every JSON field of the tools runs once
here, so that the few GPU minutes they are written for are not lost to a Python slip."""
import json
import os
import runpy
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def emu_tools(monkeypatch, tmp_path):
    import femcy_b200
    import femcy_b200.stiffnessMtrx as sm
    from emu_ctx import EmuContext

    class ToolCtx(EmuContext):
        def time_ms(self, kind):
            return 1.0

    monkeypatch.setattr(sm, "Context", ToolCtx)
    monkeypatch.chdir(tmp_path)
    for k in ("FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_MINB", "FEMCY_CG_FOLD_BARRIER", "FEMCY_SELL_SIGMA"):
        monkeypatch.delenv(k, raising=False)
    return tmp_path


def _lines(path):
    return [json.loads(l) for l in open(path) if l.strip().startswith("{")]


def test_quick_ab_dry_run(emu_tools, monkeypatch):
    monkeypatch.setenv("QAB_N4", "3")
    monkeypatch.setenv("QAB_N10", "2")
    monkeypatch.setattr(sys, "argv", ["quick_ab.py", "dry"])
    runpy.run_path(os.path.join(ROOT, "tools", "quick_ab.py"), run_name="__main__")
    out = _lines(emu_tools / "gpurun_out" / "dry_quick_ab.jsonl")
    assert out[-1]["what"] == "done"
    errors = [d for d in out if "error" in d]
    assert not errors, errors
    asm = {(d["kind"], d["variant"]) for d in out if d["what"] == "assembly"}
    assert {("C3D4", v) for v in (1, 5, 11, 21, 2, 6, 7, 8, 16, 17, 9, 10, 20, 18, 12, 13, 14, 22)} <= asm
    assert {("C3D10", v) for v in (1, 19, 6, 7, 8, 9, 10, 20, 12, 13, 15, 2)} <= asm
    assert {d["variant"] for d in out if d["what"] == "cg"} >= {"persistent", "single_reduction", "three_kernel_graph"}
    sig = {d["sigma"]: d for d in out if d["what"] == "cg_c3d10"}
    assert set(sig) == {0, 256, 1024}
    assert sig[256]["nslots"] <= sig[0]["nslots"] and sig[256]["nnzb"] == sig[0]["nnzb"]
    # the summary tool reads what the A/B tool wrote
    monkeypatch.setattr(sys, "argv", ["pick_defaults.py", str(emu_tools / "gpurun_out" / "dry_quick_ab.jsonl")])
    runpy.run_path(os.path.join(ROOT, "tools", "pick_defaults.py"), run_name="__main__")


def test_ab_variants_dry_run(emu_tools, monkeypatch, capsys):
    monkeypatch.setattr(sys, "argv", ["ab_variants.py", "C3D4", "3", "C3D10", "2"])
    runpy.run_path(os.path.join(ROOT, "tools", "ab_variants.py"), run_name="__main__")
    recs = [json.loads(l) for l in capsys.readouterr().out.splitlines() if l.strip().startswith("{")]
    by_kind = {r["kind"]: r for r in recs if "kind" in r}
    assert set(by_kind) == {"C3D4", "C3D10"}
    for kind, r in by_kind.items():
        bad = {k: v for k, v in r["assembly"].items() if "error" in v or v.get("max_rel_diff_vs_v1", 0.0) > 1e-12}
        assert not bad, (kind, bad)
    assert "v19" in by_kind["C3D10"]["assembly"] and "v15" in by_kind["C3D10"]["assembly"]
    assert "v18" in by_kind["C3D4"]["assembly"] and "v14" in by_kind["C3D4"]["assembly"]


def test_scaling_ab_dry_run(emu_tools, monkeypatch, capsys):
    """tools/scaling_ab.py (in-process A/B of the multi-GPU PCG switches), single process: every mode runs, the
    environment is restored between modes (a FEMCY_NO_P2P set by the partition survives), the JSON lines are complete."""
    import torch
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setenv("FEMCY_NO_P2P", "1")               # as partition.py sets it when peer access is missing
    monkeypatch.setattr(sys, "argv", ["scaling_ab.py", "--tag", "dry", "--n", "3", "--iters", "7", "--reps", "2"])
    runpy.run_path(os.path.join(ROOT, "tools", "scaling_ab.py"), run_name="__main__")
    assert os.environ.get("FEMCY_NO_P2P") == "1"
    assert "FEMCY_CG_VARIANT" not in os.environ and "FEMCY_CG_MULTIKERNEL" not in os.environ
    out = _lines(emu_tools / "gpurun_out" / "dry_scaling_ab.jsonl")
    assert out[0]["what"] == "setup" and out[-1]["what"] == "done"
    cg = [d for d in out if d["what"] == "cg"]
    sa = runpy.run_path(os.path.join(ROOT, "tools", "scaling_ab.py"), run_name="scaling_ab")
    assert [d["mode"] for d in cg] == sa["DEFAULT_ORDER"] and set(sa["DEFAULT_ORDER"]) == set(sa["MODES"])
    for d in cg:
        assert "error" not in d, d
        assert d["ms_per_iter_best"] > 0 and d["max_rel_diff_x_vs_first"] < 1e-9 and d["n_gpus"] == 1


def test_scaling_script_knows_every_mode_of_the_in_process_tool():
    """tools/r2_scaling_ab.sh re-runs the best in-process mode as a full bench.py: its mode -> environment table must
    agree with tools/scaling_ab.py."""
    import re
    sa = runpy.run_path(os.path.join(ROOT, "tools", "scaling_ab.py"), run_name="scaling_ab")
    sh = open(os.path.join(ROOT, "tools", "r2_scaling_ab.sh")).read()
    cases = dict(re.findall(r'^\s+([a-z0-9_]+)\)\s+echo "([^"]*)";;', sh, re.M))
    for mode, env in sa["MODES"].items():
        if mode in ("default", "multik_nccl"):
            continue
        want = sorted(f"{k}={v}" for k, v in env.items())
        assert sorted(cases.get(mode, "missing").split()) == want, mode
