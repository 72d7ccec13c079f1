"""Shortest possible A/B of the assembly / PCG variants (a few seconds of GPU time): one JSON line per measurement,
flushed immediately to gpurun_out/<tag>_quick_ab.jsonl so that a run cut off by the budget still leaves data."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femcy_b200 import Body, System_of_equations, meshgen  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r1z"
os.makedirs("gpurun_out", exist_ok=True)
fh = open(f"gpurun_out/{tag}_quick_ab.jsonl", "a")
T0 = time.time()


def emit(**kw):
    kw["t"] = round(time.time() - T0, 2)
    fh.write(json.dumps(kw) + "\n")
    fh.flush()
    os.fsync(fh.fileno())
    print(kw, flush=True)


def assembly(kind, n, variants, reps=4):
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    conn = deck.eSets[kind]
    s = System_of_equations(Body(deck.nodes, conn, deck.ELE), list(deck.materials.values())[0], False, quiet=True)
    emit(what="setup", kind=kind, n=n, ne=int(conn.shape[0]))
    for v in variants:
        try:
            s.assembly_variant = v
            ts = []
            for _ in range(reps):
                s.assemble_stiffnessMtrx()
                ts.append(round(s.ctx.time_ms(0), 4))
            emit(what="assembly", kind=kind, n=n, variant=v, ms=ts, Gelem_s=conn.shape[0] / min(ts[1:]) / 1e6)
        except Exception as e:
            emit(what="assembly", kind=kind, n=n, variant=v, error=str(e)[:200])
    return deck, s


deck, s = assembly("C3D4", int(os.environ.get("QAB_N4", "119")), [1, 5, 11, 2, 6, 7, 8, 16, 17, 9, 10, 20, 12, 13, 14, 21, 18, 22])   # TMA / bulk-copy variants last: a fault would poison the context
try:
    s.assembly_variant = 1
    s.assemble_stiffnessMtrx()
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    for name, env in (("persistent", {}), ("persistent_minb5", {"FEMCY_CG_MINB": "5"}), ("single_reduction", {"FEMCY_CG_VARIANT": "sr"}), ("single_reduction_minb5", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_MINB": "5"}), ("single_reduction_fold_barrier", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_FOLD_BARRIER": "1"}), ("three_kernel_graph", {"FEMCY_CG_MULTIKERNEL": "1"}), ("persistent_sym", {"FEMCY_CG_SYM": "1"}), ("single_reduction_sym", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1"}), ("persistent_l2persist", {"FEMCY_CG_L2_PERSIST": "1"}), ("persistent_sym_l2persist", {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "1"}), ("persistent_sym_l2matrix", {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "2"})):
        for k in ("FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_MINB", "FEMCY_CG_FOLD_BARRIER", "FEMCY_CG_SYM", "FEMCY_CG_L2_PERSIST"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms = []
        for _ in range(3):
            s.solve_by_CG(eps=1e-30, max_iter=100, check_every=100, fixed_iters=True)
            ms.append(round(s.ctx.time_ms(1) / 100, 5))
        emit(what="cg", variant=name, ms_per_iter=ms)
    for k in ("FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_MINB", "FEMCY_CG_FOLD_BARRIER", "FEMCY_CG_SYM", "FEMCY_CG_L2_PERSIST"):
        os.environ.pop(k, None)
except Exception as e:
    emit(what="cg", error=str(e)[:200])
s.close()
def cg_c3d10(sigma):
    """cfg 5's matrix: PCG ms/iteration in natural row order vs SELL-32-sigma (37 % vs ~5 % padding)."""
    os.environ["FEMCY_SELL_SIGMA"] = str(sigma)
    try:
        import ctypes as C
        deck, s = assembly("C3D10", int(os.environ.get("QAB_N10", "55")), [1, 6, 7, 8, 9, 10, 20, 12, 13, 15, 2, 19] if sigma == 0 else [1, 7, 10, 20, 15, 19])
        st = (C.c_int64 * 4)()
        s.ctx.call("femcy_pattern_stats", st)
        s.assembly_variant = 1
        s.assemble_stiffnessMtrx()
        nb = deck.neumann_bc_info[0]
        s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
        for bc in deck.dirichlet_bc_info:
            s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
        ms = []
        for _ in range(3):
            s.solve_by_CG(eps=1e-30, max_iter=100, check_every=100, fixed_iters=True)
            ms.append(round(s.ctx.time_ms(1) / 100, 5))
        emit(what="cg_c3d10", sigma=sigma, nnzb=int(st[0]), nslots=int(st[1]), ms_per_iter=ms)
        os.environ["FEMCY_CG_SYM"] = "1"
        try:
            ms = []
            for _ in range(3):
                s.solve_by_CG(eps=1e-30, max_iter=100, check_every=100, fixed_iters=True)
                ms.append(round(s.ctx.time_ms(1) / 100, 5))
            emit(what="cg_c3d10_sym", sigma=sigma, ms_per_iter=ms)
        except Exception as e:
            emit(what="cg_c3d10_sym", sigma=sigma, error=str(e)[:200])
        os.environ.pop("FEMCY_CG_SYM", None)
        s.close()
    except Exception as e:
        emit(what="cg_c3d10", sigma=sigma, error=str(e)[:200])
    os.environ.pop("FEMCY_SELL_SIGMA", None)


cg_c3d10(0)
cg_c3d10(256)
cg_c3d10(1024)
emit(what="done")
