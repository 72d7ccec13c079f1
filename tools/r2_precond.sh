#!/bin/bash
# 1 GPU: parity of the opt-in two-level preconditioner + its effect on cfg 4 (linear solve), cfg 3' (twist plate) and cfg 5
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "two_level" > gpurun_out/${tag}_tests.log 2>&1
echo "two-level tests rc=$?"; tail -15 gpurun_out/${tag}_tests.log
python /dev/stdin <<'PY' 2>&1 | tee gpurun_out/${tag}_precond.log
import sys, time, json, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from femcy_b200 import Body, System_of_equations, meshgen
def linear(kind, n=None, cells=None, lengths=None, eps=1e-8, coarse=6000):
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1) if n else meshgen.SyntheticDeck(kind, cells=cells, lengths=lengths, jitter=0.0)
    s = System_of_equations(Body(deck.nodes, deck.eSets[kind], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
    s.assemble_stiffnessMtrx()
    nb = deck.neumann_bc_info[0]; s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info: s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    out = {"dofs": s.N}
    for kind_p in ("jacobi", "two_level"):
        s.set_preconditioner(kind_p, coarse) if kind_p == "two_level" else s.set_preconditioner("jacobi")
        for rep in range(2):
            t = time.perf_counter(); s.solve_by_CG(eps=eps, max_iter=200000, check_every=16 if kind_p == "two_level" else 64); s.ctx.sync(); wall = time.perf_counter() - t
        out[kind_p] = {"iters": s.last_cg_iters, "device_ms": round(s.ctx.time_ms(1), 3), "wall_ms": round(wall * 1e3, 3)}
        out[kind_p + "_x"] = s._x.to_numpy()
    out["naggs"] = getattr(s, "n_aggregates", None)
    out["rel_diff"] = float(np.abs(out["two_level_x"] - out["jacobi_x"]).max() / np.abs(out["jacobi_x"]).max())
    del out["jacobi_x"], out["two_level_x"]
    s.close()
    return out
print("plate 44x6x66 (cfg 3' size) eps 1e-10:", json.dumps(linear("C3D4", cells=(44, 6, 66), lengths=(80., 10., 120.), eps=1e-10)), flush=True)
print("cube n=48 eps 1e-8:", json.dumps(linear("C3D4", n=48)), flush=True)
print("cfg 4 (n=119) eps 1e-8, 6000 coarse:", json.dumps(linear("C3D4", n=119)), flush=True)
print("cfg 4 (n=119) eps 1e-8, 12000 coarse:", json.dumps(linear("C3D4", n=119, coarse=12000)), flush=True)
PY
FEMCY_OPT_CG_PRECOND=1 python tools/cfg5_multi.py --tag ${tag}_two_level 2>&1 | grep '^{' | cut -c1-1200
python /dev/stdin <<'PY' 2>&1 | tee -a gpurun_out/${tag}_precond.log
import sys, time, json, os
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import run_configs as rc
for pre in ("0", "1"):
    os.environ["FEMCY_OPT_CG_PRECOND"] = pre
    t = time.time()
    out = rc.run_deck(rc.twist_deck(n_inc=1))
    print("cfg 3' twist plate 104544 C3D4, 1 increment, precond", pre, json.dumps({k: out[k] for k in ("solve_s", "increments", "cg_iterations_total", "max_abs_u")}), "wall", round(time.time() - t, 1), flush=True)
PY
