#!/bin/bash
# 1 GPU: C3D10 (cfg 5) assembly + PCG with natural and sigma-sorted rows; ncu --set full of the gradient-product gather
tag=${1:-r2i}
mkdir -p gpurun_out
cat > /tmp/c10.py <<'PY'
import sys, os, json, numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
kind, n = sys.argv[1], int(sys.argv[2])
variants = [int(x) for x in sys.argv[3].split(",")]
mode = sys.argv[4] if len(sys.argv) > 4 else "time"
deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
s = System_of_equations(Body(deck.nodes, deck.eSets[kind], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
import ctypes as C
st = (C.c_int64 * 4)(); s.ctx.call("femcy_pattern_stats", st)
out = {"kind": kind, "n": n, "sigma": os.environ.get("FEMCY_OPT_SELL_SIGMA", "0"), "nnzb": int(st[0]), "nslots": int(st[1])}
for v in variants:
    s.assembly_variant = v
    ts = []
    for _ in range(3 if mode == "ncu" else 10):
        s.assemble_stiffnessMtrx(); s.ctx.sync(); ts.append(s.ctx.time_ms(0))
    out[f"asm_v{v}_ms"] = round(float(np.median(ts[1:])), 4)
if mode == "time":
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    for name, opts in (("persist", {"cg_kernel": 2}), ("stream", {"cg_kernel": 3}), ("stream_sym", {"cg_kernel": 3, "cg_sym": 1})):
        for k, v in {"cg_kernel": 0, "cg_sym": 0, **opts}.items():
            s.ctx.set_option(k, v)
        ms = []
        for _ in range(3):
            s.solve_by_CG(eps=1e-30, max_iter=200, check_every=200, fixed_iters=True)
            ms.append(s.ctx.time_ms(1) / 200)
        out[f"cg_{name}_ms_per_iter"] = round(min(ms), 5)
        out[f"cg_{name}_phases_us"] = [round(float(x) / 200 / 1e3, 1) for x in s.ctx.cg_phase_ns()]
print(json.dumps(out), flush=True)
PY
for sg in 0 1024; do FEMCY_OPT_SELL_SIGMA=$sg python /tmp/c10.py C3D10 55 20,23 2>&1 | tail -1 | tee -a gpurun_out/${tag}_c3d10.jsonl; done
FEMCY_OPT_SELL_SIGMA=0 python /tmp/c10.py C3D4 119 20,23 2>&1 | tail -1 | tee -a gpurun_out/${tag}_c3d10.jsonl
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather_p' -c 1 \
    -o gpurun_out/${tag}_gather_p_c3d10 -f python /tmp/c10.py C3D10 55 23 ncu > gpurun_out/${tag}_ncu1.log 2>&1
FEMCY_OPT_SELL_SIGMA=1024 timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather_p' -c 1 \
    -o gpurun_out/${tag}_gather_p_c3d10_sigma -f python /tmp/c10.py C3D10 55 23 ncu > gpurun_out/${tag}_ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather_p' -c 1 \
    -o gpurun_out/${tag}_gather_p_c3d4 -f python /tmp/c10.py C3D4 119 23 ncu > gpurun_out/${tag}_ncu3.log 2>&1
ls -la gpurun_out/${tag}*
