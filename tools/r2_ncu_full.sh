#!/bin/bash
# Round-2 second call (1 GPU): ncu --set full of the winners of tools/r2_first_call.sh + a clock-sampled warm loop.
tag=${1:-r2b}
mkdir -p gpurun_out
cat > /tmp/ncu_asm.py <<'PY'
import sys, os, time, subprocess, numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
kind, n = sys.argv[1], int(sys.argv[2])
deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
s = System_of_equations(Body(deck.nodes, deck.eSets[kind], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
mode = sys.argv[5] if len(sys.argv) > 5 else "asm"
if mode == "warm":
    # clocks while a warm loop runs (why do event-timed warm loops differ from ncu's per-launch times?)
    for v in [int(x) for x in sys.argv[3].split(",")]:
        s.assembly_variant = v
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
        ts = []
        t0 = time.time()
        while time.time() - t0 < 2.0:
            s.assemble_stiffnessMtrx(); s.ctx.sync(); ts.append(s.ctx.time_ms(0))
        p.terminate(); out = p.stdout.read().strip().splitlines()
        print(kind, "variant", v, "calls", len(ts), "first5", [round(t, 3) for t in ts[:5]], "median", round(float(np.median(ts)), 4), "min", round(min(ts), 4), "smi", out[::4][:6], flush=True)
    sys.exit(0)
for v in [int(x) for x in sys.argv[3].split(",")]:
    s.assembly_variant = v
    for _ in range(reps):
        s.assemble_stiffnessMtrx()
if mode == "cg":
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    s.solve_by_CG(eps=1e-30, max_iter=4, check_every=4, fixed_iters=True)
s.ctx.sync()
PY
python /tmp/ncu_asm.py C3D4 119 5,20,21,14 1 warm > gpurun_out/${tag}_warm.log 2>&1
python /tmp/ncu_asm.py C3D10 55 20,10,1 1 warm >> gpurun_out/${tag}_warm.log 2>&1
cat gpurun_out/${tag}_warm.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_elem_geometry|k_assemble_gather' -c 6 \
    -o gpurun_out/${tag}_asm_c3d4 -f python /tmp/ncu_asm.py C3D4 119 5,20,21 1 > gpurun_out/${tag}_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_elem_geometry|k_assemble_gather' -c 2 \
    -o gpurun_out/${tag}_asm_c3d10 -f python /tmp/ncu_asm.py C3D10 55 20 1 > gpurun_out/${tag}_ncu2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_cg_persistent' -c 1 \
    -o gpurun_out/${tag}_cg_c3d4 -f python /tmp/ncu_asm.py C3D4 119 20 1 cg > gpurun_out/${tag}_ncu3.log 2>&1
FEMCY_CG_SYM=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_cg_persistent' -c 1 \
    -o gpurun_out/${tag}_cg_c3d4_sym -f python /tmp/ncu_asm.py C3D4 119 20 1 cg > gpurun_out/${tag}_ncu4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_cg_persistent' -c 1 \
    -o gpurun_out/${tag}_cg_c3d10 -f python /tmp/ncu_asm.py C3D10 55 20 1 cg > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out | tail; tail -3 gpurun_out/${tag}_ncu*.log
