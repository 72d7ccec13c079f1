#!/bin/bash
# NOT RUN YET (the GPU budget of round 2 ended first): the first GPU call of a next round.
# 1 GPU: timing + ncu of what round 2's second session added and could only check for parity --
#   * the device builders (k_facet_keys / k_facet_unique, k_node_elem_*, k_neumann, k_part_*),
#   * the consistent-tangent assembly (k_assemble_scatter_ct) on cfg 4 and cfg 5 sizes,
#   * a mesh of several sections at cfg-4 size (two materials): scatter per section vs the single-section gather,
#   * the nlgeom configurations with --tangent consistent.
tag=${1:-r3n}
mkdir -p gpurun_out
cat > /tmp/r3_run.py <<'PY'
import ctypes as C, json, sys, time
import numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
from femcy_b200.body import SectionedBody
from femcy_b200.material_zoo import LinearIsotropic
from femcy_b200.partition import Partition
out = {}
deck = meshgen.SyntheticDeck("C3D4", n=119, jitter=0.1)
conn, mat = deck.eSets["C3D4"], deck.materials["Elastic"]
s = System_of_equations(Body(deck.nodes, conn, deck.ELE), mat, True, quiet=True)
def t_asm(reps=5):
    ts = []
    for _ in range(reps):
        s.assemble_stiffnessMtrx(); s.ctx.sync(); ts.append(s.ctx.time_ms(0))
    return float(np.median(ts))
out["asm_gather_ms"] = t_asm()
s.set_tangent("consistent"); out["asm_consistent_tangent_ms"] = t_asm(); s.set_tangent("reference")
n = C.c_int64(0)
for k in range(3):
    t0 = time.time(); s.ctx.call("femcy_boundary_facets", C.byref(n)); s.ctx.sync(); out["boundary_facets_ms"] = (time.time() - t0) * 1e3
nb = deck.neumann_bc_info[0]
for k in range(3):
    t0 = time.time(); s.neumannBC(nb["face_set"], nb["traction"], nb["direction"]); s.ctx.sync(); out["neumann_ms"] = (time.time() - t0) * 1e3
t0 = time.time(); p = Partition(deck.nodes, conn, 3, 8, device=0); out["partition_rank3_of_8_ms"] = (time.time() - t0) * 1e3
s.close()
# two materials split at x = 0.5: sections (scatter-add per section)
cx = deck.nodes[conn, 0].mean(axis=1)
body = SectionedBody(deck.nodes, [(conn[cx < 0.5], deck.ELE, mat), (conn[cx >= 0.5], deck.ELE, LinearIsotropic(7.0e4, 0.33))])
s = System_of_equations(body, None, False, quiet=True)
out["asm_two_sections_scatter_ms"] = t_asm()
s.close()
print(json.dumps(out))
PY
python /tmp/r3_run.py > gpurun_out/${tag}_timings.json 2> gpurun_out/${tag}_timings.err; echo "timings rc=$?"; cat gpurun_out/${tag}_timings.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_scatter_ct|k_facet_unique|k_neumann|k_part_touch' -c 6 \
    -o gpurun_out/${tag}_new_kernels -f python /tmp/r3_run.py > gpurun_out/${tag}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_new_kernels.ncu-rep > gpurun_out/${tag}_ncu_new_kernels.md 2>/dev/null
python tools/run_configs.py gpurun_out/${tag}_configs_consistent.json --skip-big --tangent consistent > gpurun_out/${tag}_configs.log 2>&1
tail -5 gpurun_out/${tag}_configs.log
