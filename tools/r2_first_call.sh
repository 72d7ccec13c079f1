#!/bin/bash
# Round-2 opener: ONE gpurun call (1 GPU, ~10 min) that
#   1. runs the gated hardware tests of the kernels written without GPU access at the end of round 1,
#   2. A/Bs every assembly + PCG variant on both headline meshes (tools/ab_variants.py),
#   3. takes an ncu launch list + one --set full capture of the new assembly kernels.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh r2a'
tag=${1:-r2a}
mkdir -p gpurun_out
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 )) s] $1" | tee -a gpurun_out/${tag}_stages.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${tag}_smi.txt 2>&1
stamp "gated tests"
FEMCY_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q > gpurun_out/${tag}_exp_tests.log 2>&1
echo "experimental tests rc=$?" | tee -a gpurun_out/${tag}_exp_tests.log
tail -5 gpurun_out/${tag}_exp_tests.log
stamp "quick A/B"
timeout 300 python tools/quick_ab.py ${tag} > gpurun_out/${tag}_quick_ab.log 2>&1; tail -40 gpurun_out/${tag}_quick_ab.log
stamp "full A/B"
[ -n "$R2_FULL_AB" ] && timeout 500 python tools/ab_variants.py C3D4 119 C3D10 55 > gpurun_out/${tag}_ab.jsonl 2> gpurun_out/${tag}_ab.err
echo "ab rc=$?"; cat gpurun_out/${tag}_ab.jsonl | cut -c1-3000
stamp "ncu passes"
# ncu: per-launch durations of one assembly call per variant (small loop), then a full capture of the rows kernels
cat > /tmp/ncu_asm.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
kind, n = sys.argv[1], int(sys.argv[2])
deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
s = System_of_equations(Body(deck.nodes, deck.eSets[kind], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
for v in [int(x) for x in sys.argv[3].split(",")]:
    s.assembly_variant = v
    for _ in range(reps):
        s.assemble_stiffnessMtrx()
s.ctx.sync()
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c3d4.csv \
    -k regex:'k_assemble|k_elem_geometry' python /tmp/ncu_asm.py C3D4 119 1,2,5,11,6,7,8,16,17,9,10,20,12,13,14,21,18,22 > gpurun_out/${tag}_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c3d10.csv \
    -k regex:'k_assemble|k_elem_geometry|k_dsdx' python /tmp/ncu_asm.py C3D10 55 1,2,6,7,8,9,10,20,15,19 > gpurun_out/${tag}_ncu2.log 2>&1
[ -n "$R2_FULL_NCU" ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_rows|k_elem_geometry|k_assemble_gather|k_assemble_tile' -c 16 \
    -o gpurun_out/${tag}_rows_c3d4 -f python /tmp/ncu_asm.py C3D4 119 5,17,10,20,14,21,22 1 > gpurun_out/${tag}_ncu3.log 2>&1
[ -n "$R2_FULL_NCU" ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_rows|k_assemble_scatter_warp|k_assemble_scatter_pairs|k_elem_geometry4|k_assemble_tile' -c 7 \
    -o gpurun_out/${tag}_asm_c3d10 -f python /tmp/ncu_asm.py C3D10 55 1,7,15,19 1 > gpurun_out/${tag}_ncu4.log 2>&1
stamp "done"
ls -la gpurun_out | tail -20
python tools/pick_defaults.py gpurun_out/${tag}_quick_ab.jsonl gpurun_out/${tag}_ab.jsonl 2>/dev/null | tee gpurun_out/${tag}_summary.txt
for r in gpurun_out/${tag}_*.ncu-rep; do python tools/ncu_summary.py $r ${r%.ncu-rep}_summary.md > /dev/null 2>&1; done
