#!/bin/bash
# quick 1-GPU iteration: gated parity of the listed assembly variants + warm timings on cfg 4 / cfg 5
tag=${1:-r2e}; v4=${2:-20,23,24}; v10=${3:-20,23,24}
mkdir -p gpurun_out
FEMCY_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q -x -k "assembly" > gpurun_out/${tag}_exp_tests.log 2>&1
echo "experimental assembly tests rc=$?"; tail -5 gpurun_out/${tag}_exp_tests.log
python /dev/stdin $v4 $v10 <<'PY' 2>&1 | tee gpurun_out/${tag}_timing.log
import sys, os, time, subprocess, numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
for kind, n, vs in (("C3D4", 119, sys.argv[1]), ("C3D10", 55, sys.argv[2])):
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    s = System_of_equations(Body(deck.nodes, deck.eSets[kind], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
    for v in [int(x) for x in vs.split(",")]:
        s.assembly_variant = v
        ts = []
        for _ in range(12):
            s.assemble_stiffnessMtrx(); s.ctx.sync(); ts.append(s.ctx.time_ms(0))
        print(kind, "variant", v, "first", round(ts[0], 3), "median", round(float(np.median(ts[2:])), 4), "min", round(min(ts), 4), flush=True)
    s.close()
PY
