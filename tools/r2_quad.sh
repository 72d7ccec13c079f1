#!/bin/bash
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -x -k "assembly or gather or sigma" > gpurun_out/${tag}_tests.log 2>&1
echo "assembly tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
python /dev/stdin <<'PY' 2>&1 | tee gpurun_out/${tag}_timing.log
import sys, os, numpy as np
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
for sigma in ("0", "-1"):
    os.environ["FEMCY_OPT_SELL_SIGMA"] = sigma
    deck = meshgen.SyntheticDeck("C3D10", n=55, jitter=0.0)
    s = System_of_equations(Body(deck.nodes, deck.eSets["C3D10"], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
    for v in (3, 2, 1):
        s.assembly_variant = v
        ts = []
        for _ in range(10):
            s.assemble_stiffnessMtrx(); s.ctx.sync(); ts.append(s.ctx.time_ms(0))
        print("C3D10 sigma", sigma, "variant", v, "median ms", round(float(np.median(ts[2:])), 4), flush=True)
    s.close()
PY
cat > /tmp/ncu_q.py <<'PY'
import sys
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
deck = meshgen.SyntheticDeck("C3D10", n=55, jitter=0.0)
s = System_of_equations(Body(deck.nodes, deck.eSets["C3D10"], deck.ELE), list(deck.materials.values())[0], False, quiet=True)
s.assemble_stiffnessMtrx(); s.assemble_stiffnessMtrx(); s.ctx.sync()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_gather_q' -s 1 -c 1 \
    -o gpurun_out/${tag}_gather_q_c3d10 -f python /tmp/ncu_q.py > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}*
