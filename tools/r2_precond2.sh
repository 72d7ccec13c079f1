#!/bin/bash
tag=${1:-r2o}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "two_level" > gpurun_out/${tag}_tests.log 2>&1
echo "two-level tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log
python /dev/stdin <<'PY' 2>&1 | tee gpurun_out/${tag}_precond.log
import sys, time, json, os
sys.path.insert(0, "."); sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import run_configs as rc
os.environ["FEMCY_OPT_CG_PRECOND"] = "1"
t = time.time()
out = rc.run_deck(rc.twist_deck(n_inc=2))
print("cfg 3' twist plate 104544 C3D4, 2 increments, two-level", json.dumps({k: out[k] for k in ("solve_s", "increments", "cg_iterations_total", "max_abs_u")}), "wall", round(time.time() - t, 1), flush=True)
PY
