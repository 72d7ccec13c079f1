#!/bin/bash
# 1 GPU: the whole -m gpu suite, then every BASELINE.json configuration end to end (default Jacobi path, then two-level)
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/${tag}_gpu_tests.log
timeout 1500 python tools/run_configs.py gpurun_out/${tag}_configs.json 2>&1 | cut -c1-600
FEMCY_OPT_CG_PRECOND=1 timeout 900 python tools/run_configs.py gpurun_out/${tag}_configs_two_level.json 2>&1 | cut -c1-600
