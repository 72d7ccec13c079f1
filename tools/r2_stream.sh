#!/bin/bash
# 1 GPU: parity of the default (streaming) PCG kernel, then persistent-vs-streaming phase clocks at full and 1/8 size
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q -x > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_gpu_tests.log
bash tools/r2_phases.sh $tag 1 --modes persist stream0 stream1 stream2 stream3 persist_sym stream0_sym stream1_sym stream2_sym
bash tools/r2_phases.sh $tag 1 --n 60 --modes persist stream0 stream1 stream2 stream3 persist_sym stream0_sym stream1_sym stream2_sym
