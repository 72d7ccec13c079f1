#!/bin/bash
# 8-GPU call: multi-GPU parity tests, in-process A/B of every PCG switch at N=8, one default bench line with the parity leg
tag=${1:-r2d}; n=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_tests.log 2>&1
echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/${tag}_multi_tests.log
R2_SKIP_BAL=1 R2_SKIP_BENCH=1 bash tools/r2_scaling_ab.sh $tag $n
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29931 \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "bench N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n${n}.json')); print(d['value'], d['cg'], d['parity'])"
tail -5 gpurun_out/${tag}_bench_n${n}.err
