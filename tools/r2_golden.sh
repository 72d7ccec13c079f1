#!/bin/bash
# 1 GPU: write the bench parity fixture (solution samples of the eps = 1e-8 solve of cfg 4) and run the default bench line
tag=${1:-r2c}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 3 --write-parity-golden > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/${tag}_bench_n1.json; tail -3 gpurun_out/${tag}_bench_n1.err
cp tests/golden/bench_solution_samples_n119.npz gpurun_out/
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1_check.json 2>> gpurun_out/${tag}_bench_n1.err
echo "bench (check against the fixture) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n1_check.json')); print(d['parity'])"
