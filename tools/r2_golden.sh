#!/bin/bash
# 1 GPU: (re)write the bench parity fixture (solution samples of the eps = 1e-8 solve of cfg 4 from dof = 0), check it from a
# second process with a different step count, then the default bench line
tag=${1:-r2w}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 2 --warmup 3 --cg-iters 100 --no-cpu-baseline --write-parity-golden > gpurun_out/${tag}_golden_run.json 2> gpurun_out/${tag}.err
echo "golden rc=$?"; cp tests/golden/bench_solution_samples_n119.npz gpurun_out/
python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2>> gpurun_out/${tag}.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n1.json')); print(d['value'], d['cg'], d['roofline_assembly']['frac'], d['roofline']['frac'], d['roofline']['frac_dram'], d['e2e']); print(d['parity'])"
STRESS_CHUNKS=50,7 python tools/parity_stress.py 2>&1 | grep '^{'
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
tail -3 gpurun_out/${tag}.err
