#!/bin/bash
tag=${1:-r2y}; n=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29941 \
    bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "bench N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n${n}.json')); print(d['value'], d['cg'], d['roofline']['phase_us_per_iteration'], d['roofline']['ms_per_launch'], d['e2e']['value'], d['e2e']['cg_value']); print(d['parity'])"
tail -2 gpurun_out/${tag}_bench_n${n}.err
