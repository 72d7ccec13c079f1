#!/bin/bash
# N-GPU call: multi-GPU parity tests, the bench line at N (parity leg included), PCG phase clocks, cfg 5 on 1 and N GPUs
tag=${1:-r2k}; n=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_tests.log 2>&1
echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/${tag}_multi_tests.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29931 bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "bench N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n${n}.json')); print(d['value'], d['cg'], d['roofline']['phase_us_per_iteration'], d['parity'])"
if [ -n "$R2_SYM" ]; then
FEMCY_OPT_CG_SYM=1 run 29932 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n${n}_sym.json 2>> gpurun_out/${tag}_bench_n${n}.err
echo "bench N=$n cg_sym rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n${n}_sym.json')); print(d['value'], d['cg'], d['parity']['parity_ok'])"
fi
run 29933 tools/cg_phases.py --tag $tag --modes stream0 stream0_sym 2>&1 | grep '^{' | cut -c1-1200
[ -f profiles/r2k_cfg5_n1.json ] || python tools/cfg5_multi.py --tag $tag 2>&1 | grep '^{' | cut -c1-1500
run 29934 tools/cfg5_multi.py --tag $tag 2>&1 | grep '^{' | cut -c1-1500
tail -3 gpurun_out/${tag}_bench_n${n}.err
