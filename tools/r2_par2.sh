#!/bin/bash
tag=${1:-r2t}; n=${2:-2}
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for rep in 1 2; do
run 2993$rep bench.py --gpus $n --steps 2 --warmup 3 --cg-iters 100 > gpurun_out/${tag}_a$rep.json 2> gpurun_out/${tag}_a$rep.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_a$rep.json')); print('stream rep $rep', d['cg']['value'], d['parity'])"
done
FEMCY_OPT_CG_KERNEL=2 run 29935 bench.py --gpus $n --steps 2 --warmup 3 --cg-iters 100 > gpurun_out/${tag}_b.json 2> gpurun_out/${tag}_b.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_b.json')); print('legacy persistent', d['cg']['value'], d['parity'])"
FEMCY_OPT_CG_KERNEL=1 run 29936 bench.py --gpus $n --steps 2 --warmup 3 --cg-iters 100 > gpurun_out/${tag}_c.json 2> gpurun_out/${tag}_c.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c.json')); print('three-kernel', d['cg']['value'], d['parity'])"
