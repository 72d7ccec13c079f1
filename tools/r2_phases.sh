#!/bin/bash
tag=${1:-r2e}; n=${2:-2}; shift 2
mkdir -p gpurun_out
if [ "$n" = "1" ]; then python tools/cg_phases.py --tag $tag "$@" 2>&1 | grep "^{"
else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29877 tools/cg_phases.py --tag $tag "$@" 2>&1 | grep '^{' ; fi
