#!/bin/bash
n=${1:-4}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 tools/parity_stress.py 2>&1 | grep '^{'; }
export STRESS_CHUNKS=50,50
STRESS_LABEL=default run 29891
STRESS_LABEL=sigma0 FEMCY_OPT_SELL_SIGMA=0 run 29892
STRESS_LABEL=legacy_persistent FEMCY_OPT_CG_KERNEL=2 run 29893
STRESS_LABEL=three_kernel FEMCY_OPT_CG_KERNEL=1 run 29894
STRESS_LABEL=asm_scatter FEMCY_OPT_ASSEMBLY_VARIANT=1 run 29895
STRESS_LABEL=asm_thread_gather FEMCY_OPT_ASSEMBLY_VARIANT=3 run 29896
