#!/bin/bash
tag=${1:-r1m}
mkdir -p gpurun_out
run() { n=$1; mode=$2
  if [ "$mode" = "persist" ]; then export FEMCY_CG_PERSISTENT=1; else unset FEMCY_CG_PERSISTENT; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n${n}_$mode.json 2> gpurun_out/scale_${tag}_n${n}_$mode.err
}
run 8 persist; run 8 multik; run 4 persist
python - <<PY
import json
for n,mode in ((8,'persist'),(8,'multik'),(4,'persist')):
    try: d=json.load(open(f"gpurun_out/scale_${tag}_n{n}_{mode}.json"))
    except Exception as e: print(n,mode,'failed',e); continue
    print(n, mode, "asm %.2f G/s  cg it/s %.0f ms/iter %.4f launches %d" % (d["value"]/1e9, d["cg"]["value"], d["cg"]["ms_per_iter"], d["gpu_launches"]))
PY
