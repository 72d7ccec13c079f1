#!/bin/bash
# persistent vs multi-kernel CG at N = 1, 4, 8 on one box
tag=${1:-r1k}
mkdir -p gpurun_out
run() { # n mode
  n=$1; mode=$2
  if [ "$mode" = "multik" ]; then export FEMCY_CG_MULTIKERNEL=1; else unset FEMCY_CG_MULTIKERNEL; fi
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${tag}_n1_$mode.json 2> gpurun_out/scale_${tag}_n1_$mode.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n${n}_$mode.json 2> gpurun_out/scale_${tag}_n${n}_$mode.err
  fi
}
run 1 persist; run 8 persist; run 8 multik; run 4 persist; run 4 multik
python - <<PY
import json
base=json.load(open("gpurun_out/scale_${tag}_n1_persist.json"))
for n,mode in ((1,'persist'),(4,'persist'),(4,'multik'),(8,'persist'),(8,'multik')):
    try: d=json.load(open(f"gpurun_out/scale_${tag}_n{n}_{mode}.json"))
    except Exception as e: print(n,mode,'failed',e); continue
    print(n, mode, "asm x%.2f  cg it/s %.0f (x%.2f) ms/iter %.4f launches %d" % (d["value"]/base["value"], d["cg"]["value"], d["cg"]["value"]/base["cg"]["value"], d["cg"]["ms_per_iter"], d["gpu_launches"]))
PY
tail -2 gpurun_out/scale_${tag}_n8_persist.err
