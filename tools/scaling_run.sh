#!/bin/bash
# 1 -> 8 GPU strong-scaling run of bench.py on one box (what the driver does at round end).
#   gpurun --gpus 8 -- 'bash tools/scaling_run.sh r1'
tag=${1:-r1}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${tag}_n1.json 2> gpurun_out/scale_${tag}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  fi
done
FEMCY_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
  bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_${tag}_n8_nccl.json 2> gpurun_out/scale_${tag}_n8_nccl.err
python - <<PY
import json, glob
base=None
for n in (1,2,4,8,'8_nccl'):
    try:
        d=json.load(open(f"gpurun_out/scale_${tag}_n{n}.json"))
    except Exception as e:
        print(n, "failed", e); continue
    if n==1: base=d
    print(n, "asm Gelem/s %.3f (x%.2f)  cg it/s %.0f (x%.2f)  ms/iter %.4f  in-loop %s  %s" % (
        d["value"]/1e9, d["value"]/base["value"], d["cg"]["value"], d["cg"]["value"]/base["cg"]["value"],
        d["cg"]["ms_per_iter"], {k: round(v,4) for k,v in d["roofline"]["in_loop_ms"].items()}, d["config"].get("cg_exchange","")[:20]))
PY
tail -3 gpurun_out/scale_${tag}_n8.err
