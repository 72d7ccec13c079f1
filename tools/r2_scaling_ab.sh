#!/bin/bash
# Round-2: multi-GPU A/B of the PCG variants on one box (N GPUs).
#   1. tools/scaling_ab.py: ONE launch, system built once, every switch combination timed in-process (~0.2 s per mode);
#      then the same with the speed-weighted partition for the main modes;
#   2. full bench.py runs (~30-60 s each at N=8, charged N x) only for: the default, the best mode of step 1, and any modes
#      named on the command line.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scaling_ab.sh r2b 8'
#   ... 'bash tools/r2_scaling_ab.sh r2c 8 sr5 persist5 sr_late_b4'
tag=${1:-r2b}; n=${2:-8}
if [ $# -gt 2 ]; then shift 2; modes="$*"; else modes=""; fi
mkdir -p gpurun_out
envs() { case $1 in
  persist)         echo "FEMCY_CG_PERSISTENT=1";;
  persist5)        echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_MINB=5";;
  persist_late)    echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_LATE_FENCE=1";;
  persist_late_fb) echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  sr)              echo "FEMCY_CG_VARIANT=sr";;
  sr5)             echo "FEMCY_CG_VARIANT=sr FEMCY_CG_MINB=5";;
  sr_late)         echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1";;
  sr_late_fb)      echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  sr_late_b4)      echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_BLOCKS_PER_SM=4";;
  multik)          echo "FEMCY_CG_MULTIKERNEL=1";;
  persist_fb)      echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_FOLD_BARRIER=1";;
  persist_b4)      echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_BLOCKS_PER_SM=4";;
  persist_b2)      echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_BLOCKS_PER_SM=2";;
  sr_fb)           echo "FEMCY_CG_VARIANT=sr FEMCY_CG_FOLD_BARRIER=1";;
  sr_late_fb5)     echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 FEMCY_CG_MINB=5";;
  sr_late_fb_b4)   echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 FEMCY_CG_BLOCKS_PER_SM=4";;
  sr_late_fb_b2)   echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 FEMCY_CG_BLOCKS_PER_SM=2";;
  sym)             echo "FEMCY_CG_SYM=1";;
  sym_late)        echo "FEMCY_CG_SYM=1 FEMCY_CG_LATE_FENCE=1";;
  sym_late_fb)     echo "FEMCY_CG_SYM=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  sym_late_fb_b2)  echo "FEMCY_CG_SYM=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 FEMCY_CG_BLOCKS_PER_SM=2";;
  sr_sym)          echo "FEMCY_CG_VARIANT=sr FEMCY_CG_SYM=1";;
  sr_sym_late_fb)  echo "FEMCY_CG_VARIANT=sr FEMCY_CG_SYM=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  sym_l2m)         echo "FEMCY_CG_SYM=1 FEMCY_CG_L2_PERSIST=2";;
  sr_sym_l2m)      echo "FEMCY_CG_VARIANT=sr FEMCY_CG_SYM=1 FEMCY_CG_L2_PERSIST=2";;
  sr_sym_late_fb_l2m) echo "FEMCY_CG_VARIANT=sr FEMCY_CG_SYM=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 FEMCY_CG_L2_PERSIST=2";;
  persist_l2)      echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_L2_PERSIST=1";;
  persist_l2m)     echo "FEMCY_CG_PERSISTENT=1 FEMCY_CG_L2_PERSIST=2";;
  sym_l2)          echo "FEMCY_CG_SYM=1 FEMCY_CG_L2_PERSIST=1";;
  persist_bal)     echo "FEMCY_CG_PERSISTENT=1";;
  sr_late_fb_bal)  echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  *)               echo "";;
esac; }
# gated multi-GPU parity of the single-reduction kernel first (2 ranks of the box)
FEMCY_EXPERIMENTAL=1 FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_sr_tests.log 2>&1
echo "multi-gpu tests under sr+late+fb rc=$?"; tail -3 gpurun_out/${tag}_multi_sr_tests.log
FEMCY_EXPERIMENTAL=1 FEMCY_CG_SYM=1 FEMCY_CG_PERSISTENT=1 timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_sym_tests.log 2>&1
echo "multi-gpu tests under sym rc=$?"; tail -3 gpurun_out/${tag}_multi_sym_tests.log
port=$((29800+RANDOM%50))
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    tools/scaling_ab.py --tag ${tag}_n${n} > gpurun_out/${tag}_n${n}_scaling_ab.log 2>&1
echo "in-process A/B rc=$?"; grep '"what": "cg"' gpurun_out/${tag}_n${n}_scaling_ab.log | cut -c1-260
[ -z "$R2_SKIP_BAL" ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((port+1)) \
    tools/scaling_ab.py --tag ${tag}_n${n}_bal --balance measured --modes persist multik sr sr_late_fb sym > gpurun_out/${tag}_n${n}_scaling_ab_bal.log 2>&1
echo "in-process A/B (measured balance) rc=$?"; grep '"what": "cg"' gpurun_out/${tag}_n${n}_scaling_ab_bal.log | cut -c1-260
best=$(python - gpurun_out/${tag}_n${n}_scaling_ab.jsonl <<'PY'
import json, sys
best = None
try:
    for line in open(sys.argv[1]):
        d = json.loads(line)
        if d.get("what") == "cg" and "ms_per_iter_best" in d and d["mode"] not in ("default", "multik_nccl"):
            if best is None or d["ms_per_iter_best"] < best[0]:
                best = (d["ms_per_iter_best"], d["mode"])
except Exception:
    pass
print(best[1] if best else "persist")
PY
)
echo "best in-process mode: $best"
[ -n "$R2_SKIP_BENCH" ] && exit 0
modes="persist $best $modes"
modes=$(echo $modes | tr ' ' '\n' | awk '!seen[$0]++' | tr '\n' ' ')
for m in $modes; do
  extra=""; case $m in *_bal) extra="--balance measured";; esac    # rows proportional to each GPU's measured copy rate
  env $(envs $m) timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900+RANDOM%50)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline $extra > gpurun_out/${tag}_n${n}_$m.json 2> gpurun_out/${tag}_n${n}_$m.err
  echo "$m rc=$?"
done
python - $tag $n $modes <<'PY'
import json, sys
tag, n, modes = sys.argv[1], sys.argv[2], sys.argv[3:]
for mode in modes:
    try: d = json.load(open(f"gpurun_out/{tag}_n{n}_{mode}.json"))
    except Exception as e: print(mode, "failed", e); continue
    print("%-16s asm %.2f G/s  cg it/s %.0f  ms/iter %.4f  launches %d" % (mode, d["value"]/1e9, d["cg"]["value"], d["cg"]["ms_per_iter"], d["gpu_launches"]))
PY
