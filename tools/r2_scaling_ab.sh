#!/bin/bash
# Round-2: multi-GPU A/B of the PCG variants on one box (N GPUs).  Each run is a full bench.py (~30 s at N=8), so the
# default list is the five that decide the defaults; name others explicitly.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scaling_ab.sh r2b 8'
#   ... 'bash tools/r2_scaling_ab.sh r2c 8 sr5 persist5 sr_late_b4'
tag=${1:-r2b}; n=${2:-8}
if [ $# -gt 2 ]; then shift 2; modes="$*"; else modes="persist persist_late_fb sr sr_late_fb multik"; fi
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
  persist_bal)     echo "FEMCY_CG_PERSISTENT=1";;
  sr_late_fb_bal)  echo "FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1";;
  *)               echo "";;
esac; }
# gated multi-GPU parity of the single-reduction kernel first (2 ranks of the box)
FEMCY_EXPERIMENTAL=1 FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1 timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_sr_tests.log 2>&1
echo "multi-gpu tests under sr+late+fb rc=$?"; tail -3 gpurun_out/${tag}_multi_sr_tests.log
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
