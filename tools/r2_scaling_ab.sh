#!/bin/bash
# Round-2: multi-GPU A/B of the PCG variants on one box (N GPUs):  reference recurrence (persistent, default from 4
# ranks) vs single-reduction persistent (FEMCY_CG_VARIANT=sr) vs three-kernel graph.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_scaling_ab.sh r2b 8'
tag=${1:-r2b}; n=${2:-8}
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29900+RANDOM%50)) \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_n${n}_$name.json 2> gpurun_out/${tag}_n${n}_$name.err
  echo "$name rc=$?"
}
# gated multi-GPU parity of the sr variant first (2 ranks of the box)
FEMCY_EXPERIMENTAL=1 FEMCY_CG_VARIANT=sr timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/${tag}_multi_sr_tests.log 2>&1
echo "multi-gpu tests under sr rc=$?"; tail -3 gpurun_out/${tag}_multi_sr_tests.log
run persist FEMCY_CG_PERSISTENT=1
run sr FEMCY_CG_VARIANT=sr
run sr5 FEMCY_CG_VARIANT=sr FEMCY_CG_MINB=5
run persist5 FEMCY_CG_PERSISTENT=1 FEMCY_CG_MINB=5
run sr_late FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1
run sr_late_b4 FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_BLOCKS_PER_SM=4
run persist_late FEMCY_CG_PERSISTENT=1 FEMCY_CG_LATE_FENCE=1
run sr_late_fb FEMCY_CG_VARIANT=sr FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1
run persist_late_fb FEMCY_CG_PERSISTENT=1 FEMCY_CG_LATE_FENCE=1 FEMCY_CG_FOLD_BARRIER=1
run multik FEMCY_CG_MULTIKERNEL=1
python - <<PY
import json
for mode in ("persist", "persist5", "persist_late", "persist_late_fb", "sr", "sr5", "sr_late", "sr_late_b4", "sr_late_fb", "multik"):
    try: d = json.load(open("gpurun_out/${tag}_n${n}_%s.json" % mode))
    except Exception as e: print(mode, "failed", e); continue
    print(mode, "asm %.2f G/s  cg it/s %.0f  ms/iter %.4f  launches %d" % (d["value"]/1e9, d["cg"]["value"], d["cg"]["ms_per_iter"], d["gpu_launches"]))
PY
