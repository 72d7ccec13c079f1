#!/bin/bash
# 1 GPU: the full -m gpu suite, the bench line, ncu --set full of the default kernels (PCG kernel: 4 iterations per launch)
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/${tag}_gpu_tests.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench rc=$?"; cut -c1-3000 gpurun_out/${tag}_bench_n1.json; tail -3 gpurun_out/${tag}_bench_n1.err
FEMCY_OPT_CG_SYM=1 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1_sym.json 2>> gpurun_out/${tag}_bench_n1.err
echo "bench (cg_sym) rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n1_sym.json')); print(d['cg'], d['parity'])"
cat > /tmp/ncu_run.py <<'PY'
import sys
sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen
mode = sys.argv[1]
deck = meshgen.SyntheticDeck("C3D4", n=119, jitter=0.1)
s = System_of_equations(Body(deck.nodes, deck.eSets["C3D4"], deck.ELE), deck.materials["Elastic"], False, quiet=True)
s.assemble_stiffnessMtrx(); s.assemble_stiffnessMtrx()
if mode != "asm":
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    s.solve_by_CG(eps=1e-30, max_iter=4, check_every=4, fixed_iters=True)
s.ctx.sync()
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_elem_geometry|k_assemble_gather' -s 2 -c 2 \
    -o gpurun_out/${tag}_asm_gather -f python /tmp/ncu_run.py asm > gpurun_out/${tag}_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_cg_stream' -c 1 \
    -o gpurun_out/${tag}_cg_stream -f python /tmp/ncu_run.py cg > gpurun_out/${tag}_ncu2.log 2>&1
FEMCY_OPT_CG_SYM=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_cg_stream' -c 1 \
    -o gpurun_out/${tag}_cg_stream_sym -f python /tmp/ncu_run.py cg > gpurun_out/${tag}_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ls -la gpurun_out/${tag}*; tail -2 gpurun_out/${tag}_ncu*.log
