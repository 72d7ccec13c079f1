"""Quick single-GPU performance probe (development aid, not the benchmark contract -- see bench.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from femcy_b200 import Body, System_of_equations, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
kind = sys.argv[2] if len(sys.argv) > 2 else "C3D4"
reorder = {"lex": False, "morton": True}.get(sys.argv[3] if len(sys.argv) > 3 else "", False)
reps = 5
t0 = time.time()
deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
conn = deck.eSets[kind]
print(f"mesh: {conn.shape[0]} elements, {deck.nodes.shape[0]} nodes, gen {time.time() - t0:.1f}s", flush=True)
t0 = time.time()
body = Body(deck.nodes, conn, deck.ELE)
s = System_of_equations(body, list(deck.materials.values())[0], False, quiet=True, reorder=reorder)
print('element order:', 'z-curve' if s.element_perm is not None else 'as given', flush=True)
out = {"ne": int(conn.shape[0]), "nn": int(deck.nodes.shape[0]), "nnz": s.nnz, "setup_s": time.time() - t0,
       "pattern_ms": s.ctx.time_ms(2)}
import ctypes as C
st = (C.c_int64 * 4)()
s.ctx.call("femcy_pattern_stats", st)
out["nnzb"], out["nslots"], out["nslice"], out["maxw"] = [int(v) for v in st]
print(out, flush=True)
ne = conn.shape[0]
import os
for variant in [1, 2]:
    s.assembly_variant = variant
    ts = []
    for r in range(reps + 2):
        s.assemble_stiffnessMtrx()
        ts.append(s.ctx.time_ms(0))
    ms = float(np.median(ts[2:]))
    bytes_per_elem = {"C3D4": 1360, "C3D10": 7720}[kind]
    out[f"assemble_v{variant}_ms"] = ms
    out[f"assemble_v{variant}_Gelem_s"] = ne / ms / 1e6
    out[f"assemble_v{variant}_algTBs"] = ne * bytes_per_elem / ms / 1e9
    print(f"variant {variant}: {ts}", flush=True)
t0 = time.time()
nb = deck.neumann_bc_info[0]
s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
out["neumann_host_s"] = time.time() - t0
t0 = time.time()
for bc in deck.dirichlet_bc_info:
    s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
s.ctx.sync()
out["dirichlet_s"] = time.time() - t0
iters = 200
for r in range(3):
    s.solve_by_CG(eps=1e-30, max_iter=iters, check_every=iters, fixed_iters=True)
    ms = s.ctx.time_ms(1)
    print(f"cg {iters} iters: {ms:.2f} ms  -> {iters / ms * 1e3:.1f} it/s", flush=True)
N = s.N
spmv_bytes = s.nnz * 12 + N * 20
out["cg_it_s"] = iters / ms * 1e3
out["cg_alg_TBs"] = (spmv_bytes + 11 * N * 8) * iters / ms / 1e9
out["device_GB"] = s.ctx.lib.femcy_device_bytes(s.ctx.h) / 1e9
# converge for real at the reference eps
t0 = time.time()
s.solve_by_CG(eps=1e-3, max_iter=N, check_every=64)
out["cg_to_1e-3_iters"] = s.last_cg_iters
out["cg_to_1e-3_s"] = time.time() - t0
print(json.dumps(out, indent=1))
