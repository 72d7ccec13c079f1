"""Per-rank phase clock of the persistent PCG kernel (femcy_cg_phase_ns) on the cfg-4 system: where an iteration's time goes
on 1..8 GPUs -- SpMV loop, grid barrier + fold, cross-rank exchange, vector updates.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29871 \
        tools/cg_phases.py --tag r2e [--modes default sym] [--balance measured]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
NAMES = ["spmv", "barrier1", "exchange1", "update_xr", "barrier2", "exchange2", "update_d_barrier3"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2e")
    ap.add_argument("--n", type=int, default=119)
    ap.add_argument("--kind", default="C3D4")
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--balance", default="equal")
    ap.add_argument("--modes", nargs="*", default=["persist", "stream0"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    MODES = {"multik": {"cg_kernel": 1}, "persist": {"cg_kernel": 2}, "persist_sym": {"cg_kernel": 2, "cg_sym": 1}}
    for k in range(4):
        MODES[f"stream{k}"] = {"cg_kernel": 3, "cg_stream_cfg": k}
        MODES[f"stream{k}_sym"] = {"cg_kernel": 3, "cg_stream_cfg": k, "cg_sym": 1}
    MODES["default"] = {}
    BASE = {"cg_kernel": 0, "cg_sym": 0, "cg_stream_cfg": 0}
    from femcy_b200._lib import as_d, as_i32
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    deck, system, rhs, bcs, ne_global, nn_global, part = bench.build_problem(args.n, rank, world, local, balance=args.balance)
    system.rhs.from_numpy(rhs)
    system.assemble_stiffnessMtrx()
    system.ctx.call("femcy_dirichlet_linear", as_i32(bcs[0]), as_i32(bcs[1]), as_d(bcs[2]), len(bcs[0]))
    os.makedirs("gpurun_out", exist_ok=True)
    for mode in args.modes:
        for k, v in {**BASE, **MODES[mode]}.items():
            system.ctx.set_option(k, v)
        best = None
        for rep in range(3):
            if world > 1:
                dist.barrier()
            system.solve_by_CG(eps=1e-30, max_iter=args.iters, check_every=args.iters, fixed_iters=True)
            ms = system.ctx.time_ms(1)
            ph = system.ctx.cg_phase_ns() / args.iters / 1e3      # us per iteration
            if best is None or ms < best[0]:
                best = (ms, ph)
        rec = [best[0] / args.iters * 1e3] + [float(v) for v in best[1]] + [float(system.N_own), float(system.nnz)]
        if world > 1:
            t = torch.tensor(rec, dtype=torch.float64)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            allr = [a.tolist() for a in allr]
        else:
            allr = [rec]
        if rank == 0:
            out = {"mode": mode, "n_gpus": world, "iters": args.iters, "balance": args.balance,
                   "us_per_iter_max_over_ranks": max(a[0] for a in allr),
                   "per_rank_us": [{"rank": r, "event_us": round(a[0], 2), **{NAMES[i]: round(a[1 + i], 2) for i in range(7)},
                                    "dofs_own": int(a[8]), "nnz": int(a[9])} for r, a in enumerate(allr)]}
            line = json.dumps(out)
            print(line, flush=True)
            with open(f"gpurun_out/{args.tag}_cg_phases_n{world}.jsonl", "a") as fh:
                fh.write(line + "\n")
    system.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
