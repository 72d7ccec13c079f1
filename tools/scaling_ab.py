"""In-process multi-GPU A/B of the PCG variants: ONE torchrun launch builds the partitioned cfg-4 system once per rank
and then times every variant on it (the switches are environment variables that femcy_cg_solve reads at every call), so
a mode costs ~0.2 s of box time instead of a full bench.py run (~30-60 s of setup each, charged N x on an N-GPU box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29871 \
        tools/scaling_ab.py --tag r2b [--n 119] [--iters 500] [--modes persist sr ...]

One JSON line per mode on rank 0 (stdout and gpurun_out/<tag>_scaling_ab.jsonl, flushed at once): ms per iteration =
library-internal CUDA-event time of the solve (max over ranks, best and median of `--reps` solves after one warm-up) and
the largest difference of the iterate after `--iters` iterations from the first mode's (all modes run the same
recurrence or, for `sr*`, an algebraically equivalent one)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ("FEMCY_CG_PERSISTENT", "FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_MINB", "FEMCY_CG_LATE_FENCE",
        "FEMCY_CG_FOLD_BARRIER", "FEMCY_CG_BLOCKS_PER_SM", "FEMCY_NO_P2P", "FEMCY_CG_SYM", "FEMCY_CG_L2_PERSIST")
MODES = {
    "default": {},
    "persist": {"FEMCY_CG_PERSISTENT": "1"},
    "persist5": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_MINB": "5"},
    "persist_late": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_LATE_FENCE": "1"},
    "persist_fb": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_FOLD_BARRIER": "1"},
    "persist_late_fb": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1"},
    "persist_b4": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_BLOCKS_PER_SM": "4"},
    "persist_b2": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_BLOCKS_PER_SM": "2"},
    "sr": {"FEMCY_CG_VARIANT": "sr"},
    "sr5": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_MINB": "5"},
    "sr_late": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1"},
    "sr_fb": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_FOLD_BARRIER": "1"},
    "sr_late_fb": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1"},
    "sr_late_fb5": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1", "FEMCY_CG_MINB": "5"},
    "sr_late_b4": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_BLOCKS_PER_SM": "4"},
    "sr_late_fb_b4": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1", "FEMCY_CG_BLOCKS_PER_SM": "4"},
    "sr_late_fb_b2": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1", "FEMCY_CG_BLOCKS_PER_SM": "2"},
    "sym": {"FEMCY_CG_SYM": "1"},
    "sym_late": {"FEMCY_CG_SYM": "1", "FEMCY_CG_LATE_FENCE": "1"},
    "sym_late_fb": {"FEMCY_CG_SYM": "1", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1"},
    "sym_late_fb_b2": {"FEMCY_CG_SYM": "1", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1", "FEMCY_CG_BLOCKS_PER_SM": "2"},
    "sr_sym": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1"},
    "sr_sym_late_fb": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1"},
    "persist_l2": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_L2_PERSIST": "1"},
    "sym_l2": {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "1"},
    "persist_l2m": {"FEMCY_CG_PERSISTENT": "1", "FEMCY_CG_L2_PERSIST": "2"},
    "sym_l2m": {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "2"},
    "sr_sym_l2m": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "2"},
    "sr_sym_late_fb_l2m": {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1", "FEMCY_CG_LATE_FENCE": "1", "FEMCY_CG_FOLD_BARRIER": "1", "FEMCY_CG_L2_PERSIST": "2"},
    "multik": {"FEMCY_CG_MULTIKERNEL": "1"},
    "multik_nccl": {"FEMCY_CG_MULTIKERNEL": "1", "FEMCY_NO_P2P": "1"},      # halo + reductions through NCCL inside the loop
}
DEFAULT_ORDER = ["default", "persist", "multik", "sr", "persist_late", "sr_late", "persist_fb", "sr_fb", "persist_late_fb",
                 "sr_late_fb", "persist5", "sr5", "sr_late_fb5", "persist_b4", "sr_late_b4", "sr_late_fb_b4", "persist_b2",
                 "sr_late_fb_b2", "sym", "sym_late", "sym_late_fb", "sym_late_fb_b2", "sr_sym", "sr_sym_late_fb", "persist_l2", "sym_l2", "persist_l2m", "sym_l2m", "sr_sym_l2m", "sr_sym_late_fb_l2m", "multik_nccl"]


def run_modes(system, modes, iters, reps, allmax, emit, barrier):
    """time every mode on `system` (assembled, boundary conditions applied); `allmax(v)` = max of a float over the ranks."""
    xref = None
    base = {k: os.environ.get(k) for k in KEYS}        # e.g. FEMCY_NO_P2P=1 set by the partition when peer access is missing

    def reset():
        for k, v in base.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    for name in modes:
        reset()
        os.environ.update(MODES[name])
        rec = {"what": "cg", "mode": name, "env": MODES[name], "iters": iters}
        failed = 0.0
        try:
            ms = []
            for r in range(reps + 1):
                barrier()
                system.solve_by_CG(eps=1e-30, max_iter=iters, check_every=iters, fixed_iters=True)
                ms.append(system.ctx.time_ms(1))
            x = system._x.to_numpy()
            if xref is None:
                xref = x
            scale = allmax(float(np.abs(xref).max()))
            rec["max_rel_diff_x_vs_first"] = allmax(float(np.abs(x - xref).max())) / (scale if scale > 0 else 1.0)
        except Exception as exc:                       # a peer-wait timeout etc.: report, then stop (the ranks may disagree)
            rec["error"] = str(exc)[:300]
            failed = 1.0
        if allmax(failed) > 0:
            rec.setdefault("error", "another rank failed")
            emit(rec)
            break
        per = [allmax(m) / iters for m in ms[1:]]
        rec["ms_per_iter_best"], rec["ms_per_iter_median"] = min(per), float(np.median(per))
        rec["it_s"] = 1.0 / min(per) * 1e3
        rec["first_solve_ms_per_iter"] = allmax(ms[0]) / iters
        emit(rec)
    reset()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2b")
    ap.add_argument("--n", type=int, default=119)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--balance", default="equal", choices=["equal", "measured"])
    ap.add_argument("--modes", nargs="*", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from femcy_b200._lib import as_d, as_i32
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    t0 = time.time()
    deck, system, rhs, bcs, ne_global, nn_global, part = bench.build_problem(args.n, rank, world, local, balance=args.balance)
    system.rhs.from_numpy(rhs)
    system.assemble_stiffnessMtrx()
    system.ctx.call("femcy_dirichlet_linear", as_i32(bcs[0]), as_i32(bcs[1]), as_d(bcs[2]), len(bcs[0]))
    setup_s = time.time() - t0

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()

    os.makedirs("gpurun_out", exist_ok=True)
    fh = open(f"gpurun_out/{args.tag}_scaling_ab.jsonl", "a") if rank == 0 else None

    def emit(rec):
        if rank != 0:
            return
        rec["n_gpus"] = world
        line = json.dumps(rec)
        fh.write(line + "\n")
        fh.flush()
        os.fsync(fh.fileno())
        print(line, flush=True)

    emit({"what": "setup", "n": args.n, "ne": int(ne_global), "dofs": int(nn_global * 3), "balance": args.balance,
          "weights": getattr(part, "weights", None) if part is not None else None, "setup_s": round(setup_s, 1)})
    modes = args.modes or DEFAULT_ORDER
    unknown = [m for m in modes if m not in MODES]
    if unknown:
        raise SystemExit(f"unknown modes {unknown}; known: {sorted(MODES)}")
    run_modes(system, modes, args.iters, args.reps, allmax, emit, barrier)
    emit({"what": "done", "wall_s": round(time.time() - t0, 1)})
    system.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
