"""Repeat the bench's parity solve (eps = 1e-8 on cfg 4) several times on N GPUs with different launch chunkings and report the
iteration count, the true residual and the solution samples against the single-GPU fixture: a rare halo race shows up as a
solution that drifts in the low modes (1e-5) although the residual converges.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29891 tools/parity_stress.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
    deck, system, rhs, bcs, ne_global, nn_global, part = bench.build_problem(n, rank, world, local)
    orig = system.solve_by_CG
    chunks = [int(c) for c in os.environ.get("STRESS_CHUNKS", "50,7,500,50,33,50").split(",")]
    for rep, chunk in enumerate(chunks):
        def solve(eps=None, max_iter=None, check_every=None, fixed_iters=False, _c=chunk):
            return orig(eps=eps, max_iter=max_iter, check_every=_c, fixed_iters=fixed_iters)
        system.solve_by_CG = solve
        out = bench.parity_check(system, rhs, bcs, part, n, nn_global, world)
        if rank == 0:
            print(json.dumps({"label": os.environ.get("STRESS_LABEL", ""), "rep": rep, "check_every": chunk, "iters": out["iters"], "residual": out["residual_inf_rel"],
                              "diff_vs_1gpu": out.get("max_rel_diff_x_vs_single_gpu"), "ok": out["parity_ok"]}), flush=True)
    system.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
