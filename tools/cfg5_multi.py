"""BASELINE.json configs[4] (cfg 5): synthetic unit cube, 998 250 C3D10, neo-Hookean, nlgeom, one increment, on 1..8 GPUs.

    python tools/cfg5_multi.py --tag r2k                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29881 \
        tools/cfg5_multi.py --tag r2k [--sigma 1024]

Reports: assembly + internal-force time per residual evaluation, PCG iterations/s inside the Newton loop, the
Newton loop count, and -- as the parity check of the partitioned path -- max|u| and the displacement at 512 fixed global
dofs, which must agree between the 1-GPU and the N-GPU runs (gpurun_out/<tag>_cfg5_n<N>.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2k")
    ap.add_argument("--n", type=int, default=55)
    ap.add_argument("--sigma", type=int, default=0)
    ap.add_argument("--cg-eps", type=float, default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    if args.sigma:
        os.environ["FEMCY_OPT_SELL_SIGMA"] = str(args.sigma)
    from femcy_b200 import Body, System_of_equations, meshgen
    t0 = time.time()
    deck = meshgen.SyntheticDeck("C3D10", n=args.n, nlgeom=True, traction=0.01,
                                 time_incs={"ini_inc": 1., "max_time": 1., "min_inc": 1e-5, "max_inc": 1.})
    nn_global, ne_global = deck.nodes.shape[0], deck.eSets["C3D10"].shape[0]
    part = None
    if world > 1:
        from femcy_b200.partition import Communicator, Partition
        part = Partition(deck.nodes, deck.eSets["C3D10"], rank, world, device=device)
        part.comm = Communicator()
        loc = part.localize_deck(deck)
    else:
        loc = deck
    s = System_of_equations(Body(loc.nodes, loc.eSets["C3D10"], loc.ELE), list(loc.materials.values())[0], True, device=local,
                            quiet=True, partition=part, cg_eps=args.cg_eps)
    setup_s = time.time() - t0
    # device time of the hot calls, accumulated over the Newton loop
    acc = {"asm_ms": 0.0, "asm_calls": 0, "cg_ms": 0.0, "force_s": 0.0, "force_calls": 0}
    asm0, cg0, f0 = s.assemble_stiffnessMtrx, s.solve_by_CG, s.assemble_nodal_force_GN

    def asm():
        asm0(); s.ctx.sync(); acc["asm_ms"] += s.ctx.time_ms(0); acc["asm_calls"] += 1

    def cg(*a, **k):
        r = cg0(*a, **k); acc["cg_ms"] += s.ctx.time_ms(1); return r

    def force():
        s.ctx.sync(); t = time.perf_counter(); f0(); s.ctx.sync(); acc["force_s"] += time.perf_counter() - t; acc["force_calls"] += 1

    s.assemble_stiffnessMtrx, s.solve_by_CG, s.assemble_nodal_force_GN = asm, cg, force
    if world > 1:
        dist.barrier()
    t0 = time.time()
    s.solve(loc)
    s.ctx.sync()
    solve_s = time.time() - t0
    u = s.dof.to_numpy()
    umax = float(np.abs(u[: s.N_own]).max()) if s.N_own else 0.0
    idx = np.sort(np.random.default_rng(5).choice(nn_global * 3, size=512, replace=False))
    vals = np.zeros(idx.size)
    if part is None:
        vals[:] = u[idx]
    else:
        g2l = np.full(nn_global, -1, dtype=np.int64)
        g2l[part.local_to_global[: part.n_own]] = np.arange(part.n_own)
        l = g2l[idx // 3]
        own = l >= 0
        vals[own] = u[l[own] * 3 + idx[own] % 3]
    rec = [acc["asm_ms"], acc["cg_ms"], acc["force_s"], solve_s, umax]
    if world > 1:
        t = torch.from_numpy(vals); dist.all_reduce(t)
        m = torch.tensor(rec, dtype=torch.float64); dist.all_reduce(m, op=dist.ReduceOp.MAX); rec = [float(v) for v in m]
    if rank == 0:
        out = {"config": f"cfg 5: Kuhn cube n={args.n}, {ne_global} C3D10, {nn_global * 3} dofs, NeoHookean(0.4, 20), nlgeom, TRVEC 0.01, 1 increment",
               "n_gpus": world, "sigma": args.sigma, "cg_eps": args.cg_eps, "setup_s": round(setup_s, 2), "solve_s": round(rec[3], 3),
               "increments": [(float(t), bool(c), int(n)) for t, c, n in s.inc_trace],
               "residual_evaluations": acc["asm_calls"], "assembly_ms_per_call": rec[0] / max(acc["asm_calls"], 1),
               "internal_force_ms_per_call": rec[2] * 1e3 / max(acc["force_calls"], 1),
               "cg_iterations_total": int(s.cg_iters_total), "cg_ms_total": rec[1],
               "cg_iter_per_s": s.cg_iters_total / (rec[1] * 1e-3) if rec[1] > 0 else None,
               "elements_per_s_assembly": ne_global / (rec[0] / max(acc["asm_calls"], 1) * 1e-3),
               "max_abs_u": rec[4], "u_samples_checksum": float(vals.sum()), "u_samples": [float(v) for v in vals[:64]]}
        ref_path = f"gpurun_out/{args.tag}_cfg5_n1.json"
        if not os.path.exists(ref_path):
            ref_path = os.path.join(ROOT, "profiles", "r2k_cfg5_n1.json")      # the committed 1-GPU run
        if world > 1 and os.path.exists(ref_path):
            ref = json.load(open(ref_path))
            out["max_rel_diff_u_samples_vs_1gpu"] = float(np.abs(np.array(out["u_samples"]) - np.array(ref["u_samples"])).max() / ref["max_abs_u"])
            out["same_newton_trace_as_1gpu"] = out["increments"] == [tuple(x) for x in ref["increments"]] or out["increments"] == ref["increments"]
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/{args.tag}_cfg5_n{world}{'_sigma' if args.sigma else ''}.json", "w") as fh:
            json.dump(out, fh, indent=1)
        print(json.dumps({k: v for k, v in out.items() if k != "u_samples"}), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
