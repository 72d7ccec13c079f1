#!/usr/bin/env python
"""Benchmark of the FEMcy hot path on B200 (contract: see the task statement / DESIGN.md section 7).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # CPU arm: the reference's algorithm on the host cores

Workload (BASELINE.json configs[3]): synthetic unit-cube Kuhn mesh, n=119 cells/edge ->
10 110 954 C3D4 elements, 1 728 000 nodes, 5 184 000 dofs; LinearIsotropic(E=2.1e5, nu=0.3);
face x=0 clamped, TRVEC traction on x=1; fp64.  One *step* = one linear increment of the hot path:
assemble K on X+u (K.fill(0) included), impose the Dirichlet conditions, run `cg_iters` Jacobi-PCG
iterations with the convergence exit disabled.  Reported: K-assembly elements/s (`value`) and
CG iterations/s (`cg.value`), each with its HBM roofline fraction.  With N>1 the same mesh is
partitioned by elements/rows over the ranks (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "K-assembly elems/sec (value) + CG-iter/sec (cg.value) on 10M-elem C3D4; HBM GB/s vs roofline"
ASM_BYTES_PER_ELEM = 1360          # 4*4 + 2*4*3*8 + 12*12*8


# dram__bytes_read.sum + dram__bytes_write.sum of the default kernels on this workload come from the committed ncu summary
# profiles/r2_traffic.json (written by tools/ncu_traffic.py from the `ncu --set full` reports of the CURRENT defaults)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic.json")


def measured_traffic():
    try:
        with open(TRAFFIC_FILE) as fh:
            return json.load(fh)
    except Exception:
        return {}


def spmv_bytes(nnz, N):
    return nnz * 12 + N * 20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


PARITY_SAMPLES = 4096
PARITY_EPS = 1.0e-8
PARITY_TOL_X = 1.0e-6          # north_star: displacements within 1e-6 relative
PARITY_TOL_RES = 1.0e-6        # true residual max|b - K x| / max|b| of the eps = 1e-8 solve


def parity_golden_path(n):
    return os.path.join(ROOT, "tests", "golden", f"bench_solution_samples_n{n}.npz")


def parity_check(system, rhs, bcs, part, n, nn_global, world, write_golden=False):
    """Correctness of the partitioned (or single-GPU) solve of THIS workload, after the timed region: assemble, eliminate,
    solve to eps = 1e-8, then (1) the true residual max|b - K x| / max|b| recomputed with one more SpMV (owned rows, max over
    the ranks) and (2) x at PARITY_SAMPLES fixed pseudo-random global dofs against the committed single-GPU values
    (tests/golden/bench_solution_samples_n<n>.npz, written by `bench.py --gpus 1 --write-parity-golden`)."""
    import torch
    import torch.distributed as dist
    from femcy_b200._lib import VEC, as_d, as_i32
    ctx = system.ctx
    # K is assembled on X + dof (stiffnessMtrx.py:141-142): start from dof = 0, not from whatever unconverged iterate the timed
    # steps left behind -- otherwise the system solved here depends on the step count and on N (a 1e-5 effect on x)
    system.dof.fill(0.)
    system.assemble_stiffnessMtrx()
    system.rhs.from_numpy(rhs)
    ctx.call("femcy_dirichlet_linear", as_i32(bcs[0]), as_i32(bcs[1]), as_d(bcs[2]), len(bcs[0]))
    system.solve_by_CG(eps=PARITY_EPS, max_iter=20000, check_every=50)
    iters = int(system.last_cg_iters)
    n_own = system.N_own
    x = ctx.vec_get("x", system.N)
    ctx.call("femcy_spmv", VEC["x"], VEC["Ad"])            # Ad = K x on the owned rows (halo refreshed inside)
    b = ctx.vec_get("rhs", system.N)[:n_own]
    res = float(np.abs(b - ctx.vec_get("Ad", system.N)[:n_own]).max()) if n_own else 0.0
    bmax = float(np.abs(b).max()) if n_own else 0.0
    xmax = float(np.abs(x[:n_own]).max()) if n_own else 0.0
    rng = np.random.default_rng(20261017)
    idx = np.sort(rng.choice(nn_global * 3, size=min(PARITY_SAMPLES, nn_global * 3), replace=False))
    vals = np.zeros(idx.size)
    if part is None:
        vals[:] = x[idx]
    else:
        g2l = np.full(nn_global, -1, dtype=np.int64)
        g2l[part.local_to_global[: part.n_own]] = np.arange(part.n_own)
        loc = g2l[idx // 3]
        own = loc >= 0
        vals[own] = x[loc[own] * 3 + idx[own] % 3]
    if world > 1:
        t = torch.from_numpy(vals)
        dist.all_reduce(t)                                 # every sample is owned by exactly one rank
        m = torch.tensor([res, bmax, xmax], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        res, bmax, xmax = (float(v) for v in m)
    out = {"eps": PARITY_EPS, "iters": iters, "residual_inf_rel": res / bmax if bmax > 0 else None, "samples": int(idx.size),
           "max_abs_x": xmax, "checksum_x_samples": float(vals.sum())}
    gp = parity_golden_path(n)
    if write_golden and world == 1:
        np.savez(gp, idx=idx, x=vals, iters=iters, max_abs_x=xmax, n=n)
        out["golden_written"] = os.path.relpath(gp, ROOT)
    ok = out["residual_inf_rel"] is not None and out["residual_inf_rel"] <= PARITY_TOL_RES
    if os.path.exists(gp):
        g = np.load(gp)
        same = np.array_equal(g["idx"], idx)
        diff = float(np.abs(vals - g["x"]).max() / float(g["max_abs_x"])) if same else None
        out.update({"golden": os.path.relpath(gp, ROOT), "golden_iters": int(g["iters"]), "max_rel_diff_x_vs_single_gpu": diff})
        ok = ok and same and diff <= PARITY_TOL_X
    else:
        out["golden"] = None                               # no committed single-GPU values for this size: residual check only
    out["tolerances"] = {"x": PARITY_TOL_X, "residual": PARITY_TOL_RES}
    out["parity_ok"] = bool(ok)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(n, rank, world, device, jitter=0.1, balance="equal"):
    from femcy_b200 import Body, System_of_equations, meshgen
    deck = meshgen.SyntheticDeck("C3D4", n=n, jitter=jitter)
    ne_global = deck.eSets["C3D4"].shape[0]
    nn_global = deck.nodes.shape[0]
    part = None
    if world > 1:
        from femcy_b200.partition import Communicator, Partition, measure_device_bandwidth
        comm = Communicator()
        weights = None
        if balance == "measured":
            # rows proportional to each GPU's measured copy rate: every PCG iteration waits for the slowest rank
            weights = [float(w) for w in comm.allgather_object(measure_device_bandwidth(device))]
        try:
            # this rank's piece of the mesh, computed on this rank's GPU (femcy_partition)
            part = Partition(deck.nodes, deck.eSets["C3D4"], rank, world, weights=weights, device=device)
        except Exception as e:      # set-up only: the NumPy statement of the same scheme yields the same arrays; say so loudly
            print(f"[bench] rank {rank}: device partitioner failed ({e}); using the host partitioner", file=sys.stderr, flush=True)
            part = Partition(deck.nodes, deck.eSets["C3D4"], rank, world, weights=weights)
        part.weights = weights
        part.comm = comm
        deck = part.localize_deck(deck)
    body = Body(deck.nodes, deck.eSets["C3D4"], deck.ELE)
    system = System_of_equations(body, deck.materials["Elastic"], False, device=device, quiet=True, partition=part)
    nb = deck.neumann_bc_info[0]
    rhs = system.neumann_vector(nb["face_set"], nb["traction"], nb["direction"])
    bc_n = np.concatenate([np.asarray(bc["node_set"], dtype=np.int32) for bc in deck.dirichlet_bc_info])
    bc_c = np.concatenate([np.full(len(bc["node_set"]), bc["dof"], dtype=np.int32) for bc in deck.dirichlet_bc_info])
    bc_v = np.zeros(len(bc_n))
    return deck, system, rhs, (np.ascontiguousarray(bc_n), np.ascontiguousarray(bc_c), bc_v), ne_global, nn_global, part


def builder_timings(deck, system, device):
    """Wall time (ms, host clock around calls that synchronise) of the device-side builders and opt-in kernels on THIS workload,
    after the timed region and the parity leg: rows f1 / f2 of the scope table have no throughput metric of their own, so the
    numbers ride on the bench line.  Set-up work, never part of `value`; any failure is reported as text, not raised."""
    import ctypes as C
    from femcy_b200._lib import as_i32
    out = {}
    ctx = system.ctx

    def timed(name, fn):
        try:
            ctx.sync()
            t0 = time.time()
            r = fn()
            ctx.sync()
            out[name] = (time.time() - t0) * 1e3
            return r
        except Exception as e:          # noqa: BLE001 -- diagnostics only
            out[name] = "failed: " + str(e)[:200]
            return None

    out["pattern_build"] = ctx.time_ms(2)                       # femcy_build_pattern of the set-up (CUDA events)
    n = C.c_int64(0)
    timed("boundary_facets", lambda: ctx.call("femcy_boundary_facets", C.byref(n)))
    out["boundary_facets_found"] = int(n.value)
    ne, n_en = system.body.np_elements.shape
    ptr, lst = np.empty(system.body.np_nodes.shape[0] + 1, np.int32), np.empty(max(ne * n_en, 1), np.int32)
    timed("node_elements_incl_d2h", lambda: ctx.call("femcy_node_elements", as_i32(ptr), as_i32(lst)))
    nb = deck.neumann_bc_info[0]
    timed("neumann_device", lambda: system.neumannBC(nb["face_set"], nb["traction"], nb["direction"]))
    out["neumann_facets"] = len(nb["face_set"])
    t0 = time.time()
    system.neumann_vector(nb["face_set"], nb["traction"], nb["direction"])
    out["neumann_host_numpy"] = (time.time() - t0) * 1e3
    for name, variant, tangent in (("assembly_scatter", 1, 0), ("assembly_consistent_tangent", 1, 1)):
        try:
            ctx.set_option("consistent_tangent", tangent)      # (the kernel is timed on the bench mesh; nothing is solved with it)
            for _ in range(2):
                ctx.call("femcy_assemble_K", variant)
            ctx.sync()
            out[name] = ctx.time_ms(0)
        except Exception as e:          # noqa: BLE001
            out[name] = "failed: " + str(e)[:200]
        finally:
            ctx.set_option("consistent_tangent", 0)
    # opt-in consistent tangent vs the reference's modified Newton on a small neo-Hookean C3D10 problem (cfg 5's kind, 24 576
    # elements): Newton loops, PCG iterations, wall time of the whole solve
    try:
        from femcy_b200 import Body, System_of_equations, meshgen
        small = meshgen.SyntheticDeck("C3D10", n=16, nlgeom=True, traction=0.01)
        newton = {}
        for tangent in ("reference", "consistent"):
            s2 = System_of_equations(Body(small.nodes, small.eSets["C3D10"], small.ELE), small.materials["Hyperelastic, neo hooke"], True,
                                     device=device or 0, quiet=True, cg_eps=1e-8)
            s2.set_tangent(tangent)
            t0 = time.time()
            s2.solve(small)
            newton[tangent] = {"solve_s": time.time() - t0, "newton_loops": [int(l) for _, _, l in s2.inc_trace],
                               "converged": [bool(c) for _, c, _ in s2.inc_trace], "pcg_iterations": int(s2.cg_iters_total),
                               "max_abs_u": float(np.abs(s2.dof.to_numpy()).max()), "tangent_fallbacks": int(s2.tangent_fallbacks)}
            s2.close()
        out["newton_c3d10_n16"] = newton
    except Exception as e:              # noqa: BLE001
        out["newton_c3d10_n16"] = "failed: " + str(e)[:200]
    # row f4 at the bench size: the same mesh as two sections (two materials, split at x = 0.5) -- union pattern from one sort,
    # one scatter-add pass per section; compare with `assembly_scatter` (one section, same kernel) above
    try:
        from femcy_b200 import System_of_equations
        from femcy_b200.body import SectionedBody
        from femcy_b200.material_zoo import LinearIsotropic
        conn = system.body.np_elements
        left = system.body.np_nodes[conn, 0].mean(axis=1) < 0.5
        sbody = SectionedBody(system.body.np_nodes, [(conn[left], deck.ELE, system.material),
                                                     (conn[~left], deck.ELE, LinearIsotropic(modulus=7.0e4, poisson_ratio=0.33))])
        t0 = time.time()
        s3 = System_of_equations(sbody, None, False, device=device or 0, quiet=True)
        out["two_sections_setup_s"] = time.time() - t0
        out["two_sections_pattern_build"] = s3.ctx.time_ms(2)
        ts = []
        for _ in range(3):
            s3.assemble_stiffnessMtrx()
            s3.ctx.sync()
            ts.append(s3.ctx.time_ms(0))
        out["assembly_two_sections_scatter"] = float(np.median(ts))
        out["two_sections_nnz"] = int(s3.nnz)
        s3.close()
    except Exception as e:              # noqa: BLE001
        out["assembly_two_sections_scatter"] = "failed: " + str(e)[:200]
    if device is not None:
        from femcy_b200.partition import Partition
        t0 = time.time()
        try:
            p = Partition(deck.nodes, deck.eSets["C3D4"], 3, 8, device=device)
            out["partition_device_rank3_of_8"] = (time.time() - t0) * 1e3
            out["partition_local_elements"] = int(p.elements.shape[0])
        except Exception as e:      # noqa: BLE001
            out["partition_device_rank3_of_8"] = "failed: " + str(e)[:200]
    return out


# femcy_assemble_K formulations (include/femcy_b200.h); 0 = library default = gather
ASM_KERNELS = {0: "k_elem_geometry4t (TMA tensor store) + k_assemble_gather_h<3,4> (pair of lanes per block)",
               1: "cudaMemset(K) + k_assemble_scatter<3,4,1>",
               2: "k_elem_geometry4t (TMA tensor store) + k_assemble_gather_h<3,4> (pair of lanes per block)",
               3: "k_elem_geometry4t (TMA tensor store) + k_assemble_gather_p<3,4,1> (thread per block)"}
CG_KERNELS = {0: "k_cg_stream<3,16,2,2>", 1: "k_spmv_dot<3> + k_update_xr + k_update_d (CUDA graph)", 2: "k_cg_persistent<3,6>",
              3: "k_cg_stream<3,16,2,2>"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from femcy_b200._lib import VEC, as_d, as_i32
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    t_setup = time.time()
    deck, system, rhs, bcs, ne_global, nn_global, part = build_problem(args.n, rank, world, local, balance=args.balance)
    ctx = system.ctx
    stream = torch.cuda.Stream()
    ctx.call("femcy_set_stream", stream.cuda_stream)
    N_loc, N_glob = system.N, nn_global * 3
    nnz_loc = system.nnz
    system.rhs.from_numpy(rhs)
    t_setup = time.time() - t_setup
    cg_iters = args.cg_iters

    cg_launch_ms, cg_phase_us = [], []

    def one_step(ev=None):
        if ev:
            ev[0].record(stream)
        system.assemble_stiffnessMtrx()
        if ev:
            ev[1].record(stream)
        ctx.call("femcy_dirichlet_linear", as_i32(bcs[0]), as_i32(bcs[1]), as_d(bcs[2]), len(bcs[0]))
        if ev:
            ev[2].record(stream)
        system.solve_by_CG(eps=1e-30, max_iter=cg_iters, check_every=cg_iters, fixed_iters=True)
        if ev:
            ev[3].record(stream)
            # the solve has synchronised the stream: the library's own event pair around the PCG kernel launch (one
            # cooperative launch = cg_iters iterations) and the kernel's device phase clock are read without disturbing it
            cg_launch_ms.append(ctx.time_ms(1))
            cg_phase_us.append(ctx.cg_phase_ns() / 1e3 / cg_iters)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # (the Dirichlet pass edits rhs in place; with zero prescribed values it is idempotent)
        for _ in range(args.warmup):
            one_step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        l0 = ctx.launches()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for k in range(args.steps):
            one_step(evs[k])
        t1.record(stream)
        barrier()
        launches = ctx.launches() - l0
        clocks = sampler.stop() if rank == 0 else None
        total_ms = t0.elapsed_time(t1)
        asm_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
        bc_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
        cg_ms = sum(e[2].elapsed_time(e[3]) for e in evs)

        # stand-alone SpMV (the three-kernel path's k_spmv_dot, back to back): the L2-warm upper bound of the SpMV rate
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ctx.call("femcy_spmv", VEC["d"], VEC["Ad"])
        barrier()
        s0.record(stream)
        n_spmv = 20
        for _ in range(n_spmv):
            ctx.call("femcy_spmv", VEC["d"], VEC["Ad"])
        s1.record(stream)
        barrier()
        spmv_alone_ms = s0.elapsed_time(s1) / n_spmv

        # ---- end-to-end through the public API with (pinned) host buffers --------------------------
        u_host = torch.zeros(N_loc, dtype=torch.float64).pin_memory().numpy()
        rhs_host = torch.from_numpy(rhs).pin_memory().numpy()
        x_host = torch.empty(N_loc, dtype=torch.float64).pin_memory().numpy()
        import ctypes as C
        vol_total = C.c_double(0.)
        vol_host = torch.empty(system.body.np_elements.shape[0], dtype=torch.float64).pin_memory().numpy()   # fallback only
        e2e_metric = "mesh volume (8 B)"
        e2e_asm, e2e_cg = [], []
        for k in range(2 + args.steps):
            barrier()
            ta = time.perf_counter()
            ctx.call("femcy_vec_set", VEC["dof"], as_d(u_host), N_loc)                 # H2D u
            system.assemble_stiffnessMtrx()        # (fused with the geometry pass: no separate get_dsdx_and_vol needed)
            if e2e_metric == "mesh volume (8 B)":
                try:
                    ctx.call("femcy_gp_sum", 0, C.byref(vol_total))                      # D2H: the mesh volume (8 B metric)
                except Exception as exc:                                                 # insurance: never lose the run to the read-back
                    sys.stderr.write(f"femcy_gp_sum failed ({exc}); e2e reads the vol array back instead\n")
                    e2e_metric = "vol array"
            if e2e_metric == "vol array":
                ctx.call("femcy_gp_get", 0, as_d(vol_host), vol_host.size)
                vol_total.value = float(vol_host.sum())

            tb = time.perf_counter()
            ctx.call("femcy_vec_set", VEC["rhs"], as_d(rhs_host), N_loc)                # H2D rhs
            ctx.call("femcy_dirichlet_linear", as_i32(bcs[0]), as_i32(bcs[1]), as_d(bcs[2]), len(bcs[0]))
            system.solve_by_CG(eps=1e-30, max_iter=cg_iters, check_every=cg_iters, fixed_iters=True)
            ctx.call("femcy_vec_get", VEC["x"], as_d(x_host), N_loc)                    # D2H x
            tc = time.perf_counter()
            if k >= 2:
                e2e_asm.append(tb - ta)
                e2e_cg.append(tc - tb)

        parity = None
        if not args.no_parity:
            parity = parity_check(system, rhs, bcs, part, args.n, nn_global, world, write_golden=args.write_parity_golden)

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cg_kernel_ms = float(np.mean(cg_launch_ms))                      # mean duration of one PCG kernel launch (= cg_iters iterations)
    phases = np.mean(np.asarray(cg_phase_us), axis=0)                # us per iteration: spmv | b1 | x1 | xr | b2 | x2 | d+b3
    spmv_phase_ms = float(phases[0] + phases[1]) * 1e-3              # SpMV loop + the barrier that waits for its slowest block
    total_ms, asm_ms, cg_ms, bc_ms, cg_kernel_ms, spmv_phase_ms, spmv_alone_ms = map(
        maxr, (total_ms, asm_ms, cg_ms, bc_ms, cg_kernel_ms, spmv_phase_ms, spmv_alone_ms))
    e2e_asm_s, e2e_cg_s = maxr(sum(e2e_asm)), maxr(sum(e2e_cg))
    nnz_glob = nnz_loc
    if world > 1:
        t = torch.tensor([nnz_loc], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        nnz_glob = int(t.item())
    peak, peak_src = measured_peaks()
    K = args.steps
    value = ne_global * K / (asm_ms * 1e-3)
    cg_value = cg_iters * K / (cg_ms * 1e-3)
    asm_GBs = ne_global * ASM_BYTES_PER_ELEM * K / (asm_ms * 1e-3) / 1e9
    iter_bytes = spmv_bytes(nnz_glob, N_glob) + 11 * N_glob * 8                      # one PCG iteration
    cgit_GBs = iter_bytes * cg_iters * K / (cg_ms * 1e-3) / 1e9
    kern_GBs = iter_bytes * cg_iters / (cg_kernel_ms * 1e-3) / 1e9                   # the persistent kernel alone
    spmv_GBs = spmv_bytes(nnz_glob, N_glob) / (spmv_phase_ms * 1e-3) / 1e9 if spmv_phase_ms > 0 else None
    opt_kernel = int(os.environ.get("FEMCY_OPT_CG_KERNEL", "0"))
    if world > 1 and not getattr(part, "p2p", False):
        opt_kernel = 1
    opt_sym = int(os.environ.get("FEMCY_OPT_CG_SYM", "0"))
    cg_name = CG_KERNELS.get(opt_kernel, "?") + (" (upper-half SpMV)" if opt_sym else "")
    tr = measured_traffic() if (args.n == 119 and world == 1) else {}
    tr_cg = tr.get("cg_sym" if opt_sym else {0: "cg_stream", 3: "cg_stream", 2: "cg_persistent"}.get(opt_kernel, ""), {})
    tr_asm = tr.get({0: "assembly_gather", 2: "assembly_gather", 3: "assembly_gather_thread", 1: "assembly_scatter"}.get(int(system.assembly_variant), ""), {})
    cg_dram = tr_cg.get("dram_bytes_per_iteration")
    asm_dram = tr_asm.get("dram_bytes_per_assembly")
    out = {
        "metric": METRIC, "value": value, "unit": "elem/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"unit-cube Kuhn C3D4 n={args.n}: {ne_global} elements, {nn_global} nodes, "
                               f"{N_glob} dofs, nnz {nnz_glob}; step = assemble K + Dirichlet + {cg_iters} PCG iterations",
                   "elements": ne_global, "dofs": N_glob, "nnz": nnz_glob, "cg_iters_per_step": cg_iters,
                   "parallelism": f"element/row partition x{world}" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (K values 1.8 GB, connectivity+slots 0.8 GB per pass)",
                   "assembly_variant": int(system.assembly_variant), "assembly_kernels": ASM_KERNELS.get(int(system.assembly_variant)),
                   "cg_kernel": cg_name,
                   "options": {k: v for k, v in os.environ.items() if k.startswith("FEMCY_OPT_")}},
        "cg": {"value": cg_value, "unit": "iter/s", "ms_per_iter": cg_ms / (cg_iters * K),
               "algorithmic_GBs": cgit_GBs, "frac_of_peak": cgit_GBs / (peak * world)},
        "phase_ms_per_step": {"assemble": asm_ms / K, "dirichlet": bc_ms / K, "cg": cg_ms / K},
        # the dominant kernel of the step: ONE launch of the persistent PCG kernel = cg_iters whole iterations
        "roofline": {"kernel": cg_name + " (1 launch/step = %d PCG iterations: SpMV + dot, x/r update + norms, d update)" % cg_iters,
                     "bound": "hbm", "achieved": kern_GBs, "peak": peak * world, "unit": "GB/s", "frac": kern_GBs / (peak * world),
                     "traffic": None if cg_dram is None else cg_dram * cg_iters,
                     "frac_dram": None if cg_dram is None else cg_dram * cg_iters / (cg_kernel_ms * 1e-3) / 1e9 / (peak * world),
                     "traffic_source": tr_cg.get("source"),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": iter_bytes * cg_iters,
                     "ms_per_launch": cg_kernel_ms,
                     "how": "mean over the timed steps of the library's CUDA-event pair around the kernel launch; frac = algorithmic "
                            "bytes (SURVEY 8d: CSR 12 B/nnz) / time / peak, frac_dram = ncu dram bytes / time / peak (the 3x3-block "
                            "SELL-32 format moves 8.4 B/nnz, so frac > frac_dram)",
                     "phase_us_per_iteration": {k: float(v) for k, v in zip(
                         ("spmv", "barrier_fold_1", "exchange_1", "update_xr", "barrier_fold_2", "exchange_2", "update_d_barrier_3"), phases)},
                     "spmv_phase": {"algorithmic_bytes": spmv_bytes(nnz_glob, N_glob), "ms": spmv_phase_ms,
                                    "achieved": spmv_GBs, "frac": None if spmv_GBs is None else spmv_GBs / (peak * world),
                                    "standalone_k_spmv_dot_ms": spmv_alone_ms}},
        "roofline_assembly": {"kernel": ASM_KERNELS.get(int(system.assembly_variant), "variant %d" % system.assembly_variant),
                              "bound": "hbm", "achieved": asm_GBs,
                              "peak": peak * world, "unit": "GB/s", "frac": asm_GBs / (peak * world),
                              "traffic": asm_dram,
                              "frac_dram": None if asm_dram is None else asm_dram / (asm_ms / K * 1e-3) / 1e9 / (peak * world),
                              "traffic_source": tr_asm.get("source"),
                              "algorithmic_bytes_per_launch": ne_global * ASM_BYTES_PER_ELEM, "ms_per_launch": asm_ms / K},
        "e2e": {"value": ne_global * K / e2e_asm_s, "unit": "elem/s",
                "cg_value": cg_iters * K / e2e_cg_s, "cg_unit": "iter/s",
                "h2d_bytes_per_step": 2 * N_loc * 8,
                "d2h_bytes_per_step": int((8 if e2e_metric.startswith("mesh") else vol_host.size * 8) + N_loc * 8),
                "mesh_volume": vol_total.value,      # rank-local (the unit cube: 1.0 on one GPU; interface elements are integrated redundantly on several)
                "what": "assembly: H2D u -> assemble_stiffnessMtrx (geometry fused) -> D2H " + e2e_metric + "; "
                        "cg: H2D rhs -> Dirichlet + solve_by_CG -> D2H x; pinned host buffers, host clock around the calls"},
        "gpu_launches": int(launches), "clocks": clocks, "setup_s": t_setup,
        "parity": parity, "parity_ok": None if parity is None else parity["parity_ok"],
    }
    if world == 1 and not args.no_builder_timings:
        out["builders_ms"] = builder_timings(deck, system, local)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sample_n=args.cpu_sample_n, cg_iters=10, steps=1)
    if world > 1:
        out["config"]["partitioner"] = getattr(part, "built_on", "host")
        out["config"]["partition_balance"] = args.balance if getattr(part, "weights", None) is None else {"measured_GBs": part.weights}
        out["config"]["cg_exchange"] = ("NVLink peer memory (cudaIpc): halo push fused into the d update, partial dots through "
                                        "peer windows, all inside the persistent kernel" if getattr(part, "p2p", False)
                                        and not os.environ.get("FEMCY_OPT_NO_P2P") else "NCCL send/recv + all-gather")
        dist.barrier()
        dist.destroy_process_group()
    return out if rank == 0 else None


def _cpu_arm(n, cg_iters, steps, warmup):
    """The reference's algorithm (ELL rows, full-row slot scans, fp64 atomics, 8-kernel PCG with host
    scalars) restated in C/OpenMP (oracle/femcy_oracle.c) and timed on the host cores: per step
    get_dsdx_and_vol + assemble_stiffnessMtrx, then Dirichlet (untimed) and `cg_iters` PCG iterations."""
    from femcy_b200 import meshgen
    from femcy_b200.body import Body
    from femcy_b200.neumann import neumann_vector
    from oracle import c_oracle as CO, femcy_oracle as O
    deck = meshgen.SyntheticDeck("C3D4", n=n, jitter=0.1)
    nodes = deck.nodes
    conn = np.ascontiguousarray(deck.eSets["C3D4"], dtype=np.int32)
    C = O.C_linear_isotropic(2.1e5, 0.3)
    dN, w = O.elem_tables("C3D4")
    dN = np.ascontiguousarray(dN)
    ij = CO.ell_pattern(conn, nodes.shape[0], 3)
    nb = deck.neumann_bc_info[0]
    rhs = neumann_vector(Body(nodes, conn, deck.ELE), nb["face_set"], nb["traction"], nb["direction"])
    dofs = np.concatenate([bc["node_set"] * 3 + bc["dof"] for bc in deck.dirichlet_bc_info])
    u = np.zeros(nodes.size)
    spm = np.empty((ij.shape[0], ij.shape[1] - 1))
    all_cores = len(os.sched_getaffinity(0))

    def run(threads, n_warm, n_steps):
        CO.set_num_threads(threads)
        ta, tc = [], []
        for k in range(n_warm + n_steps):
            t0 = time.perf_counter()
            dsdx, vol = CO.dsdx_vol(nodes, conn, u, dN, w)
            CO.assemble_ell(conn, 3, dsdx, vol, C, ij, spm)
            t1 = time.perf_counter()
            A, b = CO.dirichlet_ell(spm, ij, dofs, rhs)
            t2 = time.perf_counter()
            CO.pcg_ell(A, ij, b, eps=1e-30, max_iter=cg_iters, fixed_iters=True)
            t3 = time.perf_counter()
            if k >= n_warm:
                ta.append(t1 - t0)
                tc.append(t3 - t2)
        return {"threads": threads, "asm_s": sum(ta), "cg_s": sum(tc)}

    # the thread count is chosen on one probe step each (all cores / half of them: the atomics of the scatter do not
    # always scale to both sockets), then the timed steps run with the better one
    probes = [run(t, 1, 1) for t in sorted({all_cores, max(1, all_cores // 2)}, reverse=True)]
    threads = min(probes, key=lambda r: r["asm_s"] + r["cg_s"])["threads"]
    best = run(threads, max(0, warmup - 2), steps)
    ne, N = conn.shape[0], ij.shape[0]
    return {"ne": ne, "N": N, "steps": steps, "cg_iters": cg_iters, **best}


def cpu_baseline(sample_n, cg_iters, steps):
    r = _cpu_arm(sample_n, cg_iters, steps, warmup=1)
    scale = r["N"] / 5184000.0
    return {"value": r["ne"] * steps / r["asm_s"], "unit": "elem/s", "cores": r["threads"], "kind": "port",
            "cg_value": cg_iters * steps / r["cg_s"] * scale, "cg_unit": "iter/s (scaled to the 5 184 000-dof size)",
            "cg_value_at_sample_size": cg_iters * steps / r["cg_s"],
            "sample": f"Kuhn cube n={sample_n}: {r['ne']} C3D4 elements, {r['N']} dofs; {steps} x (assembly + "
                      f"{cg_iters} PCG iterations); C/OpenMP restatement of the reference's Taichi kernels "
                      "(Taichi is not installable in this image)"}


def run_reference(args):
    """CPU arm: the reference's own algorithm for the path on the host cores (oracle port -- Taichi is
    not installable in this image), same metric/unit/config; each step is a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n, cg_it, K = args.ref_n, args.ref_cg_iters, args.steps
    r = _cpu_arm(n, cg_it, K, args.warmup)
    val = r["ne"] * K / r["asm_s"]
    scale = r["N"] / 5184000.0
    cgv = cg_it * K / r["cg_s"] * scale
    same = (n == args.n)
    sample = (f"Kuhn cube n={n}: {r['ne']} C3D4 elements, {r['N']} dofs per step" +
              (" (the full n=%d workload); each step = 1 assembly + %d PCG iterations" % (n, cg_it) if same else
               f" (bounded sample of the n={args.n} workload); CG iter/s scaled by the dof ratio {scale:.4f} to the 5 184 000-dof size"))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "elem/s", "n_gpus": args.gpus, "steps": K,
           "warmup": args.warmup, "ms_per_step": (r["asm_s"] + r["cg_s"]) / K * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "unit-cube Kuhn C3D4, LinearIsotropic; step = get_dsdx_and_vol + assemble_stiffnessMtrx "
                                  f"+ {cg_it} PCG iterations; " + sample},
           "cg": {"value": cgv, "unit": "iter/s"},
           "cpu_baseline": {"value": val, "unit": "elem/s", "cores": r["threads"], "kind": "port", "sample": sample,
                            "cg_value": cgv, "cg_unit": "iter/s"},
           "e2e": {"value": val, "unit": "elem/s", "cg_value": cgv, "cg_unit": "iter/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=119, help="cells per edge of the Kuhn cube (119 -> 10.1M elements)")
    ap.add_argument("--cg-iters", type=int, default=500, help="PCG iterations per step, convergence exit disabled (SURVEY 8d: 500)")
    ap.add_argument("--balance", default="equal", choices=["equal", "measured"],
                    help="multi-GPU row partition: equal node counts, or proportional to each GPU's measured copy rate")
    ap.add_argument("--cpu-sample-n", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-builder-timings", action="store_true", help="skip the post-run timing of the device builders / opt-in kernels")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run solve + residual / solution-sample check")
    ap.add_argument("--write-parity-golden", action="store_true",
                    help="(1 GPU) write tests/golden/bench_solution_samples_n<n>.npz from this run's solution")
    ap.add_argument("--ref-n", type=int, default=119, help="reference arm: cells per edge (119 = the bench workload itself)")
    ap.add_argument("--ref-cg-iters", type=int, default=20, help="reference arm: PCG iterations per step (a rate: iter/s)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # the contract is ONE JSON line on stdout: libraries that print to stdout (NCCL's version banner,
    # torchrun notices) are sent to stderr for the duration of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    sys.stdout.flush()
    if out is not None:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
        if out.get("parity_ok") is False:
            sys.stderr.write("bench.py: PARITY FAILED: %s\n" % json.dumps(out["parity"]))
            sys.exit(3)


if __name__ == "__main__":
    main()
