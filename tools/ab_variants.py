"""A/B of the assembly and PCG variants on one GPU (development aid for round 2; bench.py stays the contract).

    python tools/ab_variants.py [C3D4 119] [C3D10 55] ...        # default: both headline configurations

For every mesh: each assembly variant is checked against variant 1 (max relative difference of K) and timed
(median of 7 warm calls, CUDA events inside the library); then the PCG variants (three-kernel graph, persistent,
single-reduction persistent) run 200 fixed iterations each.  One JSON line per mesh on stdout.
Variants: 1 scatter (default) | 2 per-block gather | 3 scatter, capped registers | 4 scatter, contiguous element
ranges per warp (C3D10/CPS8; rejected r1z) | 5 gather, slice-major launch order (1-GP; default) | 6 owner-computes rows
assembly | 7 / 8 rows + L2 / L1 prefetch + staged pass 1 | 9 gather over node-sector records + staged pass 1.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femcy_b200 import Body, System_of_equations, meshgen  # noqa: E402

BYTES = {"C3D4": 1360, "C3D10": 7720}


def run(kind, n, check=True):
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
    conn = deck.eSets[kind]
    ne = conn.shape[0]
    s = System_of_equations(Body(deck.nodes, conn, deck.ELE), list(deck.materials.values())[0], False, quiet=True)
    out = {"kind": kind, "n": n, "ne": int(ne), "dofs": int(s.N), "nnz": int(s.nnz), "assembly": {}, "cg": {}}
    u = 1e-4 * np.random.default_rng(0).standard_normal(s.N)
    s.dof.from_numpy(u)
    variants = [1, 2, 3, 6, 7, 8, 9, 10, 20, 12, 13] + ([5, 11, 14, 16, 17, 21, 18, 22] if kind == "C3D4" else [15, 19])   # async-copy variants last
    ref = None
    for v in variants:
        s.assembly_variant = v
        try:
            ts = []
            for _ in range(9):
                s.assemble_stiffnessMtrx()
                ts.append(s.ctx.time_ms(0))
            ms = float(np.median(ts[2:]))
            rec = {"ms": ms, "Gelem_s": ne / ms / 1e6, "alg_TBs": ne * BYTES[kind] / ms / 1e9, "first_call_ms": ts[0]}
            if check and ne <= 200_000:          # small meshes: compare every entry; larger ones: K.x below
                K = s.csr()
                if ref is None:
                    ref = K
                rec["max_rel_diff_vs_v1"] = float(abs(K - ref).max() / abs(ref).max())
            elif check:
                # big mesh: compare K.x for a fixed x instead of exporting 230 M values
                x = np.sin(np.arange(s.N) * 0.37)
                s._x.from_numpy(x)
                from femcy_b200._lib import VEC
                s.ctx.call("femcy_spmv", VEC["x"], VEC["du"])
                y = s.du.to_numpy()
                if ref is None:
                    ref = y
                rec["max_rel_diff_Kx_vs_v1"] = float(np.abs(y - ref).max() / np.abs(ref).max())
            out["assembly"][f"v{v}"] = rec
        except Exception as e:                      # an experimental variant failing must not hide the others
            out["assembly"][f"v{v}"] = {"error": str(e)[:200]}
        print(f"# {kind} n={n} assembly v{v}: {out['assembly'][f'v{v}']}", file=sys.stderr, flush=True)
    # PCG variants on the default-assembled, Dirichlet-eliminated system
    s.assembly_variant = 1
    s.dof.fill(0.)
    s.assemble_stiffnessMtrx()
    nb = deck.neumann_bc_info[0]
    s.neumannBC(nb["face_set"], nb["traction"], nb["direction"])
    for bc in deck.dirichlet_bc_info:
        s.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
    iters = 200
    xref = None
    for name, env in (("persistent", {}), ("persistent_minb5", {"FEMCY_CG_MINB": "5"}), ("three_kernel_graph", {"FEMCY_CG_MULTIKERNEL": "1"}), ("single_reduction", {"FEMCY_CG_VARIANT": "sr"}), ("single_reduction_minb5", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_MINB": "5"}), ("single_reduction_fold_barrier", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_FOLD_BARRIER": "1"}), ("persistent_sym", {"FEMCY_CG_SYM": "1"}), ("single_reduction_sym", {"FEMCY_CG_VARIANT": "sr", "FEMCY_CG_SYM": "1"}), ("persistent_l2persist", {"FEMCY_CG_L2_PERSIST": "1"}), ("persistent_sym_l2persist", {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "1"}), ("persistent_sym_l2matrix", {"FEMCY_CG_SYM": "1", "FEMCY_CG_L2_PERSIST": "2"})):
        for k in ("FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_PERSISTENT", "FEMCY_CG_MINB", "FEMCY_CG_FOLD_BARRIER", "FEMCY_CG_SYM", "FEMCY_CG_L2_PERSIST"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            ms = []
            for _ in range(4):
                s.solve_by_CG(eps=1e-30, max_iter=iters, check_every=100, fixed_iters=True)
                ms.append(s.ctx.time_ms(1))
            x = s._x.to_numpy()
            if xref is None:
                xref = x
            m = float(np.median(ms[1:]))
            out["cg"][name] = {"ms_per_iter": m / iters, "it_s": iters / m * 1e3,
                               "max_rel_diff_x_vs_persistent": float(np.abs(x - xref).max() / np.abs(xref).max())}
        except Exception as e:
            out["cg"][name] = {"error": str(e)[:200]}
        print(f"# {kind} n={n} cg {name}: {out['cg'][name]}", file=sys.stderr, flush=True)
    for k in ("FEMCY_CG_MULTIKERNEL", "FEMCY_CG_VARIANT", "FEMCY_CG_PERSISTENT", "FEMCY_CG_MINB", "FEMCY_CG_FOLD_BARRIER", "FEMCY_CG_SYM", "FEMCY_CG_L2_PERSIST"):
        os.environ.pop(k, None)
    s.close()
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    jobs = [(args[i], int(args[i + 1])) for i in range(0, len(args) - 1, 2)] or [("C3D4", 119), ("C3D10", 55)]
    for kind, n in jobs:
        t0 = time.time()
        r = run(kind, n)
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)
