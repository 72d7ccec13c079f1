"""Run the BASELINE.json configurations end to end on one GPU and record what happened
(parity-test cases, not bench lines -- see bench.py for the headline metric).

    python tools/run_configs.py [out.json] [--skip-big] [--tangent consistent]

--tangent consistent: the nlgeom configurations (2, 3, 3', 5) run with the opt-in exact tangent (full Newton) instead of the
reference's constant-C stiffness; the record then carries `tangent` and `tangent_fallbacks` (the reference traces stay beside it).

cfg 1  elliptic membrane CPS3 (golden deck)          linear, 1 increment
cfg 2  beam CPS6 large deformation (golden deck)      nlgeom Newton, 4 increments
cfg 2' elliptic membrane CPS8 (golden deck)           quadratic quad
cfg 3  twist plate C3D4 1 116 elements (golden, first 2 increments), user rotation BC
cfg 3' synthetic twist plate 44x6x66 cells = 104 544 C3D4, nlgeom, user rotation BC, 2 increments
cfg 4  synthetic unit cube 10 110 954 C3D4, linear solve to eps = 1e-3 (reference) and 1e-8
cfg 5  synthetic unit cube 998 250 C3D10, neo-Hookean, nlgeom Newton
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from femcy_b200 import meshgen  # noqa: E402
from femcy_b200.material_zoo import LinearIsotropic  # noqa: E402
from helpers import GoldenDeck, load_golden, rel_err, system_from_deck  # noqa: E402


TANGENT = "reference"


def run_deck(deck, golden=None, **kw):
    t0 = time.time()
    s = system_from_deck(deck, **kw)
    if TANGENT == "consistent" and s.geometric_nonlinear:
        s.set_tangent("consistent")
    t_setup = time.time() - t0
    t0 = time.time()
    s.solve(deck)
    s.ctx.sync()
    t_solve = time.time() - t0
    s.compute_strain_stress()
    u = s.dof.to_numpy()
    out = {"elements": int(s.body.np_elements.shape[0]), "dofs": int(s.N), "nnz": s.nnz, "setup_s": t_setup,
           "solve_s": t_solve, "increments": [(float(t), bool(c), int(n)) for t, c, n in s.inc_trace],
           "cg_iterations_total": s.cg_iters_total, "max_abs_u": float(np.abs(u).max()),
           "max_mises": float(s.mises_stress.to_numpy().max()), "tangent": getattr(s, "tangent", "reference"),
           "tangent_fallbacks": int(getattr(s, "tangent_fallbacks", 0))}
    if golden is not None:
        out["rel_err_u_vs_reference"] = rel_err(u, golden["dof_final"])
        out["reference_trace"] = [(float(t), bool(c), int(n)) for t, c, n, _ in golden["inc_trace"]]
    s.close()
    return out


def twist_deck(cells=(44, 6, 66), lengths=(80., 10., 120.), n_inc=2):
    """cfg 3': the real twist deck's box, clamp z=Lz, rotate z=0 about the axis through (40,5) (user BC)."""
    deck = meshgen.SyntheticDeck("C3D4", cells=cells, lengths=lengths, nlgeom=True,
                                 material=LinearIsotropic(2e11, 0.3), traction=0.0)
    nodes = deck.nodes
    top = np.nonzero(np.abs(nodes[:, 2] - lengths[2]) < 1e-9)[0]
    bot = np.nonzero(np.abs(nodes[:, 2]) < 1e-9)[0]
    deck.dirichlet_bc_info = ([{"node_set": top, "dof": c, "val": 0., "user": False} for c in range(3)] +
                              [{"node_set": bot, "dof": c, "val": 0., "user": True} for c in range(3)])
    deck.neumann_bc_info = []
    deck.time_incs = {"ini_inc": 0.05, "max_time": 0.05 * n_inc, "min_inc": 1e-5, "max_inc": 0.05}
    return deck


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/configs.json"
    skip_big = "--skip-big" in sys.argv
    global TANGENT
    if "--tangent" in sys.argv:
        TANGENT = sys.argv[sys.argv.index("--tangent") + 1]
        if TANGENT not in ("reference", "consistent"):
            raise SystemExit("--tangent reference | consistent")
    res = {}
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)

    def save():
        with open(out_path, "w") as fh:
            json.dump(res, fh, indent=1)

    for name, key in (("cfg1_cps3_ellip", "cps3_ellip"), ("cfg2_cps6_beam_largedef", "cps6_beam_largedef_newton"),
                      ("cfg2p_cps8_ellip", "cps8_ellip")):
        g = load_golden(key)
        res[name] = run_deck(GoldenDeck(g), g)
        print(name, json.dumps(res[name])[:300], flush=True)
        save()
    g = load_golden("c3d4_twist_2inc")
    d = GoldenDeck(g)
    d.time_incs["max_time"] = float(g["inc_trace"][-1, 0])
    res["cfg3_twist_c3d4_2inc"] = run_deck(d, g)
    print("cfg3", json.dumps(res["cfg3_twist_c3d4_2inc"])[:300], flush=True)
    res["cfg3p_twist_104544"] = run_deck(twist_deck())
    print("cfg3p", json.dumps(res["cfg3p_twist_104544"])[:400], flush=True)
    save()
    if not skip_big:
        deck = meshgen.SyntheticDeck("C3D4", n=119, jitter=0.1)
        res["cfg4_cube_10M_eps1e-3"] = run_deck(deck)           # reference eps (N >= 1e5)
        print("cfg4", json.dumps(res["cfg4_cube_10M_eps1e-3"])[:400], flush=True)
        res["cfg4_cube_10M_eps1e-8"] = run_deck(deck, cg_eps=1e-8)
        u3 = res["cfg4_cube_10M_eps1e-3"]["max_abs_u"]
        u8 = res["cfg4_cube_10M_eps1e-8"]["max_abs_u"]
        res["cfg4_note"] = f"max|u| at eps=1e-3 differs from eps=1e-8 by {abs(u3 - u8) / u8:.2e} relative (SURVEY H5)"
        save()
        deck5 = meshgen.SyntheticDeck("C3D10", n=55, nlgeom=True, traction=0.01,
                                      time_incs={"ini_inc": 1., "max_time": 1., "min_inc": 1e-5, "max_inc": 1.})
        res["cfg5_cube_c3d10_1M_neohookean"] = run_deck(deck5)
        print("cfg5", json.dumps(res["cfg5_cube_c3d10_1M_neohookean"])[:400], flush=True)
    save()
    print("written", out_path)


if __name__ == "__main__":
    main()
