#!/usr/bin/env python
"""Tier-1 oracle: execute the UNMODIFIED reference sources under the sequential taichi shim
and dump golden vectors (TEST INFRASTRUCTURE ONLY -- runs in the build container, where
/root/reference exists; its outputs are committed under tests/golden/).

What is executed is the reference's own code:
  reader.inp_info.InpInfo            (/root/reference/reader/inp_info.py:14)
  body.Body                          (/root/reference/body.py:12)
  stiffnessMtrx.System_of_equations  (/root/reference/stiffnessMtrx.py:19)
  conjugateGradientSolver.ConjugateGradientSolver_rowMajor (/root/reference/conjugateGradientSolver.py:8)

Only patches: GUI no-op, and the reductions that use ti.atomic_max/min on a
kernel-local scalar are replaced by NumPy one-liners with the same meaning.

Usage (cwd is switched to the reference root because of its `sys.path.append("./...")` hacks):
  python oracle/run_reference.py DECK.inp OUT.npz [--cg] [--no-solve] [--max-incs N]
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FEMCY_REFERENCE", "/root/reference")


def perturbation(nodes: np.ndarray) -> np.ndarray:
    """Deterministic smooth displacement field used to exercise the X+u code paths.
    Must stay in sync with tests/helpers.py::perturbation."""
    span = float((nodes.max(axis=0) - nodes.min(axis=0)).max())
    amp = 0.02 * span
    x = (nodes - nodes.min(axis=0)) / span
    dm = nodes.shape[1]
    u = np.zeros_like(nodes)
    for c in range(dm):
        phase = 1.3 * x[:, 0] + 0.7 * x[:, 1] + (0.4 * x[:, 2] if dm == 3 else 0.0)
        u[:, c] = amp * np.sin(2.1 * phase + 0.9 * c)
    return u.reshape(-1)


def ell_to_coo(sparseIJ: np.ndarray, vals: np.ndarray):
    cnt = sparseIJ[:, 0]
    rows, cols, v = [], [], []
    for i in range(sparseIJ.shape[0]):
        c = int(cnt[i])
        rows.append(np.full(c, i, dtype=np.int64))
        cols.append(sparseIJ[i, 1:1 + c].astype(np.int64))
        v.append(vals[i, :c])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    v = np.concatenate(v)
    order = np.lexsort((cols, rows))
    return rows[order], cols[order], v[order]


def deck_info(inp, system, material):
    """Everything a test needs to re-run the deck without the .inp file: BC sets, loads, material
    parameters, and the Neumann rhs at full load as the reference's own neumannBC computes it
    (stiffnessMtrx.py:369-411)."""
    out = {}
    nodes_cat, ptr, dofs, vals, users = [], [0], [], [], []
    for bc in inp.dirichlet_bc_info:
        ns = np.array([*bc["node_set"]], dtype=np.int32)
        nodes_cat.append(ns)
        ptr.append(ptr[-1] + len(ns))
        dofs.append(bc["dof"]); vals.append(bc["val"]); users.append(bool(bc["user"]))
    out["bc_nodes"] = np.concatenate(nodes_cat) if nodes_cat else np.zeros(0, np.int32)
    out["bc_ptr"] = np.array(ptr, dtype=np.int64)
    out["bc_dof"] = np.array(dofs, dtype=np.int32)
    out["bc_val"] = np.array(vals, dtype=np.float64)
    out["bc_user"] = np.array(users, dtype=bool)
    out["n_neumann"] = np.array(len(inp.neumann_bc_info))
    for k, nbc in enumerate(inp.neumann_bc_info):
        out[f"nm{k}_facets"] = np.array(sorted(nbc["face_set"]), dtype=np.int32)
        out[f"nm{k}_traction"] = np.array(float(nbc["traction"]))
        out[f"nm{k}_direction"] = np.array(nbc["direction"], dtype=np.float64) if "direction" in nbc else np.zeros(0)
    if hasattr(material, "modulus"):
        out["mat_params"] = np.array([material.modulus, material.poisson_ratio], dtype=np.float64)
    else:
        out["mat_params"] = np.array([material.C1, material.D1], dtype=np.float64)
    # Neumann rhs at full load (last *Dsload wins, quirk B1)
    system.rhs.fill(0.0)
    for nbc in inp.neumann_bc_info:
        if "direction" in nbc:
            system.neumannBC(nbc["face_set"], load_val=nbc["traction"], load_dir=nbc["direction"])
        else:
            system.neumannBC(nbc["face_set"], load_val=nbc["traction"])
    out["rhs_neumann"] = system.rhs.to_numpy()
    system.rhs.fill(0.0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("deck")
    ap.add_argument("out")
    ap.add_argument("--cg", action="store_true", help="force the reference CG branch in solve_dof")
    ap.add_argument("--no-solve", action="store_true", help="kernel-level vectors only")
    ap.add_argument("--max-incs", type=int, default=0, help="stop after N accepted increments (0 = all)")
    ap.add_argument("--augment", action="store_true", help="only add the deck description to an existing OUT.npz")
    args = ap.parse_args()

    deck = os.path.abspath(args.deck)
    out = os.path.abspath(args.out)
    sys.path.insert(0, os.path.join(HERE, "taichi_shim"))
    os.chdir(REF)
    sys.path.insert(0, REF)

    import taichi as ti  # the shim
    import tiGadgets as tg
    import conjugateGradientSolver as cgmod
    import stiffnessMtrx as sm
    from reader.inp_info import InpInfo
    from body import Body

    # ---- patches (documented in the module docstring) ----
    sm.System_of_equations.show_window = lambda *a, **k: None
    cg_calls = {"rmax": 0}

    def _rmax(self):
        cg_calls["rmax"] += 1
        return float(np.abs(self.r.a).max())

    cgmod.ConjugateGradientSolver_rowMajor.rmax = _rmax
    tg.field_abs_max = lambda f: float(np.abs(f.a).max())
    tg.field_max = lambda f: float(f.a.max())
    tg.field_min = lambda f: float(f.a.min())
    tg.vectorField_max = lambda f: float(f.a.max())

    ti.init(arch=ti.cpu, default_fp=ti.f64)
    t0 = time.time()
    inp = InpInfo(deck)
    nodes, eSets = inp.nodes, inp.eSets
    etype = list(eSets.keys())[0]
    elements = eSets[etype]
    body = Body(nodes=nodes, elements=elements, ELE=inp.ELE)
    material = list(inp.materials.values())[0]
    system = sm.System_of_equations(body, material, inp.geometric_nonlinear)
    if args.cg:
        sm.System_of_equations.solve_dof = sm.System_of_equations.solve_by_CG

    if args.augment:
        old = dict(np.load(out))
        old.update(deck_info(inp, system, material))
        np.savez_compressed(out, **old)
        print(f"[run_reference] augmented {out}")
        return

    res = {}
    res.update(deck_info(inp, system, material))
    res["deck"] = np.array(os.path.relpath(deck, REF))
    res["elem_type"] = np.array(etype)
    res["nodes"] = np.asarray(nodes, dtype=np.float64)
    res["elements"] = np.asarray(elements, dtype=np.int32)
    res["mat_class"] = np.array(type(material).__name__)
    res["mat_type"] = np.array(material.type)
    res["C"] = np.asarray(material.C, dtype=np.float64)
    res["nlgeom"] = np.array(bool(inp.geometric_nonlinear))
    res["time_incs"] = np.array([inp.time_incs[k] for k in ("ini_inc", "max_time", "min_inc", "max_inc")])
    sparseIJ = system.sparseIJ.to_numpy().astype(np.int64)

    # ---- kernel-level vectors: K at u = 0 and at a smooth perturbation ----
    system.dof.fill(0.0)
    system.get_dsdx_and_vol()
    system.assemble_stiffnessMtrx()
    r, c, v = ell_to_coo(sparseIJ, system.sparseMtrx_rowMajor.to_numpy())
    res["K_rows"], res["K_cols"], res["K0_vals"] = r.astype(np.int32), c.astype(np.int32), v
    res["vol0"] = system.vol.to_numpy()

    u1 = perturbation(res["nodes"])
    system.dof.from_numpy(u1)
    system.get_dsdx_and_vol()
    system.assemble_stiffnessMtrx()
    _, _, v1 = ell_to_coo(sparseIJ, system.sparseMtrx_rowMajor.to_numpy())
    res["u1"] = u1
    res["K1_vals"] = v1
    res["dsdx1"] = system.dsdx.to_numpy()
    res["vol1"] = system.vol.to_numpy()

    # stress recovery at the perturbed state, both constitutive modes (a8), internal force (a9)
    system.get_deformation_gradient()
    res["F1"] = system.F.to_numpy()
    material.constitutiveOfSmallDeform(system.F, system.cauchy_stress, system.ddsdde)
    res["cauchy_small1"] = system.cauchy_stress.to_numpy()
    mises = {"planeStrain": system.get_mises_stress_planeStrain, "planeStress": system.get_mises_stress_planeStress,
             "3d": system.get_mises_stress_3d}[material.type]
    mises()
    res["mises_small1"] = system.mises_stress.to_numpy()
    system.assemble_nodal_force_GN()
    res["cauchy_large1"] = system.cauchy_stress.to_numpy()
    res["nodal_force1"] = system.nodal_force.to_numpy()
    mises()
    res["mises_large1"] = system.mises_stress.to_numpy()
    system.get_elasEng()
    res["elsEng1"] = np.array(float(system.elsEng[None]))
    system.cauchy_stress.fill(0.0)
    system.dof.fill(0.0)

    # ---- Neumann rhs at full load, then the linear Dirichlet elimination on K0 (a4, a5) ----
    import copy
    neumannBCs = copy.deepcopy(inp.neumann_bc_info)
    dirichletBCs = copy.deepcopy(inp.dirichlet_bc_info)
    for bc in dirichletBCs:
        ns = ti.field(ti.i32, shape=(len(bc["node_set"])))
        ns.from_numpy(np.array([*bc["node_set"]]))
        bc["node_set"] = ns
    system.get_dsdx_and_vol()
    system.assemble_stiffnessMtrx()
    geo = system.geometric_nonlinear
    system.geometric_nonlinear = False          # take the linear-equation Dirichlet branch
    system.time1 = 1.0
    system.impose_boundary_condition({"neumannBCs": neumannBCs, "dirichletBCs": dirichletBCs})
    system.geometric_nonlinear = geo
    system.time1 = 0.0
    _, _, vbc = ell_to_coo(sparseIJ, system.sparseMtrx_rowMajor.to_numpy())
    res["Kbc_vals"] = vbc
    res["rhs_bc"] = system.rhs.to_numpy()
    system.rhs.fill(0.0)
    system.dof.fill(0.0)

    # ---- full solve through the reference driver ----
    if not args.no_solve:
        trace = []
        resid = []
        orig_adv = sm.System_of_equations.advance_inc
        orig_norm = tg.field_norm

        class _Stop(Exception):
            pass

        def adv(self, *a, **k):
            out_ = orig_adv(self, *a, **k)
            trace.append((self.time1, float(out_[0]), float(out_[1]), float(np.abs(self.dof.a).max())))
            if args.max_incs and sum(1 for t in trace if t[1] > 0) >= args.max_incs:
                raise _Stop()
            return out_

        def norm(f):
            val = orig_norm(f)
            resid.append(float(val))
            return val

        sm.System_of_equations.advance_inc = adv
        tg.field_norm = norm
        sm.tg.field_norm = norm
        try:
            system.solve(inp, show_newton_steps=False, save2path=None)
        except _Stop:
            pass
        res["dof_final"] = system.dof.to_numpy()
        res["inc_trace"] = np.array(trace, dtype=np.float64).reshape(-1, 4)
        res["residual_trace"] = np.array(resid, dtype=np.float64)
        res["cg_rmax_calls"] = np.array(cg_calls["rmax"])
        system.compute_strain_stress()
        res["cauchy_final"] = system.cauchy_stress.to_numpy()
        res["mises_final"] = system.mises_stress.to_numpy()
        system.get_elasEng()
        res["elsEng_final"] = np.array(float(system.elsEng[None]))

    res["wall_s"] = np.array(time.time() - t0)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **res)
    print(f"[run_reference] {deck} -> {out}  ({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
