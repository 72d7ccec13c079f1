"""Headless driver: the reference's `main.py` without `input()` prompts and GUI windows.

    python -m femcy_b200.main path/to/deck.inp [--device 0] [--stress 1] [--save out.npz] [--vtk out.vtk] [--quiet]

Same sequence as `/root/reference/main.py:21-80`: read the deck, build `Body` and `System_of_equations`,
`solve`, elastic energy, strain/stress recovery, then print the figures the reference prints (max Mises at
the integration points, max |dof|, max nodal (extrapolated) Mises, and the chosen stress component).
"""
import argparse
import time

import numpy as np

from . import Body, InpInfo, System_of_equations
from .tiGadgets import field_abs_max

_STRESS_ID_2D = {0: (0, 0), 1: (1, 1), 2: (0, 1)}
_STRESS_ID_3D = {0: (0, 0), 1: (1, 1), 2: (2, 2), 3: (0, 1), 4: (2, 0), 5: (1, 2)}   # Voigt order of main.py:66-74


def run(file_name, device=0, stress_index=None, save=None, quiet=False, vtk=None):
    inp = InpInfo(file_name)
    if len(inp.sections) > 1:
        return run_sections(inp, device, save, quiet, vtk)
    body = Body(nodes=inp.nodes, elements=list(inp.eSets.values())[0], ELE=inp.ELE)
    material = list(inp.materials.values())[0]
    system = System_of_equations(body, material, inp.geometric_nonlinear, device=device, quiet=quiet)
    t0 = time.time()
    system.solve(inp, show_newton_steps=False, save2path=None)
    t1 = time.time()
    dof = system.dof.to_numpy()
    print(f"system.dof = \n{dof}, time for finite element computing is {t1 - t0} s")
    system.get_elasEng()
    print(f"total elastic energy is {float(system.elsEng)}")
    system.compute_strain_stress()
    mises = system.mises_stress.to_numpy()
    print(f"max mises_stress at integration point is {mises.max()} MPa; max dof (disp) = {field_abs_max(system.dof)}")
    # extrapolation to the element nodes + nodal averaging run on the device (femcy_extrapolate, row f3)
    nodal, nodal_mean = system.mises_stress.extrapolate_on_device(system.ELE.extrapolation_matrix(), nn=body.np_nodes.shape[0])
    print(f"max nodal mises_stress = {nodal.max()}")
    out = {"dof": dof, "mises": mises, "nodal_mises": nodal, "cauchy": system.cauchy_stress.to_numpy(),
           "elastic_energy": float(system.elsEng), "inc_trace": np.array(system.inc_trace, dtype=float)}
    if stress_index is not None:
        ids = _STRESS_ID_2D if system.dm == 2 else _STRESS_ID_3D
        i, j = ids[stress_index]
        comp = out["cauchy"][:, :, i, j]
        print(f"maximum stress[{(i, j)}] = {np.abs(comp).max()} MPa; max nodal stress{(i, j)} = "
              f"{system.ELE.extrapolate(comp).max()}")
    if save:
        np.savez_compressed(save, **out)
    if vtk:
        from .vtk import write_vtk
        write_vtk(vtk, body, point_data={"U": dof, "mises": nodal_mean},
                  cell_data={"mises_gp_mean": mises.mean(axis=1)})
    system.close()
    return out


def run_sections(inp, device=0, save=None, quiet=False, vtk=None):
    """Row f4: a deck with several element types and / or `*Solid Section` materials (the reference stops at
    `reader/inp_info.py:125-128`).  Same sequence; per-Gauss-point results come back as one array per section."""
    body, _ = inp.sectioned_body()
    system = System_of_equations(body, None, inp.geometric_nonlinear, device=device, quiet=quiet)
    t0 = time.time()
    system.solve(inp, show_newton_steps=False, save2path=None)
    dof = system.dof.to_numpy()
    print(f"system.dof = \n{dof}, time for finite element computing is {time.time() - t0} s")
    system.get_elasEng()
    print(f"total elastic energy is {float(system.elsEng)}")
    system.compute_strain_stress()
    mises = system.mises_stress.to_numpy()
    print(f"max mises_stress at integration point is {max(m.max() for m in mises)} MPa; max dof (disp) = {field_abs_max(system.dof)}")
    out = {"dof": dof, "elastic_energy": float(system.elsEng), "inc_trace": np.array(system.inc_trace, dtype=float)}
    for k, sec in enumerate(inp.sections):
        nodal, _ = system.mises_stress[k].extrapolate_on_device(sec["ELE"].extrapolation_matrix())
        print(f"section {k} ({sec['etype']}, {sec['material_name']}): {len(sec['elements'])} elements, "
              f"max mises {mises[k].max()}, max nodal mises {nodal.max()}")
        out[f"mises_{k}"], out[f"nodal_mises_{k}"] = mises[k], nodal
        out[f"cauchy_{k}"] = system.cauchy_stress[k].to_numpy()
    if save:
        np.savez_compressed(save, **out)
    if vtk:
        from .vtk import write_vtk
        write_vtk(vtk, body, point_data={"U": dof}, cell_data={"mises_gp_mean": [m.mean(axis=1) for m in mises]})
    system.close()
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("deck", help="Abaqus / CalculiX .inp file")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--stress", type=int, default=None,
                    help="stress component to report: 2-D 0:xx 1:yy 2:xy; 3-D 0:xx 1:yy 2:zz 3:xy 4:zx 5:yz")
    ap.add_argument("--save", default=None, help="write dof / stresses to this .npz")
    ap.add_argument("--vtk", default=None, help="write mesh + displacement + nodal Mises to this legacy-VTK file")
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args(argv)
    run(args.deck, args.device, args.stress, args.save, args.quiet, args.vtk)


if __name__ == "__main__":
    main()
