#!/usr/bin/env python
"""Regenerate the golden vectors in this directory from the reference itself.

Each fixture is produced by executing the unmodified reference sources
(/root/reference) under the sequential taichi shim -- see oracle/run_reference.py.
Runs only in the build container (the GPU box has no /root/reference); the .npz
outputs are committed.

  python tests/golden/make_golden.py            # all fixtures, 8 in parallel
  python tests/golden/make_golden.py cps3_ellip # a subset
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FEMCY_REFERENCE", "/root/reference")

# name -> (deck path relative to /root/reference/tests, extra runner flags)
FIXTURES = {
    "cps3_ellip": ("elliptic_membrane/element_linear/ellip_membrane_linEle_localVeryFine.inp", []),
    "cps6_ellip": ("elliptic_membrane/element_quadratic/ellip_membrane_quadritic_trig_neumann.inp", []),
    "cps4_ellip": ("elliptic_membrane/element_quadrilateral/ellip_CPS4.inp", []),
    "cps8_ellip": ("elliptic_membrane/element_quadrilateral/ellip_CPS8.inp", []),
    "cpe3_cook": ("cook_membrane/smallDef_linearEl/cookMembrane_2d_linearEl.inp", []),
    "cpe3_cook_nu4999": ("cook_membrane/smallDef_linearEl/nu0.4999/cookMembrane_2d_linearEl.inp", []),
    "cpe6_cook": ("cook_membrane/smallDef_quadEl/cook_membrane_2d.inp", []),
    "c3d4_ellip": ("elliptic_membrane/3D/linearEl/ellip_membrane_3d_linearEl.inp", []),
    "c3d10_ellip": ("elliptic_membrane/3D/quadEl/ellip_membrane_3d.inp", []),
    "c3d4_cook": ("cook_membrane/3D/smallDef_linerEl_coarse/cook_3d_linearEl_smallDef.inp", []),
    "c3d10_cook": ("cook_membrane/3D/smallDef_qualEl_coarse/cook_3d_quadEl_smallDef.inp", []),
    "cps3_dirforce_4inc": ("elliptic_membrane/directional_force/ellip_localVeryFine_directional_force.inp", []),
    "cps3_bydisp_4inc": ("elliptic_membrane/load_by_disp/ellip_membrane_localFine_dirichlet.inp", []),
    "c3d4_neohookean_newton": ("cook_membrane/3D/neo-Hookean/cook_3d_linearEl_largeDef.inp", []),
    "cpe3_cook_largedef_newton": ("cook_membrane/largeDef_linearEl/cookMembrane_2d_linearEl.inp", []),
    "cps3_dense_cg": ("elliptic_membrane/very_dense/ellip_dense_CPS3_0d04.inp", ["--cg"]),
    "c3d4_twist_2inc": ("twist/twist_plate_C3D4.inp", ["--max-incs", "2"]),
    "cps6_beam_largedef_newton": ("beam_deflection/load800_freeEnd_largeDef/beamDeflec_quadPSE_largeD_load800.inp", []),
}


AUGMENT = bool(os.environ.get("GOLDEN_AUGMENT"))


def run(name):
    deck, flags = FIXTURES[name]
    if AUGMENT:
        flags = ["--augment"]
    out = os.path.join(HERE, name + ".npz")
    log = os.path.join("/tmp", f"golden_{name}.log")
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"),
           os.path.join(REF, "tests", deck), out] + flags
    with open(log, "w") as fh:
        rc = subprocess.call(cmd, stdout=fh, stderr=subprocess.STDOUT)
    print(name, "rc =", rc, flush=True)
    return rc


if __name__ == "__main__":
    names = sys.argv[1:] or list(FIXTURES)
    with ThreadPoolExecutor(max_workers=int(os.environ.get("GOLDEN_JOBS", "6"))) as ex:
        rcs = list(ex.map(run, names))
    sys.exit(max(rcs))
