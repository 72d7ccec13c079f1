"""Static SASS instruction mix of the hot kernels (no GPU needed): cuobjdump -sass on the built library.

    python tools/sass_mix.py profiles/r1z_sass_mix.md [regex ...]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "femcy_b200", "libfemcy_b200.so")
DEFAULT = [r"k_assemble_scatterILi3ELi4ELi1ELi1", r"k_assemble_scatter_warpILi3ELi10", r"k_elem_geometryILi3ELi4",
           r"k_assemble_gatherILi3ELi4", r"k_elem_geometry4sILi3ELi4", r"k_assemble_gather4ILi3ELi4ELi1ELb0",
           r"k_assemble_gather4ILi3ELi4ELi1ELb1", r"k_assemble_rowsILi3ELi4ELi1ELi0", r"k_assemble_rowsILi3ELi4ELi1ELi1",
           r"k_assemble_rowsILi3ELi10ELi4ELi1", r"k_assemble_gather4ILi3ELi10ELi4ELb1",
           r"k_spmv_dotILi3ELb0", r"k_cg_persistentILi3", r"k_cg_streamILi3", r"k_assemble_gather_[hqp]", r"k_elem_geometry4t", r"k_update_xr", r"k_update_d_p2pILi3"]


def main(out, patterns):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            ins = m.group(1).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            funcs[cur].append(ins.split()[0])
    rows = ["# SASS instruction mix (static counts; `cuobjdump -sass femcy_b200/libfemcy_b200.so`, sm_100a)", "",
            "No tensor-core (`UTC*MMA`/`HMMA`) or TMA (`UTMALDG`/`UBLKCP`) instructions on the path: fp64 gather/scatter work "
            "(DESIGN.md section 4).  `REDG.E.ADD.F64` = fire-and-forget fp64 atomics of the scatter assembly; `LDG.E.EF*` = "
            "evict-first loads of the matrix stream; `*.STRONG.SYS` / `MEMBAR.SC.SYS` = the NVLink peer-memory exchange; "
            "`CCTL.PF*` / `LDG...LTC` = software prefetch of the rows kernels.", "",
            "| kernel | instructions | DFMA/DMUL/DADD | LDG | STG | REDG | LDS/STS | BAR | notable opcodes |", "|---|---|---|---|---|---|---|---|---|"]
    for pat in patterns:
        for name, ins in funcs.items():
            if not re.search(pat, name):
                continue
            c = collections.Counter(i.split(".")[0] for i in ins)
            full = collections.Counter(ins)
            notable = sorted(k for k in full if re.search(r"STRONG\.SYS|MEMBAR|\.EF|REDG|ATOM|CCTL|PREFETCH|UBLKCP|UTMA|LTC|LDGSTS|LDGDEPBAR", k))
            rows.append(f"| `{name[:60]}` | {len(ins)} | {c['DFMA']}/{c['DMUL']}/{c['DADD']} | {c['LDG']} | {c['STG']} | "
                        f"{c['REDG']} | {c['LDS']}/{c['STS']} | {c['BAR']} | {', '.join(notable[:8])} |")
    open(out, "w").write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:] or DEFAULT)
