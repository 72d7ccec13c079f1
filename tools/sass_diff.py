"""Do the default kernels still compile to the same instructions as in an earlier revision?  (no GPU needed)

    python tools/sass_diff.py <git-rev> [kernel-name-substring ...]

Compiles assembly.cu and cg.cu of <git-rev> and of the working tree to cubins (nvcc, sm_100a), extracts the SASS of every
kernel whose mangled name contains one of the substrings (`a|b`: either spelling, for a kernel whose template list
changed; default: the kernels on the default path) and compares the instruction text (addresses and encodings stripped).  Used before a round's GPU run to show that work on the opt-in
variants left the hardware-verified default kernels untouched."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-cubin"]
DEFAULT = ["k_elem_geometryILi3ELi4E", "k_assemble_gatherILi3ELi4E", "k_assemble_scatterILi3ELi4ELi1ELi1E",
           "k_assemble_scatter_warpILi3ELi10ELi4E", "k_dsdx_volILi3ELi4ELi1E",
           "k_cg_persistentILi3EEv|k_cg_persistentILi3ELi6EEv|k_cg_persistentILi3ELi6ELb0EEv", "k_spmv_dotILi3ELb0", "k_spmv_dotILi3ELb1", "k_update_xr", "k_update_d_p2pILi3E", "k_cg_initILi3E"]


def sass(cubin):
    txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append(re.sub(r"\s+", " ", m.group(1)).strip())
    return out


def build(src_root, tmp, tag):
    res = {}
    for f in ("assembly.cu", "cg.cu"):
        cub = os.path.join(tmp, f"{tag}_{f}.cubin")
        subprocess.check_call(["nvcc"] + FLAGS + ["-o", cub, f], cwd=os.path.join(src_root, "femcy_b200", "csrc"),
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        res.update(sass(cub))
    return res


def main(rev, pats):
    with tempfile.TemporaryDirectory() as tmp:
        old_root = os.path.join(tmp, "old")
        os.makedirs(old_root)
        ar = subprocess.run(["git", "archive", rev, "femcy_b200/csrc", "include"], cwd=ROOT, capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", old_root], input=ar, check=True)
        old, new = build(old_root, tmp, "old"), build(ROOT, tmp, "new")
    same = True
    for pat in pats:
        alts = pat.split("|")                     # a kernel whose template list changed: any of the alternatives, either side
        o = {k: v for k, v in old.items() if any(a in k for a in alts)}
        n = {k: v for k, v in new.items() if any(a in k for a in alts)}
        if not o or not n:
            print(f"{pat:44s} old: {len(o)} kernel(s), new: {len(n)} kernel(s) -- not comparable by this name")
            continue
        (ko, vo), (kn, vn) = sorted(o.items())[0], sorted(n.items())[0]
        eq = vo == vn
        same &= eq
        print(f"{pat:44s} {len(vo):6d} -> {len(vn):6d} instructions  {'IDENTICAL' if eq else 'DIFFERENT'}")
    return 0 if same else 1


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    sys.exit(main(sys.argv[1], sys.argv[2:] or DEFAULT))
