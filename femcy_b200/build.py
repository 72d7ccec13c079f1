"""Build libfemcy_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

    python -m femcy_b200.build          # or: from femcy_b200.build import build; build()

nvcc cross-compiles without a GPU.  The .so lands next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfemcy_b200.so")
SOURCES = ["core.cu", "pattern.cu", "assembly.cu", "cg.cu", "precond.cu", "bc.cu", "post.cu", "comm.cu", "topology.cu", "partition.cu"]
HEADERS = ["ctx.cuh", "elem_math.cuh", "constitutive.cuh", "device_compat.cuh", "kernel_types.cuh", "assembly_kernels.cuh", "cg_kernels.cuh", "bc_kernels.cuh", "post_kernels.cuh", "pattern_kernels.cuh", "topology_kernels.cuh", "partition_kernels.cuh",
           os.path.join("..", "..", "include", "femcy_b200.h")]
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libfemcy_b200.so")
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + FLAGS + ["-c", s, "-o", o]
            log = open(o + ".log", "w")
            procs.append((src, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log, o))
    failed = []
    for src, p, log, o in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            failed.append((src, open(o + ".log").read()))
        elif verbose:
            print(open(o + ".log").read())
    if failed:
        for src, txt in failed:
            sys.stderr.write(f"---- {src} ----\n{txt}\n")
        raise RuntimeError("nvcc failed for: " + ", ".join(s for s, _ in failed))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
