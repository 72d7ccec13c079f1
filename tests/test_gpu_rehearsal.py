"""CPU rehearsal of GPU test files: the `-m gpu` tests of the rows added last (f1 topology / Neumann, f4 sections, f2 consistent
tangent) run UNCHANGED against emu_ctx.EmuContext -- every C-ABI call answered by the product's kernel source on the CPU SIMT
emulation (tests/emu_plugin.py) -- in a subprocess, so that a slip in the test code or in the host layer cannot wait for the GPU box
to be found.  Excluded: the tests of the C library's own error messages, the 1.3 M-element mesh and the device partitioner's
`Context(device)` (covered on the emulation by tests/test_partition.py).  The GPU run stays the parity gate for the shipped binary."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gpu_test_files_pass_on_the_emulated_context():
    cmd = [sys.executable, "-m", "pytest", "test_gpu_sections.py", "test_gpu_topology.py", "test_gpu_kernels.py", "-m", "gpu", "-p", "emu_plugin",
           "-q", "-x", "-k", "(sections or topology or consistent_tangent or neumann) and not errors and not large_mesh and not device_partitioner"]
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.strip().split("\n")[-15:])
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail
