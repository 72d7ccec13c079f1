"""Run a cross-section of the emulated kernels under AddressSanitizer (TEST INFRASTRUCTURE, see simt.h).

Invoked by tests/test_simt_kernels.py::test_emulated_kernels_under_address_sanitizer as a subprocess with
LD_PRELOAD=libasan: an out-of-bounds access of any kernel on the NumPy-owned buffers, the shared-memory arrays or the
fiber stacks aborts the process with an ASan report."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import ctypes as C
import simt
simt._libs[""] = C.CDLL(sys.argv[1])
import numpy as np
from test_simt_kernels import _case, _linear_system
from oracle import femcy_oracle as O
ok = True
for kind, n in (("C3D4", 4), ("C3D10", 2), ("CPS3", 6), ("CPS8", 4)):
    nodes, conn, ELE, mat = _case(kind, n)
    u = 0.01 * np.random.default_rng(3).standard_normal(nodes.size)
    Kref = O.assemble_K(nodes, conn.astype(np.int64), u, kind, np.asarray(mat.C))
    ngp = ELE.device_tables()[0].shape[0]
    for v in (1, 2):
        pat = simt.SellPattern(conn, nodes.shape[0], dm=nodes.shape[1])
        val, _ = simt.assemble(ELE, mat, nodes, conn, u, pat, variant=v)
        err = abs(pat.to_csr(val) - Kref).max() / abs(Kref).max()
        print(kind, v, "%.1e" % err); ok &= err < 1e-12
from test_simt_kernels import _delaunay_tets
from femcy_b200.material_zoo import LinearIsotropic
dn, dc, dELE = _delaunay_tets()
dmat = LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3)
du = 1e-3 * np.random.default_rng(2).standard_normal(dn.size)
dK = O.assemble_K(dn, dc.astype(np.int64), du, "C3D4", np.asarray(dmat.C))
for sigma in (0, 64):
    for v in (1, 2):
        pat = simt.SellPattern(dc, dn.shape[0], dm=3, sigma=sigma)
        val, _ = simt.assemble(dELE, dmat, dn, dc, du, pat, variant=v)
        err = abs(pat.to_csr(val) - dK).max() / abs(dK).max()
        print("delaunay", sigma, v, "%.1e" % err); ok &= err < 1e-12
nodes, conn, K, b = _linear_system(n=4)
for nr, mode in ((1, 1), (2, 1), (1, 2), (3, 2), (2, 0)):
    systems = simt.split_system(nodes, conn, K, b, nr, 3)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=8, mode=mode)
    print("cg", nr, mode, it); ok &= rmax < 1e-8 * r0
for nr, mode in ((1, 1), (3, 2)):
    systems = simt.split_system(nodes, conn, K, b, nr, 3)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=8, mode=mode, sym=1)
    print("cg sym", nr, mode, it); ok &= rmax < 1e-8 * r0
got = simt.build_pattern(conn, nodes.shape[0], sigma=64)
print("pattern ok", got["nnzb"], "all ok", ok)
assert ok
print("ASAN_CHECK_PASSED")
