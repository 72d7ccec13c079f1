"""Kernel-logic tests on the CPU SIMT emulation (tests/simt): the SOURCE of the product's CUDA kernels
(femcy_b200/csrc/*_kernels.cuh) is compiled with g++ against a fiber-based emulation of the CUDA execution model and
checked against the oracle -- indexing, barrier structure, reduction order, the peer-window protocol of the
multi-GPU CG.  This is test infrastructure: it proves nothing about performance or the GPU memory model, and the
`-m gpu` parity tests remain the parity gate.  No product code can reach the emulation.
"""
import numpy as np
import pytest

import simt
from femcy_b200 import meshgen
from femcy_b200.element_zoo import ELEMENT_TYPES
from femcy_b200.material_zoo import LinearIsotropic, LinearIsotropicPlaneStress, NeoHookean
from oracle import femcy_oracle as O


def _mesh2d(kind, n=5, seed=0):
    """small structured 2-D meshes of every 2-D element family (unit square, jittered interior)."""
    rng = np.random.default_rng(seed)
    quad = kind in ("CPS4", "CPS8")
    order2 = kind in ("CPS6", "CPS8")
    m = 2 * n + 1 if order2 else n + 1
    xs = np.linspace(0, 1, m)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    idx = np.arange(m * m).reshape(m, m)
    nodes = np.stack([X.ravel(), Y.ravel()], axis=1)
    conn = []
    s = 2 if order2 else 1
    for i in range(n):
        for j in range(n):
            a, b, c, d = idx[s * i, s * j], idx[s * i + s, s * j], idx[s * i + s, s * j + s], idx[s * i, s * j + s]
            if kind == "CPS3":
                conn += [[a, b, c], [a, c, d]]
            elif kind == "CPS4":
                conn += [[a, b, c, d]]
            elif kind == "CPS6":
                ab, bc, ca = idx[2 * i + 1, 2 * j], idx[2 * i + 2, 2 * j + 1], idx[2 * i + 1, 2 * j + 1]
                cd, da = idx[2 * i + 1, 2 * j + 2], idx[2 * i, 2 * j + 1]
                conn += [[a, b, c, ab, bc, ca], [a, c, d, ca, cd, da]]
            else:
                ab, bc = idx[2 * i + 1, 2 * j], idx[2 * i + 2, 2 * j + 1]
                cd, da = idx[2 * i + 1, 2 * j + 2], idx[2 * i, 2 * j + 1]
                conn += [[a, b, c, d, ab, bc, cd, da]]
    conn = np.array(conn, dtype=np.int32)
    used = np.unique(conn)
    lut = -np.ones(nodes.shape[0], dtype=np.int64)
    lut[used] = np.arange(used.size)
    nodes = nodes[used] + 0.02 / n * rng.standard_normal((used.size, 2))
    return nodes, lut[conn].astype(np.int32)


_CASES = [("C3D4", 4), ("C3D10", 2), ("CPS3", 6), ("CPS4", 5), ("CPS6", 4), ("CPS8", 4)]


def _case(kind, n):
    if kind.startswith("C3D"):
        deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1 if kind == "C3D4" else 0.0)
        return deck.nodes, deck.eSets[kind], deck.ELE, list(deck.materials.values())[0]
    nodes, conn = _mesh2d(kind, n)
    return nodes, conn, ELEMENT_TYPES[kind](), LinearIsotropicPlaneStress(modulus=2.0e5, poisson_ratio=0.3)


@pytest.mark.parametrize("kind,n", _CASES)
@pytest.mark.parametrize("variant", [1, 2, 3])
def test_emulated_assembly_matches_oracle(kind, n, variant):
    """variant 1 = atomic scatter (thread / warp per element), 2 = gather (default): node-sector records -- through the TMA
    tensor-store kernel for C3D4 -- then one thread per stored block accumulating the gradient products."""
    nodes, conn, ELE, mat = _case(kind, n)
    dm = nodes.shape[1]
    rng = np.random.default_rng(3)
    u = 0.01 * rng.standard_normal(nodes.size)
    pat = simt.SellPattern(conn, nodes.shape[0], dm=dm)
    Kref = O.assemble_K(nodes, conn.astype(np.int64), u, kind, np.asarray(mat.C))
    val, vol = simt.assemble(ELE, mat, nodes, conn, u, pat, variant=variant)
    assert not np.isnan(val).any()
    K = pat.to_csr(val)
    assert abs(K - Kref).max() <= 1e-12 * abs(Kref).max()
    _, vref = O.dsdx_and_vol(nodes, conn.astype(np.int64), u, kind)
    if variant in (2, 3):      # the gather (re)computes vol in its first pass (3 = thread-per-block pass 2 for 4-GP elements too)
        assert np.abs(vol - vref).max() <= 1e-13 * np.abs(vref).max()


def _linear_system(n=5, seed=1):
    deck = meshgen.SyntheticDeck("C3D4", n=n, jitter=0.1)
    conn, nodes = deck.eSets["C3D4"], deck.nodes
    mat = deck.materials["Elastic"]
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(nodes.size), "C3D4", np.asarray(mat.C))
    b = np.random.default_rng(seed).standard_normal(nodes.size)
    dofs = np.concatenate([bc["node_set"] * 3 + bc["dof"] for bc in deck.dirichlet_bc_info])
    Kbc, rbc = O.dirichlet_linear(K, b, dofs, np.zeros(len(dofs)))
    return nodes, conn, Kbc, rbc


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["three_kernel", "persistent", "streaming"])
def test_emulated_pcg_matches_oracle(nranks, mode):
    """same iteration count as the statement-for-statement oracle PCG and the same iterate; several ranks run
    concurrently and exchange halo values / partial sums through the emulated peer windows."""
    nodes, conn, K, b = _linear_system()
    xr, itr = O.pcg(K, b, eps=1e-8)
    systems = simt.split_system(nodes, conn, K, b, nranks, 3)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=8, mode=mode)
    x = simt.gather_solution(systems, nodes.size)
    assert it == itr
    assert rmax < 1e-8 * r0
    assert np.abs(x - xr).max() <= 1e-11 * np.abs(xr).max()


def test_emulated_pcg_fixed_iterations_and_first_iterates():
    nodes, conn, K, b = _linear_system(n=4)
    for k in (1, 2, 5):
        systems = simt.split_system(nodes, conn, K, b, 1, 3)
        it, _, _ = simt.cg_solve(systems, eps=1e-30, max_iter=k, check_every=4, fixed=True, mode=1)
        assert it == k
        # oracle iterate after k iterations
        M = 1.0 / K.diagonal()
        x = np.zeros_like(b); r = b.copy(); d = M * r
        for _ in range(k):
            Ad = K @ d
            rMr = np.dot(r * M, r)
            alpha = rMr / np.dot(d, Ad)
            x = x + alpha * d
            r = r - alpha * Ad
            d = M * r + (np.dot(r * M, r) / rMr) * d
        xe = simt.gather_solution(systems, nodes.size)
        assert np.abs(xe - x).max() <= 1e-12 * np.abs(x).max()


@pytest.mark.parametrize("mode", [1, 2], ids=["persistent", "streaming"])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
@pytest.mark.parametrize("eps,check_every", [(1e-3, 1), (1e-8, 8)])
def test_emulated_pcg_with_symmetric_half_storage(nranks, eps, check_every, mode):
    """option cg_sym: the persistent kernel's SpMV streams the upper half of the matrix (suffix j >= i of every sorted
    row, ghost columns included) and scatters the transposed products with atomics.  Same stopping iterate as the
    oracle PCG up to summation order; on several ranks the interface blocks are stored by both owners, so no
    contribution crosses ranks."""
    nodes, conn, K, b = _linear_system()
    xr, itr = O.pcg(K, b, eps=eps)
    systems = simt.split_system(nodes, conn, K, b, nranks, 3)
    it, r0, rmax = simt.cg_solve(systems, eps=eps, max_iter=2000, check_every=check_every, mode=mode, sym=1)
    x = simt.gather_solution(systems, nodes.size)
    assert abs(it - itr) <= 1 and rmax < eps * r0
    tol = 1e-10 if it == itr else 10 * eps          # one iteration more or less: only the stop rule's accuracy
    assert np.abs(x - xr).max() <= tol * np.abs(xr).max()


@pytest.mark.parametrize("nranks,mode", [(1, 1), (3, 1), (2, 2), (1, 2)])
def test_emulated_symmetric_half_storage_with_sigma_sorted_rows(nranks, mode):
    """upper-half SpMV on a SELL-32-sigma pattern (positions != row nodes: the suffix j >= i, the diagonal test and
    the scatter targets all go by node id)."""
    nodes, conn, K, b = _linear_system()
    xr, itr = O.pcg(K, b, eps=1e-8)
    systems = simt.split_system(nodes, conn, K, b, nranks, 3, sigma=64)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=8, mode=mode, sym=1)
    x = simt.gather_solution(systems, nodes.size)
    assert abs(it - itr) <= 1 and rmax < 1e-8 * r0
    assert np.abs(x - xr).max() <= (1e-9 if it == itr else 1e-7) * np.abs(xr).max()


def test_emulated_symmetric_half_storage_first_iterates_and_2d():
    """fixed iteration counts (no stop rule): the iterate after k iterations equals the default kernel's to rounding;
    also a 2-dof-per-node system (plane-stress quads)."""
    nodes, conn, K, b = _linear_system(n=4)
    for k in (1, 3, 8):
        xs = []
        for sym in (0, 1):
            systems = simt.split_system(nodes, conn, K, b, 2, 3)
            it, _, _ = simt.cg_solve(systems, eps=1e-30, max_iter=k, check_every=3, fixed=True, mode=1, sym=sym)
            assert it == k
            xs.append(simt.gather_solution(systems, nodes.size))
        assert np.abs(xs[0] - xs[1]).max() <= 1e-12 * np.abs(xs[0]).max()
    nodes, conn, ELE, mat = _case("CPS4", 5)
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(nodes.size), "CPS4", np.asarray(mat.C))
    rng = np.random.default_rng(4)
    bb = rng.standard_normal(nodes.size)
    fixed = np.flatnonzero(nodes[:, 0] < nodes[:, 0].min() + 0.1 * np.ptp(nodes[:, 0]))      # the left edge (jittered nodes)
    assert fixed.size >= 3
    dofs = np.concatenate([fixed * 2, fixed * 2 + 1])
    Kbc, rbc = O.dirichlet_linear(K, bb, dofs, np.zeros(len(dofs)))
    xr, itr = O.pcg(Kbc, rbc, eps=1e-8)
    systems = simt.split_system(nodes, conn, Kbc, rbc, 1, 2)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=3000, check_every=4, mode=1, sym=1)
    assert abs(it - itr) <= 1 and rmax < 1e-8 * r0
    assert np.abs(simt.gather_solution(systems, nodes.size) - xr).max() <= 1e-6 * np.abs(xr).max()


@pytest.mark.parametrize("variant", [1, 2])
def test_emulated_assembly_on_a_partition(variant):
    """rank-local assembly of the multi-GPU path: rows of the owned nodes only, ghost columns included; interface
    elements are integrated redundantly (no communication).  Every rank's rows must equal the global matrix's."""
    from femcy_b200.partition import Partition
    deck = meshgen.SyntheticDeck("C3D4", n=4, jitter=0.1)
    nodes, conn, mat = deck.nodes, deck.eSets["C3D4"], deck.materials["Elastic"]
    u = 0.01 * np.random.default_rng(5).standard_normal(nodes.size)
    Kref = O.assemble_K(nodes, conn.astype(np.int64), u, "C3D4", np.asarray(mat.C)).tocsr()
    for rank in range(3):
        part = Partition(nodes, conn, rank, 3)
        pat = simt.SellPattern(part.elements, part.n_local, nn_own=part.n_own, dm=3)
        gd = (part.local_to_global[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        val, _ = simt.assemble(deck.ELE, mat, part.nodes, part.elements, u[gd], pat, variant=variant)
        K = pat.to_csr(val)
        Kloc = Kref[gd[: part.n_own * 3]][:, gd]
        assert abs(K - Kloc).max() <= 1e-12 * abs(Kref).max()


# ---- SELL-32-sigma (option sell_sigma) --------------------------------------------------
def test_sigma_sorting_removes_the_padding_of_quadratic_meshes():
    """C3D10 rows alternate between 65-block corner nodes and 14..42-block mid-edge nodes: ~40 % padding in natural
    order, a few % when rows are sorted by length inside windows of 256 nodes."""
    nodes, conn = meshgen.kuhn_box_c3d10(8)
    p0 = simt.SellPattern(conn, nodes.shape[0], dm=3)
    p1 = simt.SellPattern(conn, nodes.shape[0], dm=3, sigma=256)
    assert p0.nnzb == p1.nnzb
    pad0, pad1 = 1 - p0.nnzb / p0.nslots, 1 - p1.nnzb / p1.nslots
    assert pad0 > 0.3 and pad1 < 0.12, (pad0, pad1)
    assert np.array_equal(np.sort(p1.rowof[: nodes.shape[0]]), np.arange(nodes.shape[0]))


@pytest.mark.parametrize("kind,n,variant", [("C3D10", 2, 1), ("C3D10", 2, 2), ("C3D4", 4, 1), ("C3D4", 4, 2), ("CPS6", 4, 2),
                                            ("CPS8", 4, 2), ("CPS3", 6, 2)])
def test_emulated_assembly_with_sigma_sorted_rows(kind, n, variant):
    nodes, conn, ELE, mat = _case(kind, n)
    dm = nodes.shape[1]
    u = 0.01 * np.random.default_rng(3).standard_normal(nodes.size)
    pat = simt.SellPattern(conn, nodes.shape[0], dm=dm, sigma=64)
    Kref = O.assemble_K(nodes, conn.astype(np.int64), u, kind, np.asarray(mat.C))
    val, _ = simt.assemble(ELE, mat, nodes, conn, u, pat, variant=variant)
    assert not np.isnan(val).any()
    assert abs(pat.to_csr(val) - Kref).max() <= 1e-12 * abs(Kref).max()


@pytest.mark.parametrize("nranks,mode", [(1, 0), (1, 1), (2, 0), (3, 1), (2, 2), (1, 2)])
def test_emulated_pcg_with_sigma_sorted_rows(nranks, mode):
    nodes, conn, K, b = _linear_system()
    xr, itr = O.pcg(K, b, eps=1e-8)
    systems = simt.split_system(nodes, conn, K, b, nranks, 3, sigma=64)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=8, mode=mode)
    x = simt.gather_solution(systems, nodes.size)
    assert it == itr
    assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()


@pytest.mark.parametrize("sigma", [0, 64])
@pytest.mark.parametrize("mode", [0, 1])
def test_emulated_dirichlet_matches_oracle(sigma, mode):
    """k_bc_mark + k_bc_apply (bc.cu) against the oracle's sequential restatement of stiffnessMtrx.py:279-341."""
    deck = meshgen.SyntheticDeck("C3D4", n=4, jitter=0.1)
    nodes, conn, mat = deck.nodes, deck.eSets["C3D4"], deck.materials["Elastic"]
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(nodes.size), "C3D4", np.asarray(mat.C))
    rng = np.random.default_rng(2)
    b = rng.standard_normal(nodes.size)
    bn = np.concatenate([bc["node_set"] for bc in deck.dirichlet_bc_info]).astype(np.int32)
    bc_ = np.concatenate([np.full(len(bc["node_set"]), bc["dof"]) for bc in deck.dirichlet_bc_info]).astype(np.int32)
    bv = 0.01 * rng.standard_normal(bn.size)
    dofs = bn.astype(np.int64) * 3 + bc_
    pat = simt.SellPattern(conn, nodes.shape[0], dm=3, sigma=sigma)
    val = pat.from_csr(K)
    target = b.copy()
    simt.dirichlet(pat, val, target, bn, bc_, bv, mode)
    if mode == 0:
        Kr, rr = O.dirichlet_linear(K, b, dofs, bv)
    else:
        Kr, rr = O.dirichlet_newton(K, b, dofs)
    assert abs(pat.to_csr(val) - Kr).max() <= 1e-13 * abs(K).max()
    assert np.abs(target - rr).max() <= 1e-12 * np.abs(rr).max()


# ---- stress recovery / internal force / energy (post.cu) against the reference's own goldens ----------------
_POST_DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "cpe3_cook", "cpe6_cook", "c3d4_ellip",
               "c3d10_ellip", "c3d4_neohookean_newton"]


@pytest.mark.parametrize("name", _POST_DECKS)
def test_emulated_stress_and_force_kernels_match_reference_goldens(name):
    """the same sequence as tests/test_gpu_parity.py::test_geometry_and_stress_kernels, on the emulator."""
    from helpers import abs_err_scaled, load_golden, make_element, make_material, rel_err
    g = load_golden(name)
    ELE, mat = make_element(g), make_material(g)
    p = simt.Post(ELE, mat, g["nodes"], g["elements"], g["u1"])
    assert rel_err(p.deformation_gradient(), g["F1"]) < 1e-13
    assert rel_err(p.constitutive(False), g["cauchy_small1"]) < 1e-12
    assert abs_err_scaled(p.mises(), g["mises_small1"], np.abs(g["cauchy_small1"]).max()) < 1e-12
    f = p.internal_force()
    assert rel_err(p.cauchy, g["cauchy_large1"]) < 1e-12
    assert rel_err(f, g["nodal_force1"].reshape(-1)) < 1e-11
    assert rel_err(p.vol, g["vol1"]) < 1e-12 and rel_err(p.dsdx, g["dsdx1"]) < 1e-12
    assert abs_err_scaled(p.mises(), g["mises_large1"], np.abs(g["cauchy_large1"]).max()) < 1e-12
    p.deformation_gradient()
    _, tot = p.energy()
    assert abs(tot - float(g["elsEng1"])) <= 1e-11 * abs(float(g["elsEng1"]))


# ---- pattern build kernels (pattern.cu) against the NumPy statement of the layout ---------------------------------
@pytest.mark.parametrize("kind,n,sigma,own", [("C3D4", 4, 0, 1.0), ("C3D4", 4, 64, 1.0), ("C3D10", 2, 0, 1.0), ("C3D10", 3, 64, 1.0),
                                              ("CPS6", 4, 32, 1.0), ("C3D4", 4, 0, 0.6), ("C3D4", 4, 32, 0.6)])
def test_emulated_pattern_build_matches_layout_statement(kind, n, sigma, own):
    """k_elem_keys ... k_entry_slots, k_sigma_keys/k_rowpos in the order build_from_keys runs them (CUB sorts replaced
    by std::stable_sort) == tests/simt.SellPattern, array for array;
    own < 1: only the first rows are owned (rank-local pattern of the multi-GPU path)."""
    nodes, conn, ELE, mat = _case(kind, n)
    nn = nodes.shape[0]
    nn_own = int(nn * own)
    ref = simt.SellPattern(conn, nn, nn_own=nn_own, dm=nodes.shape[1], sigma=sigma)
    got = simt.build_pattern(conn, nn, nn_own=nn_own, sigma=sigma)
    assert (got["nnzb"], got["nslots"], got["nslice"], got["max_row_blocks"]) == (ref.nnzb, ref.nslots, ref.nslice, ref.max_row_blocks)
    for k in ("blkptr", "slice_ptr", "colidx", "diag_slot", "slot_beg", "slot_end", "elem_slot", "ent_list"):
        assert np.array_equal(got[k], getattr(ref, k)), k
    if sigma:
        assert np.array_equal(got["rowof"], ref.rowof) and np.array_equal(got["rowpos"][:nn_own], ref.rowpos[:nn_own])


def test_emulated_gp_sum():
    """femcy_gp_sum's kernel (k_weighted_sum without weights): the e2e leg of bench.py reads the mesh volume back."""
    import ctypes as C
    L = simt.lib()
    a = np.random.default_rng(0).random(5000)
    partials, ticket, out = np.zeros(64), np.zeros(8, dtype=np.uint32), np.zeros(1)
    L.emu_gp_sum(simt._p(a, C.c_double), C.c_int64(a.size), simt._p(partials, C.c_double), simt._p(ticket, C.c_uint32),
                 simt._p(out, C.c_double))
    assert abs(out[0] - a.sum()) <= 1e-12 * a.sum()
    assert ticket[0] == 0


def test_emulation_suite_under_shuffled_thread_schedule():
    """re-run a cross-section of this file with SIMT_SHUFFLE (fibers visited in random order every scheduling round):
    a kernel that only works because lower thread ids happen to run first -- i.e. a missing barrier -- fails here."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SIMT_SHUFFLE="12345")
    sel = ("test_emulated_assembly_matches_oracle or test_emulated_dirichlet or persistent-2 or streaming-2 "
           "or test_emulated_gp_sum or test_emulated_pcg_with_symmetric_half_storage")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", sel, "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_emulated_kernels_under_address_sanitizer():
    """every assembly variant on four element families, the persistent PCG kernels on 1-3 ranks and the pattern build,
    with the emulation library compiled with -fsanitize=address: an out-of-bounds index in a kernel aborts with a report."""
    import os
    import shutil
    import subprocess
    import sys
    gcc = shutil.which("gcc")
    asan = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() if gcc else ""
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = os.path.join(simt.HERE, "_build", "libfemcy_simt_asan.so")
    deps = [os.path.join(simt.HERE, f) for f in ("simt.h", "emu_entry.cpp")] + \
           [os.path.join(simt.CSRC, f) for f in os.listdir(simt.CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-fsanitize=address", "-fno-omit-frame-pointer",
                               "-I", simt.HERE, "-o", so, os.path.join(simt.HERE, "emu_entry.cpp"), "-lpthread"])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, os.path.join(simt.HERE, "asan_check.py"), so], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ASAN_CHECK_PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("nranks,cg_mode", [(2, 1), (3, 2), (4, 2)])
def test_emulated_partitioned_pipeline_matches_global_solve(nranks, cg_mode):
    """the numerical pipeline of the multi-GPU path, kernels only: every emulated rank assembles the rows of its owned nodes
    (library default: slice-major gather; interface elements redundantly, no communication), eliminates the Dirichlet dofs
    it sees (owned rows, owned + ghost columns, non-zero prescribed values), then the ranks run the persistent PCG kernel
    concurrently over the peer windows; the gathered solution must be the global oracle solve."""
    from femcy_b200.body import Body
    from femcy_b200.neumann import neumann_vector
    from femcy_b200.partition import Partition
    deck = meshgen.SyntheticDeck("C3D4", n=5, jitter=0.1)
    nodes, conn, mat = deck.nodes, deck.eSets["C3D4"], deck.materials["Elastic"]
    N = nodes.size
    nb = deck.neumann_bc_info[0]
    rhs = neumann_vector(Body(nodes, conn, deck.ELE), nb["face_set"], nb["traction"], nb["direction"])
    bn = np.concatenate([bc["node_set"] for bc in deck.dirichlet_bc_info]).astype(np.int64)
    bc_ = np.concatenate([np.full(len(bc["node_set"]), bc["dof"]) for bc in deck.dirichlet_bc_info]).astype(np.int64)
    bv = 1e-3 * np.sin(np.arange(bn.size))                       # non-zero prescribed displacements
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(N), "C3D4", np.asarray(mat.C))
    Kbc, rbc = O.dirichlet_linear(K, rhs, bn * 3 + bc_, bv)
    x_ref, it_ref = O.pcg(Kbc, rbc, eps=1e-9)

    parts = [Partition(nodes, conn, r, nranks) for r in range(nranks)]
    systems = []
    for p in parts:
        sysm = simt.RankSystem.__new__(simt.RankSystem)
        sysm.part, sysm.dm = p, 3
        sysm.pat = simt.SellPattern(p.elements, p.n_local, nn_own=p.n_own, dm=3)
        gd = (p.local_to_global[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
        sysm.val, _ = simt.assemble(deck.ELE, mat, p.nodes, p.elements, np.zeros(p.n_local * 3), sysm.pat, variant=2)
        b = np.zeros(p.n_local * 3)
        b[: p.n_own * 3] = rhs[gd[: p.n_own * 3]]
        loc = p.global_to_local[bn]                              # Dirichlet nodes present on this rank (owned or ghost)
        keep = loc >= 0
        simt.dirichlet(sysm.pat, sysm.val, b, loc[keep], bc_[keep], bv[keep], 0)
        sysm.b = b
        sysm.gdofs_own = gd[: p.n_own * 3]
        sysm.vecs = {k: np.zeros(p.n_local * 3) for k in "xrdMA"}
        sysm.scal, sysm.partials = np.zeros(64), np.zeros(4096)
        sysm.ticket, sysm.window = np.zeros(8, dtype=np.uint32), np.zeros(simt.WINDOW_WORDS, dtype=np.uint64)
        systems.append(sysm)
    for s_ in systems:
        s_.plan(systems)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-9, max_iter=5000, check_every=8, mode=cg_mode)
    x = simt.gather_solution(systems, N)
    assert abs(it - it_ref) <= 1
    assert np.abs(x - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    assert np.abs(x[bn * 3 + bc_] - bv).max() <= 1e-8 * np.abs(bv).max()      # identity rows, solved to the PCG tolerance


# ---- unstructured meshes (irregular valence: ragged rows, uneven tiles) --------------------------------------------
def _delaunay_tets(npts=260, seed=0):
    """random points in the unit cube, Delaunay tetrahedra with the reference's orientation (det > 0), slivers removed."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((npts, 3))
    tets = Delaunay(pts).simplices.astype(np.int64)
    ELE = ELEMENT_TYPES["C3D4"]()
    _, vol = O.dsdx_and_vol(pts, tets, np.zeros(pts.size), "C3D4")
    flip = vol[:, 0] < 0
    tets[flip] = tets[flip][:, [1, 0, 2, 3]]
    _, vol = O.dsdx_and_vol(pts, tets, np.zeros(pts.size), "C3D4")
    assert (vol > 0).all()
    keep = vol[:, 0] > 1e-7
    tets = tets[keep]
    used = np.unique(tets)
    lut = -np.ones(npts, dtype=np.int64)
    lut[used] = np.arange(used.size)
    return pts[used], lut[tets].astype(np.int32), ELE


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("sigma", [0, 64])
def test_emulated_assembly_on_a_delaunay_mesh(variant, sigma):
    """every C3D4 assembly variant on an unstructured mesh (node valence 4..40: ragged rows, uneven element tiles),
    natural and sigma-sorted row order."""
    nodes, conn, ELE = _delaunay_tets()
    mat = LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3)
    u = 1e-3 * np.random.default_rng(2).standard_normal(nodes.size)
    pat = simt.SellPattern(conn, nodes.shape[0], dm=3, sigma=sigma)
    assert pat.max_row_blocks > 20 and np.diff(pat.blkptr).min() < 10          # genuinely ragged
    Kref = O.assemble_K(nodes, conn.astype(np.int64), u, "C3D4", np.asarray(mat.C))
    val, _ = simt.assemble(ELE, mat, nodes, conn, u, pat, variant=variant)
    assert not np.isnan(val).any()
    assert abs(pat.to_csr(val) - Kref).max() <= 1e-12 * abs(Kref).max()


@pytest.mark.parametrize("nranks,mode", [(1, 1), (3, 2), (2, 2)])
def test_emulated_pcg_on_a_delaunay_mesh(nranks, mode):
    nodes, conn, ELE = _delaunay_tets(npts=200, seed=3)
    mat = LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3)
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(nodes.size), "C3D4", np.asarray(mat.C))
    fixed = np.flatnonzero(nodes[:, 0] < 0.15)
    dofs = (fixed[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    b = np.random.default_rng(4).standard_normal(nodes.size)
    Kbc, rbc = O.dirichlet_linear(K, b, dofs, np.zeros(dofs.size))
    xr, itr = O.pcg(Kbc, rbc, eps=1e-8)
    systems = simt.split_system(nodes, conn, Kbc, rbc, nranks, 3, sigma=32)
    it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=5000, check_every=8, mode=mode)
    x = simt.gather_solution(systems, nodes.size)
    # sliver tetrahedra make this system ill-conditioned: the stopping iteration moves by a few with the summation order
    # of the SpMV (SELL slices vs CSR rows), so the checks are the stop rule itself and the residual of the returned x
    assert abs(it - itr) <= max(3, itr // 20)
    assert rmax < 1e-8 * r0
    assert np.abs(rbc - Kbc @ x).max() <= 2e-8 * np.abs(rbc).max()
    assert np.abs(x - xr).max() <= 1e-4 * np.abs(xr).max()


@pytest.mark.parametrize("sigma", [0, 64])
def test_emulated_pattern_build_on_a_delaunay_mesh(sigma):
    nodes, conn, _ = _delaunay_tets()
    nn = nodes.shape[0]
    ref = simt.SellPattern(conn, nn, dm=3, sigma=sigma)
    got = simt.build_pattern(conn, nn, sigma=sigma)
    for k in ("blkptr", "slice_ptr", "colidx", "diag_slot", "slot_beg", "slot_end", "elem_slot", "ent_list"):
        assert np.array_equal(got[k], getattr(ref, k)), k
    if sigma:
        assert np.array_equal(got["rowof"], ref.rowof)


@pytest.mark.parametrize("mode", [1, 2], ids=["persistent", "streaming"])
@pytest.mark.parametrize("nranks,check_every", [(2, 1), (2, 32), (3, 8), (4, 8)])
def test_emulated_pcg_consecutive_solves_share_windows_and_sequence_numbers(nranks, check_every, mode):
    """three solves in a row on the SAME rank systems: the peer windows, the halo flags and the exchange tags continue from
    S_SEQ of the solve before (block 0's stop decision costs one more d update + halo push + SpMV after the deciding iteration:
    every rank must leave the loop with flags == S_SEQ, or the next solve reads stale halos / matches stale tags)"""
    nodes, conn, K, b = _linear_system()
    b2 = np.roll(b, 7) * 0.5 + b
    b2[K.diagonal() == 1.0] = 0.0
    refs = [O.pcg(K, rhs, eps=1e-8) for rhs in (b, b2, b)]
    systems = simt.split_system(nodes, conn, K, b, nranks, 3)
    for rep, rhs in enumerate((b, b2, b)):
        for s_old, s_new in zip(systems, simt.split_system(nodes, conn, K, rhs, nranks, 3)):
            s_old.b[:] = s_new.b
        it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=check_every, mode=mode)
        x = simt.gather_solution(systems, nodes.size)
        assert it == refs[rep][1] and rmax < 1e-8 * r0
        assert np.abs(x - refs[rep][0]).max() <= 1e-11 * np.abs(refs[rep][0]).max()
        seqs = {float(s.scal[11]) for s in systems}                       # S_SEQ
        assert len(seqs) == 1
        assert all(int(s.window[r]) == int(next(iter(seqs))) for s in systems for r in range(nranks))      # every halo flag == S_SEQ


@pytest.mark.parametrize("mode", [1, 2], ids=["persistent", "streaming"])
def test_emulated_pcg_bench_sequence_on_several_ranks(mode):
    """bench.py's order of solves under a partition: fixed-iteration steps (exit disabled), then the parity solve to eps = 1e-8,
    then fixed steps again -- all on the same windows"""
    nodes, conn, K, b = _linear_system()
    xr, itr = O.pcg(K, b, eps=1e-8)
    for nranks in (2, 4):
        systems = simt.split_system(nodes, conn, K, b, nranks, 3)
        for _ in range(2):
            assert simt.cg_solve(systems, eps=1e-30, max_iter=20, check_every=20, fixed=True, mode=mode)[0] == 20
        it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=2000, check_every=32, mode=mode)
        x = simt.gather_solution(systems, nodes.size)
        assert it == itr and np.abs(x - xr).max() <= 1e-11 * np.abs(xr).max()
        assert simt.cg_solve(systems, eps=1e-30, max_iter=20, check_every=20, fixed=True, mode=mode)[0] == 20


@pytest.mark.parametrize("mode", [1, 2], ids=["persistent", "streaming"])
def test_emulated_pcg_on_eight_ranks(mode):
    """the 8-GPU configuration of the scaling run: 8 emulated ranks (interior ranks have two neighbours, every rank exchanges
    partial sums with all seven others), two solves in a row; same iteration count and iterate as the oracle PCG"""
    nodes, conn, K, b = _linear_system(n=9)
    xr, itr = O.pcg(K, b, eps=1e-8)
    systems = simt.split_system(nodes, conn, K, b, 8, 3)
    for _ in range(2):
        it, r0, rmax = simt.cg_solve(systems, eps=1e-8, max_iter=3000, check_every=16, mode=mode)
        x = simt.gather_solution(systems, nodes.size)
        assert it == itr and rmax < 1e-8 * r0
        assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()
