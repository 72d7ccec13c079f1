"""`System_of_equations`: the reference's FE core object, re-hosted on the B200 library.

Same constructor, method names and state attributes as `/root/reference/stiffnessMtrx.py:19-845`
so the driver code reads like the reference's (`main.py:26-41`); every hot method is one call
into `libfemcy_b200.so`:

  get_dsdx_and_vol            :132  -> femcy_get_dsdx_and_vol
  assemble_stiffnessMtrx      :161  -> femcy_assemble_K       (geometry fused, no row scans)
  dirichletBC_linearEquations :279  -> femcy_dirichlet_linear
  dirichletBC_forNewtonMethod :310  -> femcy_dirichlet_val + femcy_dirichlet_newton
  neumannBC                   :369  -> host NumPy (as in the reference), vectorised, one upload
  solve_dof / solve_by_CG     :272/:254 -> femcy_cg_solve (always the CUDA PCG, see `solve_dof`)
  compute_strain_stress       :436  -> femcy_deformation_gradient / _strain / _constitutive / _mises
  assemble_nodal_force_GN     :609  -> femcy_internal_force
  get_elasEng                 :592  -> femcy_elastic_energy
  solve / advance_inc         :647/:714 -> host control flow, transcribed decision for decision
                                     (incl. the increment quirks SURVEY H11 / App. B14-B15)

State lives on the GPU; attributes such as `dof`, `rhs`, `mises_stress` are field objects with
`to_numpy()/from_numpy()/fill()/copy_from()` like Taichi fields.  There is no CPU fallback.
"""
import copy
import ctypes as C
import os
import time
from typing import Tuple

import numpy as np

from . import tiGadgets as tg
from . import user_defined as ud
from ._lib import Context, as_d, as_i32
from .body import Body, SectionedBody
from .fields import DeviceGPArray, DeviceVector, HostField, SectionedGPField
from .neumann import neumann_vector

# below this many dofs the reference calls a direct solver (stiffnessMtrx.py:272-276); we always
# run the CUDA PCG and use a tight tolerance there so the answer is direct-solve quality
_DIRECT_SIZE = 1e5


class _DeviceTopology:
    """Row f1: the topology queries of `Body` answered by the library (topology.cu) while the system's context lives."""

    def __init__(self, system):
        self.system = system

    def alive(self):
        return getattr(self.system.ctx, "h", True) is not None and not self.system.sectioned

    def boundary_facets(self):
        s = self.system
        n = C.c_int64(0)
        s.ctx.call("femcy_boundary_facets", C.byref(n))
        ele, kid = np.empty(max(n.value, 1), dtype=np.int32), np.empty(max(n.value, 1), dtype=np.int32)
        s.ctx.call("femcy_get_boundary_facets", as_i32(ele), as_i32(kid))
        ele, kid = ele[: n.value], kid[: n.value]
        if s.element_perm is not None:                 # device element k is the caller's element perm[k]
            ele = np.asarray(s.element_perm)[ele]
            order = np.argsort(kid.astype(np.int64) * s.body.np_elements.shape[0] + ele, kind="stable")
            ele, kid = ele[order], kid[order]
        return ele, kid

    def node_elements(self):
        s = self.system
        ne, n_en = s.body.np_elements.shape
        ptr, lst = np.empty(s.body.np_nodes.shape[0] + 1, dtype=np.int32), np.empty(max(ne * n_en, 1), dtype=np.int32)
        s.ctx.call("femcy_node_elements", as_i32(ptr), as_i32(lst))
        lst = lst[: ne * n_en]
        if s.element_perm is not None:
            lst = np.asarray(s.element_perm)[lst]
            for i in range(len(ptr) - 1):               # (reordered runs only: keep every node's list ascending)
                lst[ptr[i]:ptr[i + 1]].sort()
        return ptr, lst


class System_of_equations:
    def __init__(self, body: Body, material, geometric_nonlinear: bool, device: int = 0,
                 cg_eps: float = None, assembly_variant: int = 0, quiet: bool = False,
                 partition=None, reorder=False):
        self.dm = body.dm
        self.geometric_nonlinear = geometric_nonlinear
        self.body = body
        self.elements, self.nodes = body.elements, body.nodes
        self.ELE = body.ELE
        # row f4: a SectionedBody carries one (element kind, material) per section; `material` may then be None (the
        # sections' own), one material for all sections, or a list with one per section
        self.sectioned = isinstance(body, SectionedBody) and len(body.parts) > 1
        if isinstance(body, SectionedBody):
            mats = list(material) if isinstance(material, (list, tuple)) else [material] * len(body.parts)
            mats = [m if m is not None else bm for m, bm in zip(mats, body.materials)]
            if len(mats) != len(body.parts) or any(m is None for m in mats):
                raise ValueError("every section needs a material")
            self.materials = mats
            material = mats[0]
        else:
            self.materials = [material]
        if self.sectioned and (partition is not None or reorder):
            raise NotImplementedError("a mesh of several sections runs on one GPU, in the caller's element order")
        self.material = material
        self.C = material.C
        self.quiet = quiet
        self.cg_eps = cg_eps
        # 0 = the library's default (gather), 1 = atomic scatter, 2 = gather (include/femcy_b200.h: femcy_assemble_K);
        # FEMCY_OPT_ASSEMBLY_VARIANT selects one for A/B runs without touching caller code
        self.assembly_variant = assembly_variant or int(os.environ.get("FEMCY_OPT_ASSEMBLY_VARIANT", "0"))
        self.partition = partition
        self.comm = None if partition is None else partition.comm

        nn, ne = body.np_nodes.shape[0], body.np_elements.shape[0]
        n_en = body.np_elements.shape[1]
        nn_own = nn if partition is None else partition.n_own
        self.N = nn * self.dm
        self.N_own = nn_own * self.dm
        # size of the GLOBAL system: every rank must take the same solver decisions (eps, iteration bound),
        # as the reference does on its single global dof count (stiffnessMtrx.py:272-276)
        self.N_global = self.N if partition is None else int(partition.nn_global) * self.dm
        self.ctx = ctx = Context(device)
        nodes = np.ascontiguousarray(body.np_nodes, dtype=np.float64)
        # optional device-side element order (Z-order curve of the centroids); host-visible element
        # numbering is unchanged (per-element outputs are permuted back on read).  OFF by default:
        # measured on B200 it makes the atomic scatter 1.3-1.7x SLOWER (spatially compact elements run
        # concurrently and collide on the same K entries; profiles/r1_notes.md) -- kept for experiments.
        self.element_perm = None
        if reorder is True:
            from .meshgen import locality_order
            self.element_perm = locality_order(body.np_nodes, body.np_elements)
        conn = body.np_elements if self.element_perm is None else body.np_elements[self.element_perm]
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        ctx.call("femcy_set_mesh", self.dm, nn, nn_own, as_d(nodes), ne, n_en, as_i32(conn))
        dN, w = self.ELE.device_tables()
        self.n_gp = len(w)
        ctx.call("femcy_set_element", self.n_gp, as_d(dN), as_d(w))
        self._upload_material()
        # row f1: facet tables of the element kind -> Neumann vector and boundary facets on the device
        kn, fw, fn, fN, fdN = self.ELE.device_facet_tables()
        ctx.call("femcy_set_facet_tables", kn.shape[0], kn.shape[1], fw.shape[1], as_i32(kn), as_d(fw), as_d(fn), as_d(fN), as_d(fdN))
        if not self.sectioned:
            body._device_topology = _DeviceTopology(self)
        self.parts = [body]
        if self.sectioned:
            self.parts = body.parts
            for k in range(1, len(body.parts)):
                part = body.parts[k]
                conn_k = np.ascontiguousarray(part.np_elements, dtype=np.int32)
                sec = C.c_int(0)
                ctx.call("femcy_add_section", conn_k.shape[0], conn_k.shape[1], as_i32(conn_k), C.byref(sec))
                assert sec.value == k
                dN_k, w_k = part.ELE.device_tables()
                ctx.call("femcy_set_element", len(w_k), as_d(dN_k), as_d(w_k))
                self._upload_material(k)
            ctx.call("femcy_select_section", 0)
        nnz = C.c_int64(0)
        ctx.call("femcy_build_pattern", C.byref(nnz))
        self.nnz = int(nnz.value)
        if partition is not None:
            partition.install(ctx)

        # ---- fields (names of stiffnessMtrx.py:33-66,95,113) ----
        self.rhs = DeviceVector(ctx, "rhs", self.N)
        self.dof = DeviceVector(ctx, "dof", self.N)
        self.nodal_force = DeviceVector(ctx, "nodal_force", self.N)
        self.residual_nodal_force = DeviceVector(ctx, "residual", self.N)
        self.du = DeviceVector(ctx, "du", self.N)
        self.dof_old = DeviceVector(ctx, "dof_old", self.N)
        self._x = DeviceVector(ctx, "x", self.N)
        g, d = self.n_gp, self.dm

        def gp_field(name, tail):
            """[ne, n_gp, *tail] device array; several sections: one per section behind a SectionedGPField"""
            if not self.sectioned:
                return DeviceGPArray(ctx, name, (ne, g) + tail(n_en), self.element_perm)
            return SectionedGPField([DeviceGPArray(ctx, name, (p.np_elements.shape[0], p.ELE.n_gp) + tail(p.ELE.n_en), None, section=k)
                                     for k, p in enumerate(self.parts)])

        self.F = gp_field("F", lambda m: (d, d))
        self.cauchy_stress = gp_field("cauchy", lambda m: (d, d))
        self.strain = gp_field("strain", lambda m: (d, d))
        self.mises_stress = gp_field("mises", lambda m: ())
        self.elsEngDens = gp_field("energy", lambda m: ())
        self.dsdx = gp_field("dsdx", lambda m: (m, d))
        self.vol = gp_field("vol", lambda m: ())
        self.elsEng = HostField(np.zeros(()))
        self.visualize_field = HostField(np.zeros((ne, g)))
        self.nodal_vals = HostField(np.zeros((ne, n_en)))

        self.time0 = 0.
        self.time1 = 0.
        self.dt = 0.
        self.compiled = False
        self.last_cg_iters = 0
        self.cg_iters_total = 0
        self.preconditioner = "jacobi"
        if os.environ.get("FEMCY_OPT_CG_PRECOND") == "1":          # A/B hook of the tools (see _lib.Context)
            self.set_preconditioner("two_level")

    # ------------------------------------------------------------------------------------------
    def _say(self, *a):
        if not self.quiet:
            print(*a)

    def _upload_material(self, section=None):
        """hand the material (of one section; None = section 0 / the only one) to the library"""
        m = self.materials[section or 0]
        Cm = np.ascontiguousarray(np.asarray(m.C, dtype=np.float64))
        p = np.ascontiguousarray(m.device_params(), dtype=np.float64)
        if section is not None and self.sectioned:
            self.ctx.call("femcy_select_section", int(section))
        self.ctx.call("femcy_set_material", int(m.kind), as_d(p), len(p), as_d(Cm), Cm.shape[0])

    def _sync_ghosts(self):
        """multi-GPU: refresh the ghost entries of `dof` from their owners (no-op on one GPU)."""
        if self.partition is not None and self.partition.nranks > 1:
            from ._lib import VEC
            self.ctx.call("femcy_halo_exchange", VEC["dof"])

    def _field_norm(self, f):
        """tiGadgets.field_norm (RMS over ALL dofs of the global mesh)."""
        if self.partition is None or self.partition.nranks == 1:
            return tg.field_norm(f)
        sumsq = float(self.ctx.norms(f.name)[2])
        tot, _ = self.comm.allreduce_sum_max(sumsq, 0.0)
        return float((tot / (self.partition.nn_global * self.dm)) ** 0.5)

    def ddsdde_init(self):
        """ddsdde is the constant tangent C at every Gauss point (stiffnessMtrx.py:124-129): the
        kernels read C from the constant bank instead of a 288 B/GP field."""
        for k in range(len(self.materials) if self.sectioned else 1):
            self._upload_material(k if self.sectioned else None)

    # ---- hot kernels ----------------------------------------------------------------------------
    def get_dsdx_and_vol(self):
        self.ctx.call("femcy_get_dsdx_and_vol")

    def assemble_stiffnessMtrx(self):
        self.ctx.call("femcy_assemble_K", int(self.assembly_variant))

    assemble_stiffnessMtrx_faster = assemble_stiffnessMtrx

    def assemble_sparseMtrx(self):
        self.assemble_stiffnessMtrx()

    # ---- matrix export (comparison / interoperability) -------------------------------------------
    def csr(self):
        """scipy CSR copy of the current K (owned rows; sorted columns)."""
        import scipy.sparse as sp
        rp = np.empty(self.N_own + 1, dtype=np.int32)
        ci = np.empty(self.nnz, dtype=np.int32)
        v = np.empty(self.nnz, dtype=np.float64)
        self.ctx.call("femcy_get_csr_pattern", as_i32(rp), as_i32(ci))
        self.ctx.call("femcy_get_K_csr_values", as_d(v))
        return sp.csr_matrix((v, ci, rp), shape=(self.N_own, self.N))

    @property
    def sparseIJ(self):
        """The reference's ELL index array [N, W+1] (count in column 0, -1 padding;
        stiffnessMtrx.py:78-89), exported from the device pattern on demand."""
        K = self.csr()
        cnt = np.diff(K.indptr)
        W = int(cnt.max())
        ij = -np.ones((K.shape[0], W + 1), dtype=np.int32)
        ij[:, 0] = cnt
        pos = np.arange(K.nnz) - np.repeat(K.indptr[:-1], cnt)
        ij[np.repeat(np.arange(K.shape[0]), cnt), pos + 1] = K.indices
        return HostField(ij, dtype=np.int32)

    @property
    def sparseMtrx_rowMajor(self):
        K = self.csr()
        cnt = np.diff(K.indptr)
        out = np.zeros((K.shape[0], int(cnt.max())))
        pos = np.arange(K.nnz) - np.repeat(K.indptr[:-1], cnt)
        out[np.repeat(np.arange(K.shape[0]), cnt), pos] = K.data
        return HostField(out)

    # ---- preconditioner (row f2; opt-in: the reference's Jacobi stays the default) ------------------
    def set_preconditioner(self, kind="jacobi", max_coarse_unknowns=6000):
        """"jacobi": the reference's diagonal preconditioner (conjugateGradientSolver.py:48-51).  "two_level": Chebyshev-
        Jacobi smoothing + a rigid-body-mode coarse space over geometric node aggregates (csrc/precond.cu) -- one GPU only;
        same stopping rule, 15-50x fewer iterations on elasticity meshes; the iterates are not the reference's."""
        if kind == "jacobi":
            self.ctx.set_option("cg_precond", 0)
        elif kind == "two_level":
            if self.partition is not None and self.partition.nranks > 1:
                raise RuntimeError("the two-level preconditioner is single-GPU only")
            from .precond import geometric_aggregates
            agg, nagg = geometric_aggregates(self.body.np_nodes, max_coarse_unknowns)
            self.ctx.call("femcy_set_aggregates", nagg, as_i32(np.ascontiguousarray(agg)))
            self.ctx.set_option("cg_precond", 1)
            self.n_aggregates = nagg
        else:
            raise ValueError("preconditioner: 'jacobi' or 'two_level'")
        self.preconditioner = kind

    # ---- tangent (row f2; opt-in: the reference's constant-C stiffness stays the default) ---------------------------
    def set_tangent(self, kind="reference"):
        """"reference": K = sum B^T C B vol with the constant C of the material on the current configuration -- what the
        reference assembles at every Newton step (ddsdde is never updated, material_zoo/neo_hookean.py:62-64), i.e. a modified
        Newton iteration.  "consistent": the exact linearisation of the internal force (material + geometric stiffness, the
        tangent differentiated from the constitutive law itself, csrc/assembly_kernels.cuh: k_assemble_scatter_ct) -- full
        Newton: fewer loops to the same converged solution; the iterates are not the reference's."""
        if kind not in ("reference", "consistent"):
            raise ValueError("tangent: 'reference' or 'consistent'")
        if kind == "consistent" and not self.geometric_nonlinear:
            # a linear analysis IS its stiffness matrix: another K is another answer (neo-Hookean: the exact tangent at rest has
            # half the shear stiffness of the reference's C), not a faster way to the same one
            raise ValueError("the consistent tangent is for geometrically non-linear (Newton) analyses")
        self.ctx.set_option("consistent_tangent", 1 if kind == "consistent" else 0)
        self.tangent = kind
        self.tangent_fallbacks = 0

    # ---- linear solves ---------------------------------------------------------------------------
    def solve_by_CG(self, eps=None, max_iter=None, check_every=None, fixed_iters=False, _retry=False):
        """ConjugateGradientSolver_rowMajor.re_init()+solve() on the device
        (stiffnessMtrx.py:254-269, conjugateGradientSolver.py:32-127)."""
        b = "residual" if self.geometric_nonlinear else "rhs"
        if eps is None:
            eps = self.cg_eps
        if eps is None:
            eps = 1.0e-3 if self.N_global >= _DIRECT_SIZE else 1.0e-10
        if max_iter is None:
            # reference bound: b.shape[0] iterations (conjugateGradientSolver.py:109); the
            # direct-solve-quality mode gets more room
            max_iter = self.N_global if self.N_global >= _DIRECT_SIZE else 50 * self.N_global + 1000
        if check_every is None:
            check_every = 32
        it, r0, r1 = C.c_int64(0), C.c_double(0.), C.c_double(0.)
        from ._lib import VEC
        self.ctx.call("femcy_cg_solve", VEC[b], float(eps), int(max_iter), int(check_every), 1 if fixed_iters else 0,
                      C.byref(it), C.byref(r0), C.byref(r1))
        self.last_cg_iters = int(it.value)
        self.cg_iters_total += self.last_cg_iters
        self.last_cg_residuals = (r0.value, r1.value)
        self.last_cg_breakdown = bool(self.ctx.cg_breakdown())
        if (getattr(self, "tangent", "reference") == "consistent" and self.geometric_nonlinear and not fixed_iters and not _retry
                and (self.last_cg_breakdown or not (r1.value < eps * r0.value)) and getattr(self, "_newton_bcs", None) is not None):
            # the exact tangent is not positive definite at this state (e.g. a St. Venant-Kirchhoff solid under compression):
            # CG has no solution for it.  This increment continues with the reference's constant-C stiffness, which always is
            # (`solve` switches back once an increment has converged).
            self._say("\033[31;1m the consistent tangent is not positive definite here: this increment continues with the reference tangent \033[0m")
            self.tangent_fallbacks += 1
            self._tangent_on_hold = True
            self.ctx.set_option("consistent_tangent", 0)
            self.assemble_stiffnessMtrx()
            for bc in self._newton_bcs:
                self.dirichletBC_forNewtonMethod_kernel(nodeSet=bc["node_set"], dm_specified=bc["dof"], sval=bc["val"])
            return self.solve_by_CG(eps, max_iter, check_every, fixed_iters, _retry=True)
        if self.last_cg_breakdown:
            # K not positive definite / NaN (a diverged Newton step): no meaningful solution exists for CG.  The reference's
            # driver recovers from such steps through its NaN test (stiffnessMtrx.py:790-793): hand it NaN
            self._say("\033[31;1m PCG broke down (K not positive definite or NaN) after {} iterations \033[0m".format(it.value))
            self._x.fill(float("nan"))
        if not fixed_iters and not (r1.value < eps * r0.value) and r0.value > 0:
            self._say(f"\033[31;1m PCG stopped after {it.value} iterations with max|r|/max|r0| = "
                      f"{r1.value / r0.value:.3e} (target {eps:.1e}) \033[0m")
        if not self.geometric_nonlinear:
            self.dof.copy_from(self._x)                      # self.dof = self.PCG.x   (:264)
        else:
            tg.c_equals_a_minus_b(self.dof, self.dof, self._x)  # dof -= x            (:267)
        self._sync_ghosts()
        return self._x

    def solve_by_scipy(self):
        """The reference's host direct-solve branch (stiffnessMtrx.py:219-251), kept callable for
        users who ask for it explicitly.  It is never selected by `solve_dof`."""
        import scipy.sparse.linalg as sl
        if self.partition is not None:
            raise RuntimeError("solve_by_scipy is single-GPU only")
        K = self.csr().tocsr()
        b = self.residual_nodal_force if self.geometric_nonlinear else self.rhs
        self._x.from_numpy(sl.spsolve(K, b.to_numpy()))
        if not self.geometric_nonlinear:
            self.dof.copy_from(self._x)
        else:
            tg.c_equals_a_minus_b(self.dof, self.dof, self._x)
        return self._x

    def solve_dof(self):
        """The reference switches to scipy below 1e5 dofs (:272-276); this path always runs the
        CUDA PCG -- with the reference's eps (1e-3) at and above that size, and a tight eps below
        it where the reference result is a direct solve."""
        return self.solve_by_CG()

    # ---- boundary conditions ---------------------------------------------------------------------
    @staticmethod
    def _bc_arrays(nodeSet, dm_specified, sval=None):
        nodes = np.ascontiguousarray(nodeSet.to_numpy() if hasattr(nodeSet, "to_numpy") else nodeSet, dtype=np.int32).reshape(-1)
        comps = np.full(nodes.size, int(dm_specified), dtype=np.int32)
        vals = None if sval is None else np.full(nodes.size, float(sval), dtype=np.float64)
        return nodes, comps, vals

    def dirichletBC_linearEquations(self, nodeSet, dm_specified: int, sval: float):
        n, c, v = self._bc_arrays(nodeSet, dm_specified, sval)
        # a node listed twice in one set is applied once (same value): drop repeats
        n, idx = np.unique(n, return_index=True)
        self.ctx.call("femcy_dirichlet_linear", as_i32(np.ascontiguousarray(n)), as_i32(np.ascontiguousarray(c[idx])),
                      as_d(np.ascontiguousarray(v[idx])), n.size)

    def dirichletBC_forNewtonMethod(self, dirichletBCs):
        for bc in dirichletBCs:
            self.dirichletBC_dof(bc["node_set"], bc["dof"], bc["val"], bc["user"], self.time1)
            self.dirichletBC_forNewtonMethod_kernel(nodeSet=bc["node_set"], dm_specified=bc["dof"], sval=bc["val"])

    def dirichletBC_forNewtonMethod_kernel(self, nodeSet, dm_specified: int, sval: float):
        n, c, _ = self._bc_arrays(nodeSet, dm_specified)
        n, idx = np.unique(n, return_index=True)
        self.ctx.call("femcy_dirichlet_newton", as_i32(np.ascontiguousarray(n)), as_i32(np.ascontiguousarray(c[idx])), n.size)

    def dirichletBC_dof(self, nodeSet, dm_specified: int, sval: float, user: bool, time: float):
        if not user:
            self.dirichletBC_val(nodeSet, dm_specified, sval)
        else:
            ns = nodeSet.to_numpy() if hasattr(nodeSet, "to_numpy") else nodeSet
            ud.user_dirichletBC(self.dof, ns, self.dm, dm_specified, self.body.np_nodes, time)

    def dirichletBC_val(self, nodeSet, dm_specified: int, sval: float):
        n, c, v = self._bc_arrays(nodeSet, dm_specified, sval)
        self.ctx.call("femcy_dirichlet_val", as_i32(n), as_i32(c), as_d(v), n.size)

    def neumann_vector(self, load_facets, load_val: float, load_dir=np.array([])):
        """Consistent nodal loads of a traction on a set of boundary facets (host NumPy, as in the
        reference: stiffnessMtrx.py:386-411); see femcy_b200/neumann.py."""
        if self.sectioned:
            from .neumann import neumann_vector_sections
            return neumann_vector_sections(self.body, load_facets, load_val, load_dir)
        return neumann_vector(self.body, load_facets, load_val, load_dir)

    def _facet_pairs(self, load_facets):
        """(element, facet key index) int32 arrays of a loaded surface: carried by the set itself (meshgen.FacetSet, the
        reader's FaceSet) or looked up among the boundary facets, which the device finds (Body.boundary_arrays)."""
        if hasattr(load_facets, "kid"):
            ele, kid = load_facets.ele, load_facets.kid
        else:
            facets = np.array(sorted(load_facets), dtype=np.int64) if not isinstance(load_facets, np.ndarray) else load_facets
            if facets.size == 0:
                return np.zeros(0, np.int32), np.zeros(0, np.int32)
            ele, kid = self.body.locate_boundary_facets(facets)
        ele = np.asarray(ele, dtype=np.int64)
        if self.element_perm is not None:
            inv = np.empty(len(self.element_perm), dtype=np.int64)
            inv[np.asarray(self.element_perm)] = np.arange(len(self.element_perm))
            ele = inv[ele]
        return np.ascontiguousarray(ele, dtype=np.int32), np.ascontiguousarray(kid, dtype=np.int32)

    def neumannBC(self, load_facets, load_val: float, load_dir=np.array([])):
        """rhs is refreshed at every call, so only the last *Dsload of a deck acts (:384, quirk B1).  One section: the
        vector is integrated on the device (femcy_neumann, row f1); several sections: host NumPy, one upload."""
        if self.sectioned:
            self.rhs.from_numpy(self.neumann_vector(load_facets, load_val, load_dir))
            return
        ele, kid = self._facet_pairs(load_facets)
        d = np.ascontiguousarray(np.asarray(load_dir, dtype=np.float64).reshape(-1)[: self.dm])
        self.ctx.call("femcy_neumann", ele.size, as_i32(ele), as_i32(kid), float(load_val), as_d(d) if d.size else None)

    def impose_boundary_condition(self, boundary_conditions: dict):
        for nbc in boundary_conditions["neumannBCs"]:
            if "direction" in nbc:
                self.neumannBC(nbc["face_set"], load_val=nbc["traction"], load_dir=nbc["direction"])
            else:
                self.neumannBC(nbc["face_set"], load_val=nbc["traction"])
        if not self.geometric_nonlinear:
            for bc in boundary_conditions["dirichletBCs"]:
                self.dirichletBC_linearEquations(bc["node_set"], bc["dof"], bc["val"])
        else:
            for bc in boundary_conditions["dirichletBCs"]:
                self.dirichletBC_dof(bc["node_set"], bc["dof"], bc["val"], bc["user"], self.time1)

    # ---- strain / stress ---------------------------------------------------------------------------
    def get_deformation_gradient(self):
        self.ctx.call("femcy_deformation_gradient")

    def get_strain_smallDeformation(self):
        self.ctx.call("femcy_strain", 0)

    def get_strain_largeDeformation(self):
        self.ctx.call("femcy_strain", 1)

    def get_mises_stress_planeStress(self):
        self.ctx.call("femcy_mises")

    get_mises_stress_planeStrain = get_mises_stress_planeStress
    get_mises_stress_3d = get_mises_stress_planeStress

    def compute_strain_stress(self):
        self._sync_ghosts()
        self.get_deformation_gradient()
        if not self.geometric_nonlinear:
            self.get_strain_smallDeformation()
            self.material.constitutiveOfSmallDeform(self.F, self.cauchy_stress, None)
        else:
            self.get_strain_largeDeformation()   # stress was computed by the last internal-force pass
        self.ctx.call("femcy_mises")

    def get_elasEng(self):
        tot = C.c_double(0.)
        self.get_deformation_gradient()
        self.ctx.call("femcy_elastic_energy", C.byref(tot))
        if self.partition is not None and self.partition.nranks > 1:
            # interface elements are integrated redundantly by the neighbouring ranks: count every element once (on
            # its primary rank) and sum over the ranks.  Post-processing, not hot: the masked sum runs on the host.
            prim = np.asarray(self.partition.elem_primary, dtype=bool)
            e = self.elsEngDens.to_numpy()[prim] * self.vol.to_numpy()[prim]
            total, _ = self.comm.allreduce_sum_max(float(e.sum()), 0.0)
            tot.value = total
        self.elsEng[...] = tot.value
        return tot.value

    def assemble_nodal_force_GN(self):
        self.ctx.call("femcy_internal_force")

    # ---- increment / Newton driver (host control flow of stiffnessMtrx.py:647-822) ------------------
    def solve(self, inp, show_newton_steps: bool = False, save2path: str = None):
        max_inc = inp.time_incs["max_inc"]
        min_inc = inp.time_incs["min_inc"]
        max_time = inp.time_incs["max_time"]
        self.dt = inp.time_incs["ini_inc"]

        neumannBCs = copy.deepcopy(inp.neumann_bc_info)
        dirichletBCs = copy.deepcopy(inp.dirichlet_bc_info)
        for bc in dirichletBCs:
            bc["node_set"] = HostField(np.array([*bc["node_set"]]), dtype=np.int32)
        boundary_conditions = {"neumannBCs": neumannBCs, "dirichletBCs": dirichletBCs}
        self.inc_trace = []

        kinc = -1
        while self.time1 < max_time:
            kinc += 1
            self.time1 = min(self.time0 + self.dt, max_time)
            self._say("\033[40;33;1m >>>>>>>>>>>>>>>>>>>>>>>>>>>>>>"
                      ">>>>> kinc = {}, time0 = {}, dt = {} \033[0m".format(kinc, self.time0, self.dt))
            load_ratio = self.time1 / max_time
            for i, nbc in enumerate(neumannBCs):
                nbc["traction"] = inp.neumann_bc_info[i]["traction"] * load_ratio
            for i, bc in enumerate(dirichletBCs):
                bc["val"] = inp.dirichlet_bc_info[i]["val"] * load_ratio
            converged, newton_loop = self.advance_inc(inp, boundary_conditions, show_newton_steps, save2path)
            self.inc_trace.append((self.time1, bool(converged), int(newton_loop)))
            if not converged:
                self.time1 = self.time0
                self.dt /= 4.
                self.dof.copy_from(self.dof_old)
                kinc -= 1
                if self.dt < min_inc:
                    self._say("\033[31;1m allowable minimum dt is reached, "
                              "Newton's method not converges, solution is not found. \033[0m")
                    break
                continue
            if getattr(self, "_tangent_on_hold", False):          # opt-in consistent tangent: back on after a converged increment
                self._tangent_on_hold = False
                self.ctx.set_option("consistent_tangent", 1)
            if newton_loop <= 8:
                self.dt = min(self.dt * 1.5, max_inc)
            self.dof_old.copy_from(self.dof)
            self.time0 = self.time1

    def _residual(self, boundary_conditions):
        """f_int and K at the current dofs, residual = f_int - rhs with the Dirichlet rows fixed,
        returns RMS(residual) (the block repeated at stiffnessMtrx.py:720-723,756-759,779-783)."""
        self._sync_ghosts()
        self.assemble_nodal_force_GN()
        self.assemble_stiffnessMtrx()
        tg.c_equals_a_minus_b(self.residual_nodal_force, self.nodal_force, self.rhs)
        self.dirichletBC_forNewtonMethod(boundary_conditions["dirichletBCs"])
        self._newton_bcs = boundary_conditions["dirichletBCs"]      # (a consistent-tangent step may have to re-eliminate)
        return self._field_norm(self.residual_nodal_force)

    def advance_inc(self, inp, boundary_conditions: dict, show_newton_steps: bool = False,
                    save2path: str = None, window=None) -> Tuple[bool, int]:
        geometric_nonlinear = inp.geometric_nonlinear
        t0 = time.time()
        self._sync_ghosts()
        self.get_dsdx_and_vol()
        self.assemble_stiffnessMtrx()
        if not self.compiled:
            self.compiled = True
        self._say("time for assemble (launch) is {} s".format(time.time() - t0))

        self.impose_boundary_condition(boundary_conditions)

        if not geometric_nonlinear:
            self.solve_dof()
            return True, 0

        pre_residual = self._residual(boundary_conditions)
        if not hasattr(self, "ini_residual"):
            self.ini_residual = pre_residual          # captured once, reused by later increments (B4)
        self._say("\033[40;33;1m initial residual_nodal_force = {} \033[0m".format(self.ini_residual))

        if self.ini_residual < 1.e-9:
            self._say("\033[32;1m good! nonlinear converge! \033[0m")
            # the reference falls through to `return True, newton_loop` with newton_loop unbound
            # here (UnboundLocalError, :767-822); a zero-load increment is reported as converged
            return True, 0
        newton_loop = -1
        while pre_residual / (self.ini_residual + 1.e-30) >= 0.01:
            newton_loop += 1
            if newton_loop >= 24:
                return False, newton_loop
            du = self.solve_dof()                     # dof = dof - K^-1 residual

            residual = self._residual(boundary_conditions)
            if np.isnan(residual):
                self._say("NaN occurs, automatically recompute with smaller time step")
                return False, newton_loop
            self._say("\033[40;33;1m newton_loop = {}, residual_nodal_force = {} \033[0m".format(newton_loop, residual))

            # residual falling: keep going along du (at most 10 extra steps)
            relax_loop = -1
            relaxation = 1.
            while 0.1 * pre_residual < residual < pre_residual:
                new_residual = residual
                relax_loop += 1
                if relax_loop >= 10:
                    break
                tg.a_equals_b_plus_c_mul_d(self.dof, self.dof, -relaxation, du)
                residual = self._residual(boundary_conditions)
                if residual > new_residual:
                    tg.a_equals_b_plus_c_mul_d(self.dof, self.dof, +relaxation, du)
                    residual = self._residual(boundary_conditions)
                    relaxation *= 0.5

            # residual growing: back off (at most 2 halvings)
            relax_loop = -1
            relaxation = 0.5
            while residual > pre_residual:
                relax_loop += 1
                if relax_loop >= 2:
                    break
                tg.a_equals_b_plus_c_mul_d(self.dof, self.dof, (1. - relaxation), du)
                tg.field_multiply(du, relaxation)
                residual = self._residual(boundary_conditions)

            pre_residual = residual
        return True, newton_loop

    # ---- housekeeping ---------------------------------------------------------------------------------
    def close(self):
        self.ctx.close()
