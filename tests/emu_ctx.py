"""Stand-in for `femcy_b200._lib.Context` that answers the C-ABI calls by running the product's CUDA KERNEL SOURCE on
the CPU SIMT emulation (tests/simt) -- TEST INFRASTRUCTURE ONLY, like fake_ctx.FakeContext (from which it inherits the
vector plumbing).  Real host code (`System_of_equations`) + real kernel code run end to end on the CPU, with the library's
default choices mirrored: assembly variant 0 -> the gather (csrc/assembly.cu: launch_assemble); PCG -> the streaming
persistent cooperative kernel (csrc/cg.cu, one GPU).
Nothing under femcy_b200/ knows about this file; the tests monkeypatch `femcy_b200.stiffnessMtrx.Context`."""
import ctypes as C

import numpy as np

import simt
from fake_ctx import FakeContext, SectionedFakeContext, _arr, _set, _VNAME


class _OneRank:
    """what simt.cg_solve needs from a rank-local system"""

    def __init__(self, pat, dm, val, b, N):
        self.pat, self.dm, self.val = pat, dm, val
        self.b = np.ascontiguousarray(b, dtype=np.float64)
        self.vecs = {k: np.zeros(N) for k in "xrdMA"}
        self.scal = np.zeros(64)
        self.partials = np.zeros(4096)
        self.ticket = np.zeros(8, dtype=np.uint32)
        self.window = np.zeros(simt.WINDOW_WORDS, dtype=np.uint64)


class EmuContext(SectionedFakeContext):
    assembly_log = None
    options = None
    # row f4: what the library parks per section (fake_ctx.SectionedFakeContext) + the emulation's own per-section state
    _PER_SECTION = SectionedFakeContext._PER_SECTION + ("_dN", "_w", "_shape", "_kind", "_tab", "_slots")
    _slots = None

    def __init__(self, device=0):
        super().__init__(device)
        import os
        from femcy_b200._lib import OPTIONS
        for name in OPTIONS:          # _lib.Context applies FEMCY_OPT_<NAME> of the process environment once, at creation
            v = os.environ.get("FEMCY_OPT_" + name.upper())
            if v is not None:
                self._femcy_set_option(name, int(v))

    def _femcy_set_element(self, n_gp, dN, w):
        super()._femcy_set_element(n_gp, dN, w)
        self._dN = _arr(dN, n_gp * self.n_en * self.dm).copy()
        self._w = _arr(w, n_gp).copy()
        self._shape = (self.n_gp, self.n_en, self.dm)

    def _femcy_set_material(self, kind, params, nparams, Cm, n_v):
        super()._femcy_set_material(kind, params, nparams, Cm, n_v)
        self._kind = int(kind)
        self._tab = simt.make_tables_raw(self._dN, self._w, self.C, self.params)

    def _femcy_build_pattern(self, nnz_ref):
        if self.sections:
            return self._build_pattern_sections(nnz_ref)
        self._sigma = int((self.options or {}).get("sell_sigma", -1))       # option of the next build, as in pattern.cu
        if self._sigma == -1:                                                # automatic: sigma-sort when natural order pads > 15 %
            nat = simt.SellPattern(self.conn, self.nn, dm=self.dm, sigma=0)
            self._sigma = 1024 if (self.nn >= 4096 and nat.nslots > 1.15 * nat.nnzb) else 0
        self.spat = simt.SellPattern(self.conn, self.nn, dm=self.dm, sigma=self._sigma)
        self.val = self.spat.val_zeros()
        _set(nnz_ref, self.spat.nnzb * self.dm * self.dm)

    def _build_pattern_sections(self, nnz_ref):
        """pattern.cu: build_pattern_sections -- the pattern kernels over the keys of all sections, per-section slots"""
        self._park()
        o = simt.build_pattern_sections([S["conn"] for S in self.sections], self.nn)
        self._layout = o
        for S, slots in zip(self.sections, o["elem_slot_sections"]):
            S["_slots"] = slots
        self._load(self.cur)
        self.spat = _LayoutPattern(o, self.nn, self.dm)
        self.val = self.spat.val_zeros()
        _set(nnz_ref, o["nnzb"] * self.dm * self.dm)

    def _femcy_pattern_stats(self, out4):
        for i, v in enumerate((self.spat.nnzb, self.spat.nslots, self.spat.nslice, self.spat.max_row_blocks)):
            out4[i] = v

    @property
    def K(self):
        K = self.spat.to_csr(self.val).tocsr()
        K.sort_indices()
        return K

    @K.setter
    def K(self, value):      # the oracle-backed handlers of FakeContext are all overridden below
        if value is not None:
            raise AttributeError("EmuContext keeps K in the device layout")

    def _femcy_get_dsdx_and_vol(self):
        def one():
            self.gp["dsdx"], self.gp["vol"] = simt.dsdx_and_vol_raw(self._tab, self._shape, self.nodes, self.conn, self.vec["dof"])
        self._for_sections(one)

    def _femcy_assemble_K(self, variant):
        v = int(variant)
        if self.sections:                       # assembly.cu: one zero-fill, then one scatter pass per section
            from femcy_b200._lib import FemcyError
            if v not in (0, 1):
                raise FemcyError("a mesh of several sections assembles by scatter-add (variant 0 or 1)")
            total = self.spat.val_zeros()

            def one():
                if self._tab is None:
                    raise FemcyError("set_element and set_material for every section first")
                pat = simt.SectionPattern(self._layout, self._slots, self.dm, self.nn)
                ct = (self.options or {}).get("consistent_tangent", 0)
                val, _, _ = simt.assemble_raw(self._tab, self._shape, self.nodes, self.conn, self.vec["dof"], pat,
                                              variant=4 if ct else 1, knob=self._kind if ct else 0)
                total[:] += val
            self._for_sections(one)
            self.val = total
            return
        if (self.options or {}).get("consistent_tangent", 0) and v in (0, 1):      # assembly.cu: launch_assemble_ct
            self.val, _, _ = simt.assemble_raw(self._tab, self._shape, self.nodes, self.conn, self.vec["dof"], self.spat, variant=4, knob=self._kind)
            return
        if v == 0:
            v = 2
        if self.assembly_log is not None:
            self.assembly_log.append(v)
        self.val, vol, _ = simt.assemble_raw(self._tab, self._shape, self.nodes, self.conn, self.vec["dof"], self.spat, variant=v)
        if v != 1:
            self.gp["vol"] = vol              # the gather (re)computes vol in its first pass

    # ---- row f1: the topology.cu kernels on the emulation ------------------------------------------------------------------
    def _topology(self):
        t = simt.Topology.__new__(simt.Topology)
        t.nodes, t.conn = np.ascontiguousarray(self.nodes), np.ascontiguousarray(self.conn, dtype=np.int32)
        f = self.ft
        t.tabs = (np.ascontiguousarray(f["keys"], dtype=np.int32), f["w"], f["normal"], f["N"], f["dN"])
        t.nkeys, t.width = f["keys"].shape
        t.nfp = f["w"].shape[1]
        t.nn, t.dm = t.nodes.shape
        t.ne, t.n_en = t.conn.shape
        return t

    def _femcy_boundary_facets(self, count_ref):
        self.bnd = self._topology().boundary_facets()
        _set(count_ref, len(self.bnd[0]))

    def _femcy_node_elements(self, ptr, lst):
        p, l = self._topology().node_elements()
        _arr(ptr, self.nn + 1, np.int32)[:] = p
        if l.size:
            _arr(lst, l.size, np.int32)[:] = l

    def _femcy_neumann(self, nf, ele, kid, traction, direction):
        from femcy_b200._lib import FemcyError
        e, k = _arr(ele, nf, np.int32), _arr(kid, nf, np.int32)
        if nf and (e.min() < 0 or e.max() >= self.conn.shape[0] or k.min() < 0 or k.max() >= self.ft["keys"].shape[0]):
            raise FemcyError("femcy_neumann: facet (element, key) out of range")
        d = None if direction is None else _arr(direction, self.dm)
        self.vec["rhs"][:] = self._topology().neumann(e, k, traction, d)

    def _femcy_set_option(self, key, value):
        from femcy_b200._lib import FemcyError, OPTIONS
        k = key.decode() if isinstance(key, bytes) else key
        if k not in OPTIONS:
            raise FemcyError(f"femcy_set_option: unknown option '{k}'")
        if self.options is None:
            self.options = {}
        self.options[k] = int(value)

    def set_option(self, name, value):
        self._femcy_set_option(name, value)

    def _check_bc(self, nodes, comps, n):
        """bc.cu: upload_bc validates the host lists before any kernel runs"""
        from femcy_b200._lib import FemcyError
        nd, cp = _arr(nodes, n, np.int32), _arr(comps, n, np.int32)
        if n and (nd.min() < 0 or nd.max() >= self.nn or cp.min() < 0 or cp.max() >= self.dm):
            raise FemcyError("Dirichlet node/component out of range")

    def _femcy_dirichlet_val(self, nodes, comps, vals, n):
        self._check_bc(nodes, comps, n)
        super()._femcy_dirichlet_val(nodes, comps, vals, n)

    def _femcy_dirichlet_linear(self, nodes, comps, vals, n):
        self._check_bc(nodes, comps, n)
        if n:
            simt.dirichlet(self.spat, self.val, self.vec["rhs"], _arr(nodes, n, np.int32), _arr(comps, n, np.int32), _arr(vals, n), 0)

    def _femcy_dirichlet_newton(self, nodes, comps, n):
        self._check_bc(nodes, comps, n)
        if n:
            simt.dirichlet(self.spat, self.val, self.vec["residual"], _arr(nodes, n, np.int32), _arr(comps, n, np.int32), np.zeros(n), 1)

    def _post(self):
        p = simt.Post(None, None, self.nodes, self.conn, self.vec["dof"], raw=(self._tab, self._shape, self._kind))
        p.F[...] = self.gp["F"]
        p.cauchy[...] = self.gp["cauchy"]
        p.vol[...] = self.gp["vol"]
        return p

    def _femcy_deformation_gradient(self):
        def one():
            self.gp["F"] = self._post().deformation_gradient().copy()
        self._for_sections(one)

    def _femcy_constitutive(self, large):
        def one():
            self.gp["cauchy"] = self._post().constitutive(bool(large)).copy()
        self._for_sections(one)

    def _femcy_strain(self, large):
        def one():
            self.gp["strain"] = self._post().strain(bool(large))
        self._for_sections(one)

    def _femcy_mises(self):
        def one():
            self.gp["mises"] = self._post().mises()
        self._for_sections(one)

    def _femcy_internal_force(self):
        total = np.zeros(self.N)              # post.cu: one zero-fill, then one scatter-add pass per section

        def one():
            p = self._post()
            total[:] += p.internal_force()
            self.gp["cauchy"], self.gp["F"], self.gp["vol"], self.gp["dsdx"] = p.cauchy.copy(), p.F.copy(), p.vol.copy(), p.dsdx.copy()
        self._for_sections(one)
        self.vec["nodal_force"][:] = total

    def _femcy_elastic_energy(self, tot_ref):
        tot = [0.0]

        def one():
            p = self._post()
            self.gp["energy"], t = p.energy()
            tot[0] += t
        self._for_sections(one)
        _set(tot_ref, tot[0])

    def _femcy_spmv(self, x_sel, y_sel):
        self.vec[_VNAME[y_sel]][:] = self.K @ self.vec[_VNAME[x_sel]]

    def _femcy_cg_solve(self, b_sel, eps, max_iter, check_every, fixed, it_ref, r0_ref, r1_ref):
        sysm = _OneRank(self.spat, self.dm, self.val, self.vec[_VNAME[b_sel]], self.N)
        it, r0, r1 = simt.cg_solve([sysm], eps=float(eps), max_iter=int(max_iter), check_every=int(check_every),
                                   fixed=bool(fixed), mode={0: 2, 1: 0, 2: 1, 3: 2}[int((self.options or {}).get("cg_kernel", 0))],
                                   sym=int((self.options or {}).get("cg_sym", 0)))
        self._breakdown = bool(sysm.scal[7] == 2.0)         # S_DONE == 2 (kernel_types.cuh): femcy_cg_breakdown
        self.vec["x"][:] = sysm.vecs["x"]
        for k, name in (("r", "r"), ("d", "d"), ("M", "M"), ("A", "Ad")):
            self.vec[name][:] = sysm.vecs[k]
        if it_ref is not None:
            _set(it_ref, it)
        if r0_ref is not None:
            _set(r0_ref, r0)
        if r1_ref is not None:
            _set(r1_ref, r1)

    _breakdown = False

    def cg_breakdown(self):
        return self._breakdown

    # ---- ConjugateGradientSolver_rowMajor drop-in: the reference's ELL arrays -> scalar SELL-32 (femcy_cg_from_ell) ----
    def _femcy_cg_from_ell(self, N, W, spm, ij):
        import scipy.sparse as sp
        from femcy_b200._lib import VEC
        N, W = int(N), int(W)
        vals = _arr(spm, N * W).reshape(N, W)
        idx = _arr(ij, N * (W + 1), np.int32).reshape(N, W + 1)
        cnt = idx[:, 0]
        mask = np.arange(W)[None, :] < cnt[:, None]
        rows = np.repeat(np.arange(N), cnt)
        K = sp.csr_matrix((vals[mask], (rows, idx[:, 1:][mask])), shape=(N, N))
        K.sum_duplicates()
        self.dm, self.nn, self.N, self.N_own = 1, N, N, N
        # a "mesh" of N one-dof nodes: pattern straight from the matrix (one pseudo-element per stored entry pair)
        coo = K.tocoo()
        self.spat = _pattern_from_coo(coo.row, coo.col, N)
        self.val = self.spat.from_csr(K)
        for k in VEC:
            self.vec[k] = np.zeros(N)


class _LayoutPattern:
    """the SellPattern members the context uses (Dirichlet kernels, PCG, CSR export), from the layout arrays of an emulated
    multi-section pattern build (natural row order)"""

    def __init__(self, o, nn, dm):
        self.dm, self.nn, self.nn_own = dm, nn, nn
        self.nnzb, self.nslice, self.nslots, self.max_row_blocks = o["nnzb"], o["nslice"], o["nslots"], o["max_row_blocks"]
        self.blkptr, self.slice_ptr, self.colidx, self.diag_slot = o["blkptr"], o["slice_ptr"], o["colidx"], o["diag_slot"]
        slots = np.flatnonzero(self.colidx >= 0)
        sl = np.searchsorted(self.slice_ptr, slots, side="right") - 1
        brow = sl * 32 + (slots & 31)
        order = np.lexsort((self.colidx[slots], brow))
        self.brow, self.bcol, self.bslot = brow[order].astype(np.int64), self.colidx[slots][order].astype(np.int64), slots[order]
        self.rowof = self.rowpos = None
        self.sigma = 0
        self._o = o

    val_zeros = simt.SellPattern.val_zeros
    to_csr = simt.SellPattern.to_csr
    from_csr = simt.SellPattern.from_csr


def _pattern_from_coo(rows, cols, N):
    """SellPattern (dm = 1) of an arbitrary sparse matrix: 2-node pseudo-elements (row, col) reproduce exactly its blocks
    when only the (a=0, b=1) entry of each pseudo-element is kept; simpler: build the layout arrays directly."""
    pat = simt.SellPattern.__new__(simt.SellPattern)
    order = np.lexsort((cols, rows))
    brow, bcol = rows[order].astype(np.int64), cols[order].astype(np.int64)
    nnzb = brow.size
    blkptr = np.searchsorted(brow, np.arange(N + 1)).astype(np.int32)
    rowlen = np.diff(blkptr)
    nslice = (N + 31) // 32
    padded = np.zeros(nslice * 32, dtype=np.int64)
    padded[:N] = rowlen
    w = padded.reshape(nslice, 32).max(axis=1)
    slice_ptr = np.zeros(nslice + 1, dtype=np.int32)
    slice_ptr[1:] = np.cumsum(w * 32)
    k = np.arange(nnzb) - blkptr[brow]
    bslot = slice_ptr[brow // 32] + k * 32 + (brow % 32)
    nslots = int(slice_ptr[-1])
    colidx = np.full(nslots, -1, dtype=np.int32)
    colidx[bslot] = bcol
    diag_slot = np.full(N, -1, dtype=np.int32)
    d = brow == bcol
    diag_slot[brow[d]] = bslot[d]
    pat.dm, pat.nn, pat.nn_own, pat.nnzb, pat.nslice, pat.nslots = 1, N, N, nnzb, nslice, nslots
    pat.max_row_blocks = int(w.max()) if nslice else 0
    pat.blkptr, pat.slice_ptr, pat.colidx, pat.diag_slot = blkptr, slice_ptr, colidx, diag_slot
    pat.brow, pat.bcol, pat.bslot = brow, bcol, bslot
    pat.rowof = pat.rowpos = None
    pat.sigma = 0
    return pat

