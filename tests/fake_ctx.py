"""CPU stand-in for `femcy_b200._lib.Context` (TEST INFRASTRUCTURE ONLY).

It answers the C-ABI calls that `System_of_equations` makes with the NumPy oracle
(`oracle/femcy_oracle.py`) and a direct sparse solve, so that the *host-side* logic of the product --
the increment / Newton driver transcribed from the reference (`stiffnessMtrx.py:647-822`), the boundary
condition plumbing, the Neumann assembly, the multi-increment quirks -- can be exercised by the
`-m "not gpu"` suite against the golden traces of the reference's own run.  Nothing under `femcy_b200/`
knows about this file; the tests monkeypatch `femcy_b200.stiffnessMtrx.Context` with `FakeContext`.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sl

from femcy_b200._lib import GP, VEC
from oracle import femcy_oracle as O

_FAMILY = {(2, 3): "tri3", (2, 6): "tri6", (2, 4): "quad4", (2, 8): "quad8", (3, 4): "tet4", (3, 10): "tet10"}
_MAT = {0: "LinearIsotropic", 1: "LinearIsotropicPlaneStrain", 2: "LinearIsotropicPlaneStress", 3: "NeoHookean"}
_VNAME = {v: k for k, v in VEC.items()}


def _arr(ptr, n, dtype=np.float64):
    """numpy view of a ctypes pointer produced by as_d / as_i32."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),))


def _set(ref, value):
    ref._obj.value = value       # ctypes.byref(...) object


class FakeContext:
    def __init__(self, device=0):
        self.device = device
        self.vec = {}
        self.gp = {}
        self.K = None
        self.n_launch = 0
        self.lib = None

    # ---- the Context surface used by the host code ----------------------------------------------
    def close(self):
        pass

    def sync(self):
        pass

    def launches(self):
        return self.n_launch

    def time_ms(self, kind):
        return 0.0

    def vec_get(self, which, n):
        name = which if isinstance(which, str) else _VNAME[which]
        return self.vec[name][: int(n)].copy()

    def vec_set(self, which, arr):
        name = which if isinstance(which, str) else _VNAME[which]
        a = np.asarray(arr, dtype=np.float64).reshape(-1)
        self.vec[name][: a.size] = a

    def gp_get(self, which, shape):
        return self.gp[which].reshape(shape).copy()

    def gp_set(self, which, arr):
        self.gp[which] = np.asarray(arr, dtype=np.float64).copy()

    def norms(self, which):
        name = which if isinstance(which, str) else _VNAME[which]
        v = self.vec[name][: self.N_own]
        s = float(np.sum(v * v))
        return np.array([(s / v.size) ** 0.5, float(np.abs(v).max()) if v.size else 0.0, s])

    def call(self, name, *a):
        self.n_launch += 1
        return getattr(self, "_" + name)(*a)

    def cg_breakdown(self):
        return False

    def _femcy_extrapolate(self, which, comp, E, elem_nodal, node_mean):
        """the library's femcy_extrapolate, stated in NumPy: nodal = E . Gauss-point values, mean over adjacent elements"""
        name = {v: k for k, v in GP.items()}[int(which)]
        ne, n_en = self.conn.shape
        a = np.asarray(self.gp[name], dtype=np.float64).reshape(ne, -1)
        n_gp = self.n_gp
        ncomp = a.shape[1] // n_gp
        vals = a.reshape(ne, n_gp, ncomp)[:, :, int(comp)]
        Em = _arr(E, n_en * n_gp).reshape(n_en, n_gp)
        out = vals @ Em.T
        if elem_nodal is not None:
            _arr(elem_nodal, ne * n_en)[:] = out.reshape(-1)
        if node_mean is not None:
            s_ = np.bincount(self.conn.reshape(-1), weights=out.reshape(-1), minlength=self.nn)
            c_ = np.bincount(self.conn.reshape(-1), minlength=self.nn)
            _arr(node_mean, self.nn)[:] = s_ / np.maximum(c_, 1)

    # ---- row f1: topology + Neumann vector (topology.cu), stated in NumPy from the uploaded facet tables -------------
    def _femcy_set_facet_tables(self, nkeys, width, nfp, key_nodes, w, normals, N, dN):
        self.ft = {"keys": _arr(key_nodes, nkeys * width, np.int32).reshape(nkeys, width).astype(np.int64).copy(),
                   "w": _arr(w, nkeys * nfp).reshape(nkeys, nfp).copy(),
                   "normal": _arr(normals, nkeys * nfp * self.dm).reshape(nkeys, nfp, self.dm).copy(),
                   "N": _arr(N, nkeys * nfp * width).reshape(nkeys, nfp, width).copy(),
                   "dN": _arr(dN, nkeys * nfp * self.n_en * self.dm).reshape(nkeys, nfp, self.n_en, self.dm).copy()}

    def _femcy_boundary_facets(self, count_ref):
        keys = self.ft["keys"]
        ne = self.conn.shape[0]
        facs = np.concatenate([np.sort(self.conn[:, k], axis=1) for k in keys])
        _, inverse, counts = np.unique(facs, axis=0, return_inverse=True, return_counts=True)
        single = np.nonzero(counts[inverse.reshape(-1)] == 1)[0]
        self.bnd = ((single % ne).astype(np.int32), (single // ne).astype(np.int32))
        _set(count_ref, len(single))

    def _femcy_get_boundary_facets(self, ele, kid):
        n = len(self.bnd[0])
        if n:
            _arr(ele, n, np.int32)[:] = self.bnd[0]
            _arr(kid, n, np.int32)[:] = self.bnd[1]

    def _femcy_node_elements(self, ptr, lst):
        ne, n_en = self.conn.shape
        flat = self.conn.reshape(-1)
        order = np.argsort(flat * ne + np.repeat(np.arange(ne), n_en), kind="stable")
        _arr(lst, ne * n_en, np.int32)[:] = np.repeat(np.arange(ne), n_en)[order]
        p = np.zeros(self.nn + 1, dtype=np.int64)
        np.cumsum(np.bincount(flat, minlength=self.nn), out=p[1:])
        _arr(ptr, self.nn + 1, np.int32)[:] = p

    def _femcy_neumann(self, nf, ele, kid, traction, direction):
        """stiffnessMtrx.py:386-411 from the uploaded tables"""
        rhs = self.vec["rhs"]
        rhs[:] = 0.0
        if nf == 0:
            return
        ele, kid = _arr(ele, nf, np.int32).astype(np.int64), _arr(kid, nf, np.int32).astype(np.int64)
        d = None if direction is None else _arr(direction, self.dm).copy()
        T, dm = self.ft, self.dm
        for e, k in zip(ele, kid):
            kn = T["keys"][k]
            X = self.nodes[self.conn[e]]
            if dm == 2:
                size = np.linalg.norm(X[kn[0]] - X[kn[1]])
            else:
                size = 0.5 * np.linalg.norm(np.cross(X[kn[1]] - X[kn[0]], X[kn[2]] - X[kn[0]]))
            for p in range(T["w"].shape[1]):
                if d is None:
                    n = T["normal"][k, p] @ np.linalg.inv(X.T @ T["dN"][k, p])
                    n = n / (np.linalg.norm(n) + 1.e-30)
                else:
                    n = d
                flux = traction * n * size * T["w"][k, p]
                for q, a in enumerate(kn):
                    rhs[self.conn[e, a] * dm: self.conn[e, a] * dm + dm] += flux * T["N"][k, p, q]

    def _femcy_set_aggregates(self, nagg, agg):
        self.n_aggregates = int(nagg)

    def cg_phase_ns(self):
        return np.full(7, 1000.0)

    def set_option(self, name, value):
        pass

    # ---- handlers ---------------------------------------------------------------------------------
    def _femcy_set_mesh(self, dm, nn, nn_own, nodes, ne, n_en, conn):
        self.dm, self.nn, self.ne, self.n_en = int(dm), int(nn), int(ne), int(n_en)
        self.N = self.N_own = self.nn * self.dm
        self.nodes = _arr(nodes, nn * dm).reshape(nn, dm).copy()
        self.conn = _arr(conn, ne * n_en, np.int32).reshape(ne, n_en).astype(np.int64)
        self.etype = _FAMILY[(self.dm, self.n_en)]

    def _femcy_set_element(self, n_gp, dN, w):
        self.n_gp = int(n_gp)
        dn = _arr(dN, n_gp * self.n_en * self.dm).reshape(n_gp, self.n_en, self.dm)
        dn_ref, w_ref = O.elem_tables(self.etype)
        assert np.allclose(dn, dn_ref, atol=1e-15) and np.allclose(_arr(w, n_gp), w_ref, atol=1e-16)
        for k in VEC:
            self.vec[k] = np.zeros(self.N)
        dd = self.dm * self.dm
        self.gp = {"vol": np.zeros((self.ne, n_gp)), "F": np.zeros((self.ne, n_gp, self.dm, self.dm)),
                   "cauchy": np.zeros((self.ne, n_gp, self.dm, self.dm)), "mises": np.zeros((self.ne, n_gp)),
                   "energy": np.zeros((self.ne, n_gp)), "strain": np.zeros((self.ne, n_gp, self.dm, self.dm)),
                   "dsdx": np.zeros((self.ne, n_gp, self.n_en, self.dm))}

    def _femcy_set_material(self, kind, params, nparams, Cm, n_v):
        self.mat_class = _MAT[int(kind)]
        self.params = tuple(_arr(params, nparams).tolist())
        self.C = _arr(Cm, n_v * n_v).reshape(n_v, n_v).copy()

    def _femcy_build_pattern(self, nnz_ref):
        rows, cols = O.pattern(self.conn, self.nn, self.dm)
        self.pat = (rows, cols)
        _set(nnz_ref, len(rows))

    def _femcy_get_csr_pattern(self, rp, ci):
        K = self.K.tocsr()
        K.sort_indices()
        _arr(rp, self.N + 1, np.int32)[:] = K.indptr
        _arr(ci, K.nnz, np.int32)[:] = K.indices

    def _femcy_get_K_csr_values(self, v):
        K = self.K.tocsr()
        K.sort_indices()
        _arr(v, K.nnz)[:] = K.data

    def _femcy_vec_fill(self, which, val):
        self.vec[_VNAME[which]][:] = val

    def _femcy_vec_copy(self, dst, src):
        self.vec[_VNAME[dst]][:] = self.vec[_VNAME[src]]

    def _femcy_vec_lincomb(self, dst, a, alpha, b):
        self.vec[_VNAME[dst]][:] = self.vec[_VNAME[a]] + alpha * self.vec[_VNAME[b]]

    def _femcy_vec_scale(self, which, s):
        self.vec[_VNAME[which]] *= s

    def _femcy_get_dsdx_and_vol(self):
        self.gp["dsdx"], self.gp["vol"] = O.dsdx_and_vol(self.nodes, self.conn, self.vec["dof"], self.etype)

    def _femcy_assemble_K(self, variant):
        K = O.assemble_K(self.nodes, self.conn, self.vec["dof"], self.etype, self.C)
        # keep the full structural pattern (explicit zeros) like the device matrix
        rows, cols = self.pat
        self.K = sp.csr_matrix((O.csr_on_pattern(K, rows, cols), (rows, cols)), shape=(self.N, self.N))

    def _bc(self, nodes, comps, n):
        return _arr(nodes, n, np.int32).astype(np.int64) * self.dm + _arr(comps, n, np.int32)

    def _eliminate(self, dofs):
        flag = np.zeros(self.N)
        flag[dofs] = 1.0
        D = sp.diags(1.0 - flag)
        return (D @ self.K @ D + sp.diags(flag)).tocsr()

    def _femcy_dirichlet_linear(self, nodes, comps, vals, n):
        if n == 0:
            return
        dofs = self._bc(nodes, comps, n)
        v = np.zeros(self.N)
        v[dofs] = _arr(vals, n)
        rhs = self.vec["rhs"]
        corr = self.K @ v
        free = np.ones(self.N, dtype=bool)
        free[dofs] = False
        rhs[free] -= corr[free]
        rhs[dofs] = v[dofs]
        self.K = self._eliminate(dofs)

    def _femcy_dirichlet_newton(self, nodes, comps, n):
        if n == 0:
            return
        dofs = self._bc(nodes, comps, n)
        self.vec["residual"][dofs] = 0.0
        self.K = self._eliminate(dofs)

    def _femcy_dirichlet_val(self, nodes, comps, vals, n):
        if n:
            self.vec["dof"][self._bc(nodes, comps, n)] = _arr(vals, n)

    def _femcy_deformation_gradient(self):
        self.gp["F"] = O.deformation_gradient(self.nodes, self.conn, self.vec["dof"], self.etype)

    def _femcy_constitutive(self, large):
        self.gp["cauchy"] = O.cauchy_stress(self.gp["F"], self.mat_class, self.params, self.C, bool(large))

    def _femcy_strain(self, large):
        F = self.gp["F"]
        I = np.eye(self.dm)
        Ft = np.swapaxes(F, -1, -2)
        self.gp["strain"] = (Ft @ F - I) / 2.0 if large else (F + Ft) / 2.0 - I

    def _femcy_mises(self):
        mt = {"LinearIsotropicPlaneStrain": "planeStrain", "LinearIsotropicPlaneStress": "planeStress"}.get(self.mat_class, "3d")
        self.gp["mises"] = O.mises(self.gp["cauchy"], mt, self.params[1])

    def _femcy_internal_force(self):
        f, sig, F = O.internal_force(self.nodes, self.conn, self.vec["dof"], self.etype, self.mat_class, self.params, self.C)
        self.vec["nodal_force"][:] = f
        self.gp["cauchy"], self.gp["F"] = sig, F
        self.gp["dsdx"], self.gp["vol"] = O.dsdx_and_vol(self.nodes, self.conn, self.vec["dof"], self.etype)

    def _femcy_elastic_energy(self, tot_ref):
        _set(tot_ref, 0.0)     # not needed by the driver tests

    def _femcy_cg_solve(self, b_sel, eps, max_iter, check_every, fixed, it_ref, r0_ref, r1_ref):
        b = self.vec[_VNAME[b_sel]]
        x = sl.spsolve(self.K.tocsc(), b)        # the reference's own choice below 1e5 dofs (:219-251)
        self.vec["x"][:] = x
        r = b - self.K @ x
        _set(it_ref, 1)
        _set(r0_ref, float(np.abs(b).max()))
        _set(r1_ref, float(np.abs(r).max()))


class SectionedFakeContext(FakeContext):
    """FakeContext + the section calls of row f4 (femcy_add_section / femcy_select_section), stated the way the library
    implements them: the per-section attributes always describe the SELECTED section, the others are parked; the
    all-section calls (pattern, assembly, geometry, stress recovery, internal force) loop over the sections."""
    _PER_SECTION = ("ne", "n_en", "n_gp", "conn", "etype", "mat_class", "params", "C", "gp")

    def __init__(self, device=0):
        super().__init__(device)
        self.sections = []
        self.cur = 0

    def _park(self):
        if self.sections:
            self.sections[self.cur] = {k: getattr(self, k, None) for k in self._PER_SECTION}

    def _load(self, s):
        for k, v in self.sections[s].items():
            setattr(self, k, v)
        self.cur = s

    def _for_sections(self, fn):
        if not self.sections:
            return fn()
        keep = self.cur
        for s in range(len(self.sections)):
            self._park()
            self._load(s)
            fn()
        self._park()
        self._load(keep)

    def _femcy_set_mesh(self, *a):
        self.sections, self.cur = [], 0
        super()._femcy_set_mesh(*a)

    def _femcy_set_element(self, n_gp, dN, w):
        vec = self.vec
        super()._femcy_set_element(n_gp, dN, w)
        if self.cur != 0 and vec:
            self.vec = vec                 # the named vectors belong to the mesh, not to the section

    def _femcy_add_section(self, ne, n_en, conn, sec_ref):
        if not self.sections:
            self.sections = [None]
            self.cur = 0
        self._park()
        if (self.dm, int(n_en)) not in _FAMILY:
            from femcy_b200._lib import FemcyError
            raise FemcyError("unsupported (dm, n_en) element shape")
        S = {k: None for k in self._PER_SECTION}
        S.update({"ne": int(ne), "n_en": int(n_en), "n_gp": 0, "gp": {}, "etype": _FAMILY[(self.dm, int(n_en))],
                  "conn": _arr(conn, ne * n_en, np.int32).reshape(ne, n_en).astype(np.int64)})
        self.sections.append(S)
        self._load(len(self.sections) - 1)
        if sec_ref is not None:
            sec_ref._obj.value = self.cur

    def _femcy_select_section(self, s):
        from femcy_b200._lib import FemcyError
        n = len(self.sections) or 1
        if not 0 <= int(s) < n:
            raise FemcyError("femcy_select_section: no such section")
        if self.sections and int(s) != self.cur:
            self._park()
            self._load(int(s))

    def _femcy_section_count(self):
        return len(self.sections) or 1

    def _femcy_build_pattern(self, nnz_ref):
        if not self.sections:
            return super()._femcy_build_pattern(nnz_ref)
        self._park()
        keys = []
        for S in self.sections:
            r, c = O.pattern(S["conn"], self.nn, self.dm)
            keys.append(r.astype(np.int64) * self.N + c)
        key = np.unique(np.concatenate(keys))
        self.pat = (key // self.N, key % self.N)
        _set(nnz_ref, len(key))

    def _femcy_assemble_K(self, variant):
        if not self.sections:
            return super()._femcy_assemble_K(variant)
        from femcy_b200._lib import FemcyError
        if int(variant) not in (0, 1):
            raise FemcyError("a mesh of several sections assembles by scatter-add (variant 0 or 1)")
        rows, cols = self.pat
        total = sp.csr_matrix((self.N, self.N))
        self._park()
        for S in self.sections:
            total = total + O.assemble_K(self.nodes, S["conn"], self.vec["dof"], S["etype"], S["C"])
        self.K = sp.csr_matrix((O.csr_on_pattern(total.tocsr(), rows, cols), (rows, cols)), shape=(self.N, self.N))

    def _femcy_get_dsdx_and_vol(self):
        self._for_sections(super()._femcy_get_dsdx_and_vol)

    def _femcy_deformation_gradient(self):
        self._for_sections(super()._femcy_deformation_gradient)

    def _femcy_constitutive(self, large):
        self._for_sections(lambda: FakeContext._femcy_constitutive(self, large))

    def _femcy_strain(self, large):
        self._for_sections(lambda: FakeContext._femcy_strain(self, large))

    def _femcy_mises(self):
        self._for_sections(super()._femcy_mises)

    def _femcy_internal_force(self):
        if not self.sections:
            return super()._femcy_internal_force()
        total = np.zeros(self.N)

        def one():
            FakeContext._femcy_internal_force(self)
            total[:] += self.vec["nodal_force"]
        self._for_sections(one)
        self.vec["nodal_force"][:] = total

    def _femcy_elastic_energy(self, tot_ref):
        tot = [0.0]

        def one():
            F = O.deformation_gradient(self.nodes, self.conn, self.vec["dof"], self.etype)
            sig = O.cauchy_stress(F, self.mat_class, self.params, self.C, False)
            I = np.eye(self.dm)
            eps_ = (F + np.swapaxes(F, -1, -2)) / 2.0 - I
            _, vol = O.dsdx_and_vol(self.nodes, self.conn, self.vec["dof"], self.etype)
            self.gp["energy"] = 0.5 * np.sum(sig * eps_, axis=(-2, -1))
            self.gp["vol"] = vol
            tot[0] += float(np.sum(self.gp["energy"] * vol))
        self._for_sections(one)
        _set(tot_ref, tot[0])
