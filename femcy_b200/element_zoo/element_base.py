"""Element plugin surface of the B200 path.

Mirrors the abstract interface of the reference (`/root/reference/element_zoo/element_base.py:9-53`:
shapeFunc, dshape_dnat, shapeFunc_pyscope, dshape_dnat_pyscope, globalNormal, strainMtrx, getMesh,
extrapolate) and its data attributes (dm, gaussPoints, gaussWeights, integPointNum_eachFacet,
facet_natural_coos, facet_point_weights, facet_natural_normals, inp_surface_num).

Difference in kind: in the reference the `ti.func` members are inlined into JIT kernels.  Here an
element is *data*: the CUDA kernels are templates over (dm, n_en, n_gp) and receive
dN/dxi evaluated at the Gauss points plus the weights (`device_tables()`), so a new isoparametric
element with a standard Voigt B matrix needs no new CUDA code beyond a template instantiation.
All methods below are host-side NumPy.
"""
import abc

import numpy as np

from ..fields import HostField


class ElementBase(abc.ABC):
    dm: int          # spatial dimension
    n_en: int        # nodes per element

    # ---- to be provided by each element -------------------------------------------------
    @abc.abstractmethod
    def shapeFunc_pyscope(self, natCoo):
        """N_a(xi): array [n_en]."""

    @abc.abstractmethod
    def dshape_dnat_pyscope(self, natCoo):
        """dN_a/dxi_k: array [n_en, dm]."""

    # ---- reference-compatible aliases (were ti.func in the reference) ----------------------
    def shapeFunc(self, natCoo):
        return self.shapeFunc_pyscope(natCoo)

    def dshape_dnat(self, natCoo):
        return self.dshape_dnat_pyscope(natCoo)

    def _finish_init(self, gauss_points, gauss_weights):
        self.gaussPoints = HostField(gauss_points)
        self.gaussWeights = HostField(gauss_weights)
        self.gaussPoints_visualize = self.gaussPoints
        self.integPointNum_eachFacet = len(next(iter(self.facet_point_weights.values())))
        self.n_gp = int(self.gaussPoints.shape[0])

    # ---- tables consumed by the CUDA templates ----------------------------------------------
    def device_tables(self):
        """(dNdxi [n_gp, n_en, dm], weights [n_gp]) as contiguous float64 arrays."""
        gps = np.asarray(self.gaussPoints)
        dn = np.stack([np.asarray(self.dshape_dnat_pyscope(gp), dtype=np.float64) for gp in gps])
        return np.ascontiguousarray(dn), np.ascontiguousarray(np.asarray(self.gaussWeights, dtype=np.float64))

    # ---- B matrix (strainMtrx), Voigt rows 2-D [xx,yy,xy], 3-D [xx,yy,zz,xy,zx,yz] -----------
    def strainMtrx(self, dsdx):
        """B(grad N), shape (n_v, n_en*dm); same layout as e.g.
        /root/reference/element_zoo/element_linear_tetrahedral.py:137-177."""
        g = np.asarray(dsdx, dtype=np.float64)
        n_en, dm = g.shape
        if dm == 2:
            B = np.zeros((3, n_en * 2))
            B[0, 0::2] = g[:, 0]
            B[1, 1::2] = g[:, 1]
            B[2, 0::2] = g[:, 1]
            B[2, 1::2] = g[:, 0]
        else:
            B = np.zeros((6, n_en * 3))
            B[0, 0::3] = g[:, 0]
            B[1, 1::3] = g[:, 1]
            B[2, 2::3] = g[:, 2]
            B[3, 0::3] = g[:, 1]
            B[3, 1::3] = g[:, 0]
            B[4, 0::3] = g[:, 2]
            B[4, 2::3] = g[:, 0]
            B[5, 1::3] = g[:, 2]
            B[5, 2::3] = g[:, 1]
        return B

    # ---- Neumann helper -------------------------------------------------------------------------
    def globalNormal(self, nodes, facet, integPointId=0):
        """Unit outward normal and (facet size x point weight) at one facet integration point.

        Same construction as the reference (e.g. element_linear_triangular.py:88-120,
        element_linear_tetrahedral.py:98-134): the natural-space normal is pushed forward with
        (dx/dxi)^-1 evaluated at the facet point and normalised; the facet size is the distance of
        its first two (sorted) nodes in 2-D and the area of the triangle of its first three
        (sorted) nodes in 3-D."""
        nodes = np.asarray(nodes, dtype=np.float64)
        key = tuple(sorted(facet))
        nat = self.facet_natural_coos[key][integPointId]
        dxdn = nodes.T @ self.dshape_dnat_pyscope(nat)
        n = np.asarray(self.facet_natural_normals[key][integPointId], dtype=np.float64) @ np.linalg.inv(dxdn)
        n = n / (np.linalg.norm(n) + 1.e-30)
        if self.dm == 2:
            size = np.linalg.norm(nodes[key[0]] - nodes[key[1]])
        else:
            size = 0.5 * np.linalg.norm(np.cross(nodes[key[1]] - nodes[key[0]], nodes[key[2]] - nodes[key[0]]))
        return n, size * self.facet_point_weights[key][integPointId]

    # ---- vectorised facet data for large meshes (synthetic decks) -------------------------------
    def facet_point_table(self, key):
        """(natural coords [npt, dm], weights [npt], shape values [npt, n_en]) of one facet key."""
        nat = np.asarray(self.facet_natural_coos[key], dtype=np.float64)
        w = np.asarray(self.facet_point_weights[key], dtype=np.float64)
        N = np.stack([self.shapeFunc_pyscope(p) for p in nat])
        return nat, w, N

    def device_facet_tables(self):
        """The facet data of this element kind as the flat arrays `femcy_set_facet_tables` takes (row f1):
        key_nodes [nkeys, width] int32, w [nkeys, nfp], normals [nkeys, nfp, dm], N [nkeys, nfp, width] (shape functions of the
        facet's own nodes at the facet points), dN [nkeys, nfp, n_en, dm]."""
        keys = self.element_facets()
        key_nodes = np.ascontiguousarray(keys, dtype=np.int32)
        w, nrm, N, dN = [], [], [], []
        for key in keys:
            nat, wk, Nk = self.facet_point_table(key)
            w.append(wk)
            nrm.append(np.asarray(self.facet_natural_normals[key], dtype=np.float64))
            N.append(Nk[:, list(key)])
            dN.append(np.stack([np.asarray(self.dshape_dnat_pyscope(pt), dtype=np.float64) for pt in nat]))
        return (key_nodes, np.ascontiguousarray(w, dtype=np.float64), np.ascontiguousarray(nrm, dtype=np.float64),
                np.ascontiguousarray(N, dtype=np.float64), np.ascontiguousarray(dN, dtype=np.float64))

    # ---- post-processing ---------------------------------------------------------------------------
    def extrapolation_matrix(self):
        """[n_en, n_gp] matrix taking Gauss-point values to (per-element, un-averaged) nodal values.
        Default: constant (single Gauss point) -- the linear elements of the reference do exactly
        this (element_linear_triangular.py `extrapolate`)."""
        if self.n_gp == 1:
            return np.ones((self.n_en, 1))
        raise NotImplementedError

    def extrapolate(self, internal_vals, nodal_vals=None):
        if getattr(internal_vals, "extrapolate_on_device", None) is not None and len(internal_vals.shape) == 2:
            # a scalar per-Gauss-point device field (mises, energy, vol): extrapolated by the library (femcy_extrapolate)
            out = internal_vals.extrapolate_on_device(self.extrapolation_matrix())[0]
        else:
            vals = internal_vals.to_numpy() if hasattr(internal_vals, "to_numpy") else np.asarray(internal_vals)
            out = vals @ self.extrapolation_matrix().T
        if nodal_vals is not None:
            if hasattr(nodal_vals, "from_numpy"):
                nodal_vals.from_numpy(out)
            else:
                nodal_vals[...] = out
        return out

    def element_facets(self):
        """Facet keys (sorted local node tuples) in the order of facet_natural_coos."""
        return list(self.facet_natural_coos.keys())

    def getMesh(self, elements):
        """Surface facets of the mesh (sorted global node tuples owned by exactly one element).
        The reference's getMesh also builds render triangles; rendering is out of scope here."""
        elements = np.asarray(elements)
        keys = self.element_facets()
        allf = np.concatenate([np.sort(elements[:, list(k)], axis=1) for k in keys])
        owner = np.tile(np.arange(len(elements)), len(keys))
        uniq, inv, cnt = np.unique(allf, axis=0, return_inverse=True, return_counts=True)
        face2ele = {}
        for f, e in zip(map(tuple, allf.tolist()), owner.tolist()):
            face2ele.setdefault(f, set()).add(e)
        surfaces = uniq[cnt == 1]
        return uniq, face2ele, surfaces
