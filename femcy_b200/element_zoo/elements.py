"""The six element plugins of the reference, restated as data + NumPy.

Class names, node ordering (Abaqus), Gauss rules, facet conventions and `inp_surface_num` follow
the reference so the `.inp` front end and the Neumann assembly behave identically:

  Element_linear_triangular        CPE3/CPS3   /root/reference/element_zoo/element_linear_triangular.py:19-145
  Element_quadratic_triangular     CPE6/CPS6   /root/reference/element_zoo/element_quadratic_triangular.py:21-185
  Element_linear_quadrilateral     CPE4/CPS4   /root/reference/element_zoo/element_linear_quadrilateral.py:16-164
  Element_quadratic_quadrilateral  CPE8/CPS8   /root/reference/element_zoo/element_quadratic_quadrilateral.py:17-211
  Element_linear_tetrahedral       C3D4        /root/reference/element_zoo/element_linear_tetrahedral.py:22-177
  Element_quadratic_tetrahedral    C3D10       /root/reference/element_zoo/element_quadratic_tetrahedral.py:27-250

Facet tables are generated from the natural coordinates of the element's nodes instead of being
typed in; the one irregularity of the reference (CPS8/CPE8 sub-facets (0,7)/(3,7) carry each
other's corner point) is reproduced on purpose and marked below.
"""
import numpy as np

from .element_base import ElementBase

_S2 = 2.0 ** 0.5 / 2.0


def _tables(node_nat, facets, normals, corner_w, mid_w, n_corner):
    """facet key -> point list / weights / normals.  A facet's points are the natural coordinates
    of its own nodes: corner nodes first (weight corner_w), then mid-side nodes (weight mid_w)."""
    coos, wts, nrm = {}, {}, {}
    for key, normal in zip(facets, normals):
        corners = [n for n in key if n < n_corner]
        mids = [n for n in key if n >= n_corner]
        pts = [list(map(float, node_nat[n])) for n in corners + mids]
        coos[key] = pts
        wts[key] = [corner_w] * len(corners) + [mid_w] * len(mids)
        nrm[key] = [list(normal)] * len(pts)
    return coos, wts, nrm


# =============================================================================================
class Element_linear_triangular(ElementBase):
    """3-node triangle; N = [xi, eta, 1-xi-eta]; 1 Gauss point (1/3,1/3), w = 1/2."""
    dm, n_en = 2, 3

    def __init__(self):
        # one integration point per edge, at the edge midpoint, weight 1
        self.facet_natural_coos = {(0, 1): [[0.5, 0.5]], (1, 2): [[0., 0.5]], (0, 2): [[0.5, 0.]]}
        self.facet_point_weights = {k: [1.] for k in self.facet_natural_coos}
        self.facet_natural_normals = {(0, 1): [[_S2, _S2]], (1, 2): [[-1., 0.]], (0, 2): [[0., -1.]]}
        self.inp_surface_num = [((0, 1),), ((1, 2),), ((2, 0),)]
        self._finish_init([[1. / 3., 1. / 3.]], [0.5])

    def shapeFunc_pyscope(self, nc):
        return np.array([nc[0], nc[1], 1. - nc[0] - nc[1]])

    def dshape_dnat_pyscope(self, nc):
        return np.array([[1., 0.], [0., 1.], [-1., -1.]])


# =============================================================================================
class Element_quadratic_triangular(ElementBase):
    """6-node triangle, mid-side nodes 3=(0,1) 4=(1,2) 5=(2,0); 3 Gauss points, w = 1/6."""
    dm, n_en = 2, 6
    _NAT = np.array([[1., 0.], [0., 1.], [0., 0.], [.5, .5], [0., .5], [.5, 0.]])

    def __init__(self):
        facets = [(0, 3), (1, 3), (1, 4), (2, 4), (2, 5), (0, 5)]
        normals = [(1., 1.), (1., 1.), (-1., 0.), (-1., 0.), (0., -1.), (0., -1.)]
        coos, wts, nrm = _tables(self._NAT, facets, normals, 0.5, 0.5, 3)
        # the reference lists the mid-side point first; the order is immaterial (equal weights)
        self.facet_natural_coos = {k: v[::-1] for k, v in coos.items()}
        self.facet_point_weights, self.facet_natural_normals = wts, nrm
        self.inp_surface_num = [((0, 3), (3, 1)), ((1, 4), (4, 2)), ((2, 5), (5, 0))]
        self._finish_init([[2. / 3., 1. / 6.], [1. / 6., 2. / 3.], [1. / 6., 1. / 6.]], [1. / 6.] * 3)

    @staticmethod
    def _L(nc):
        return np.array([nc[0], nc[1], 1. - nc[0] - nc[1]])

    def shapeFunc_pyscope(self, nc):
        L = self._L(nc)
        return np.array([L[0] * (2. * L[0] - 1.), L[1] * (2. * L[1] - 1.), L[2] * (2. * L[2] - 1.),
                         4. * L[0] * L[1], 4. * L[1] * L[2], 4. * L[2] * L[0]])

    def dshape_dnat_pyscope(self, nc):
        L = self._L(nc)
        dL = np.array([[1., 0.], [0., 1.], [-1., -1.]])          # dL_i/d(xi,eta)
        out = np.zeros((6, 2))
        for i in range(3):
            out[i] = (4. * L[i] - 1.) * dL[i]
        for m, (i, j) in enumerate([(0, 1), (1, 2), (2, 0)]):
            out[3 + m] = 4. * (L[i] * dL[j] + L[j] * dL[i])
        return out

    def extrapolation_matrix(self):
        return _affine_extrapolation(self._NAT, np.asarray(self.gaussPoints))


# =============================================================================================
class Element_linear_quadrilateral(ElementBase):
    """4-node quad on [-1,1]^2; 2x2 Gauss at +-1/sqrt(3) ordered like the nodes, w = 1."""
    dm, n_en = 2, 4
    _NAT = np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.]])

    def __init__(self):
        facets = [(0, 1), (1, 2), (2, 3), (0, 3)]
        normals = [(0., -1.), (1., 0.), (0., 1.), (-1., 0.)]
        self.facet_natural_coos, self.facet_point_weights, self.facet_natural_normals = _tables(
            self._NAT, facets, normals, 0.5, 0.5, 4)
        self.inp_surface_num = [((0, 1),), ((1, 2),), ((2, 3),), ((0, 3),)]
        t = 1. / 3. ** 0.5
        self._finish_init(self._NAT * t, [1., 1., 1., 1.])

    def shapeFunc_pyscope(self, nc):
        sx, sy = self._NAT[:, 0], self._NAT[:, 1]
        return (1. + sx * nc[0]) * (1. + sy * nc[1]) / 4.

    def dshape_dnat_pyscope(self, nc):
        sx, sy = self._NAT[:, 0], self._NAT[:, 1]
        return np.stack([sx * (1. + sy * nc[1]) / 4., sy * (1. + sx * nc[0]) / 4.], axis=1)

    def extrapolation_matrix(self):
        return _bilinear_extrapolation(self._NAT)


# =============================================================================================
class Element_quadratic_quadrilateral(ElementBase):
    """8-node serendipity quad, mid-side nodes 4..7 on edges (0,1),(1,2),(2,3),(3,0)."""
    dm, n_en = 2, 8
    _NAT = np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.], [0., -1.], [1., 0.], [0., 1.], [-1., 0.]])

    def __init__(self):
        facets = [(0, 4), (1, 4), (1, 5), (2, 5), (2, 6), (3, 6), (0, 7), (3, 7)]
        normals = [(0., -1.), (0., -1.), (1., 0.), (1., 0.), (0., 1.), (0., 1.), (-1., 0.), (-1., 0.)]
        coos, wts, nrm = _tables(self._NAT, facets, normals, 0.5, 0.5, 4)
        # REFERENCE QUIRK, kept for parity: on the left edge the two sub-facets carry each other's
        # corner point (element_quadratic_quadrilateral.py:40), so a load on face S4 puts nothing
        # on the corner nodes.
        coos[(0, 7)][0], coos[(3, 7)][0] = [-1., 1.], [-1., -1.]
        self.facet_natural_coos, self.facet_point_weights, self.facet_natural_normals = coos, wts, nrm
        self.inp_surface_num = [((0, 4), (1, 4)), ((1, 5), (2, 5)), ((2, 6), (3, 6)), ((0, 7), (3, 7))]
        t = 1. / 3. ** 0.5
        self._finish_init(self._NAT[:4] * t, [1., 1., 1., 1.])

    def shapeFunc_pyscope(self, nc):
        x, y = nc[0], nc[1]
        sx, sy = self._NAT[:4, 0], self._NAT[:4, 1]
        corner = (1. + sx * x) * (1. + sy * y) * (sx * x + sy * y - 1.) / 4.
        mid = np.array([(1. - x * x) * (1. - y) / 2., (1. - y * y) * (1. + x) / 2.,
                        (1. - x * x) * (1. + y) / 2., (1. - y * y) * (1. - x) / 2.])
        return np.concatenate([corner, mid])

    def dshape_dnat_pyscope(self, nc):
        x, y = nc[0], nc[1]
        sx, sy = self._NAT[:4, 0], self._NAT[:4, 1]
        dcx = sx * (1. + sy * y) * (2. * sx * x + sy * y) / 4.
        dcy = sy * (1. + sx * x) * (2. * sy * y + sx * x) / 4.
        dmx = np.array([-x * (1. - y), (1. - y * y) / 2., -x * (1. + y), -(1. - y * y) / 2.])
        dmy = np.array([-(1. - x * x) / 2., -y * (1. + x), (1. - x * x) / 2., -y * (1. - x)])
        return np.stack([np.concatenate([dcx, dmx]), np.concatenate([dcy, dmy])], axis=1)

    def extrapolation_matrix(self):
        return _bilinear_extrapolation(self._NAT)


# =============================================================================================
class Element_linear_tetrahedral(ElementBase):
    """4-node tet; N = [zeta, xi, 1-xi-eta-zeta, eta]; 1 Gauss point (1/4,1/4,1/4), w = 1/6."""
    dm, n_en = 3, 4
    _NAT = np.array([[0., 0., 1.], [1., 0., 0.], [0., 0., 0.], [0., 1., 0.]])
    _FACETS = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
    _NORMALS = [(0., 0., -1.), (-1., 0., 0.), (1., 1., 1.), (0., -1., 0.)]

    def __init__(self):
        # one point per face: the face centroid, weight 1
        self.facet_natural_coos = {k: [list(self._NAT[list(k)].mean(axis=0))] for k in self._FACETS}
        self.facet_point_weights = {k: [1.] for k in self._FACETS}
        self.facet_natural_normals = {k: [list(n)] for k, n in zip(self._FACETS, self._NORMALS)}
        self.inp_surface_num = [((0, 1, 2),), ((0, 1, 3),), ((1, 2, 3),), ((0, 2, 3),)]
        self._finish_init([[0.25, 0.25, 0.25]], [1. / 6.])

    def shapeFunc_pyscope(self, nc):
        return np.array([nc[2], nc[0], 1. - nc[0] - nc[1] - nc[2], nc[1]])

    def dshape_dnat_pyscope(self, nc):
        return np.array([[0., 0., 1.], [1., 0., 0.], [-1., -1., -1.], [0., 1., 0.]])


# =============================================================================================
class Element_quadratic_tetrahedral(ElementBase):
    """10-node tet; mid-edge nodes 4=(0,1) 5=(1,2) 6=(2,0) 7=(0,3) 8=(1,3) 9=(2,3); 4 Gauss points."""
    dm, n_en = 3, 10
    _EDGES = [(0, 1), (1, 2), (2, 0), (0, 3), (3, 1), (2, 3)]
    _dL = np.array([[0., 0., 1.], [1., 0., 0.], [-1., -1., -1.], [0., 1., 0.]])   # d(nc_i)/d(xi,eta,zeta)

    def __init__(self, gauss_points_count=4):
        corner = Element_linear_tetrahedral._NAT
        self._NAT = np.concatenate([corner, [(corner[i] + corner[j]) / 2. for i, j in self._EDGES]])
        facets = [(1, 2, 3, 5, 8, 9), (0, 2, 3, 6, 7, 9), (0, 1, 3, 4, 7, 8), (0, 1, 2, 4, 5, 6)]
        # 6 points per face: the 3 corners (w = 1/12) and the 3 mid-edge nodes (w = 1/4)
        self.facet_natural_coos, self.facet_point_weights, self.facet_natural_normals = _tables(
            self._NAT, facets, Element_linear_tetrahedral._NORMALS, 1. / 12., 1. / 4., 4)
        self.inp_surface_num = [((0, 1, 2, 4, 5, 6),), ((0, 1, 3, 4, 7, 8),), ((1, 2, 3, 5, 8, 9),), ((0, 2, 3, 6, 7, 9),)]
        a, b = 0.585410196624968, 0.138196601125010
        self._finish_init([[a, b, b], [b, a, b], [b, b, a], [b, b, b]], [1. / 24.] * 4)

    @staticmethod
    def _L(nc):
        return np.array([nc[2], nc[0], 1. - nc[0] - nc[1] - nc[2], nc[1]])

    def shapeFunc_pyscope(self, nc):
        L = self._L(nc)
        return np.concatenate([L * (2. * L - 1.), [4. * L[i] * L[j] for i, j in self._EDGES]])

    def dshape_dnat_pyscope(self, nc):
        L = self._L(nc)
        out = np.zeros((10, 3))
        for i in range(4):
            out[i] = (4. * L[i] - 1.) * self._dL[i]
        for m, (i, j) in enumerate(self._EDGES):
            out[4 + m] = 4. * (L[i] * self._dL[j] + L[j] * self._dL[i])
        return out

    def extrapolation_matrix(self):
        return _affine_extrapolation(self._NAT, np.asarray(self.gaussPoints))


# ---- Gauss-point -> node extrapolation operators (post-processing only) -----------------------
def _affine_extrapolation(node_nat, gauss_nat):
    """Simplex elements: the affine function through the Gauss-point values evaluated at the nodes
    (same numbers as the tables in element_quadratic_triangular.py:295-302 and
    element_quadratic_tetrahedral.py:322-338)."""
    A_g = np.hstack([np.ones((len(gauss_nat), 1)), gauss_nat])
    A_n = np.hstack([np.ones((len(node_nat), 1)), node_nat])
    return A_n @ np.linalg.inv(A_g)


def _bilinear_extrapolation(node_nat):
    """Quads: bilinear interpolant of the 2x2 Gauss values, evaluated at sqrt(3) x node coordinates
    (element_linear_quadrilateral.py:228-238, element_quadratic_quadrilateral.py:287-301)."""
    s = 3. ** 0.5
    corner = Element_linear_quadrilateral._NAT
    out = np.zeros((len(node_nat), 4))
    for n, (x, y) in enumerate(node_nat * s):
        out[n] = (1. + corner[:, 0] * x) * (1. + corner[:, 1] * y) / 4.
    return out


ELEMENT_TYPES = {
    "CPE3": Element_linear_triangular, "CPS3": Element_linear_triangular,
    "CPE4": Element_linear_quadrilateral, "CPS4": Element_linear_quadrilateral,
    "CPS6": Element_quadratic_triangular, "CPE6": Element_quadratic_triangular,
    "CPS8": Element_quadratic_quadrilateral, "CPE8": Element_quadratic_quadrilateral,
    "C3D4": Element_linear_tetrahedral, "C3D10": Element_quadratic_tetrahedral,
}
