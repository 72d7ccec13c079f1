"""Tier-2 oracle: vectorised NumPy/SciPy fp64 restatement of the reference's hot-path kernels.

TEST INFRASTRUCTURE ONLY.  Nothing under femcy_b200/ imports this file; it is used by tests/,
by __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the *checker*.

Pinning: every function here is validated in tests/test_oracle.py against golden vectors produced
by executing the unmodified reference sources (tests/golden/*.npz, made by
oracle/run_reference.py) -- K at u=0 and at a perturbed state, dsdx/vol, F, Cauchy stress (small
and large), Mises, internal force, Dirichlet-eliminated K/rhs, and CG iteration counts.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import numpy as np
import scipy.sparse as sp

A_, B_ = 0.585410196624968, 0.138196601125010


# ---- element tables ----------------------------------------------------------------------------
def _dn_tri3(p):      # element_zoo/element_linear_triangular.py:66-73
    return np.array([[1., 0.], [0., 1.], [-1., -1.]])


def _dn_tri6(p):      # element_zoo/element_quadratic_triangular.py:89-100
    x, y = p
    z = 1. - x - y
    return np.array([[4 * x - 1, 0.], [0., 4 * y - 1], [1 - 4 * z, 1 - 4 * z],
                     [4 * y, 4 * x], [-4 * y, 4 * (z - y)], [4 * (z - x), -4 * x]])


def _dn_quad4(p):     # element_zoo/element_linear_quadrilateral.py:77-85
    x, y = p
    return np.array([[-(1 - y), -(1 - x)], [(1 - y), -(1 + x)], [(1 + y), (1 + x)], [-(1 + y), (1 - x)]]) / 4.


def _dn_quad8(p):     # element_zoo/element_quadratic_quadrilateral.py:81-108
    x, y = p
    return np.array([
        [-(1 - y) * (-2 * x - y) / 4, -(1 - x) * (-2 * y - x) / 4],
        [(1 - y) * (2 * x - y) / 4, -(1 + x) * (-2 * y + x) / 4],
        [(1 + y) * (2 * x + y) / 4, (1 + x) * (2 * y + x) / 4],
        [-(1 + y) * (-2 * x + y) / 4, (1 - x) * (2 * y - x) / 4],
        [-x * (1 - y), -(1 - x * x) / 2], [(1 - y * y) / 2, -y * (1 + x)],
        [-x * (1 + y), (1 - x * x) / 2], [-(1 - y * y) / 2, -y * (1 - x)]])


def _dn_tet4(p):      # element_zoo/element_linear_tetrahedral.py:74-82
    return np.array([[0., 0., 1.], [1., 0., 0.], [-1., -1., -1.], [0., 1., 0.]])


def _dn_tet10(p):     # element_zoo/element_quadratic_tetrahedral.py:108-126
    n0, n1, n3 = p[2], p[0], p[1]
    n2 = 1. - p[0] - p[1] - p[2]
    return np.array([[0, 0, 4 * n0 - 1], [4 * n1 - 1, 0, 0], [1 - 4 * n2] * 3, [0, 4 * n3 - 1, 0],
                     [4 * n0, 0, 4 * n1], [4 * (n2 - n1), -4 * n1, -4 * n1], [-4 * n0, -4 * n0, 4 * (n2 - n0)],
                     [0, 4 * n0, 4 * n3], [4 * n3, 4 * n1, 0], [-4 * n3, 4 * (n2 - n3), -4 * n3]], dtype=float)


_T = 1. / 3. ** 0.5
_ELEMS = {
    # family: (dN function, gauss points, weights)     (gauss tables: same files, __init__)
    "tri3": (_dn_tri3, [[1 / 3, 1 / 3]], [0.5]),
    "tri6": (_dn_tri6, [[2 / 3, 1 / 6], [1 / 6, 2 / 3], [1 / 6, 1 / 6]], [1 / 6] * 3),
    "quad4": (_dn_quad4, [[-_T, -_T], [_T, -_T], [_T, _T], [-_T, _T]], [1.] * 4),
    "quad8": (_dn_quad8, [[-_T, -_T], [_T, -_T], [_T, _T], [-_T, _T]], [1.] * 4),
    "tet4": (_dn_tet4, [[0.25, 0.25, 0.25]], [1 / 6]),
    "tet10": (_dn_tet10, [[A_, B_, B_], [B_, A_, B_], [B_, B_, A_], [B_, B_, B_]], [1 / 24] * 4),
}
FAMILY = {"CPS3": "tri3", "CPE3": "tri3", "CPS6": "tri6", "CPE6": "tri6", "CPS4": "quad4", "CPE4": "quad4",
          "CPS8": "quad8", "CPE8": "quad8", "C3D4": "tet4", "C3D10": "tet10"}


def elem_tables(etype):
    fn, gps, w = _ELEMS[FAMILY.get(etype, etype)]
    return np.stack([fn(np.array(g, dtype=float)) for g in gps]), np.array(w, dtype=float)


# ---- materials (tangent C) ------------------------------------------------------------------------
def C_linear_isotropic(E, nu):          # material_zoo/linear_isotropic.py:17-33
    G = E / 2. / (1. + nu)
    c00 = E * (1. - nu) / (1. + nu) / (1. - 2. * nu)
    c01 = E * nu / (1. + nu) / (1. - 2. * nu)
    C = np.zeros((6, 6))
    C[:3, :3] = c01
    C[[0, 1, 2], [0, 1, 2]] = c00
    C[[3, 4, 5], [3, 4, 5]] = G
    return C


def C_plane_strain(E, nu):              # material_zoo/linear_isotropic_plane_strain.py:17-28
    G = E / 2. / (1. + nu)
    t1 = E / (1. + nu)
    t2 = nu / (abs(1. - 2. * nu) + 1.e-30)
    return np.array([[t1 * (1 + t2), t1 * t2, 0.], [t1 * t2, t1 * (1 + t2), 0.], [0., 0., G]])


def C_plane_stress(E, nu):              # material_zoo/linear_isotropic_plane_stress.py:17-20
    G = E / 2. / (1. + nu)
    c00 = E / (1. - nu ** 2)
    return np.array([[c00, c00 * nu, 0.], [c00 * nu, c00, 0.], [0., 0., G]])


def C_neo_hookean(C1, D1):              # material_zoo/neo_hookean.py:22-42
    vol = np.zeros((6, 6))
    vol[:3, :3] = 1.
    return 4. * C1 * np.eye(6) + 2. * D1 * vol


# ---- geometry (a2) ---------------------------------------------------------------------------------
def dsdx_and_vol(nodes, elements, u, etype):
    """stiffnessMtrx.py:132-150: grad N and vol at every Gauss point on the configuration X+u."""
    dN, w = elem_tables(etype)
    dm = nodes.shape[1]
    x = (nodes + u.reshape(-1, dm))[elements]                    # [ne, n_en, dm]
    J = np.einsum("eai,gak->egik", x, dN)                        # localNodes^T @ dsdn
    Ji = np.linalg.inv(J)
    dsdx = np.einsum("gak,egkj->egaj", dN, Ji)
    vol = np.linalg.det(J) * w[None, :]
    return dsdx, vol


def B_matrix(dsdx):
    """strainMtrx of all six elements (e.g. element_linear_tetrahedral.py:137-177):
    [..., n_en, dm] -> [..., n_v, n_en*dm]."""
    n_en, dm = dsdx.shape[-2:]
    lead = dsdx.shape[:-2]
    if dm == 2:
        B = np.zeros(lead + (3, n_en * 2))
        B[..., 0, 0::2] = dsdx[..., 0]
        B[..., 1, 1::2] = dsdx[..., 1]
        B[..., 2, 0::2] = dsdx[..., 1]
        B[..., 2, 1::2] = dsdx[..., 0]
    else:
        B = np.zeros(lead + (6, n_en * 3))
        B[..., 0, 0::3] = dsdx[..., 0]
        B[..., 1, 1::3] = dsdx[..., 1]
        B[..., 2, 2::3] = dsdx[..., 2]
        B[..., 3, 0::3] = dsdx[..., 1]
        B[..., 3, 1::3] = dsdx[..., 0]
        B[..., 4, 0::3] = dsdx[..., 2]
        B[..., 4, 2::3] = dsdx[..., 0]
        B[..., 5, 1::3] = dsdx[..., 2]
        B[..., 5, 2::3] = dsdx[..., 1]
    return B


# ---- assembly (a3) -----------------------------------------------------------------------------------
def element_dofs(elements, dm):
    return (elements[:, :, None] * dm + np.arange(dm)[None, None, :]).reshape(elements.shape[0], -1)


def assemble_K(nodes, elements, u, etype, C, chunk=200000):
    """stiffnessMtrx.py:161-186: K = sum_e sum_g B^T C B vol scattered by global dof; CSR with the
    full co-element pattern (structural zeros kept), sorted columns."""
    dm = nodes.shape[1]
    N = nodes.shape[0] * dm
    ne = elements.shape[0]
    K = None
    for s in range(0, ne, chunk):
        el = elements[s:s + chunk]
        dsdx, vol = dsdx_and_vol(nodes, el, u, etype)
        B = B_matrix(dsdx)
        Ke = np.einsum("egpi,pq,egqj,eg->eij", B, C, B, vol, optimize=True)
        ed = element_dofs(el, dm)
        n_edof = ed.shape[1]
        rows = np.repeat(ed, n_edof, axis=1).reshape(-1)
        cols = np.tile(ed, (1, n_edof)).reshape(-1)
        part = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(N, N)).tocsr()
        K = part if K is None else K + part
    K.sum_duplicates()
    K.sort_indices()
    return K


def pattern(elements, nn, dm):
    """body.py:182-194 + stiffnessMtrx.py:84-88: the set of (row, col) dof pairs, as sorted COO."""
    n_en = elements.shape[1]
    i = np.repeat(elements, n_en, axis=1).reshape(-1).astype(np.int64)
    j = np.tile(elements, (1, n_en)).reshape(-1).astype(np.int64)
    key = np.unique(i * nn + j)
    bi, bj = key // nn, key % nn
    r = (bi[:, None, None] * dm + np.arange(dm)[None, :, None]) + np.zeros((1, 1, dm), dtype=np.int64)
    c = (bj[:, None, None] * dm + np.arange(dm)[None, None, :]) + np.zeros((1, dm, 1), dtype=np.int64)
    rows, cols = r.reshape(-1), c.reshape(-1)
    order = np.lexsort((cols, rows))
    return rows[order], cols[order]


def csr_on_pattern(K, rows, cols):
    """values of K on a given sorted (rows, cols) pattern (structural zeros become 0.0)."""
    return np.asarray(K[rows, cols]).reshape(-1)


# ---- Dirichlet (a4), sequential semantics -------------------------------------------------------------
def dirichlet_linear(K, rhs, dofs, vals):
    """stiffnessMtrx.py:279-307 applied dof after dof (the intended sequential result)."""
    K = K.tolil(copy=True)
    rhs = rhs.copy()
    for i, v in zip(dofs, vals):
        col = K[:, i].toarray().ravel()
        nz = list(K.rows[i])                # symmetric pattern: rows that hold column i
        for j in nz:
            rhs[j] -= v * col[j]
        rhs[i] = v
        for j in nz:
            K[i, j] = 0.
            K[j, i] = 0.
        K[i, i] = 1.
    return K.tocsr(), rhs


def dirichlet_newton(K, residual, dofs):
    """stiffnessMtrx.py:317-341."""
    K = K.tolil(copy=True)
    residual = residual.copy()
    for i in dofs:
        residual[i] = 0.
        for j in list(K.rows[i]):
            K[i, j] = 0.
            K[j, i] = 0.
        K[i, i] = 1.
    return K.tocsr(), residual


# ---- PCG (a7) -----------------------------------------------------------------------------------------
def pcg(K, b, eps=1.0e-3, max_iter=None, trace=False):
    """conjugateGradientSolver.py:103-127, statement for statement (Jacobi M = 1/diag, stop when
    max|r| < eps*max|r0|, test after the d update).  Returns x, iterations[, per-iteration max|r|]."""
    M = 1. / K.diagonal()
    x = np.zeros_like(b)
    r = b.copy()
    d = M * r
    r0 = np.abs(r).max()
    hist = []
    n = b.shape[0] if max_iter is None else max_iter
    it = 0
    for i in range(n):
        Ad = K @ d
        rMr = np.dot(r * M, r)
        alpha = rMr / np.dot(d, Ad)
        x = x + alpha * d
        r = r - alpha * Ad
        beta = np.dot(r * M, r) / rMr
        d = M * r + beta * d
        rmax = np.abs(r).max()
        it = i + 1
        if trace:
            hist.append(rmax)
        if rmax < eps * r0:
            break
    return (x, it, np.array(hist)) if trace else (x, it)


# ---- deformation gradient, stress, Mises (a8) -----------------------------------------------------------
def deformation_gradient(nodes, elements, u, etype):
    """stiffnessMtrx.py:532-556: F = I + u_e^T grad_X N (reference configuration)."""
    dm = nodes.shape[1]
    dsdX, _ = dsdx_and_vol(nodes, elements, np.zeros_like(u), etype)
    ue = u.reshape(-1, dm)[elements]
    return np.einsum("eai,egaj->egij", ue, dsdX) + np.eye(dm)


def _sw(A):
    return np.swapaxes(A, -1, -2)


def _voigt(E):
    return np.stack([E[..., 0, 0], E[..., 1, 1], E[..., 2, 2], 2 * E[..., 0, 1], 2 * E[..., 2, 0], 2 * E[..., 1, 2]], -1)


def _unvoigt(s):
    return np.stack([np.stack([s[..., 0], s[..., 3], s[..., 4]], -1), np.stack([s[..., 3], s[..., 1], s[..., 5]], -1),
                     np.stack([s[..., 4], s[..., 5], s[..., 2]], -1)], -2)


def cauchy_stress(F, mat_class, params, C, large):
    """constitutiveOfSmallDeform / constitutiveOfLargeDeform of the four material classes:
    linear_isotropic.py:35-76, linear_isotropic_plane_strain.py:44-86,
    linear_isotropic_plane_stress.py:36-96, neo_hookean.py:44-77."""
    dm = F.shape[-1]
    I = np.eye(dm)
    if mat_class == "NeoHookean":
        C1, D1 = params
        J = np.linalg.det(F)[..., None, None]
        return 2. * C1 / J * (F @ _sw(F) - I) + 2. * D1 * (J - 1.) * I
    if mat_class == "LinearIsotropic":
        E = (F + _sw(F)) / 2. - I if not large else (_sw(F) @ F - I) / 2.
        S = _unvoigt(_voigt(E) @ C.T)
        return S if not large else F @ S @ _sw(F) / np.linalg.det(F)[..., None, None]
    if mat_class == "LinearIsotropicPlaneStrain":
        E = (F + _sw(F)) / 2. - I if not large else (_sw(F) @ F - I) / 2.
        ev = np.stack([E[..., 0, 0], E[..., 1, 1], E[..., 0, 1] + E[..., 1, 0]], -1)
        s = ev @ C.T
        S = np.stack([np.stack([s[..., 0], s[..., 2]], -1), np.stack([s[..., 2], s[..., 1]], -1)], -2)
        return S if not large else F @ S @ _sw(F) / np.linalg.det(F)[..., None, None]
    if mat_class == "LinearIsotropicPlaneStress":
        Em, nu = params
        G = Em / 2. / (1. + nu)
        c00 = Em / (1. - nu ** 2)
        c01 = c00 * nu
        C6 = np.zeros((6, 6))
        C6[0, 0] = C6[1, 1] = c00
        C6[0, 1] = C6[1, 0] = c01
        C6[3, 3] = G
        F3 = np.zeros(F.shape[:-2] + (3, 3))
        F3[..., :2, :2] = F
        F3[..., 2, 2] = -nu / (1. - nu) * (F[..., 0, 0] + F[..., 1, 1] - 2.) + 1.
        I3 = np.eye(3)
        E = (F3 + _sw(F3)) / 2. - I3 if not large else (_sw(F3) @ F3 - I3) / 2.
        S = _unvoigt(_voigt(E) @ C6.T)
        if large:
            S = F3 @ S @ _sw(F3) / np.linalg.det(F3)[..., None, None]
        return S[..., :2, :2]
    raise ValueError(mat_class)


def mises(sigma, mat_type, nu=0.):
    """stiffnessMtrx.py:457-501."""
    s = np.zeros(sigma.shape[:-2] + (3, 3))
    dm = sigma.shape[-1]
    s[..., :dm, :dm] = sigma
    if mat_type == "planeStrain":
        s[..., 2, 2] = nu * (sigma[..., 0, 0] + sigma[..., 1, 1])
    dev = s - np.eye(3) * (np.trace(s, axis1=-2, axis2=-1) / 3.)[..., None, None]
    return (1.5 * np.sum(dev * dev, axis=(-2, -1))) ** 0.5


# ---- internal force (a9) ---------------------------------------------------------------------------------
def internal_force(nodes, elements, u, etype, mat_class, params, C):
    """stiffnessMtrx.py:609-644: f_int[node] = sum_e sum_g grad_x N[nid] . sigma . vol with
    sigma = constitutiveOfLargeDeform(F) and grad/vol on the current configuration."""
    dm = nodes.shape[1]
    F = deformation_gradient(nodes, elements, u, etype)
    sig = cauchy_stress(F, mat_class, params, C, True)
    dsdx, vol = dsdx_and_vol(nodes, elements, u, etype)
    fe = np.einsum("egaj,egji,eg->eai", dsdx, sig, vol)
    f = np.zeros(nodes.shape[0] * dm)
    np.add.at(f, element_dofs(elements, dm).reshape(-1), fe.reshape(-1))
    return f, sig, F


# ---- consistent tangent (row f2, opt-in; NOT what the reference assembles) -------------------------------------------
def spatial_tangent(F, mat_class, params, C, h=1.0e-6):
    """A[..., i, m, j, n] = (1/J) d tau_im / dh [(I + h e_j e_n^T) F] - sigma_in delta_mj with tau = det(F) sigma(F): the
    exact linearisation of the reference's internal force (stiffnessMtrx.py:609-644) with respect to the nodal
    displacements, d f_a,i = sum_b grad N_a,m A_imjn grad N_b,n vol d u_b,j.  The reference itself keeps ddsdde constant
    (neo_hookean.py:62-64 is commented out).  Central differences of `cauchy_stress`; checked against finite differences of
    `internal_force` in tests/test_oracle.py."""
    dm = F.shape[-1]
    J = np.linalg.det(F)
    sig = cauchy_stress(F, mat_class, params, C, True)
    A = np.zeros(F.shape[:-2] + (dm,) * 4)
    I = np.eye(dm)
    for j in range(dm):
        for n in range(dm):
            l = np.zeros((dm, dm))
            l[j, n] = 1.0
            Fp, Fm = (I + h * l) @ F, (I - h * l) @ F
            tp = np.linalg.det(Fp)[..., None, None] * cauchy_stress(Fp, mat_class, params, C, True)
            tm = np.linalg.det(Fm)[..., None, None] * cauchy_stress(Fm, mat_class, params, C, True)
            A[..., :, :, j, n] = (tp - tm) / (2. * h) / J[..., None, None]
            for i in range(dm):
                A[..., i, j, j, n] -= sig[..., i, n]
    return 0.5 * (A + np.moveaxis(A, (-4, -3, -2, -1), (-2, -1, -4, -3)))       # major symmetry (hyperelastic laws)


def assemble_K_consistent(nodes, elements, u, etype, mat_class, params, C):
    """K_ab,ij = sum_g grad N_a,m A_imjn grad N_b,n vol on the current configuration (material + geometric stiffness)."""
    dm = nodes.shape[1]
    F = deformation_gradient(nodes, elements, u, etype)
    A = spatial_tangent(F, mat_class, params, C)
    g, vol = dsdx_and_vol(nodes, elements, u, etype)
    ne, n_en = elements.shape
    Ke = np.einsum("egam,egimjn,egbn,eg->eaibj", g, A, g, vol, optimize=True).reshape(ne, n_en * dm, n_en * dm)
    ed = element_dofs(elements, dm)
    rows = np.repeat(ed, n_en * dm, axis=1).reshape(-1)
    cols = np.tile(ed, (1, n_en * dm)).reshape(-1)
    K = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(nodes.size, nodes.size)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def field_norm(f):
    """tiGadgets.py:29-37 (an RMS)."""
    return float((np.sum(f ** 2) / f.size) ** 0.5)


# ---- synthetic meshes -------------------------------------------------------------------
def kuhn_cube(n, lengths=(1., 1., 1.), cells=None):
    """Kuhn triangulation of a box: cells^3 hexahedra, 6 tets each, node order fixed so that
    det[x1-x2, x3-x2, x0-x2] > 0 (the reference's C3D4 orientation).
    Independent restatement used to cross-check femcy_b200.meshgen."""
    nx, ny, nz = (n, n, n) if cells is None else cells
    xs = np.linspace(0, lengths[0], nx + 1)
    ys = np.linspace(0, lengths[1], ny + 1)
    zs = np.linspace(0, lengths[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, k):
        return (k * (ny + 1) + j) * (nx + 1) + i

    k_, j_, i_ = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i_, j_, k_ = i_.ravel(), j_.ravel(), k_.ravel()
    import itertools
    tets = []
    for perm in itertools.permutations(range(3)):
        off = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            off[s + 1] = off[s]
            off[s + 1, ax] += 1
        v = [nid(i_ + o[0], j_ + o[1], k_ + o[2]) for o in off]
        tets.append(np.stack(v, axis=1))
    tets = np.stack(tets, axis=1).reshape(-1, 4)
    x = nodes[tets]
    det = np.linalg.det(np.stack([x[:, 1] - x[:, 2], x[:, 3] - x[:, 2], x[:, 0] - x[:, 2]], axis=2))
    flip = det < 0
    tets[flip, 0], tets[flip, 1] = tets[flip, 1].copy(), tets[flip, 0].copy()
    return nodes, tets
