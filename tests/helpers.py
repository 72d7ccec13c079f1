"""Shared helpers for the test-suite (golden loading, material reconstruction, tolerances)."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def golden_names():
    # (bench_*.npz are fixtures of bench.py's parity leg, topology_*.npz those of the reference's Body queries: not decks)
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(f).startswith(("bench_", "topology_")))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def material_params(g):
    """(class name, (p0, p1)) of a golden file: the constructor parameters recovered from C
    (the goldens store the tangent, not E/nu)."""
    mc = str(g["mat_class"])
    C = g["C"]
    if "mat_params" in g.files:
        return mc, (float(g["mat_params"][0]), float(g["mat_params"][1]))
    if mc == "NeoHookean":
        return mc, (C[3, 3] / 4., C[0, 1] / 2.)
    if mc == "LinearIsotropic":
        G, lam = C[3, 3], C[0, 1]
        nu = lam / (2. * (lam + G))
        return mc, (2. * G * (1. + nu), nu)
    if mc == "LinearIsotropicPlaneStrain":
        r = C[0, 1] / C[0, 0]          # nu/(1-nu)
        nu = r / (1. + r)
        return mc, (2. * C[2, 2] * (1. + nu), nu)
    if mc == "LinearIsotropicPlaneStress":
        nu = C[0, 1] / C[0, 0]
        return mc, (2. * C[2, 2] * (1. + nu), nu)
    raise ValueError(mc)


def make_material(g):
    from femcy_b200 import material_zoo as mz
    mc, p = material_params(g)
    return getattr(mz, mc)(*p)


def make_element(g):
    from femcy_b200.element_zoo import ELEMENT_TYPES
    return ELEMENT_TYPES[str(g["elem_type"])]()


def perturbation(nodes):
    """Same displacement field as oracle/run_reference.py::perturbation."""
    span = float((nodes.max(axis=0) - nodes.min(axis=0)).max())
    amp = 0.02 * span
    x = (nodes - nodes.min(axis=0)) / span
    dm = nodes.shape[1]
    u = np.zeros_like(nodes)
    for c in range(dm):
        phase = 1.3 * x[:, 0] + 0.7 * x[:, 1] + (0.4 * x[:, 2] if dm == 3 else 0.0)
        u[:, c] = amp * np.sin(2.1 * phase + 0.9 * c)
    return u.reshape(-1)


def deck_path(g):
    """Absolute path of the deck a golden was made from, if the reference tree is present."""
    p = os.path.join(os.environ.get("FEMCY_REFERENCE", "/root/reference"), str(g["deck"]))
    return p if os.path.exists(p) else None


class GoldenDeck:
    """InpInfo-shaped deck rebuilt from a golden file (the .inp decks live in /root/reference and do
    not travel to the GPU box; the goldens carry nodes, connectivity, sets, loads, material, increments)."""

    def __init__(self, g):
        self.nodes = g["nodes"]
        et = str(g["elem_type"])
        self.eSets = {et: g["elements"].astype(np.int64)}
        self.ELE = make_element(g)
        self.materials = {"m": make_material(g)}
        self.geometric_nonlinear = bool(g["nlgeom"])
        ti = g["time_incs"]
        self.time_incs = {"ini_inc": float(ti[0]), "max_time": float(ti[1]), "min_inc": float(ti[2]), "max_inc": float(ti[3])}
        self.dirichlet_bc_info = []
        for k in range(len(g["bc_ptr"]) - 1):
            ns = g["bc_nodes"][g["bc_ptr"][k]:g["bc_ptr"][k + 1]].astype(np.int64)
            self.dirichlet_bc_info.append({"node_set": ns, "dof": int(g["bc_dof"][k]), "val": float(g["bc_val"][k]),
                                           "user": bool(g["bc_user"][k])})
        self.neumann_bc_info = []
        for k in range(int(g["n_neumann"])):
            item = {"face_set": set(map(tuple, g[f"nm{k}_facets"].tolist())), "traction": float(g[f"nm{k}_traction"])}
            if g[f"nm{k}_direction"].size:
                item["direction"] = g[f"nm{k}_direction"]
            self.neumann_bc_info.append(item)


def system_from_deck(deck, **kw):
    from femcy_b200 import Body, System_of_equations
    body = Body(deck.nodes, list(deck.eSets.values())[0], deck.ELE)
    kw.setdefault("quiet", True)
    return System_of_equations(body, list(deck.materials.values())[0], deck.geometric_nonlinear, **kw)


def abs_err_scaled(a, b, scale):
    """max|a-b| / scale -- for quantities that are small differences of large ones (Mises at nu~0.5)."""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / scale)


# ---- row f4: meshes of several sections -----------------------------------------------------------------------------
def sectioned_K(deck, u):
    """NumPy-oracle stiffness of a `meshgen.SectionedDeck` at displacement u: the sum of the sections' matrices."""
    from oracle import femcy_oracle as O
    import scipy.sparse as sp
    K = None
    nn, dm = deck.nodes.shape
    N = nn * dm
    keys = []
    for s in deck.sections:
        Kp = O.assemble_K(deck.nodes, s["elements"], u, s["etype"], np.asarray(s["material"].C))
        K = Kp if K is None else K + Kp
        r, c = O.pattern(s["elements"], nn, dm)
        keys.append(r.astype(np.int64) * N + c)
    # the full structural pattern (union of the sections' couplings), exact zeros kept -- like the device matrix
    key = np.unique(np.concatenate(keys))
    rows, cols = key // N, key % N
    K = sp.csr_matrix((O.csr_on_pattern(K.tocsr(), rows, cols), (rows, cols)), shape=(N, N))
    K.sort_indices()
    return K


def sectioned_direct_solution(deck, rhs):
    """oracle solution of the linear deck: sequential Dirichlet elimination + SuperLU"""
    import scipy.sparse.linalg as sl
    from oracle import femcy_oracle as O
    dm = deck.nodes.shape[1]
    K = sectioned_K(deck, np.zeros(deck.nodes.size))
    dofs = np.concatenate([bc["node_set"] * dm + bc["dof"] for bc in deck.dirichlet_bc_info])
    Kb, rb = O.dirichlet_linear(K, rhs, dofs, np.zeros(len(dofs)))
    return sl.spsolve(Kb.tocsc(), rb), K


def material_oracle_args(mat):
    """(class name, params, C) as the oracle's stress functions take them"""
    name = type(mat).__name__
    p = (float(mat.C1), float(mat.D1)) if name == "NeoHookean" else (float(mat.modulus), float(mat.poisson_ratio))
    return name, p, np.asarray(mat.C, dtype=np.float64)
