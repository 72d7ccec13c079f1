"""Synthetic benchmark meshes and decks -- the reference has no generator.

  kuhn_box_c3d4   : box of nx*ny*nz cells, 6 Kuhn tetrahedra per cell (the permutation paths along
                    the cell diagonal), node order per tet fixed so det[x1-x2, x3-x2, x0-x2] > 0
                    (the C3D4 convention of element_linear_tetrahedral.py:74-82).
  kuhn_box_c3d10  : same tets with mid-edge nodes in the Abaqus edge order
                    0-1, 1-2, 2-0, 0-3, 1-3, 2-3 (element_quadratic_tetrahedral.py:89-105).
  SyntheticDeck   : an object with the fields of `InpInfo` (reader/inp_info.py) so that
                    `System_of_equations.solve(deck)` runs unchanged: clamp on face x=0, a TRVEC
                    traction on face x=Lx.

Nodes are numbered lexicographically (x fastest); elements cell by cell, so consecutive elements
touch neighbouring rows of K (L2-friendly scatter) and x-slabs / z-slabs are contiguous id ranges.
"""
import itertools

import numpy as np

from .element_zoo import Element_linear_tetrahedral, Element_quadratic_tetrahedral
from .material_zoo import LinearIsotropic, NeoHookean

_PERMS = list(itertools.permutations(range(3)))


_SPREAD = None


def morton_key(i, j, k, bits=10):
    """Interleave the low `bits` (<= 10) bits of three non-negative integer arrays (Z-order key)."""
    global _SPREAD
    if _SPREAD is None:
        t = np.arange(1024, dtype=np.int64)
        sp = np.zeros(1024, dtype=np.int64)
        for b in range(10):
            sp |= ((t >> b) & 1) << (3 * b)
        _SPREAD = sp
    m = (1 << bits) - 1
    return _SPREAD[np.asarray(i) & m] | (_SPREAD[np.asarray(j) & m] << 1) | (_SPREAD[np.asarray(k) & m] << 2)


def locality_order(nodes, elements, bits=10):
    """Permutation that sorts elements along a Z-order curve of their centroids.  Consecutive elements
    then touch a compact set of K rows, which keeps the scatter assembly's working set inside L2
    whatever the mesh size (lexicographic plane-by-plane orders revisit a row one whole plane later)."""
    dm = nodes.shape[1]
    n_corner = min(elements.shape[1], dm + 1 if elements.shape[1] in (3, 4, 6, 10) else 4)
    c = np.zeros((elements.shape[0], dm))
    for a in range(n_corner):
        c += nodes[elements[:, a]]
    c /= n_corner
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = np.minimum(((c - lo) * ((2 ** bits) / np.maximum(hi - lo, 1e-300))).astype(np.int64), 2 ** bits - 1)
    key = morton_key(q[:, 0], q[:, 1], q[:, 2] if dm == 3 else np.zeros(q.shape[0], dtype=np.int64), bits)
    return np.argsort(key, kind="stable")


def _grid_nodes(nx, ny, nz, lengths):
    xs = np.linspace(0., lengths[0], nx + 1)
    ys = np.linspace(0., lengths[1], ny + 1)
    zs = np.linspace(0., lengths[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def kuhn_box_c3d4(n=None, cells=None, lengths=(1., 1., 1.), jitter=0.0, seed=0, dtype=np.int32):
    nx, ny, nz = (n, n, n) if cells is None else cells
    nodes = _grid_nodes(nx, ny, nz, lengths)
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (k * sy + j * sx + i).ravel().astype(np.int64)
    strides = np.array([1, sx, sy], dtype=np.int64)
    conn = np.empty((base.size, 6, 4), dtype=np.int64)
    for t, perm in enumerate(_PERMS):
        v = [base]
        for ax in perm:
            v.append(v[-1] + strides[ax])
        # path vertices v0..v3; orientation depends on the permutation parity only
        parity = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b]) % 2
        order = (0, 1, 2, 3) if parity == 0 else (1, 0, 2, 3)
        for c, o in enumerate(order):
            conn[:, t, c] = v[o]
    conn = conn.reshape(-1, 4)
    # make det[x1-x2, x3-x2, x0-x2] > 0 for every tet (checked on the first cell, all cells congruent)
    x = nodes[conn[:6]]
    det = np.linalg.det(np.stack([x[:, 1] - x[:, 2], x[:, 3] - x[:, 2], x[:, 0] - x[:, 2]], axis=2))
    flip = np.tile(det < 0, base.size)
    conn[flip, 0], conn[flip, 1] = conn[flip, 1].copy(), conn[flip, 0].copy()
    if jitter > 0.:
        rng = np.random.default_rng(seed)
        h = min(lengths[0] / nx, lengths[1] / ny, lengths[2] / nz)
        interior = np.all((nodes > 1e-12) & (nodes < np.array(lengths) - 1e-12), axis=1)
        nodes[interior] += jitter * h * rng.uniform(-1., 1., size=(int(interior.sum()), 3))
    return nodes, conn.astype(dtype)


def kuhn_box_c3d10(n=None, cells=None, lengths=(1., 1., 1.), dtype=np.int32):
    nx, ny, nz = (n, n, n) if cells is None else cells
    _, corner = kuhn_box_c3d4(cells=(nx, ny, nz), lengths=lengths, dtype=np.int64)
    fine = _grid_nodes(2 * nx, 2 * ny, 2 * nz, lengths)
    # corner node (i,j,k) of the coarse grid sits at (2i,2j,2k) of the fine grid
    cs = np.array([1, nx + 1, (nx + 1) * (ny + 1)])
    fs = np.array([1, 2 * nx + 1, (2 * nx + 1) * (2 * ny + 1)])
    ck = corner // cs[2]
    cj = (corner % cs[2]) // cs[1]
    ci = corner % cs[1]
    fine_id = 2 * ci * fs[0] + 2 * cj * fs[1] + 2 * ck * fs[2]      # [ne,4]
    edges = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
    mids = [(fine_id[:, a] + fine_id[:, b]) // 2 for a, b in edges]   # midpoint index = mean of fine indices
    conn = np.concatenate([fine_id, np.stack(mids, axis=1)], axis=1)
    return fine, conn.astype(dtype)


class FacetSet:
    """A loaded surface given directly as (sorted facet node ids, owning element, local facet key
    index): lets `System_of_equations.neumann_vector` skip the all-facets boundary search."""

    def __init__(self, facets, ele, kid):
        self.facets, self.ele, self.kid = facets, ele, kid

    def __len__(self):
        return len(self.facets)


def face_facets(nodes, conn, ELE, axis, value, tol=1e-9):
    """Facets of the mesh lying in the plane x_axis == value -> FacetSet."""
    on = np.abs(nodes[:, axis] - value) < tol
    keys = ELE.element_facets()
    out_f, out_e, out_k = [], [], []
    cand = np.nonzero(on[conn[:, :4]].sum(axis=1) >= 3)[0]      # elements with >= 3 corner nodes on the plane
    c = conn[cand]
    for kid, key in enumerate(keys):
        corners = [a for a in key if a < 4]
        sel = np.all(on[c[:, corners]], axis=1)
        if sel.any():
            out_f.append(np.sort(c[sel][:, list(key)].astype(np.int64), axis=1))
            out_e.append(cand[sel])
            out_k.append(np.full(int(sel.sum()), kid, dtype=np.int64))
    if not out_f:
        return FacetSet(np.zeros((0, len(keys[0])), dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    return FacetSet(np.concatenate(out_f), np.concatenate(out_e), np.concatenate(out_k))


class SyntheticDeck:
    """InpInfo-shaped description of the unit-box benchmark problems:
    all components fixed on x=0, TRVEC traction `traction` along `direction` on x=Lx."""

    def __init__(self, kind="C3D4", n=8, cells=None, lengths=(1., 1., 1.), material=None, nlgeom=False,
                 traction=1.0, direction=(0., 1., 0.), jitter=0.0, time_incs=None):
        if kind == "C3D4":
            self.nodes, conn = kuhn_box_c3d4(n, cells, lengths, jitter)
            self.ELE = Element_linear_tetrahedral()
            mat = material or LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3)
            mkey = "Elastic"
        elif kind == "C3D10":
            self.nodes, conn = kuhn_box_c3d10(n, cells, lengths)
            self.ELE = Element_quadratic_tetrahedral()
            mat = material or NeoHookean(C1=0.4, D1=20.)
            mkey = "Hyperelastic, neo hooke"
        else:
            raise ValueError(kind)
        self.kind = kind
        self.eSets = {kind: conn}
        self.materials = {mkey: mat}
        self.geometric_nonlinear = bool(nlgeom)
        self.time_incs = time_incs or {"ini_inc": 1., "max_time": 1., "min_inc": 1e-5, "max_inc": 1.}
        fixed = np.nonzero(np.abs(self.nodes[:, 0]) < 1e-9)[0].astype(np.int64)
        self.node_sets = {"fixed": fixed}
        self.ele_sets, self.face_sets = {}, {}
        self.dirichlet_bc_info = [{"node_set": fixed, "dof": c, "val": 0., "user": False} for c in range(3)]
        loaded = face_facets(self.nodes, conn, self.ELE, 0, lengths[0])
        self.face_sets["loaded"] = loaded
        self.neumann_bc_info = [{"face_set": loaded, "traction": float(traction), "direction": np.array(direction, dtype=float)}]


def write_inp(deck, path, set_name="Set-fixed", surf_name="Surf-load"):
    """Write a SyntheticDeck as an Abaqus-style `.inp` that `reader.InpInfo` (and the reference reader,
    /root/reference/reader/inp_info.py) parses back to the same problem: *Node / *Element, an assembly-level
    *Nset for the clamp, one internal *Elset + *Surface entry per loaded face number, *Material, *Step/*Static,
    *Boundary and a TRVEC *Dsload.  Used to exercise the `.inp` front end at benchmark sizes."""
    kind = deck.kind
    conn = deck.eSets[kind]
    nodes = deck.nodes
    fs = deck.face_sets["loaded"]
    keys = [tuple(sorted(k)) for k in deck.ELE.element_facets()]
    snum = {}
    for f, subs in enumerate(deck.ELE.inp_surface_num):
        snum[tuple(sorted(subs[0]))] = f + 1
    mat = list(deck.materials.values())[0]
    with open(path, "w") as fh:
        fh.write("*Heading\n** synthetic deck written by femcy_b200.meshgen.write_inp\n*Part, name=Part-1\n*End Part\n")
        fh.write("*Assembly, name=Assembly\n*Instance, name=Part-1-1, part=Part-1\n*Node\n")
        ids = np.arange(1, nodes.shape[0] + 1)
        np.savetxt(fh, np.column_stack([ids, nodes]), fmt=["%d"] + ["%.17g"] * nodes.shape[1], delimiter=", ")
        fh.write(f"*Element, type={kind}\n")
        np.savetxt(fh, np.column_stack([np.arange(1, conn.shape[0] + 1), conn + 1]), fmt="%d", delimiter=", ")
        fh.write("*End Instance\n")

        def write_ids(vals):
            vals = np.asarray(vals, dtype=np.int64) + 1
            for i in range(0, len(vals), 16):
                fh.write(", ".join(str(v) for v in vals[i:i + 16]) + "\n")

        fh.write(f"*Nset, nset={set_name}, instance=Part-1-1\n")
        write_ids(deck.node_sets["fixed"])
        surf_lines = []
        for kid in np.unique(fs.kid):
            s = snum[keys[int(kid)]]
            name = f"_{surf_name}_S{s}"
            fh.write(f"*Elset, elset={name}, internal, instance=Part-1-1\n")
            write_ids(np.sort(fs.ele[fs.kid == kid]))
            surf_lines.append(f"{name}, S{s}\n")
        fh.write(f"*Surface, type=ELEMENT, name={surf_name}\n")
        fh.writelines(surf_lines)
        fh.write("*End Assembly\n*Material, name=Material-1\n")
        if type(mat).__name__ == "NeoHookean":
            fh.write(f"*Hyperelastic, neo hooke\n{float(mat.C1)!r}, {float(1.0 / mat.D1)!r}\n")
        else:
            fh.write(f"*Elastic\n{float(mat.modulus)!r}, {float(mat.poisson_ratio)!r}\n")
        t = deck.time_incs
        fh.write(f"*Step, name=Step-1, nlgeom={'YES' if deck.geometric_nonlinear else 'NO'}\n*Static\n")
        fh.write(f"{t['ini_inc']!r}, {t['max_time']!r}, {t['min_inc']!r}, {t['max_inc']!r}\n")
        fh.write("*Boundary\n")
        for c in range(nodes.shape[1]):
            fh.write(f"{set_name}, {c + 1}, {c + 1}\n")
        nb = deck.neumann_bc_info[0]
        d = nb["direction"]
        fh.write(f"*Dsload\n{surf_name}, TRVEC, {float(nb['traction'])!r}, {float(d[0])!r}, {float(d[1])!r}, {float(d[2])!r}\n*End Step\n")


# ---------------------------------------------------------------------------------------------------------------
# Row f4: synthetic meshes of several sections (element kinds / materials), SURVEY section 8f item 4
def _plate_grid(nx, ny, lengths, quadratic):
    """nodes of an nx*ny-cell rectangle; quadratic: the (2nx+1)*(2ny+1) fine grid (cell centres stay unused by the
    8-node quads and are dropped by the caller's renumbering)"""
    mx, my = (2 * nx, 2 * ny) if quadratic else (nx, ny)
    xs = np.linspace(0., lengths[0], mx + 1)
    ys = np.linspace(0., lengths[1], my + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    return np.column_stack([X.reshape(-1), Y.reshape(-1)]), mx + 1


def sectioned_plate(nx=8, ny=4, lengths=(2., 1.), quadratic=False):
    """Rectangle whose left half is meshed with quadrilaterals (CPS4 / CPS8) and whose right half with triangles
    (CPS3 / CPS6, each cell split along its diagonal) -- a conforming mesh of two element kinds.
    Returns nodes [nn, 2] and [(etype, connectivity)], counter-clockwise node order, Abaqus mid-side order."""
    if nx % 2:
        raise ValueError("nx must be even (the kinds meet at x = Lx/2)")
    nodes, stride = _plate_grid(nx, ny, lengths, quadratic)
    s = 2 if quadratic else 1
    quads, tris = [], []
    for j in range(ny):
        for i in range(nx):
            def nid(a, b):
                return (s * j + b) * stride + (s * i + a)
            if not quadratic:
                c = [nid(0, 0), nid(1, 0), nid(1, 1), nid(0, 1)]
                if i < nx // 2:
                    quads.append(c)
                else:
                    tris.append([c[0], c[1], c[2]])
                    tris.append([c[0], c[2], c[3]])
            else:
                c = [nid(0, 0), nid(2, 0), nid(2, 2), nid(0, 2)]
                if i < nx // 2:
                    quads.append(c + [nid(1, 0), nid(2, 1), nid(1, 2), nid(0, 1)])
                else:
                    tris.append([c[0], c[1], c[2], nid(1, 0), nid(2, 1), nid(1, 1)])
                    tris.append([c[0], c[2], c[3], nid(1, 1), nid(1, 2), nid(0, 1)])
    quads, tris = np.array(quads, dtype=np.int64), np.array(tris, dtype=np.int64)
    # drop the unused fine-grid nodes (centres of the 8-node quads) and renumber
    used = np.zeros(nodes.shape[0], dtype=bool)
    used[quads.reshape(-1)] = True
    used[tris.reshape(-1)] = True
    lut = np.cumsum(used) - 1
    names = ("CPS8", "CPS6") if quadratic else ("CPS4", "CPS3")
    return nodes[used], [(names[0], lut[quads]), (names[1], lut[tris])]


def sectioned_bar(n=4, lengths=(2., 1., 1.), mixed=False):
    """Box of Kuhn tetrahedra split at x = Lx/2.  mixed = False: C3D4 everywhere (two sections = two materials);
    mixed = True: quadratic C3D10 on the left, and on the right linear C3D4 of HALF the size on the same nodes -- every
    6-node face of the interface is covered by 4 linear faces, so no node hangs."""
    if n % 2:
        raise ValueError("n must be even (the sections meet at x = Lx/2)")
    half = 0.5 * lengths[0]
    if not mixed:
        nodes, conn = kuhn_box_c3d4(cells=(n, n // 2, n // 2), lengths=lengths, dtype=np.int64)
        cx = nodes[conn, 0].mean(axis=1)
        return nodes, [("C3D4", conn[cx < half]), ("C3D4", conn[cx > half])]
    cells = (n, n // 2, n // 2)
    fine, c10 = kuhn_box_c3d10(cells=cells, lengths=lengths, dtype=np.int64)
    fine4, c4 = kuhn_box_c3d4(cells=tuple(2 * c for c in cells), lengths=lengths, dtype=np.int64)
    assert np.allclose(fine, fine4)
    left = fine[c10[:, :4], 0].mean(axis=1) < half
    right = fine[c4, 0].mean(axis=1) > half
    return fine, [("C3D10", c10[left]), ("C3D4", c4[right])]


class SectionedDeck:
    """InpInfo-shaped deck over a mesh of several sections: `sections` as `reader.InpInfo.read_sections` returns them,
    clamp on x = 0, TRVEC traction on x = Lx; `body()` builds the `SectionedBody`."""

    def __init__(self, kind="plate_linear", n=8, materials=None, nlgeom=False, traction=1.0, direction=None, time_incs=None):
        from .element_zoo import ELEMENT_TYPES
        from .material_zoo import LinearIsotropicPlaneStress
        if kind in ("plate_linear", "plate_quadratic"):
            self.lengths = (2., 1.)
            self.nodes, parts = sectioned_plate(n, max(n // 2, 1), self.lengths, quadratic=(kind == "plate_quadratic"))
            mats = materials or [LinearIsotropicPlaneStress(modulus=2.1e5, poisson_ratio=0.3),
                                 LinearIsotropicPlaneStress(modulus=7.0e4, poisson_ratio=0.33)]
        elif kind in ("bar_bimaterial", "bar_mixed"):
            self.lengths = (2., 1., 1.)
            self.nodes, parts = sectioned_bar(n, self.lengths, mixed=(kind == "bar_mixed"))
            if nlgeom:
                mats = materials or [NeoHookean(C1=0.4, D1=20.), NeoHookean(C1=1.2, D1=8.)]
            else:
                mats = materials or [LinearIsotropic(modulus=2.1e5, poisson_ratio=0.3), LinearIsotropic(modulus=7.0e4, poisson_ratio=0.33)]
        else:
            raise ValueError(kind)
        self.kind = kind
        dm = self.nodes.shape[1]
        self.sections = [{"etype": t, "elements": c, "ELE": ELEMENT_TYPES[t](), "material": m, "material_name": "Material-%d" % (k + 1)}
                         for k, ((t, c), m) in enumerate(zip(parts, mats))]
        self.eSets = {}
        for s in self.sections:                                   # the reader's view: one array per element TYPE
            self.eSets[s["etype"]] = np.concatenate([self.eSets[s["etype"]], s["elements"]]) if s["etype"] in self.eSets else s["elements"]
        self.ELE = self.sections[0]["ELE"]
        self.materials = {s["material_name"]: s["material"] for s in self.sections}
        self.geometric_nonlinear = bool(nlgeom)
        self.time_incs = time_incs or {"ini_inc": 1., "max_time": 1., "min_inc": 1e-5, "max_inc": 1.}
        fixed = np.nonzero(np.abs(self.nodes[:, 0]) < 1e-9)[0].astype(np.int64)
        self.node_sets = {"fixed": fixed}
        self.dirichlet_bc_info = [{"node_set": fixed, "dof": c, "val": 0., "user": False} for c in range(dm)]
        body = self.body()
        loaded = set()
        self.loaded_by_section = []                              # [(section, element rows, local facet key index)]
        for k, part in enumerate(body.parts):
            facs, ele, kid = part.boundary_arrays()
            on = np.all(np.abs(self.nodes[facs, 0] - self.lengths[0]) < 1e-9, axis=1)
            loaded.update(map(tuple, facs[on].tolist()))
            self.loaded_by_section.append((k, ele[on], kid[on]))
        if direction is None:
            direction = (0., 1., 0.)
        self.face_sets = {"loaded": loaded}
        self.neumann_bc_info = [{"face_set": loaded, "traction": float(traction), "direction": np.array(direction, dtype=float)}]

    def body(self):
        from .body import SectionedBody
        return SectionedBody(self.nodes, [(s["elements"], s["ELE"], s["material"]) for s in self.sections])


def write_inp_sections(deck, path, set_name="Set-fixed", surf_name="Surf-load"):
    """Write a SectionedDeck as an Abaqus-style `.inp`: one `*Element` block per element type with deck-wide element
    labels, one instance-level `*Elset` + `*Solid Section` per section, named `*Material` blocks, the clamp and the loaded
    surface (elements looked up by label across the types).  `reader.InpInfo` parses it back to the same sections."""
    nodes = deck.nodes
    dm = nodes.shape[1]
    label0 = 1
    labels = []
    with open(path, "w") as fh:
        fh.write("*Heading\n** synthetic multi-section deck written by femcy_b200.meshgen.write_inp_sections\n")
        fh.write("*Part, name=Part-1\n*End Part\n*Assembly, name=Assembly\n*Instance, name=Part-1-1, part=Part-1\n*Node\n")
        np.savetxt(fh, np.column_stack([np.arange(1, nodes.shape[0] + 1), nodes]), fmt=["%d"] + ["%.17g"] * dm, delimiter=", ")
        for s in deck.sections:
            conn = s["elements"]
            lab = np.arange(label0, label0 + conn.shape[0])
            labels.append(lab)
            label0 += conn.shape[0]
            fh.write(f"*Element, type={s['etype']}\n")
            np.savetxt(fh, np.column_stack([lab, conn + 1]), fmt="%d", delimiter=", ")
        for k, s in enumerate(deck.sections):
            fh.write(f"*Elset, elset=Set-sec{k + 1}, generate\n{labels[k][0]}, {labels[k][-1]}, 1\n")
            fh.write(f"** Section: Section-{k + 1}\n*Solid Section, elset=Set-sec{k + 1}, material={s['material_name']}\n,\n")
        fh.write("*End Instance\n")

        def write_ids(vals):
            vals = np.asarray(vals, dtype=np.int64)
            for i in range(0, len(vals), 16):
                fh.write(", ".join(str(v) for v in vals[i:i + 16]) + "\n")

        fh.write(f"*Nset, nset={set_name}, instance=Part-1-1\n")
        write_ids(deck.node_sets["fixed"] + 1)
        surf_lines = []
        for k, ele, kid in deck.loaded_by_section:
            ELE = deck.sections[k]["ELE"]
            keys = [tuple(sorted(q)) for q in ELE.element_facets()]
            snum = {}
            for f, subs in enumerate(ELE.inp_surface_num):
                for sub in subs:
                    snum[tuple(sorted(sub))] = f + 1
            for s_no in sorted({snum[keys[int(q)]] for q in kid}):
                rows = np.unique(ele[[snum[keys[int(q)]] == s_no for q in kid]])
                name = f"_{surf_name}_{k + 1}_S{s_no}"
                fh.write(f"*Elset, elset={name}, internal, instance=Part-1-1\n")
                write_ids(labels[k][rows])
                surf_lines.append(f"{name}, S{s_no}\n")
        fh.write(f"*Surface, type=ELEMENT, name={surf_name}\n")
        fh.writelines(surf_lines)
        fh.write("*End Assembly\n")
        for s in deck.sections:
            mat = s["material"]
            fh.write(f"*Material, name={s['material_name']}\n")
            if type(mat).__name__ == "NeoHookean":
                fh.write(f"*Hyperelastic, neo hooke\n{float(mat.C1)!r}, {float(1.0 / mat.D1)!r}\n")
            else:
                fh.write(f"*Elastic\n{float(mat.modulus)!r}, {float(mat.poisson_ratio)!r}\n")
        t = deck.time_incs
        fh.write(f"*Step, name=Step-1, nlgeom={'YES' if deck.geometric_nonlinear else 'NO'}\n*Static\n")
        fh.write(f"{t['ini_inc']!r}, {t['max_time']!r}, {t['min_inc']!r}, {t['max_inc']!r}\n*Boundary\n")
        for c in range(dm):
            fh.write(f"{set_name}, {c + 1}, {c + 1}\n")
        nb = deck.neumann_bc_info[0]
        d = list(nb["direction"]) + [0., 0., 0.]
        fh.write(f"*Dsload\n{surf_name}, TRVEC, {float(nb['traction'])!r}, {float(d[0])!r}, {float(d[1])!r}, {float(d[2])!r}\n*End Step\n")
