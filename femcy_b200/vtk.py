"""Result export: legacy-VTK unstructured grid (`.vtk`, ASCII) + nodal averaging of Gauss-point fields.

The reference shows results only in its Taichi GGUI window and lists file export as future work
(`/root/reference/README.md:130`); the GUI is out of scope here, so this is the way results
leave the process: `write_vtk(path, body, point_data={...}, cell_data={...})` writes the mesh with Abaqus node order
mapped to VTK cell types (VTK's quadratic tetrahedron numbers its last two mid-edge nodes like Abaqus, so C3D10 needs
no permutation), `nodal_average(body, nodal_vals)` turns the per-element nodal values produced by `ELE.extrapolate`
(`element_base.py`, reference `element_*.py: extrapolate`) into one value per node (mean over the adjacent elements,
which is what the reference's renderer effectively shows by overdrawing)."""
import numpy as np

# Abaqus element -> (VTK cell type id, node permutation or None)
_VTK_CELL = {
    (2, 3): 5,     # VTK_TRIANGLE
    (2, 6): 22,    # VTK_QUADRATIC_TRIANGLE   (mid-side nodes 3=(0,1) 4=(1,2) 5=(2,0): same order)
    (2, 4): 9,     # VTK_QUAD
    (2, 8): 23,    # VTK_QUADRATIC_QUAD       (4=(0,1) 5=(1,2) 6=(2,3) 7=(3,0): same order)
    (3, 4): 10,    # VTK_TETRA
    (3, 10): 24,   # VTK_QUADRATIC_TETRA      (4=(0,1) 5=(1,2) 6=(2,0) 7=(0,3) 8=(1,3) 9=(2,3): same order)
}


def nodal_average(body, nodal_vals):
    """[ne, n_en] per-element nodal values (e.g. `ELE.extrapolate(mises)`) -> [nn] mean over adjacent elements."""
    conn = body.np_elements
    vals = np.asarray(nodal_vals.to_numpy() if hasattr(nodal_vals, "to_numpy") else nodal_vals, dtype=np.float64)
    if vals.shape != conn.shape:
        raise ValueError(f"nodal values {vals.shape} do not match the connectivity {conn.shape}")
    nn = body.np_nodes.shape[0]
    s = np.bincount(conn.reshape(-1), weights=vals.reshape(-1), minlength=nn)
    c = np.bincount(conn.reshape(-1), minlength=nn)
    return s / np.maximum(c, 1)


def _write_array(fh, a, per_line=9):
    flat = np.asarray(a).reshape(-1)
    fmt = "%d" if np.issubdtype(flat.dtype, np.integer) else "%.17g"
    for i in range(0, flat.size, per_line):
        fh.write(" ".join(fmt % v for v in flat[i:i + per_line]) + "\n")


def write_vtk(path, body, point_data=None, cell_data=None, title="femcy_b200 result"):
    """Write mesh + fields.  point_data: name -> [nn] scalar, [nn, dm] / [nn*dm] vector (padded to 3 components);
    cell_data: name -> [ne] scalar (e.g. the Gauss-point mean of Mises).  A `SectionedBody` (several element kinds /
    materials, row f4) is written as one grid: the cells of its sections one after the other, each with its own VTK type;
    a cell field may then be given as a list with one array per section."""
    nodes = body.np_nodes
    nn, dm = nodes.shape
    parts = getattr(body, "parts", None) or [body]
    conns = [p.np_elements for p in parts]
    ctypes_ = []
    for c in conns:
        ctype = _VTK_CELL.get((dm, c.shape[1]))
        if ctype is None:
            raise ValueError(f"no VTK cell type for a {dm}-D element with {c.shape[1]} nodes")
        ctypes_.append(ctype)
    ne = sum(c.shape[0] for c in conns)
    with open(path, "w") as fh:
        fh.write(f"# vtk DataFile Version 3.0\n{title}\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        fh.write(f"POINTS {nn} double\n")
        p3 = np.zeros((nn, 3))
        p3[:, :dm] = nodes
        np.savetxt(fh, p3, fmt="%.17g")
        fh.write(f"CELLS {ne} {sum(c.shape[0] * (c.shape[1] + 1) for c in conns)}\n")
        for c in conns:
            np.savetxt(fh, np.column_stack([np.full(c.shape[0], c.shape[1], dtype=np.int64), c]), fmt="%d")
        fh.write(f"CELL_TYPES {ne}\n")
        _write_array(fh, np.concatenate([np.full(c.shape[0], t, dtype=np.int64) for c, t in zip(conns, ctypes_)]), per_line=32)
        if point_data:
            fh.write(f"POINT_DATA {nn}\n")
            for name, a in point_data.items():
                a = np.asarray(a.to_numpy() if hasattr(a, "to_numpy") else a, dtype=np.float64)
                if a.size == nn:
                    fh.write(f"SCALARS {name} double 1\nLOOKUP_TABLE default\n")
                    _write_array(fh, a)
                elif a.size == nn * dm:
                    v3 = np.zeros((nn, 3))
                    v3[:, :dm] = a.reshape(nn, dm)
                    fh.write(f"VECTORS {name} double\n")
                    np.savetxt(fh, v3, fmt="%.17g")
                else:
                    raise ValueError(f"point field {name}: {a.shape} is neither [nn] nor [nn, dm]")
        if cell_data:
            fh.write(f"CELL_DATA {ne}\n")
            for name, a in cell_data.items():
                a = a.to_numpy() if hasattr(a, "to_numpy") else a
                if isinstance(a, (list, tuple)):                  # one array per section
                    a = np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in a])
                a = np.asarray(a, dtype=np.float64)
                if a.size != ne:
                    raise ValueError(f"cell field {name}: {a.shape} is not [ne]")
                fh.write(f"SCALARS {name} double 1\nLOOKUP_TABLE default\n")
                _write_array(fh, a)
    return path


def read_vtk(path):
    """Minimal reader of what `write_vtk` writes (tests / post-processing without a VTK install)."""
    with open(path) as fh:
        tok = fh.read().split("\n")
    out = {"point_data": {}, "cell_data": {}}
    i = 4
    section = None
    count = 0

    def take(n_numbers, dtype=float):
        nonlocal i
        vals = []
        while len(vals) < n_numbers:
            vals.extend(tok[i].split())
            i += 1
        return np.array(vals[:n_numbers], dtype=dtype)

    while i < len(tok):
        line = tok[i].strip()
        i += 1
        if not line:
            continue
        w = line.split()
        if w[0] == "POINTS":
            out["points"] = take(int(w[1]) * 3).reshape(-1, 3)
        elif w[0] == "CELLS":
            flat = take(int(w[2]), dtype=np.int64)
            cells, pos = [], 0
            while pos < flat.size:                               # cells of several kinds: ragged
                cells.append(flat[pos + 1:pos + 1 + int(flat[pos])])
                pos += 1 + int(flat[pos])
            widths = {len(c) for c in cells}
            out["cells"] = np.array(cells) if len(widths) == 1 else cells
        elif w[0] == "CELL_TYPES":
            out["cell_types"] = take(int(w[1]), dtype=np.int64)
        elif w[0] in ("POINT_DATA", "CELL_DATA"):
            section, count = ("point_data" if w[0] == "POINT_DATA" else "cell_data"), int(w[1])
        elif w[0] == "SCALARS":
            i += 1                                    # LOOKUP_TABLE line
            out[section][w[1]] = take(count)
        elif w[0] == "VECTORS":
            out[section][w[1]] = take(count * 3).reshape(-1, 3)
    return out
