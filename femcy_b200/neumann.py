"""Neumann boundary condition: consistent nodal loads of a surface traction (row a5).

The reference computes these on the host, facet by facet in Python
(`/root/reference/stiffnessMtrx.py:369-411`, with `ELE.globalNormal`, `facet_natural_coos`,
`facet_point_weights`, `shapeFunc_pyscope` and `Body.get_boundary`).  Same arithmetic here,
vectorised over all facets that share a local facet key:

    rhs[node*dm + i] += t * (n or dir)_i * size * w_p * N_node(xi_p)

for every loaded facet, every node of the facet and every facet integration point p, where n is
the unit outward normal pushed forward from the natural normal with (dx/dxi)^-1 at xi_p, and
`size` is the facet length (2-D) or the area of the triangle of its first three nodes (3-D).
Loads act on the initial geometry (dead loads, quirk B3).
"""
import numpy as np


def neumann_vector(body, load_facets, load_val: float, load_dir=np.array([])):
    ELE, dm = body.ELE, body.dm
    rhs = np.zeros(body.np_nodes.shape[0] * dm)
    if hasattr(load_facets, "kid"):       # meshgen.FacetSet: owner elements already known
        ele, kid = load_facets.ele, load_facets.kid
        if len(ele) == 0:
            return rhs
    else:
        facets = np.array(sorted(load_facets), dtype=np.int64) if not isinstance(load_facets, np.ndarray) else load_facets
        if facets.size == 0:
            return rhs
        ele, kid = body.locate_boundary_facets(facets)
    keys = ELE.element_facets()
    load_dir = np.asarray(load_dir, dtype=np.float64)
    for k in np.unique(kid):
        key = keys[k]
        sel = np.nonzero(kid == k)[0]
        conn = body.np_elements[ele[sel]]                      # [nf, n_en]
        X = body.np_nodes[conn]                                # [nf, n_en, dm]
        nat, w, N = ELE.facet_point_table(key)
        normals = np.asarray(ELE.facet_natural_normals[key], dtype=np.float64)
        if dm == 2:
            size = np.linalg.norm(X[:, key[0]] - X[:, key[1]], axis=1)
        else:
            size = 0.5 * np.linalg.norm(np.cross(X[:, key[1]] - X[:, key[0]], X[:, key[2]] - X[:, key[0]]), axis=1)
        for p in range(len(w)):
            if load_dir.size == 0:
                dxdn = np.einsum("fai,ak->fik", X, ELE.dshape_dnat_pyscope(nat[p]))
                n = np.einsum("k,fkj->fj", normals[p], np.linalg.inv(dxdn))
                n /= (np.linalg.norm(n, axis=1, keepdims=True) + 1.e-30)
                flux = load_val * n * (size * w[p])[:, None]
            else:
                flux = load_val * load_dir[None, :dm] * (size * w[p])[:, None]
            for a in key:                                       # facet nodes only
                idx = conn[:, a, None] * dm + np.arange(dm)[None, :]
                np.add.at(rhs, idx, flux * N[p, a])
    return rhs


def neumann_vector_sections(body, load_facets, load_val: float, load_dir=np.array([])):
    """Row f4: the same consistent loads on a mesh of several sections (`body.SectionedBody`).  Every loaded facet is
    looked up among the boundary facets of each section (a facet has the node count of its own element kind) and loads
    the section that owns it; the sections' vectors add up."""
    rhs = np.zeros(body.np_nodes.shape[0] * body.dm)
    facets = sorted(load_facets) if not isinstance(load_facets, np.ndarray) else [tuple(f) for f in load_facets.tolist()]
    if len(facets) == 0:
        return rhs
    found = np.zeros(len(facets), dtype=np.int64)
    by_width = {}
    for i, f in enumerate(facets):
        by_width.setdefault(len(f), []).append(i)
    for part in body.parts:
        width = len(part.ELE.element_facets()[0])
        idx = np.asarray(by_width.get(width, []), dtype=np.int64)
        if idx.size == 0:
            continue
        q = np.array([facets[i] for i in idx], dtype=np.int64)
        ele, _ = part.locate_boundary_facets(q, missing_ok=True)
        hit = ele >= 0
        if hit.any():
            rhs += neumann_vector(part, q[hit], load_val, load_dir)
            found[idx[hit]] += 1
    if np.any(found == 0):
        raise KeyError("a loaded facet is not on the boundary of any section of the mesh")
    if np.any(found > 1):
        raise KeyError("a loaded facet lies between two sections (not a boundary facet of the mesh)")
    return rhs
