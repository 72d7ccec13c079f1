"""Mesh container + topology queries.

Keeps the data members and query names of the reference's `Body`
(`/root/reference/body.py:12-35,165-234`): `nodes`, `elements`, `np_nodes`, `np_elements`, `dm`,
`ELE`, `get_nodeEles()`, `get_coElement_nodes()`, `get_boundary()` (+ `boundary`, `facetDic`,
`node2boundary`, `boundaryNodes`).  Rendering (`show`, `show2d`, colour bars; body.py:38-162,237-292)
is out of scope.

The reference builds these with Python loops over Python sets (minutes at 1e5 elements); here
they are vectorised NumPy sorts so the same queries stay usable on multi-million-element meshes.
The sparsity pattern itself is NOT taken from here any more -- it is built on the GPU from the
connectivity (`femcy_build_pattern`); `get_coElement_nodes` stays for API compatibility and for
cross-checking that pattern in the tests.
"""
import numpy as np

from .fields import HostField


def _rows_as_keys(a):
    """View each row of a contiguous int64 2-D array as one opaque key (for unique / searchsorted)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a.view([("", a.dtype)] * a.shape[1]).reshape(-1)


class Body:
    def __init__(self, nodes: np.ndarray, elements: np.ndarray, ELE) -> None:
        self.np_nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.np_elements = np.ascontiguousarray(elements, dtype=np.int64)
        self.nodes = HostField(self.np_nodes)
        self.elements = HostField(self.np_elements, dtype=np.int32)
        self.dm = int(self.np_nodes.shape[1])
        self.ELE = ELE
        # row f1: a System_of_equations built on this body answers the topology queries below from the device
        # (femcy_boundary_facets / femcy_node_elements: one radix sort each); without one they are NumPy sorts
        self._device_topology = None

    def _device(self):
        d = self._device_topology
        return d if d is not None and d.alive() else None

    # ---- node -> elements ------------------------------------------------------------------
    def node_element_csr(self):
        """(ptr [nn+1], elems [ne*n_en]) : elements incident to every node, ascending."""
        if not hasattr(self, "_ne_csr") and self._device() is not None:
            ptr, lst = self._device().node_elements()
            self._ne_csr = (ptr.astype(np.int64), lst.astype(np.int64))
        if not hasattr(self, "_ne_csr"):
            ne, n_en = self.np_elements.shape
            flat = self.np_elements.reshape(-1)
            owner = np.repeat(np.arange(ne, dtype=np.int64), n_en)
            order = np.argsort(flat, kind="stable")
            ptr = np.zeros(self.np_nodes.shape[0] + 1, dtype=np.int64)
            np.cumsum(np.bincount(flat, minlength=self.np_nodes.shape[0]), out=ptr[1:])
            self._ne_csr = (ptr, owner[order])
        return self._ne_csr

    def get_nodeEles(self, redo=False):
        if not hasattr(self, "nodeEles") or redo:
            if redo and hasattr(self, "_ne_csr"):
                del self._ne_csr
            ptr, el = self.node_element_csr()
            self.nodeEles = [sorted(set(el[ptr[i]:ptr[i + 1]].tolist())) for i in range(len(ptr) - 1)]
        return self.nodeEles

    # ---- node -> co-element nodes (the reference's sparsity pattern source) ---------------------
    def coelement_csr(self):
        """(ptr [nn+1], cols) : sorted distinct nodes sharing an element with each node."""
        if not hasattr(self, "_co_csr"):
            conn = self.np_elements
            n_en = conn.shape[1]
            nn = self.np_nodes.shape[0]
            i = np.repeat(conn, n_en, axis=1).reshape(-1)
            j = np.tile(conn, (1, n_en)).reshape(-1)
            key = np.unique(i * nn + j)
            rows, cols = key // nn, key % nn
            ptr = np.zeros(nn + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=nn), out=ptr[1:])
            self._co_csr = (ptr, cols)
        return self._co_csr

    def get_coElement_nodes(self, redo=False):
        if not hasattr(self, "coElement_nodes") or redo:
            if redo and hasattr(self, "_co_csr"):
                del self._co_csr
            ptr, cols = self.coelement_csr()
            self.coElement_nodes = [cols[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]
        return self.coElement_nodes

    # ---- boundary facets ----------------------------------------------------------------------------
    def boundary_arrays(self):
        """All element facets that belong to exactly one element:
        (facet_nodes [nb, k] sorted global ids, element [nb], local_key_index [nb])."""
        if not hasattr(self, "_bnd") and self._device() is not None:
            # the device finds the facets owned by one element; their sorted node tuples are read off the connectivity
            ele, kid = self._device().boundary_facets()
            ele, kid = ele.astype(np.int64), kid.astype(np.int64)
            keys = np.asarray(self.ELE.element_facets(), dtype=np.int64)
            facs = np.sort(np.take_along_axis(self.np_elements[ele], keys[kid], axis=1), axis=1) if len(ele) else \
                np.zeros((0, keys.shape[1]), dtype=np.int64)
            self._bnd = (facs, ele, kid)
        if not hasattr(self, "_bnd"):
            keys = self.ELE.element_facets()
            conn = self.np_elements
            facs = np.concatenate([np.sort(conn[:, list(k)], axis=1) for k in keys])
            ele = np.tile(np.arange(conn.shape[0], dtype=np.int64), len(keys))
            kid = np.repeat(np.arange(len(keys), dtype=np.int64), conn.shape[0])
            # group equal facets: lexsort over the (few) node columns is several times faster than sorting opaque
            # structured keys (5.2 M facets of a 1.3 M-element mesh: 4.0 s -> about 1 s)
            order = np.lexsort(tuple(facs[:, c] for c in range(facs.shape[1] - 1, -1, -1)))
            srt = facs[order]
            first = np.ones(len(srt), dtype=bool)
            first[1:] = np.any(srt[1:] != srt[:-1], axis=1)
            last = np.ones(len(srt), dtype=bool)
            last[:-1] = first[1:]
            single = order[first & last]
            single.sort()
            self._bnd = (facs[single], ele[single], kid[single])
            self._all_facets = (facs, ele)
        return self._bnd

    def get_boundary(self, redo=False):
        if not hasattr(self, "boundary") or redo:
            if redo and hasattr(self, "_bnd"):
                del self._bnd
            facs, ele, _ = self.boundary_arrays()
            self.boundary = {tuple(f): int(e) for f, e in zip(facs.tolist(), ele.tolist())}
            if not hasattr(self, "_all_facets"):
                keys = self.ELE.element_facets()
                conn = self.np_elements
                self._all_facets = (np.concatenate([np.sort(conn[:, list(k)], axis=1) for k in keys]),
                                    np.tile(np.arange(conn.shape[0], dtype=np.int64), len(keys)))
            allf, alle = self._all_facets
            facetDic = {}
            for f, e in zip(map(tuple, allf.tolist()), alle.tolist()):
                facetDic.setdefault(f, []).append(e)
            self.facetDic = facetDic
            node2boundary = {}
            for f in self.boundary:
                for n in f:
                    node2boundary.setdefault(n, set()).add(f)
            self.node2boundary = node2boundary
            self.boundaryNodes = set(node2boundary.keys())
        return self.boundary

    def locate_boundary_facets(self, facets, missing_ok=False):
        """For an array [nf, k] of sorted global facet node ids: (element, local facet key index).
        missing_ok: facets that are not boundary facets of this body get element -1 instead of a KeyError."""
        facs, ele, kid = self.boundary_arrays()
        facets = np.asarray(facets, dtype=np.int64)
        if facets.ndim != 2 or facets.shape[1] != facs.shape[1] or len(facs) == 0:
            if missing_ok:
                return np.full(len(facets), -1, dtype=np.int64), np.full(len(facets), -1, dtype=np.int64)
            raise KeyError("a loaded facet is not on the boundary of the mesh")
        bk = _rows_as_keys(facs)
        order = np.argsort(bk, kind="stable")
        q = _rows_as_keys(np.sort(facets, axis=1))
        pos = np.searchsorted(bk[order], q)
        pos = np.clip(pos, 0, len(order) - 1)
        hit = order[pos]
        found = bk[hit] == q
        if not np.all(found):
            if not missing_ok:
                raise KeyError("a loaded facet is not on the boundary of the mesh")
            return np.where(found, ele[hit], -1), np.where(found, kid[hit], -1)
        return ele[hit], kid[hit]


class SectionedBody(Body):
    """Row f4: one node set carrying several SECTIONS -- element sets with their own element kind and material (Abaqus
    `*Solid Section`).  The reference accepts a single element kind and the first material only
    (`/root/reference/reader/inp_info.py:125-128`, `main.py:24`).

        body = SectionedBody(nodes, [(elements_0, ELE_0, material_0), (elements_1, ELE_1, material_1), ...])
        system = System_of_equations(body, None, nlgeom)        # the materials travel with the sections

    `parts[i]` is a plain `Body` over the shared nodes; the `Body` attributes of this object (`np_elements`, `ELE`,
    topology queries) describe section 0, so code written for one section keeps reading something sensible."""

    def __init__(self, nodes, sections):
        if len(sections) < 1:
            raise ValueError("SectionedBody needs at least one section")
        first = sections[0]
        super().__init__(nodes, first[0], first[1])
        self.parts = [Body(self.np_nodes, s[0], s[1]) for s in sections]
        self.materials = [s[2] if len(s) > 2 else None for s in sections]
        for p in self.parts:
            if p.np_elements.ndim != 2 or p.np_elements.shape[1] != p.ELE.n_en:
                raise ValueError("a section's connectivity does not match its element kind")
            if p.ELE.dm != self.dm:
                raise ValueError("all sections of a mesh must have the dimension of its nodes")
