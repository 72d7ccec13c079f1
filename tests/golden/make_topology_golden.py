#!/usr/bin/env python
"""Golden vectors of the reference's TOPOLOGY code (row f1): `Body.get_boundary` (`/root/reference/body.py:197-234`) and
`Body.get_nodeEles` (`:165-179`) executed UNMODIFIED under the sequential taichi shim (oracle/taichi_shim), for a handful of the
reference's own decks.  Run in the build container (needs /root/reference); the output is committed:

    python tests/golden/make_topology_golden.py        # writes tests/golden/topology_reference.npz

For every deck: the boundary dict as (sorted facet node tuples [nb, width], owning element [nb]) in lexicographic facet order, and
nodeEles as a CSR pair (ptr [nn+1], elements ascending per node).  The deck's nodes / connectivity come from the goldens the
kernel tests already use (same reader run), so nothing but this file is needed on the GPU box."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FEMCY_REFERENCE", "/root/reference")
DECKS = ["cps3_ellip", "cps6_ellip", "cps4_ellip", "cps8_ellip", "c3d4_ellip", "c3d10_ellip", "c3d4_cook"]


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "taichi_shim"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    import taichi as ti  # the shim
    from reader.inp_info import InpInfo
    from body import Body
    ti.init(arch=ti.cpu, default_fp=ti.f64)
    out = {}
    for name in DECKS:
        g = np.load(os.path.join(HERE, name + ".npz"))
        inp = InpInfo(os.path.join(REF, str(g["deck"])))
        elements = list(inp.eSets.values())[0]
        assert np.array_equal(elements, g["elements"]) and np.allclose(inp.nodes, g["nodes"])
        body = Body(nodes=inp.nodes, elements=elements, ELE=inp.ELE)
        bnd = body.get_boundary()                                   # the reference's own loop over Python dicts
        facets = sorted(bnd.keys())
        out[name + "_facets"] = np.array(facets, dtype=np.int32)
        out[name + "_owner"] = np.array([bnd[f] for f in facets], dtype=np.int32)
        ne = body.get_nodeEles()                                    # lists built from Python sets: order not defined -> sorted
        ptr = np.zeros(len(ne) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(v) for v in ne])
        out[name + "_ne_ptr"] = ptr
        out[name + "_ne_list"] = np.concatenate([np.sort(np.array(v, dtype=np.int32)) for v in ne])
        out[name + "_boundary_nodes"] = np.array(sorted(body.boundaryNodes), dtype=np.int32)
        print(name, len(facets), "boundary facets,", int(ptr[-1]), "node-element pairs")
    np.savez_compressed(os.path.join(HERE, "topology_reference.npz"), **out)


if __name__ == "__main__":
    main()
