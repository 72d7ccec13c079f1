"""Element / row partition of a mesh over the GPUs of one box.

The reference is single-device; this layer is new.  Scheme:

  * nodes are split into `nranks` contiguous chunks of a 1-D ordering (coordinate along the
    longest bounding-box axis, ties by id) -> slabs; a rank OWNS the matrix rows of its nodes;
  * a rank's local mesh = every element that touches one of its nodes (interface elements are
    integrated redundantly by both neighbours, so assembly needs NO communication and every owned
    row is complete);
  * local node numbering = owned nodes first (ascending global id), then ghost nodes grouped by
    owner rank (ascending global id inside a group), so each peer's ghosts are one contiguous range
    and the send list of the owner has the same order;
  * per CG iteration the ghost entries of the direction vector and the partial dot products are
    exchanged through NVLink peer memory from inside the CG kernels (cudaIpc-mapped windows, see
    `_install_p2p` and femcy_b200/csrc/cg.cu); the NCCL path (grouped ncclSend/ncclRecv + all-gather,
    femcy_b200/csrc/comm.cu) is the fallback and serves the one-off exchanges outside the iteration.

`Partition(..., device=<gpu>)` (or `ctx=<Context>`) computes the piece on the rank's own GPU (femcy_partition,
csrc/partition.cu: one sort of the node coordinates, flag / scan / compact kernels); without either it is the host-side NumPy
statement of the same scheme below, which the world_size-2 gloo tests on CPU cover and which the device result must equal
array for array (tests/test_partition.py on the emulated kernels, tests/test_gpu_topology.py on the GPU).
"""
import os

import numpy as np


def slab_axis(nodes, axis=None):
    """coordinate axis of the slabs: the longest bounding-box axis; a cube takes its last axis"""
    if axis is not None:
        return int(axis)
    ext = nodes.max(axis=0) - nodes.min(axis=0)
    axis = int(np.argmax(ext))
    # prefer the slowest-varying axis of a lexicographically numbered box (keeps ids contiguous)
    if np.allclose(ext, ext[0]):
        axis = nodes.shape[1] - 1
    return axis


def chunk_bounds(nn, nranks, weights=None):
    """bounds [nranks+1] of the ranks' chunks of the sorted node order: equal counts, or proportional to `weights`"""
    if weights is None:
        return ((np.arange(nranks + 1) * nn) // nranks).astype(np.int64)
    w = np.asarray(weights, dtype=np.float64)
    if w.shape != (nranks,) or not np.all(w > 0):
        raise ValueError("weights: one positive number per rank")
    bounds = np.concatenate([[0], np.floor(np.cumsum(w) / w.sum() * nn + 1e-9).astype(np.int64)])
    bounds[-1] = nn
    return bounds.astype(np.int64)


def node_owners(nodes, nranks, axis=None, weights=None):
    """owner rank of every node: contiguous chunks along one coordinate axis -- equal counts, or counts proportional
    to `weights` (one positive number per rank, e.g. each GPU's measured memory bandwidth: every PCG iteration waits
    for the slowest rank, and the GPUs of one box differ by up to 20 % in SpMV rate, profiles/r1_notes.md)."""
    nn = nodes.shape[0]
    if nranks == 1:
        return np.zeros(nn, dtype=np.int32)
    axis = slab_axis(nodes, axis)
    order = np.lexsort((np.arange(nn), nodes[:, axis]))
    owner = np.empty(nn, dtype=np.int32)
    bounds = chunk_bounds(nn, nranks, weights)
    for r in range(nranks):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


class Partition:
    """The piece of a global mesh that rank `rank` of `nranks` works on."""

    def __init__(self, nodes, elements, rank, nranks, axis=None, owner=None, weights=None, device=None, ctx=None):
        self.rank, self.nranks = int(rank), int(nranks)
        elements = np.asarray(elements)
        nn = nodes.shape[0]
        self.nn_global, self.ne_global, self.dm = nn, elements.shape[0], nodes.shape[1]
        self.built_on = "host"
        if (device is not None or ctx is not None) and owner is None:
            self._build_on_device(nodes, elements, axis, weights, device, ctx)
            return
        self.owner = node_owners(nodes, nranks, axis, weights) if owner is None else np.asarray(owner, dtype=np.int32)
        own_e = self.owner[elements]                                   # [ne, n_en]
        touches = (own_e == rank).any(axis=1)
        self.elem_ids = np.nonzero(touches)[0]                          # local elements (global ids)
        # an element is "primary" on the rank owning its first node: used to report each element once
        self.elem_primary = own_e[self.elem_ids, 0] == rank
        loc_conn_g = elements[self.elem_ids]
        owned = np.nonzero(self.owner == rank)[0]
        used = np.unique(loc_conn_g)
        ghosts = used[self.owner[used] != rank]
        ghosts = ghosts[np.lexsort((ghosts, self.owner[ghosts]))]       # grouped by owner, ascending id
        self.n_own = int(owned.size)
        self.local_to_global = np.concatenate([owned, ghosts]).astype(np.int64)
        self.n_local = int(self.local_to_global.size)
        g2l = np.full(nn, -1, dtype=np.int64)
        g2l[self.local_to_global] = np.arange(self.n_local)
        self.global_to_local = g2l
        self.nodes = np.ascontiguousarray(nodes[self.local_to_global])
        self.elements = np.ascontiguousarray(g2l[loc_conn_g]).astype(np.int32)

        # ---- halo plan -------------------------------------------------------------------------
        gown = self.owner[ghosts]
        self.peers, self.recv_ptr, self.send_ptr = [], [0], [0]
        recv_nodes, send_nodes = [], []
        # what do I send?  my owned nodes that sit in an element touching a peer's node.
        mixed = (own_e != own_e[:, :1]).any(axis=1)
        em, om = elements[mixed], own_e[mixed]
        for p in range(nranks):
            if p == rank:
                continue
            rn = ghosts[gown == p]
            has_p = (om == p).any(axis=1)
            cand = em[has_p][om[has_p] == rank]
            sn = np.unique(cand)
            if rn.size == 0 and sn.size == 0:
                continue
            self.peers.append(p)
            recv_nodes.append(g2l[rn])
            send_nodes.append(g2l[sn])
            self.recv_ptr.append(self.recv_ptr[-1] + rn.size)
            self.send_ptr.append(self.send_ptr[-1] + sn.size)
        self.recv_nodes = np.concatenate(recv_nodes).astype(np.int32) if recv_nodes else np.zeros(0, np.int32)
        self.send_nodes = np.concatenate(send_nodes).astype(np.int32) if send_nodes else np.zeros(0, np.int32)

    def _build_on_device(self, nodes, elements, axis, weights, device, ctx):
        """femcy_partition on this rank's GPU (csrc/partition.cu); same attributes as the NumPy path above"""
        import ctypes as C
        from ._lib import Context, as_d, as_i32, as_i64
        own_ctx = ctx is None
        if own_ctx:
            ctx = Context(int(device))
        try:
            nn, dm = nodes.shape
            ne, n_en = elements.shape
            ax = slab_axis(nodes, axis) if self.nranks > 1 else 0
            bounds = np.ascontiguousarray(chunk_bounds(nn, self.nranks, weights))
            nd = np.ascontiguousarray(nodes, dtype=np.float64)
            el = np.ascontiguousarray(elements, dtype=np.int32)
            sizes = np.zeros(6, dtype=np.int64)
            ctx.call("femcy_partition", dm, nn, as_d(nd), ne, n_en, as_i32(el), self.rank, self.nranks, ax, as_i64(bounds), as_i64(sizes))
            n_own, n_local, ne_local, npeers, n_send, n_recv = (int(v) for v in sizes)
            self.owner = np.empty(max(nn, 1), dtype=np.int32)
            self.elem_ids = np.empty(max(ne_local, 1), dtype=np.int64)
            prim = np.empty(max(ne_local, 1), dtype=np.uint8)
            self.local_to_global = np.empty(max(n_local, 1), dtype=np.int64)
            loc_el = np.empty(max(ne_local * n_en, 1), dtype=np.int32)
            loc_nd = np.empty(max(n_local * dm, 1), dtype=np.float64)
            peers = np.zeros(8, dtype=np.int32)
            sp, rp = np.zeros(9, dtype=np.int64), np.zeros(9, dtype=np.int64)
            sn, rn = np.empty(max(n_send, 1), dtype=np.int32), np.empty(max(n_recv, 1), dtype=np.int32)
            ctx.call("femcy_partition_get", as_i32(self.owner), as_i64(self.elem_ids), prim.ctypes.data_as(C.POINTER(C.c_ubyte)),
                     as_i64(self.local_to_global), as_i32(loc_el), as_d(loc_nd), as_i32(peers), as_i64(sp), as_i32(sn), as_i64(rp), as_i32(rn))
        finally:
            if own_ctx:
                ctx.close()
        self.owner = self.owner[:nn]
        self.elem_ids = self.elem_ids[:ne_local]
        self.elem_primary = prim[:ne_local].astype(bool)
        self.local_to_global = self.local_to_global[:n_local]
        self.n_own, self.n_local = n_own, n_local
        g2l = np.full(nn, -1, dtype=np.int64)
        g2l[self.local_to_global] = np.arange(n_local)
        self.global_to_local = g2l
        self.nodes = np.ascontiguousarray(loc_nd[: n_local * dm].reshape(n_local, dm))
        self.elements = np.ascontiguousarray(loc_el[: ne_local * n_en].reshape(ne_local, n_en))
        self.peers = [int(p) for p in peers[:npeers]]
        self.send_ptr, self.recv_ptr = [int(v) for v in sp[: npeers + 1]], [int(v) for v in rp[: npeers + 1]]
        self.send_nodes, self.recv_nodes = sn[:n_send].copy(), rn[:n_recv].copy()
        self.built_on = "device"

    # ---- deck localisation -------------------------------------------------------------------------
    def localize_nodes(self, node_ids):
        """global node ids -> local ids of those present on this rank (owned or ghost)."""
        l = self.global_to_local[np.asarray(node_ids, dtype=np.int64)]
        return l[l >= 0]

    def localize_deck(self, deck):
        """A deck with local nodes/elements/sets; loads restricted to facets of local elements."""
        from .meshgen import FacetSet
        import copy
        loc = copy.copy(deck)
        kind = list(deck.eSets.keys())[0]
        loc.nodes = self.nodes
        loc.eSets = {kind: self.elements}
        loc.dirichlet_bc_info = [dict(bc, node_set=self.localize_nodes(bc["node_set"])) for bc in deck.dirichlet_bc_info]
        loc.neumann_bc_info = []
        e_g2l = np.full(self.ne_global, -1, dtype=np.int64)
        e_g2l[self.elem_ids] = np.arange(self.elem_ids.size)
        gbody = None
        for nbc in deck.neumann_bc_info:
            fs = nbc["face_set"]
            if not hasattr(fs, "kid"):
                # a reader-style face set (sorted global-node tuples): locate the owner elements once on the
                # global mesh, then treat it like a meshgen.FacetSet
                from .body import Body
                if gbody is None:
                    gbody = Body(deck.nodes, deck.eSets[kind], deck.ELE)
                facets = np.array(sorted(fs), dtype=np.int64).reshape(-1, len(deck.ELE.element_facets()[0]))
                ele, kid = gbody.locate_boundary_facets(facets) if len(facets) else (np.zeros(0, np.int64), np.zeros(0, np.int64))
                fs = FacetSet(facets, ele, kid)
            keep = e_g2l[fs.ele] >= 0
            lfs = FacetSet(self.global_to_local[fs.facets[keep]], e_g2l[fs.ele[keep]], fs.kid[keep])
            loc.neumann_bc_info.append(dict(nbc, face_set=lfs))
        return loc

    # ---- device installation -------------------------------------------------------------------------
    def install(self, ctx, comm=None):
        """Create the NCCL communicator of the ctx (once per process group) and upload the halo plan."""
        from ._lib import as_i32, as_i64
        import ctypes as C
        if comm is None:
            comm = getattr(self, "comm", None)
        if comm is None:
            raise RuntimeError("Partition.install needs a Communicator (femcy_b200.partition.Communicator)")
        self.comm = comm
        uid = comm.unique_id(ctx)
        buf = (C.c_char * 128).from_buffer_copy(uid)
        ctx.call("femcy_comm_init", self.rank, self.nranks, C.cast(buf, C.c_void_p), comm.nccl_path.encode())
        peers = np.ascontiguousarray(self.peers, dtype=np.int32)
        sp = np.ascontiguousarray(self.send_ptr, dtype=np.int64)
        rp = np.ascontiguousarray(self.recv_ptr, dtype=np.int64)
        sn = np.ascontiguousarray(self.send_nodes, dtype=np.int32)
        rn = np.ascontiguousarray(self.recv_nodes, dtype=np.int32)
        ctx.call("femcy_set_halo", len(self.peers), as_i32(peers), as_i64(sp), as_i32(sn), as_i64(rp), as_i32(rn))
        if self.nranks > 1 and not os.environ.get("FEMCY_NO_P2P_SETUP"):
            self._install_p2p(ctx, comm)

    def _install_p2p(self, ctx, comm):
        """NVLink peer-memory path of the CG loop: exchange cudaIpc handles of every rank's
        {flag window, direction vector d} and tell each rank where its boundary nodes live in the
        neighbours' numbering.  Falls back to the NCCL path (with a note) if IPC is not available."""
        import ctypes as C
        from ._lib import FemcyError, as_i64
        buf = (C.c_char * 128)()
        try:
            ctx.call("femcy_p2p_export", C.cast(buf, C.c_void_p))
            mine = bytes(buf.raw)
            ok = True
        except FemcyError as e:      # pragma: no cover - depends on the box
            mine, ok = bytes(128), False
            self.p2p_error = str(e)
        # every rank publishes, for each of its peers, where that peer's ghosts start in ITS numbering
        starts = {int(p): int(self.recv_nodes[self.recv_ptr[k]]) if self.recv_ptr[k + 1] > self.recv_ptr[k] else -1
                  for k, p in enumerate(self.peers)}
        gathered = comm.allgather_object((ok, mine, starts))
        self.p2p = all(g[0] for g in gathered)
        if not self.p2p:
            return
        blob = b"".join(g[1] for g in gathered)
        remote_start = np.ascontiguousarray([gathered[p][2].get(self.rank, -1) for p in self.peers], dtype=np.int64)
        if (remote_start < 0).any() and len(self.send_nodes):
            for k, p in enumerate(self.peers):
                if remote_start[k] < 0 and self.send_ptr[k + 1] > self.send_ptr[k]:
                    raise RuntimeError("asymmetric halo plan")
        hb = (C.c_char * len(blob)).from_buffer_copy(blob)
        try:
            ctx.call("femcy_p2p_import", C.cast(hb, C.c_void_p), as_i64(remote_start))
            imported = True
        except FemcyError as e:      # pragma: no cover
            imported = False
            self.p2p_error = str(e)
        self.p2p = all(comm.allgather_object(imported))
        if not self.p2p:
            ctx.set_option("no_p2p", 1)          # consistent choice on every rank

    def gather_global(self, local_vec, comm):
        """Assemble the global nodal vector from every rank's owned entries (host side)."""
        dm = self.dm
        mine = np.ascontiguousarray(local_vec[: self.n_own * dm])
        parts = comm.allgather_object((self.local_to_global[: self.n_own], mine))
        out = np.zeros(self.nn_global * dm)
        for ids, vals in parts:
            out.reshape(-1, dm)[ids] = vals.reshape(-1, dm)
        return out


def find_nccl_library():
    """Path of the NCCL shared object torch itself uses (so the process holds one NCCL)."""
    try:
        import nvidia.nccl as nn_
        p = os.path.join(os.path.dirname(nn_.__file__), "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return "libnccl.so.2"


class Communicator:
    """Thin wrapper over torch.distributed for the host-side plumbing (bootstrap of the NCCL unique
    id, object gathers, scalar reductions of Newton-level norms).  The data path of the CG loop
    does not go through here -- it is NCCL called from the C library."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        self.nccl_path = find_nccl_library()

    def unique_id(self, ctx):
        import ctypes as C
        obj = [None]
        if self.rank == 0:
            buf = (C.c_char * 128)()
            rc = ctx.lib.femcy_comm_unique_id(self.nccl_path.encode(), C.cast(buf, C.c_void_p))
            if rc != 0:
                raise RuntimeError(f"femcy_comm_unique_id failed ({rc})")
            obj = [bytes(buf.raw)]
        self.dist.broadcast_object_list(obj, src=0, group=self.group)
        return obj[0]

    def allgather_object(self, obj):
        out = [None] * self.nranks
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def allreduce_sum_max(self, s, m):
        """(sum over ranks of s, max over ranks of m) for python floats."""
        parts = self.allgather_object((float(s), float(m)))
        return sum(p[0] for p in parts), max(p[1] for p in parts)

    def barrier(self):
        self.dist.barrier(group=self.group)


def measure_device_bandwidth(device, nbytes=1 << 28, reps=5):
    """device-to-device copy rate of one GPU in GB/s (read + write), timed with CUDA events: the weight of a rank for
    `node_owners(..., weights=)`.  Uses torch for memory and events (plumbing); needs a CUDA device."""
    import torch
    with torch.cuda.device(device):
        a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        b = torch.empty_like(a)
        for _ in range(2):
            b.copy_(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        del a, b
    return 2.0 * nbytes / ms / 1e6

