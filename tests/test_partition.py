"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (no GPU needed):
the element/row partition, the halo plan and the rank-ordered reduction are exercised with a
host stand-in for the device kernels (oracle CSR matrices), and must reproduce the global SpMV,
dot product and Neumann vector."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, emulated_device_partitioner=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from femcy_b200 import meshgen
    from femcy_b200.body import Body
    from femcy_b200.neumann import neumann_vector
    from femcy_b200.partition import Communicator, Partition
    from oracle import femcy_oracle as O

    deck = meshgen.SyntheticDeck("C3D4", n=5, jitter=0.1)
    nodes, conn = deck.nodes, deck.eSets["C3D4"]
    C = np.asarray(deck.materials["Elastic"].C)
    if emulated_device_partitioner:        # femcy_partition's orchestration + kernels on the emulation (see EmuPartitionCtx below)
        sys.path.insert(0, os.path.join(root, "tests"))
        part = Partition(nodes, conn, rank, world, ctx=EmuPartitionCtx())
        assert part.built_on == "device"
    else:
        part = Partition(nodes, conn, rank, world)
    comm = Communicator()
    dm = 3
    # local matrix: rows of owned nodes, columns local (owned + ghost)
    Kloc = O.assemble_K(part.nodes, part.elements.astype(np.int64), np.zeros(part.nodes.size), "C3D4", C)
    Kown = Kloc[: part.n_own * dm]
    rng = np.random.default_rng(7)
    xg = rng.standard_normal(nodes.size)
    x = np.zeros(part.n_local * dm)
    x.reshape(-1, dm)[: part.n_own] = xg.reshape(-1, dm)[part.local_to_global[: part.n_own]]   # ghosts unknown

    # halo exchange following the plan (gloo isend/irecv as the stand-in for ncclSend/ncclRecv)
    reqs, bufs = [], []
    for k, p in enumerate(part.peers):
        sn = part.send_nodes[part.send_ptr[k]:part.send_ptr[k + 1]]
        rn = part.recv_nodes[part.recv_ptr[k]:part.recv_ptr[k + 1]]
        sb = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, dm)[sn]))
        rb = torch.empty((len(rn), dm), dtype=torch.float64)
        reqs.append(dist.isend(sb, p))
        reqs.append(dist.irecv(rb, p))
        bufs.append((rn, rb))
    for r in reqs:
        r.wait()
    for rn, rb in bufs:
        x.reshape(-1, dm)[rn] = rb.numpy()
    assert np.array_equal(x.reshape(-1, dm), xg.reshape(-1, dm)[part.local_to_global])

    y_own = Kown @ x
    y = part.gather_global(np.concatenate([y_own, np.zeros((part.n_local - part.n_own) * dm)]), comm)
    dot_parts = comm.allgather_object(float(x[: part.n_own * dm] @ y_own))
    dot = sum(dot_parts)                      # folded in rank order on every rank

    # Neumann: owned entries of the local deck's rhs equal the global rhs
    loc = part.localize_deck(deck)
    lbody = Body(loc.nodes, loc.eSets["C3D4"], loc.ELE)
    nb = loc.neumann_bc_info[0]
    rhs_loc = neumann_vector(lbody, nb["face_set"], nb["traction"], nb["direction"])
    rhs = part.gather_global(rhs_loc, comm)
    fixed_loc = loc.dirichlet_bc_info[0]["node_set"]
    n_fixed_owned = int((fixed_loc < part.n_own).sum())
    tot_fixed = sum(comm.allgather_object(n_fixed_owned))
    n_primary = sum(comm.allgather_object(int(part.elem_primary.sum())))
    if rank == 0:
        np.savez(os.path.join(out_dir, "res.npz"), y=y, dot=dot, rhs=rhs, tot_fixed=tot_fixed, n_primary=n_primary)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("emulated_device_partitioner", [False, True], ids=["numpy_partitioner", "device_partitioner_emulated"])
def test_partition_halo_and_reductions_gloo(tmp_path, emulated_device_partitioner):
    import torch.multiprocessing as mp
    from femcy_b200 import meshgen
    from femcy_b200.body import Body
    from femcy_b200.neumann import neumann_vector
    from oracle import femcy_oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), emulated_device_partitioner), nprocs=2, join=True)
    res = np.load(tmp_path / "res.npz")
    deck = meshgen.SyntheticDeck("C3D4", n=5, jitter=0.1)
    nodes, conn = deck.nodes, deck.eSets["C3D4"]
    K = O.assemble_K(nodes, conn.astype(np.int64), np.zeros(nodes.size), "C3D4", np.asarray(deck.materials["Elastic"].C))
    xg = np.random.default_rng(7).standard_normal(nodes.size)
    yref = K @ xg
    assert np.abs(res["y"] - yref).max() < 1e-12 * np.abs(yref).max()
    assert abs(float(res["dot"]) - xg @ yref) < 1e-11 * abs(xg @ yref)
    body = Body(nodes, conn, deck.ELE)
    nb = deck.neumann_bc_info[0]
    rhs = neumann_vector(body, nb["face_set"], nb["traction"], nb["direction"])
    assert np.abs(res["rhs"] - rhs).max() < 1e-15
    assert int(res["tot_fixed"]) == len(deck.node_sets["fixed"])
    assert int(res["n_primary"]) == conn.shape[0]


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_invariants(nranks):
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition
    nodes, conn = meshgen.kuhn_box_c3d4(cells=(4, 3, 9))
    parts = [Partition(nodes, conn, r, nranks) for r in range(nranks)]
    assert sum(p.n_own for p in parts) == nodes.shape[0]
    owned = np.concatenate([p.local_to_global[: p.n_own] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(nodes.shape[0]))
    assert sum(int(p.elem_primary.sum()) for p in parts) == conn.shape[0]
    for a in parts:
        # every node of every local element is local
        assert a.elements.min() >= 0 and a.elements.max() < a.n_local
        for k, pr in enumerate(a.peers):
            b = parts[pr]
            kb = b.peers.index(a.rank)
            sent = a.local_to_global[a.send_nodes[a.send_ptr[k]:a.send_ptr[k + 1]]]
            recv = b.local_to_global[b.recv_nodes[b.recv_ptr[kb]:b.recv_ptr[kb + 1]]]
            assert np.array_equal(sent, recv)      # same nodes, same order on both sides


def test_localize_reader_style_deck():
    """A deck with reader-style face sets (sets of node tuples, as InpInfo produces) can be partitioned:
    the owned entries of the per-rank Neumann vectors add up to the global one."""
    from helpers import GoldenDeck, load_golden
    from femcy_b200.body import Body
    from femcy_b200.neumann import neumann_vector
    from femcy_b200.partition import Partition
    g = load_golden("c3d10_cook")
    deck = GoldenDeck(g)
    kind = "C3D10"
    gbody = Body(deck.nodes, deck.eSets[kind], deck.ELE)
    nb = deck.neumann_bc_info[0]
    ref = neumann_vector(gbody, nb["face_set"], nb["traction"], nb.get("direction", np.array([])))
    total = np.zeros_like(ref)
    for r in range(3):
        part = Partition(deck.nodes, deck.eSets[kind], r, 3)
        loc = part.localize_deck(deck)
        lb = Body(loc.nodes, loc.eSets[kind], loc.ELE)
        lnb = loc.neumann_bc_info[0]
        v = neumann_vector(lb, lnb["face_set"], lnb["traction"], lnb.get("direction", np.array([])))
        own = part.local_to_global[: part.n_own]
        total.reshape(-1, 3)[own] = v.reshape(-1, 3)[: part.n_own]
        # Dirichlet sets are mapped to local ids (owned + ghost)
        for bc, lbc in zip(deck.dirichlet_bc_info, loc.dirichlet_bc_info):
            assert set(part.local_to_global[lbc["node_set"]]) <= set(np.asarray(bc["node_set"]).tolist())
    assert np.abs(total - ref).max() < 1e-13 * np.abs(ref).max()


def test_weighted_node_partition():
    """node_owners(weights=): contiguous slabs with node counts proportional to the per-rank weights (GPU speed)."""
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition, node_owners
    nodes, conn = meshgen.kuhn_box_c3d4(8)
    nn = nodes.shape[0]
    w = [1.0, 0.8, 1.2, 1.0]
    owner = node_owners(nodes, 4, weights=w)
    cnt = np.bincount(owner, minlength=4)
    assert cnt.sum() == nn
    assert np.all(np.abs(cnt / nn - np.array(w) / sum(w)) < 0.02)
    # slabs stay contiguous along the partition axis, and equal weights reproduce the unweighted partition
    z = nodes[:, 2]
    for r in range(3):
        assert z[owner == r].max() <= z[owner == r + 1].min() + 1e-12
    assert np.array_equal(node_owners(nodes, 4, weights=[1, 1, 1, 1]), node_owners(nodes, 4))
    parts = [Partition(nodes, conn, r, 4, weights=w) for r in range(4)]
    assert sum(p.n_own for p in parts) == nn
    with pytest.raises(ValueError):
        node_owners(nodes, 4, weights=[1.0, 0.0, 1.0, 1.0])


# ---- device partitioner (csrc/partition.cu) --------------------------------------------------------------------------------
class EmuPartitionCtx:
    """answers femcy_partition / femcy_partition_get by running partition_build -- the product's orchestration and kernels --
    over the emulation backend (tests/simt); TEST INFRASTRUCTURE"""

    def call(self, name, *a):
        import simt
        from fake_ctx import _arr
        if name == "femcy_partition":
            dm, nn, nodes, ne, n_en, el, rank, nranks, ax, bounds, sizes = a
            nd = _arr(nodes, nn * dm).reshape(nn, dm)
            e = _arr(el, ne * n_en, np.int32).reshape(ne, n_en)
            r = self.r = simt.partition(nd, e, rank, nranks, ax, _arr(bounds, nranks + 1, np.int64))
            _arr(sizes, 6, np.int64)[:] = [r["n_own"], r["n_local"], len(r["elem_ids"]), len(r["peers"]), len(r["send_nodes"]), len(r["recv_nodes"])]
            return 0
        assert name == "femcy_partition_get"
        r = self.r
        order = ["owner", "elem_ids", "elem_primary", "local_to_global", "elements", "nodes", "peers", "send_ptr", "send_nodes", "recv_ptr", "recv_nodes"]
        types = [np.int32, np.int64, np.uint8, np.int64, np.int32, np.float64, np.int32, np.int64, np.int32, np.int64, np.int32]
        for ptr, key, dt in zip(a, order, types):
            arr = np.asarray(r[key]).astype(dt).reshape(-1)
            if arr.size:
                _arr(ptr, arr.size, dt)[:] = arr
        return 0


def assert_same_partition(a, b):
    for k in ("owner", "elem_ids", "elem_primary", "local_to_global", "global_to_local", "nodes", "elements", "send_nodes", "recv_nodes"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.peers == b.peers and list(a.send_ptr) == list(b.send_ptr) and list(a.recv_ptr) == list(b.recv_ptr)
    assert (a.n_own, a.n_local) == (b.n_own, b.n_local)


@pytest.mark.parametrize("kind,n", [("C3D4", 5), ("C3D10", 3)])
@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_emulated_device_partitioner_equals_the_numpy_statement(kind, n, nranks):
    """every array of the rank's piece -- owners, local elements, numbering, coordinates, halo plan -- for every rank, with
    equal and with weighted chunks"""
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition
    deck = meshgen.SyntheticDeck(kind, n=n, jitter=0.1) if kind == "C3D4" else meshgen.SyntheticDeck(kind, n=n)
    nodes, conn = deck.nodes, deck.eSets[kind]
    for rank in range(nranks):
        dev = Partition(nodes, conn, rank, nranks, ctx=EmuPartitionCtx())
        assert dev.built_on == "device"
        assert_same_partition(Partition(nodes, conn, rank, nranks), dev)
        w = np.linspace(1.0, 2.0, nranks)
        assert_same_partition(Partition(nodes, conn, rank, nranks, weights=w, axis=0), Partition(nodes, conn, rank, nranks, weights=w, axis=0, ctx=EmuPartitionCtx()))


def test_emulated_device_partitioner_on_a_plane_mesh_with_signed_coordinates():
    """2-D quadrilaterals, coordinates from -1 to 1 (the sortable image of negative doubles and of -0.0)"""
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition
    nodes, parts = meshgen.sectioned_plate(8, 4, (2., 1.))
    nodes = nodes - np.array([1.0, 0.5])
    nodes[np.abs(nodes) < 1e-15] = -0.0
    conn = parts[0][1]                      # the quadrilateral half: nodes of the other half are owned but unused
    for rank in range(3):
        assert_same_partition(Partition(nodes, conn, rank, 3), Partition(nodes, conn, rank, 3, ctx=EmuPartitionCtx()))


def test_emulated_device_partitioner_edge_cases():
    """more ranks than node planes (ranks that own almost nothing), an unstructured Delaunay mesh, coordinates with many ties along
    the slab axis and nodes no element uses"""
    from scipy.spatial import Delaunay
    from femcy_b200 import meshgen
    from femcy_b200.partition import Partition
    deck = meshgen.SyntheticDeck("C3D4", n=1)
    for nranks in (2, 8):
        for rank in range(nranks):
            assert_same_partition(Partition(deck.nodes, deck.eSets["C3D4"], rank, nranks),
                                  Partition(deck.nodes, deck.eSets["C3D4"], rank, nranks, ctx=EmuPartitionCtx()))
    rng = np.random.default_rng(0)
    pts = rng.random((300, 2))
    tri = Delaunay(pts).simplices.astype(np.int64)
    for rank in range(5):
        assert_same_partition(Partition(pts, tri, rank, 5), Partition(pts, tri, rank, 5, ctx=EmuPartitionCtx()))
    pts2 = np.concatenate([pts, rng.random((20, 2))])
    pts2[:, 0] = np.round(pts2[:, 0], 1)
    for rank in range(3):
        assert_same_partition(Partition(pts2, tri, rank, 3, axis=0), Partition(pts2, tri, rank, 3, axis=0, ctx=EmuPartitionCtx()))
