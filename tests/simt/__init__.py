"""CPU SIMT emulation harness (TEST INFRASTRUCTURE ONLY -- see simt.h).

`lib()` compiles tests/simt/emu_entry.cpp (which includes the product's kernel headers under
FEMCY_SIMT_EMU) with g++ and returns the ctypes handle; `SellPattern` is a NumPy restatement of the
node-block SELL-32 layout documented in femcy_b200/csrc/kernel_types.cuh and built on the device by
femcy_b200/csrc/pattern.cu -- the kernels under test consume exactly these arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "femcy_b200", "csrc")
SO = os.path.join(HERE, "_build", "libfemcy_simt.so")


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(HERE, f) for f in ("simt.h", "emu_entry.cpp")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


_flavour = ""          # "" or "kt8": second build with FEMCY_TILE_KT=8 (multi-pass path of the tile kernel)
_libs = {}


def use_flavour(name):
    """switch the emulation library used by the helpers below ('' = product constants, 'kt8' = FEMCY_TILE_KT=8)."""
    global _flavour
    _flavour = name


def lib():
    if _flavour in _libs:
        return _libs[_flavour]
    so = SO if not _flavour else SO.replace(".so", f"_{_flavour}.so")
    stale = (not os.path.exists(so)) or _stale() or os.path.getmtime(so) < os.path.getmtime(SO if os.path.exists(SO) else __file__)
    if stale:
        os.makedirs(os.path.dirname(so), exist_ok=True)
        extra = ["-DFEMCY_TILE_KT=8"] if _flavour == "kt8" else []
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I", HERE] + extra + ["-o", so,
               os.path.join(HERE, "emu_entry.cpp"), "-lpthread"]
        subprocess.check_call(cmd)
    _libs[_flavour] = C.CDLL(so)
    return _libs[_flavour]


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class ElemTables(C.Structure):
    _fields_ = [("dN", C.c_double * (4 * 10 * 3)), ("w", C.c_double * 4), ("C", C.c_double * 36), ("mat", C.c_double * 4)]


def make_tables(ELE, material):
    dN, w = ELE.device_tables()
    t = ElemTables()
    flat = dN.reshape(-1)
    for i, v in enumerate(flat):
        t.dN[i] = v
    for i, v in enumerate(w):
        t.w[i] = v
    Cm = np.asarray(material.C, dtype=np.float64).reshape(-1)
    for i, v in enumerate(Cm):
        t.C[i] = v
    for i, v in enumerate(np.asarray(material.device_params(), dtype=np.float64)[:4]):
        t.mat[i] = v
    return t


class SellPattern:
    """node-block SELL-32 pattern of a mesh (rows = the first nn_own nodes, columns = all nodes)."""

    def __init__(self, conn, nn, nn_own=None, dm=3, sigma=0, rb_shift=5):
        conn = np.asarray(conn, dtype=np.int64)
        ne, n_en = conn.shape
        nn_own = nn if nn_own is None else nn_own
        self.dm, self.nn, self.nn_own, self.ne, self.n_en = dm, nn, nn_own, ne, n_en
        P = n_en * n_en
        rows = np.repeat(conn, n_en, axis=1).reshape(-1)          # entry t = e*P + a*n_en + b -> row conn[e,a]
        cols = np.tile(conn, (1, n_en)).reshape(-1)               #                              col conn[e,b]
        ids = np.arange(ne * P, dtype=np.int64)
        valid = rows < nn_own
        key = rows[valid] * nn + cols[valid]
        order = np.argsort(key, kind="stable")
        skey = key[order]
        sids = ids[valid][order]
        n_ent = skey.size
        head = np.ones(n_ent, dtype=bool)
        head[1:] = skey[1:] != skey[:-1]
        bfirst = np.flatnonzero(head)
        nnzb = bfirst.size
        brow = (skey[bfirst] // nn).astype(np.int64)
        bcol = (skey[bfirst] % nn).astype(np.int64)
        blkptr = np.searchsorted(brow, np.arange(nn_own + 1)).astype(np.int32)
        rowlen = np.diff(blkptr)
        nslice = (nn_own + 31) // 32
        # SELL-32-sigma (pattern.cu, option sell_sigma): rows ordered by descending block count inside windows of
        # sigma consecutive nodes (stable); rowof[pos] = row at position pos, rowpos = inverse; sigma = 0: identity
        self.sigma = sigma
        if sigma:
            key = ((np.arange(nn_own) // sigma).astype(np.int64) << 8) | (255 - np.minimum(rowlen, 255))
            rowof_real = np.argsort(key, kind="stable").astype(np.int32)
            self.rowof = np.full(max(nslice * 32, 1), -1, dtype=np.int32)
            self.rowof[:nn_own] = rowof_real
            self.rowpos = np.empty(max(nn_own, 1), dtype=np.int32)
            self.rowpos[rowof_real] = np.arange(nn_own, dtype=np.int32)
            rowpos = self.rowpos[:nn_own].astype(np.int64)
        else:
            self.rowof = self.rowpos = None
            rowof_real = np.arange(nn_own)
            rowpos = np.arange(nn_own, dtype=np.int64)
        padded = np.zeros(nslice * 32, dtype=np.int64)
        padded[:nn_own] = rowlen[rowof_real]
        w = padded.reshape(nslice, 32).max(axis=1)
        slice_ptr = np.zeros(nslice + 1, dtype=np.int32)
        slice_ptr[1:] = np.cumsum(w * 32)
        nslots = int(slice_ptr[-1])
        k = np.arange(nnzb) - blkptr[brow]
        bpos = rowpos[brow]
        bslot = slice_ptr[bpos // 32] + k * 32 + (bpos % 32)
        colidx = np.full(nslots, -1, dtype=np.int32)
        colidx[bslot] = bcol
        diag_slot = np.full(nn_own, -1, dtype=np.int32)
        d = brow == bcol
        diag_slot[brow[d]] = bslot[d]
        slot_beg = np.zeros(nslots, dtype=np.int32)
        slot_end = np.zeros(nslots, dtype=np.int32)
        slot_beg[bslot] = bfirst
        slot_end[bslot] = np.append(bfirst[1:], n_ent)
        blk_of = np.cumsum(head) - 1
        elem_slot = np.full(ne * P, -1, dtype=np.int32)
        elem_slot[sids] = bslot[blk_of]
        self.nnzb, self.nslice, self.nslots, self.n_ent = nnzb, nslice, nslots, n_ent
        self.max_row_blocks = int(w.max()) if nslice else 0
        self.blkptr, self.slice_ptr, self.colidx, self.diag_slot = blkptr, slice_ptr, colidx, diag_slot
        self.slot_beg, self.slot_end = slot_beg, slot_end
        self.ent_list = sids.astype(np.uint32)
        self.elem_slot = elem_slot
        self.brow, self.bcol, self.bslot = brow, bcol, bslot
        # node -> element incidence lists of the owned rows (pattern.cu: femcy_build_incidence): entries e*n_en + a
        # grouped by node, ascending element id
        flat = conn.reshape(-1)
        key = np.where(flat < nn_own, flat, nn_own)
        order = np.argsort(key, kind="stable")
        self.inc_list = order.astype(np.uint32)
        self.inc_ptr = np.searchsorted(key[order], np.arange(nn_own + 1)).astype(np.int32)
        # per-slice element tiles (pattern.cu: femcy_build_tiles): distinct elements touching the rows of each slice,
        # ascending, and for every contribution entry (ent_list order) its (tile index << 8 | a*n_en + b)
        own = flat < nn_own
        pos_of = rowpos if sigma else np.arange(nn_own, dtype=np.int64)
        self.rb_shift = rb_shift                       # rows per tile block = 2^rb_shift (5: slices, 3: 8-row blocks)
        nblk = (nslice * 32) >> rb_shift
        sl = pos_of[flat[own]] >> rb_shift
        el = (np.arange(ne * n_en) // n_en)[own]
        pairs = np.unique(sl * (1 << 32) + el)
        t_slice, t_elem = pairs >> 32, pairs & 0xffffffff
        self.tile_elems = t_elem.astype(np.uint32) if t_elem.size else np.zeros(1, dtype=np.uint32)
        self.tile_ptr = np.searchsorted(t_slice, np.arange(nblk + 1)).astype(np.int32)
        self.n_tile = int(t_elem.size)
        self.max_tile = int(np.diff(self.tile_ptr).max()) if nblk else 0
        ent_e, ent_p = sids // P, sids % P
        eslot = elem_slot[sids].astype(np.int64)
        ent_slice = np.searchsorted(slice_ptr, eslot, side="right") - 1
        ent_blk = (ent_slice * 32 + ((eslot - slice_ptr[ent_slice]) & 31)) >> rb_shift
        lidx = np.searchsorted(pairs, ent_blk * (1 << 32) + ent_e) - self.tile_ptr[ent_blk]
        self.ent_tile = ((lidx << 8) | ent_p).astype(np.uint32) if n_ent else np.zeros(1, dtype=np.uint32)

    def val_zeros(self):
        return np.zeros(self.nslots * self.dm * self.dm, dtype=np.float64)

    def to_csr(self, val):
        """scipy CSR (scalar) of the block values in the plane layout."""
        import scipy.sparse as sp
        dm, dm2 = self.dm, self.dm * self.dm
        r, c, v = [], [], []
        slot = self.bslot
        for i in range(dm):
            for j in range(dm):
                q = i * dm + j
                idx = (((slot >> 5) * dm2 + q) << 5) + (slot & 31)
                r.append(self.brow * dm + i)
                c.append(self.bcol * dm + j)
                v.append(val[idx])
        return sp.csr_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))),
                             shape=(self.nn_own * dm, self.nn * dm))

    def from_csr(self, K):
        """plane-layout values of a scipy matrix with this pattern."""
        dm, dm2 = self.dm, self.dm * self.dm
        K = K.tocsr()
        val = self.val_zeros()
        slot = self.bslot
        for i in range(dm):
            for j in range(dm):
                q = i * dm + j
                idx = (((slot >> 5) * dm2 + q) << 5) + (slot & 31)
                val[idx] = np.asarray(K[self.brow * dm + i, self.bcol * dm + j]).reshape(-1)
        return val


class EmuAsm(C.Structure):
    _fields_ = [("dm", C.c_int), ("n_en", C.c_int), ("n_gp", C.c_int), ("tab", C.POINTER(ElemTables)),
                ("nodes", C.POINTER(C.c_double)), ("dof", C.POINTER(C.c_double)), ("elems", C.POINTER(C.c_int32)),
                ("elem_slot", C.POINTER(C.c_int32)), ("ne", C.c_int64), ("slice_ptr", C.POINTER(C.c_int32)),
                ("nslice", C.c_int64), ("slot_beg", C.POINTER(C.c_int32)), ("slot_end", C.POINTER(C.c_int32)),
                ("ent_list", C.POINTER(C.c_uint32)), ("max_row_blocks", C.c_int), ("val", C.POINTER(C.c_double)),
                ("nslots", C.c_int64), ("vol", C.POINTER(C.c_double)), ("dsdx", C.POINTER(C.c_double)),
                ("egeo", C.POINTER(C.c_double)), ("variant", C.c_int), ("chunk_warps", C.c_int),
                ("inc_ptr", C.POINTER(C.c_int32)), ("inc_list", C.POINTER(C.c_uint32)), ("egeo4", C.POINTER(C.c_double)),
                ("nn_own", C.c_int64), ("rowof", C.POINTER(C.c_int32)), ("tile_ptr", C.POINTER(C.c_int32)),
                ("tile_elems", C.POINTER(C.c_uint32)), ("ent_tile", C.POINTER(C.c_uint32)), ("max_tile", C.c_int)]


def make_tables_raw(dN, w, Cm, params):
    """ElemTables from the arrays femcy_set_element / femcy_set_material receive."""
    t = ElemTables()
    for i, v in enumerate(np.asarray(dN, dtype=np.float64).reshape(-1)):
        t.dN[i] = v
    for i, v in enumerate(np.asarray(w, dtype=np.float64).reshape(-1)):
        t.w[i] = v
    for i, v in enumerate(np.asarray(Cm, dtype=np.float64).reshape(-1)):
        t.C[i] = v
    for i, v in enumerate(np.asarray(params, dtype=np.float64).reshape(-1)[:4]):
        t.mat[i] = v
    return t


def assemble(ELE, material, nodes, conn, dof, pat, variant=1, knob=0):
    """run the product's assembly kernels (variant as in femcy_assemble_K) on the emulator; returns (val, vol)."""
    dN, w = ELE.device_tables()
    val, vol, _ = assemble_raw(make_tables(ELE, material), dN.shape, nodes, conn, dof, pat, variant, knob)
    return val, vol


def dsdx_and_vol_raw(tab, shape, nodes, conn, dof):
    """k_dsdx_vol (femcy_get_dsdx_and_vol) on the emulator -> (dsdx [ne,g,a,d], vol [ne,g])."""
    n_gp, n_en, dm = shape
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    conn32 = np.ascontiguousarray(conn, dtype=np.int32)
    dof = np.ascontiguousarray(dof, dtype=np.float64)
    ne = conn32.shape[0]
    vol = np.zeros(ne * n_gp)
    dsdx = np.zeros(ne * n_gp * n_en * dm)
    a = EmuAsm(dm, n_en, n_gp, C.pointer(tab), _p(nodes, C.c_double), _p(dof, C.c_double), _p(conn32, C.c_int32))
    a.ne, a.vol, a.dsdx = ne, _p(vol, C.c_double), _p(dsdx, C.c_double)
    assert lib().emu_get_dsdx_and_vol(C.byref(a)) == 0
    return dsdx.reshape(ne, n_gp, n_en, dm), vol.reshape(ne, n_gp)


def assemble_raw(tab, shape, nodes, conn, dof, pat, variant=1, knob=0):
    """as `assemble`, from an ElemTables struct and (n_gp, n_en, dm); returns (val, vol, dsdx)."""
    L = lib()
    assert L.emu_sizeof_tables() == C.sizeof(ElemTables)
    n_gp, n_en, dm = shape
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    conn32 = np.ascontiguousarray(conn, dtype=np.int32)
    dof = np.ascontiguousarray(dof, dtype=np.float64)
    ne = conn32.shape[0]
    val = pat.val_zeros()
    val[:] = np.nan if variant in (2, 3) else 0.0      # the atomic-free variants write every slot (no zero-fill needed)
    vol = np.zeros(ne * n_gp)
    dsdx = np.zeros(ne * n_gp * n_en * dm)
    egeo = np.zeros(ne * (n_en * dm + 1))
    egeo4 = np.zeros(ne * n_en * n_gp * 4)
    keep = [tab, nodes, conn32, dof, val, vol, dsdx, egeo, egeo4]
    a = EmuAsm(dm, n_en, n_gp, C.pointer(tab), _p(nodes, C.c_double), _p(dof, C.c_double), _p(conn32, C.c_int32),
               _p(pat.elem_slot, C.c_int32), ne, _p(pat.slice_ptr, C.c_int32), pat.nslice, _p(pat.slot_beg, C.c_int32),
               _p(pat.slot_end, C.c_int32), _p(pat.ent_list, C.c_uint32), pat.max_row_blocks, _p(val, C.c_double),
               pat.nslots, _p(vol, C.c_double), _p(dsdx, C.c_double), _p(egeo, C.c_double), variant, knob,
               _p(pat.inc_ptr, C.c_int32), _p(pat.inc_list, C.c_uint32), _p(egeo4, C.c_double), pat.nn_own,
               _p(pat.rowof, C.c_int32), _p(pat.tile_ptr, C.c_int32), _p(pat.tile_elems, C.c_uint32), _p(pat.ent_tile, C.c_uint32),
               pat.max_tile)
    rc = L.emu_assemble_K(C.byref(a))
    assert rc == 0, rc
    del keep
    return val, vol.reshape(ne, n_gp), dsdx.reshape(ne, n_gp, n_en, dm)


# ---- PCG ------------------------------------------------------------------------------------------------
MAX_RANKS = 8
WINDOW_WORDS = 7 * MAX_RANKS


class EmuCG(C.Structure):
    _fields_ = [("dm", C.c_int), ("nn_own", C.c_int64), ("nn", C.c_int64), ("nslice", C.c_int64),
                ("slice_ptr", C.POINTER(C.c_int32)), ("colidx", C.POINTER(C.c_int32)), ("diag_slot", C.POINTER(C.c_int32)),
                ("val", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double)),
                ("x", C.POINTER(C.c_double)), ("r", C.POINTER(C.c_double)), ("d", C.POINTER(C.c_double)),
                ("M", C.POINTER(C.c_double)), ("Ad", C.POINTER(C.c_double)),
                ("scal", C.POINTER(C.c_double)), ("partials", C.POINTER(C.c_double)), ("ticket", C.POINTER(C.c_uint32)),
                ("eps", C.c_double), ("max_iter", C.c_int64), ("check_every", C.c_int), ("fixed", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int),
                ("d_of", C.POINTER(C.c_double) * MAX_RANKS), ("win_of", C.POINTER(C.c_uint64) * MAX_RANKS),
                ("bflag", C.POINTER(C.c_ubyte)), ("push_ptr", C.POINTER(C.c_int32)), ("push_peer", C.POINTER(C.c_int32)),
                ("push_ridx", C.POINTER(C.c_int32)), ("bnodes", C.POINTER(C.c_int32)), ("n_bnodes", C.c_int64),
                ("slice_order", C.POINTER(C.c_int32)), ("slice_ghost", C.POINTER(C.c_ubyte)),
                ("persistent_grid", C.c_int), ("iters_out", C.c_int64), ("r0_out", C.c_double), ("rmax_out", C.c_double),
                ("variant", C.c_int), ("rowof", C.POINTER(C.c_int32)), ("late_fence", C.c_int), ("fold_bar", C.c_int), ("sym", C.c_int)]


class RankSystem:
    """One rank's share of a linear system K x = b: rows of its owned nodes in the SELL-32 layout plus the
    peer-memory push plan -- a NumPy restatement of Partition.install/_install_p2p + femcy_p2p_import
    (femcy_b200/partition.py:132-190, femcy_b200/csrc/comm.cu:225-306)."""

    def __init__(self, part, K_global, b_global, dm, sigma=0):
        self.part, self.dm = part, dm
        n_own, n_loc = part.n_own, part.n_local
        self.pat = SellPattern(part.elements, n_loc, nn_own=n_own, dm=dm, sigma=sigma)
        gdofs = (part.local_to_global[:, None] * dm + np.arange(dm)[None, :]).reshape(-1)
        Kloc = K_global.tocsr()[gdofs[: n_own * dm]][:, gdofs]
        self.val = self.pat.from_csr(Kloc)
        self.b = np.zeros(n_loc * dm)
        self.b[: n_own * dm] = b_global[gdofs[: n_own * dm]]
        self.gdofs_own = gdofs[: n_own * dm]
        self.vecs = {k: np.zeros(n_loc * dm) for k in "xrdMA"}
        self.scal = np.zeros(64)
        self.partials = np.zeros(4096)
        self.ticket = np.zeros(8, dtype=np.uint32)
        self.window = np.zeros(WINDOW_WORDS, dtype=np.uint64)

    def plan(self, systems):
        """push plan towards the peers (after every rank's RankSystem exists)."""
        p = self.part
        n_own = p.n_own
        cnt = np.zeros(n_own + 1, dtype=np.int32)
        np.add.at(cnt, p.send_nodes + 1, 1)
        cnt = np.cumsum(cnt).astype(np.int32)
        fill = cnt[:-1].copy()
        n_send = len(p.send_nodes)
        ppeer = np.zeros(max(n_send, 1), dtype=np.int32)
        pridx = np.zeros(max(n_send, 1), dtype=np.int32)
        bf = np.zeros(max(n_own, 1), dtype=np.uint8)
        for k, peer in enumerate(p.peers):
            q = systems[peer].part
            kk = q.peers.index(p.rank)
            remote_start = int(q.recv_nodes[q.recv_ptr[kk]]) if q.recv_ptr[kk + 1] > q.recv_ptr[kk] else -1
            for t in range(p.send_ptr[k], p.send_ptr[k + 1]):
                nd = p.send_nodes[t]
                o = fill[nd]
                fill[nd] += 1
                ppeer[o] = peer
                pridx[o] = remote_start + (t - p.send_ptr[k])
                bf[nd] = 1
        self.push_ptr, self.push_peer, self.push_ridx, self.bflag = cnt, ppeer, pridx, bf
        self.bnodes = np.flatnonzero(bf[:n_own]).astype(np.int32)
        if self.bnodes.size == 0:
            self.bnodes = np.zeros(1, dtype=np.int32)
            self.n_bnodes = 0
        else:
            self.n_bnodes = int(self.bnodes.size)
        pat = self.pat
        gh = np.zeros(max(pat.nslice, 1), dtype=np.uint8)
        for s in range(pat.nslice):
            gh[s] = np.any(pat.colidx[pat.slice_ptr[s]: pat.slice_ptr[s + 1]] >= n_own)
        self.slice_ghost = gh
        self.slice_order = np.concatenate([np.flatnonzero(gh[: pat.nslice] == 0), np.flatnonzero(gh[: pat.nslice] == 1)]).astype(np.int32)


def cg_solve(systems, eps=1e-3, max_iter=1000, check_every=8, fixed=False, mode=0, persistent_grid=3, variant=0, late_fence=0, fold_bar=0, sym=0):
    """Run the product's PCG kernels on the emulator, one concurrent 'rank' per entry of `systems`.
    mode 0 = three kernels per iteration, 1 = persistent cooperative kernel.  Returns (iters, r0, rmax) of rank 0;
    the solution of every rank is in systems[r].vecs['x'][:n_own*dm]."""
    L = lib()
    n = len(systems)
    arr = (EmuCG * n)()
    for r, s in enumerate(systems):
        c = arr[r]
        pat = s.pat
        c.dm, c.nn_own, c.nn, c.nslice = s.dm, pat.nn_own, pat.nn, pat.nslice
        c.slice_ptr, c.colidx, c.diag_slot = _p(pat.slice_ptr, C.c_int32), _p(pat.colidx, C.c_int32), _p(pat.diag_slot, C.c_int32)
        c.val, c.b = _p(s.val, C.c_double), _p(s.b, C.c_double)
        c.x, c.r, c.d, c.M, c.Ad = (_p(s.vecs[k], C.c_double) for k in "xrdMA")
        c.scal, c.partials, c.ticket = _p(s.scal, C.c_double), _p(s.partials, C.c_double), _p(s.ticket, C.c_uint32)
        c.eps, c.max_iter, c.check_every, c.fixed = eps, max_iter, check_every, int(fixed)
        c.rank, c.nranks = r, n
        for q, t in enumerate(systems):
            c.d_of[q] = _p(t.vecs["d"], C.c_double)
            c.win_of[q] = _p(t.window, C.c_uint64)
        if n > 1:
            c.bflag, c.push_ptr, c.push_peer, c.push_ridx = _p(s.bflag, C.c_ubyte), _p(s.push_ptr, C.c_int32), _p(s.push_peer, C.c_int32), _p(s.push_ridx, C.c_int32)
            c.bnodes, c.n_bnodes = _p(s.bnodes, C.c_int32), s.n_bnodes
            c.slice_order, c.slice_ghost = _p(s.slice_order, C.c_int32), _p(s.slice_ghost, C.c_ubyte)
        c.persistent_grid = persistent_grid
        c.variant = variant
        c.rowof = _p(pat.rowof, C.c_int32)
        c.late_fence = late_fence
        c.fold_bar = fold_bar
        c.sym = sym
    rc = L.emu_cg_solve(arr, n, mode)
    assert rc == 0, f"emu_cg_solve rc={rc}"
    return int(arr[0].iters_out), float(arr[0].r0_out), float(arr[0].rmax_out)


def split_system(nodes, conn, K, b, nranks, dm, sigma=0):
    """RankSystems of a global system for `nranks` emulated ranks (element/row partition of femcy_b200.partition)."""
    from femcy_b200.partition import Partition
    parts = [Partition(nodes, conn, r, nranks) for r in range(nranks)]
    systems = [RankSystem(p, K, b, dm, sigma) for p in parts]
    if nranks > 1:
        for s in systems:
            s.plan(systems)
    return systems


def gather_solution(systems, N):
    x = np.zeros(N)
    for s in systems:
        x[s.gdofs_own] = s.vecs["x"][: s.gdofs_own.size]
    return x


def dirichlet(pat, val, target, nodes, comps, vals, mode):
    """bc.cu's bc_apply on the emulator (in place on val / target).  mode 0 = linear equations, 1 = Newton."""
    L = lib()
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    comps = np.ascontiguousarray(comps, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    flag = np.zeros(pat.nn * pat.dm, dtype=np.uint8)
    valfull = np.zeros(pat.nn * pat.dm)
    rc = L.emu_dirichlet(pat.dm, C.c_int64(pat.nn_own), C.c_int64(pat.nslice), _p(pat.slice_ptr, C.c_int32),
                         _p(pat.colidx, C.c_int32), _p(val, C.c_double), _p(pat.rowof, C.c_int32), _p(nodes, C.c_int32),
                         _p(comps, C.c_int32), _p(vals, C.c_double), C.c_int64(nodes.size), _p(flag, C.c_ubyte),
                         _p(valfull, C.c_double), _p(target, C.c_double), mode)
    assert rc == 0
    assert not flag.any()


# ---- stress recovery / internal force / energy (post.cu) --------------------------------------------------------
class EmuPost(C.Structure):
    _fields_ = [("dm", C.c_int), ("n_en", C.c_int), ("n_gp", C.c_int), ("kind", C.c_int), ("tab", C.POINTER(ElemTables)),
                ("nodes", C.POINTER(C.c_double)), ("dof", C.POINTER(C.c_double)), ("elems", C.POINTER(C.c_int32)),
                ("ne", C.c_int64), ("nn_own", C.c_int64), ("F", C.POINTER(C.c_double)), ("cauchy", C.POINTER(C.c_double)),
                ("vol", C.POINTER(C.c_double)), ("dsdx", C.POINTER(C.c_double)), ("force", C.POINTER(C.c_double)),
                ("out", C.POINTER(C.c_double)), ("partials", C.POINTER(C.c_double)), ("ticket", C.POINTER(C.c_uint32)),
                ("total", C.POINTER(C.c_double))]


class Post:
    """post.cu's kernels on the emulator for one mesh + material (fields as in femcy_ctx)."""

    def __init__(self, ELE, material, nodes, conn, dof, raw=None):
        self.L = lib()
        if raw is None:
            self.tab = make_tables(ELE, material)
            dN, w = ELE.device_tables()
            self.n_gp, self.n_en, self.dm = dN.shape
            kind = int(material.kind)
        else:                                   # (ElemTables, (n_gp, n_en, dm), material kind id)
            self.tab, (self.n_gp, self.n_en, self.dm), kind = raw
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.dof = np.ascontiguousarray(dof, dtype=np.float64)
        ne, g, d = self.conn.shape[0], self.n_gp, self.dm
        self.ne = ne
        self.F = np.zeros((ne, g, d, d)); self.cauchy = np.zeros((ne, g, d, d)); self.vol = np.zeros((ne, g))
        self.dsdx = np.zeros((ne, g, self.n_en, d)); self.force = np.zeros(self.nodes.size)
        self.partials = np.zeros(1024); self.ticket = np.zeros(8, dtype=np.uint32); self.total = np.zeros(1)
        self.kind = kind

    def _args(self, out=None):
        return EmuPost(self.dm, self.n_en, self.n_gp, self.kind, C.pointer(self.tab), _p(self.nodes, C.c_double),
                       _p(self.dof, C.c_double), _p(self.conn, C.c_int32), self.ne, self.nodes.shape[0], _p(self.F, C.c_double),
                       _p(self.cauchy, C.c_double), _p(self.vol, C.c_double), _p(self.dsdx, C.c_double),
                       _p(self.force, C.c_double), _p(out, C.c_double), _p(self.partials, C.c_double),
                       _p(self.ticket, C.c_uint32), _p(self.total, C.c_double))

    def deformation_gradient(self):
        a = self._args()
        assert self.L.emu_deformation_gradient(C.byref(a)) == 0
        return self.F

    def constitutive(self, large):
        a = self._args()
        assert self.L.emu_per_gp(C.byref(a), 0, int(large)) == 0
        return self.cauchy

    def strain(self, large):
        out = np.zeros_like(self.F)
        a = self._args(out)
        assert self.L.emu_per_gp(C.byref(a), 1, int(large)) == 0
        return out

    def mises(self):
        out = np.zeros((self.ne, self.n_gp))
        a = self._args(out)
        assert self.L.emu_per_gp(C.byref(a), 2, 0) == 0
        return out

    def energy(self):
        out = np.zeros((self.ne, self.n_gp))
        a = self._args(out)
        assert self.L.emu_per_gp(C.byref(a), 3, 1) == 0
        return out, float(self.total[0])

    def internal_force(self):
        self.force[:] = 0.0
        a = self._args()
        assert self.L.emu_internal_force(C.byref(a)) == 0
        return self.force


# ---- pattern build (pattern.cu kernels; CUB sorts replaced by std::stable_sort) ------------------------------------
class EmuPattern(C.Structure):
    _fields_ = [("elems", C.POINTER(C.c_int32)), ("ne", C.c_int64), ("n_en", C.c_int), ("nn", C.c_int64), ("nn_own", C.c_int64),
                ("sigma", C.c_int), ("cap_slots", C.c_int64), ("stats", C.c_int64 * 4),
                ("blkptr", C.POINTER(C.c_int32)), ("slice_ptr", C.POINTER(C.c_int32)), ("colidx", C.POINTER(C.c_int32)),
                ("diag_slot", C.POINTER(C.c_int32)), ("slot_beg", C.POINTER(C.c_int32)), ("slot_end", C.POINTER(C.c_int32)),
                ("elem_slot", C.POINTER(C.c_int32)), ("ent_list", C.POINTER(C.c_uint32)), ("n_ent", C.c_int64),
                ("rowof", C.POINTER(C.c_int32)), ("rowpos", C.POINTER(C.c_int32)), ("inc_ptr", C.POINTER(C.c_int32)),
                ("inc_list", C.POINTER(C.c_uint32)), ("tile_ptr", C.POINTER(C.c_int32)), ("tile_elems", C.POINTER(C.c_uint32)),
                ("ent_tile", C.POINTER(C.c_uint32)), ("n_tile", C.c_int64), ("max_tile", C.c_int), ("rb_shift", C.c_int),
                ("nsec", C.c_int), ("elems_s", C.POINTER(C.c_int32) * 8), ("ne_s", C.c_int64 * 8), ("n_en_s", C.c_int * 8)]


def build_pattern(conn, nn, nn_own=None, sigma=0, rb_shift=5):
    """the product's pattern-build kernels on the emulator; returns a dict of the arrays femcy_build_pattern /
    femcy_build_incidence leave on the device."""
    conn32 = np.ascontiguousarray(conn, dtype=np.int32)
    ne, n_en = conn32.shape
    nn_own = nn if nn_own is None else nn_own
    total = ne * n_en * n_en
    nslice = (nn_own + 31) // 32
    cap = total + 64 * max(nslice, 1) * 32
    o = {"blkptr": np.zeros(nn_own + 1, np.int32), "slice_ptr": np.zeros(nslice + 1, np.int32), "colidx": np.zeros(cap, np.int32),
         "diag_slot": np.zeros(max(nn_own, 1), np.int32), "slot_beg": np.zeros(cap, np.int32), "slot_end": np.zeros(cap, np.int32),
         "elem_slot": np.zeros(max(total, 1), np.int32), "ent_list": np.zeros(max(total, 1), np.uint32),
         "rowof": np.zeros(max(nslice * 32, 1), np.int32), "rowpos": np.zeros(max(nn_own, 1), np.int32),
         "inc_ptr": np.zeros(nn_own + 1, np.int32), "inc_list": np.zeros(max(ne * n_en, 1), np.uint32),
         "tile_ptr": np.zeros(((nslice * 32) >> rb_shift) + 1, np.int32), "tile_elems": np.zeros(max(ne * n_en, 1), np.uint32),
         "ent_tile": np.zeros(max(total, 1), np.uint32)}
    p = EmuPattern(_p(conn32, C.c_int32), ne, n_en, nn, nn_own, sigma, cap)
    for k, a in o.items():
        setattr(p, k, _p(a, C.c_uint32 if a.dtype == np.uint32 else C.c_int32))
    p.rb_shift = rb_shift
    rc = lib().emu_build_pattern(C.byref(p))
    assert rc == 0, rc
    o["nnzb"], o["nslots"], o["nslice"], o["max_row_blocks"] = (int(v) for v in p.stats)
    o["n_ent"] = int(p.n_ent)
    for k in ("colidx", "slot_beg", "slot_end"):
        o[k] = o[k][: o["nslots"]]
    o["ent_list"] = o["ent_list"][: o["n_ent"]]
    o["n_tile"], o["max_tile"] = int(p.n_tile), int(p.max_tile)
    o["tile_elems"] = o["tile_elems"][: o["n_tile"]]
    o["ent_tile"] = o["ent_tile"][: o["n_ent"]]
    return o


# ---- row f4: a mesh of several sections (pattern.cu: build_pattern_sections; assembly.cu: one scatter pass per section) ----
class SectionPattern:
    """what assemble_raw needs from a pattern, for ONE section of a multi-section mesh: the union layout + the slots of
    this section's element-local blocks"""

    def __init__(self, o, elem_slot, dm, nn_own):
        self.dm, self.nn_own = dm, nn_own
        self.elem_slot = np.ascontiguousarray(elem_slot, dtype=np.int32)
        self.slice_ptr, self.nslice, self.nslots, self.max_row_blocks = o["slice_ptr"], o["nslice"], o["nslots"], o["max_row_blocks"]
        self.slot_beg = self.slot_end = np.zeros(1, np.int32)
        self.ent_list = np.zeros(1, np.uint32)
        self.inc_ptr = np.zeros(1, np.int32)
        self.inc_list = self.tile_elems = self.ent_tile = np.zeros(1, np.uint32)
        self.tile_ptr = np.zeros(1, np.int32)
        self.rowof, self.max_tile = None, 0

    def val_zeros(self):
        return np.zeros(max(self.nslots, 1) * self.dm * self.dm)


def build_pattern_sections(conns, nn, nn_own=None):
    """the pattern-build kernels over the concatenated keys of several sections (natural row order); returns the layout
    dict of `build_pattern` with `elem_slot` split per section."""
    conns = [np.ascontiguousarray(c, dtype=np.int32) for c in conns]
    nn_own = nn if nn_own is None else nn_own
    counts = [c.shape[0] * c.shape[1] * c.shape[1] for c in conns]
    total = sum(counts)
    nslice = (nn_own + 31) // 32
    cap = total + 64 * max(nslice, 1) * 32
    o = {"blkptr": np.zeros(nn_own + 1, np.int32), "slice_ptr": np.zeros(nslice + 1, np.int32), "colidx": np.zeros(cap, np.int32),
         "diag_slot": np.zeros(max(nn_own, 1), np.int32), "slot_beg": np.zeros(cap, np.int32), "slot_end": np.zeros(cap, np.int32),
         "elem_slot": np.zeros(max(total, 1), np.int32), "ent_list": np.zeros(max(total, 1), np.uint32),
         "rowof": np.zeros(max(nslice * 32, 1), np.int32), "rowpos": np.zeros(max(nn_own, 1), np.int32)}
    p = EmuPattern(None, 0, 0, nn, nn_own, 0, cap)
    for k, a in o.items():
        setattr(p, k, _p(a, C.c_uint32 if a.dtype == np.uint32 else C.c_int32))
    p.rb_shift = 5
    p.nsec = len(conns)
    for s, c in enumerate(conns):
        p.elems_s[s] = _p(c, C.c_int32)
        p.ne_s[s] = c.shape[0]
        p.n_en_s[s] = c.shape[1]
    rc = lib().emu_build_pattern(C.byref(p))
    assert rc == 0, rc
    o["nnzb"], o["nslots"], o["nslice"], o["max_row_blocks"] = (int(v) for v in p.stats)
    o["colidx"] = o["colidx"][: o["nslots"]]
    offs = np.concatenate([[0], np.cumsum(counts)])
    o["elem_slot_sections"] = [o["elem_slot"][offs[s]:offs[s + 1]].copy() for s in range(len(conns))]
    return o


def sell_to_csr(o, val, nn_own, nn, dm):
    """scipy CSR of a node-block SELL-32 matrix given by the layout arrays (natural row order)"""
    import scipy.sparse as sp
    slots = np.flatnonzero(o["colidx"] >= 0)
    sl = np.searchsorted(o["slice_ptr"], slots, side="right") - 1
    brow = sl * 32 + (slots & 31)
    bcol = o["colidx"][slots].astype(np.int64)
    dm2 = dm * dm
    rows, cols, vals = [], [], []
    for i in range(dm):
        for j in range(dm):
            rows.append(brow * dm + i)
            cols.append(bcol * dm + j)
            vals.append(val[((slots >> 5) * dm2 + (i * dm + j)) * 32 + (slots & 31)])
    K = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nn_own * dm, nn * dm))
    K.sort_indices()
    return K


# ---- row f1: topology builders + Neumann vector (topology.cu kernels) -------------------------------------------------------
class EmuTopo(C.Structure):
    _fields_ = [("elems", C.POINTER(C.c_int32)), ("ne", C.c_int64), ("n_en", C.c_int), ("nn", C.c_int64), ("dm", C.c_int),
                ("nodes", C.POINTER(C.c_double)), ("nkeys", C.c_int), ("width", C.c_int), ("nfp", C.c_int),
                ("key_nodes", C.POINTER(C.c_int32)), ("w", C.POINTER(C.c_double)), ("normal", C.POINTER(C.c_double)),
                ("N", C.POINTER(C.c_double)), ("dN", C.POINTER(C.c_double)),
                ("b_elem", C.POINTER(C.c_int32)), ("b_kid", C.POINTER(C.c_int32)), ("n_boundary", C.c_int64),
                ("ne_ptr", C.POINTER(C.c_int32)), ("ne_list", C.POINTER(C.c_int32)),
                ("nf", C.c_int64), ("f_elem", C.POINTER(C.c_int32)), ("f_kid", C.POINTER(C.c_int32)), ("traction", C.c_double),
                ("has_dir", C.c_int), ("dir", C.c_double * 3), ("rhs", C.POINTER(C.c_double))]


class Topology:
    """the topology.cu entry points on the emulator for one (nodes, connectivity, element kind)"""

    def __init__(self, ELE, nodes, conn):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.tabs = ELE.device_facet_tables()
        kn, w, nrm, N, dN = self.tabs
        self.nkeys, self.width = kn.shape
        self.nfp = w.shape[1]
        self.nn, self.dm = self.nodes.shape
        self.ne, self.n_en = self.conn.shape

    def _args(self):
        kn, w, nrm, N, dN = self.tabs
        return EmuTopo(_p(self.conn, C.c_int32), self.ne, self.n_en, self.nn, self.dm, _p(self.nodes, C.c_double), self.nkeys,
                       self.width, self.nfp, _p(kn, C.c_int32), _p(w, C.c_double), _p(nrm, C.c_double), _p(N, C.c_double), _p(dN, C.c_double))

    def boundary_facets(self):
        a = self._args()
        be, bk = np.zeros(max(self.ne * self.nkeys, 1), np.int32), np.zeros(max(self.ne * self.nkeys, 1), np.int32)
        a.b_elem, a.b_kid = _p(be, C.c_int32), _p(bk, C.c_int32)
        assert lib().emu_boundary_facets(C.byref(a)) == 0
        n = int(a.n_boundary)
        return be[:n], bk[:n]

    def node_elements(self):
        a = self._args()
        ptr, lst = np.zeros(self.nn + 1, np.int32), np.zeros(max(self.ne * self.n_en, 1), np.int32)
        a.ne_ptr, a.ne_list = _p(ptr, C.c_int32), _p(lst, C.c_int32)
        assert lib().emu_node_elements(C.byref(a)) == 0
        return ptr, lst[: self.ne * self.n_en]

    def neumann(self, ele, kid, traction, direction=None):
        a = self._args()
        fe, fk = np.ascontiguousarray(ele, dtype=np.int32), np.ascontiguousarray(kid, dtype=np.int32)
        rhs = np.full(self.nn * self.dm, np.nan)
        a.nf, a.f_elem, a.f_kid, a.traction, a.rhs = fe.size, _p(fe, C.c_int32), _p(fk, C.c_int32), float(traction), _p(rhs, C.c_double)
        if direction is not None and len(direction):
            a.has_dir = 1
            for i in range(self.dm):
                a.dir[i] = float(direction[i])
        assert lib().emu_neumann(C.byref(a)) == 0
        return rhs


# ---- device partitioner (partition.cu: partition_build over the emulation backend) -------------------------------------------
class EmuPartition(C.Structure):
    _fields_ = [("dm", C.c_int), ("nn", C.c_int64), ("nodes", C.POINTER(C.c_double)), ("ne", C.c_int64), ("n_en", C.c_int),
                ("elems", C.POINTER(C.c_int32)), ("rank", C.c_int), ("nranks", C.c_int), ("axis", C.c_int), ("bounds", C.POINTER(C.c_int64)),
                ("sizes", C.c_int64 * 6), ("owner", C.POINTER(C.c_int32)), ("elem_ids", C.POINTER(C.c_int64)), ("primary", C.POINTER(C.c_ubyte)),
                ("l2g", C.POINTER(C.c_int64)), ("loc_elems", C.POINTER(C.c_int32)), ("loc_nodes", C.POINTER(C.c_double)),
                ("peers", C.POINTER(C.c_int32)), ("send_ptr", C.POINTER(C.c_int64)), ("send_nodes", C.POINTER(C.c_int32)),
                ("recv_ptr", C.POINTER(C.c_int64)), ("recv_nodes", C.POINTER(C.c_int32))]


def partition(nodes, elements, rank, nranks, axis, bounds):
    """femcy_partition + femcy_partition_get on the emulator: dict of the arrays femcy_b200.partition.Partition keeps"""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    conn = np.ascontiguousarray(elements, dtype=np.int32)
    bounds = np.ascontiguousarray(bounds, dtype=np.int64)
    nn, dm = nodes.shape
    ne, n_en = conn.shape
    o = {"owner": np.zeros(max(nn, 1), np.int32), "elem_ids": np.zeros(max(ne, 1), np.int64), "primary": np.zeros(max(ne, 1), np.uint8),
         "l2g": np.zeros(max(nn, 1), np.int64), "loc_elems": np.zeros(max(ne * n_en, 1), np.int32), "loc_nodes": np.zeros(max(nn * dm, 1)),
         "peers": np.zeros(8, np.int32), "send_ptr": np.zeros(9, np.int64), "send_nodes": np.zeros(max(nn, 1), np.int32),
         "recv_ptr": np.zeros(9, np.int64), "recv_nodes": np.zeros(max(nn, 1), np.int32)}
    ct = {np.dtype(np.int32): C.c_int32, np.dtype(np.int64): C.c_int64, np.dtype(np.uint8): C.c_ubyte, np.dtype(np.float64): C.c_double}
    p = EmuPartition(dm, nn, _p(nodes, C.c_double), ne, n_en, _p(conn, C.c_int32), rank, nranks, axis, _p(bounds, C.c_int64))
    for k, a in o.items():
        setattr(p, k, _p(a, ct[a.dtype]))
    rc = lib().emu_partition(C.byref(p))
    assert rc == 0, rc
    n_own, n_local, ne_local, npeers, n_send, n_recv = (int(v) for v in p.sizes)
    return {"n_own": n_own, "n_local": n_local, "owner": o["owner"][:nn], "elem_ids": o["elem_ids"][:ne_local],
            "elem_primary": o["primary"][:ne_local].astype(bool), "local_to_global": o["l2g"][:n_local],
            "elements": o["loc_elems"][: ne_local * n_en].reshape(ne_local, n_en), "nodes": o["loc_nodes"][: n_local * dm].reshape(n_local, dm),
            "peers": o["peers"][:npeers].tolist(), "send_ptr": o["send_ptr"][: npeers + 1].tolist(), "recv_ptr": o["recv_ptr"][: npeers + 1].tolist(),
            "send_nodes": o["send_nodes"][:n_send], "recv_nodes": o["recv_nodes"][:n_recv]}
