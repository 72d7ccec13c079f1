"""ctypes wrapper of oracle/libfemcy_oracle.so -- the C/OpenMP restatement of the reference's
arch=cpu hot path (TEST INFRASTRUCTURE / CPU baseline only; see femcy_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libfemcy_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(LIB)
        _lib.oracle_pcg_ell.restype = C.c_int64
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads():
    return int(load().oracle_num_threads())


def set_num_threads(n):
    load().oracle_set_num_threads(C.c_int(int(n)))


def ell_pattern(elements, nn, dm):
    """sparseIJ [N, W+1] as the reference builds it (count, columns, -1 padding;
    stiffnessMtrx.py:78-89) -- columns sorted here, the reference's follow Python-set order."""
    n_en = elements.shape[1]
    i = np.repeat(elements, n_en, axis=1).reshape(-1).astype(np.int64)
    j = np.tile(elements, (1, n_en)).reshape(-1).astype(np.int64)
    key = np.unique(i * nn + j)
    bi, bj = key // nn, key % nn
    cnt = np.bincount(bi, minlength=nn)
    W = int(cnt.max()) * dm
    ptr = np.zeros(nn + 1, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    pos = np.arange(len(bi)) - ptr[bi]
    ij = -np.ones((nn * dm, W + 1), dtype=np.int32)
    for r in range(dm):
        ij[bi * dm + r, 0] = cnt[bi] * dm
        for c in range(dm):
            ij[bi * dm + r, 1 + pos * dm + c] = bj * dm + c
    return ij


def dsdx_vol(nodes, elements, u, dN, w):
    lib = load()
    ne, n_en = elements.shape
    dm = nodes.shape[1]
    n_gp = len(w)
    dsdx = np.empty((ne, n_gp, n_en, dm))
    vol = np.empty((ne, n_gp))
    lib.oracle_dsdx_vol(C.c_int(dm), C.c_int(n_en), C.c_int(n_gp), C.c_int64(ne), _p(nodes, C.c_double), _p(u, C.c_double),
                        _p(elements, C.c_int32), _p(dN, C.c_double), _p(w, C.c_double), _p(dsdx, C.c_double), _p(vol, C.c_double))
    return dsdx, vol


def assemble_ell(elements, dm, dsdx, vol, Cmat, ij, spm=None):
    lib = load()
    ne, n_en = elements.shape
    n_gp = vol.shape[1]
    N, W1 = ij.shape
    if spm is None:
        spm = np.empty((N, W1 - 1))
    lib.oracle_assemble_ell(C.c_int(dm), C.c_int(n_en), C.c_int(n_gp), C.c_int64(ne), C.c_int64(N), C.c_int(W1 - 1),
                            _p(elements, C.c_int32), _p(dsdx, C.c_double), _p(vol, C.c_double),
                            _p(np.ascontiguousarray(Cmat, dtype=np.float64), C.c_double), _p(ij, C.c_int32), _p(spm, C.c_double))
    return spm


def pcg_ell(spm, ij, b, eps=1e-3, max_iter=None, fixed_iters=False):
    lib = load()
    N, W = spm.shape
    x = np.empty(N)
    r0, r1 = C.c_double(0.), C.c_double(0.)
    it = lib.oracle_pcg_ell(C.c_int64(N), C.c_int(W), _p(spm, C.c_double), _p(ij, C.c_int32), _p(b, C.c_double),
                            _p(x, C.c_double), C.c_double(eps), C.c_int64(N if max_iter is None else max_iter),
                            C.c_int(1 if fixed_iters else 0), C.byref(r0), C.byref(r1))
    return x, int(it), r0.value, r1.value


def dirichlet_ell(spm, ij, dofs, rhs, vals=None):
    """dirichletBC_linearEquations on the ELL arrays (stiffnessMtrx.py:279-307), vectorised:
    rhs[j] -= val_i*K[j,i]; rhs[i] = val_i; zero row i and column i; K[i,i] = 1."""
    dofs = np.asarray(dofs, dtype=np.int64)
    vals = np.zeros(len(dofs)) if vals is None else np.asarray(vals, dtype=np.float64)
    N, W = spm.shape
    flag = np.zeros(N, dtype=bool)
    flag[dofs] = True
    vfull = np.zeros(N)
    vfull[dofs] = vals
    cols = ij[:, 1:]
    valid = cols >= 0
    colflag = np.zeros_like(valid)
    colflag[valid] = flag[cols[valid]]
    contrib = np.where(colflag, spm * vfull[np.where(valid, cols, 0)], 0.0).sum(axis=1)
    rhs = rhs - np.where(flag, 0.0, contrib)
    rhs[dofs] = vals
    spm[colflag] = 0.0
    spm[flag, :] = 0.0
    diag = (cols == np.arange(N)[:, None]) & flag[:, None]
    spm[diag] = 1.0
    return spm, rhs
