"""ctypes binding of libfemcy_b200.so (the C-ABI declared in include/femcy_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is created, the call raises.  Build with `python -m femcy_b200.build`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfemcy_b200.so")

_lib = None

c_ctx = C.c_void_p
P_d = C.POINTER(C.c_double)
P_i32 = C.POINTER(C.c_int32)
P_i64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/femcy_b200.h one to one
SIGNATURES = {
    "femcy_create": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "femcy_destroy": (None, [c_ctx]),
    "femcy_last_error": (C.c_char_p, [c_ctx]),
    "femcy_version": (C.c_char_p, []),
    "femcy_set_stream": (C.c_int, [c_ctx, C.c_void_p]),
    "femcy_sync": (C.c_int, [c_ctx]),
    "femcy_device_bytes": (C.c_int64, [c_ctx]),
    "femcy_set_mesh": (C.c_int, [c_ctx, C.c_int, C.c_int64, C.c_int64, P_d, C.c_int64, C.c_int, P_i32]),
    "femcy_set_element": (C.c_int, [c_ctx, C.c_int, P_d, P_d]),
    "femcy_set_material": (C.c_int, [c_ctx, C.c_int, P_d, C.c_int, P_d, C.c_int]),
    "femcy_add_section": (C.c_int, [c_ctx, C.c_int64, C.c_int, P_i32, C.POINTER(C.c_int)]),
    "femcy_select_section": (C.c_int, [c_ctx, C.c_int]),
    "femcy_section_count": (C.c_int, [c_ctx]),
    "femcy_build_pattern": (C.c_int, [c_ctx, P_i64]),
    "femcy_get_csr_pattern": (C.c_int, [c_ctx, P_i32, P_i32]),
    "femcy_get_K_csr_values": (C.c_int, [c_ctx, P_d]),
    "femcy_set_K_csr_values": (C.c_int, [c_ctx, P_d]),
    "femcy_pattern_stats": (C.c_int, [c_ctx, P_i64]),
    "femcy_vec_set": (C.c_int, [c_ctx, C.c_int, P_d, C.c_int64]),
    "femcy_vec_get": (C.c_int, [c_ctx, C.c_int, P_d, C.c_int64]),
    "femcy_vec_fill": (C.c_int, [c_ctx, C.c_int, C.c_double]),
    "femcy_vec_copy": (C.c_int, [c_ctx, C.c_int, C.c_int]),
    "femcy_vec_lincomb": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_double, C.c_int]),
    "femcy_vec_scale": (C.c_int, [c_ctx, C.c_int, C.c_double]),
    "femcy_vec_norms": (C.c_int, [c_ctx, C.c_int, P_d]),
    "femcy_vec_devptr": (C.c_void_p, [c_ctx, C.c_int]),
    "femcy_gp_get": (C.c_int, [c_ctx, C.c_int, P_d, C.c_int64]),
    "femcy_gp_set": (C.c_int, [c_ctx, C.c_int, P_d, C.c_int64]),
    "femcy_gp_sum": (C.c_int, [c_ctx, C.c_int, P_d]),
    "femcy_get_dsdx_and_vol": (C.c_int, [c_ctx]),
    "femcy_assemble_K": (C.c_int, [c_ctx, C.c_int]),
    "femcy_dirichlet_linear": (C.c_int, [c_ctx, P_i32, P_i32, P_d, C.c_int64]),
    "femcy_dirichlet_newton": (C.c_int, [c_ctx, P_i32, P_i32, C.c_int64]),
    "femcy_dirichlet_val": (C.c_int, [c_ctx, P_i32, P_i32, P_d, C.c_int64]),
    "femcy_set_facet_tables": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, P_i32, P_d, P_d, P_d, P_d]),
    "femcy_boundary_facets": (C.c_int, [c_ctx, P_i64]),
    "femcy_get_boundary_facets": (C.c_int, [c_ctx, P_i32, P_i32]),
    "femcy_node_elements": (C.c_int, [c_ctx, P_i32, P_i32]),
    "femcy_neumann": (C.c_int, [c_ctx, C.c_int64, P_i32, P_i32, C.c_double, P_d]),
    "femcy_deformation_gradient": (C.c_int, [c_ctx]),
    "femcy_constitutive": (C.c_int, [c_ctx, C.c_int]),
    "femcy_strain": (C.c_int, [c_ctx, C.c_int]),
    "femcy_mises": (C.c_int, [c_ctx]),
    "femcy_internal_force": (C.c_int, [c_ctx]),
    "femcy_elastic_energy": (C.c_int, [c_ctx, P_d]),
    "femcy_extrapolate": (C.c_int, [c_ctx, C.c_int, C.c_int, P_d, P_d, P_d]),
    "femcy_cg_solve": (C.c_int, [c_ctx, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_int, P_i64, P_d, P_d]),
    "femcy_spmv": (C.c_int, [c_ctx, C.c_int, C.c_int]),
    "femcy_cg_from_ell": (C.c_int, [c_ctx, C.c_int64, C.c_int, P_d, P_i32]),
    "femcy_partition": (C.c_int, [c_ctx, C.c_int, C.c_int64, P_d, C.c_int64, C.c_int, P_i32, C.c_int, C.c_int, C.c_int, P_i64, P_i64]),
    "femcy_partition_get": (C.c_int, [c_ctx, P_i32, P_i64, C.POINTER(C.c_ubyte), P_i64, P_i32, P_d, P_i32, P_i64, P_i32, P_i64, P_i32]),
    "femcy_comm_init": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_char_p]),
    "femcy_comm_unique_id": (C.c_int, [C.c_char_p, C.c_void_p]),
    "femcy_set_halo": (C.c_int, [c_ctx, C.c_int, P_i32, P_i64, P_i32, P_i64, P_i32]),
    "femcy_halo_exchange": (C.c_int, [c_ctx, C.c_int]),
    "femcy_p2p_export": (C.c_int, [c_ctx, C.c_void_p]),
    "femcy_p2p_import": (C.c_int, [c_ctx, C.c_void_p, P_i64]),
    "femcy_last_time_ms": (C.c_int, [c_ctx, C.c_int, P_d]),
    "femcy_cg_phase_ns": (C.c_int, [c_ctx, P_d]),
    "femcy_set_option": (C.c_int, [c_ctx, C.c_char_p, C.c_int]),
    "femcy_cg_breakdown": (C.c_int, [c_ctx]),
    "femcy_set_aggregates": (C.c_int, [c_ctx, C.c_int64, P_i32]),
    "femcy_launch_count": (C.c_int64, [c_ctx]),
}

VEC = {"dof": 0, "rhs": 1, "residual": 2, "nodal_force": 3, "du": 4, "dof_old": 5,
       "x": 6, "r": 7, "d": 8, "M": 9, "Ad": 10}
GP = {"vol": 0, "dsdx": 1, "F": 2, "cauchy": 3, "mises": 4, "strain": 5, "energy": 6}


OPTIONS = ("cg_kernel", "cg_sym", "cg_profile", "cg_stream_cfg", "no_graph", "no_p2p", "sell_sigma", "cg_precond", "consistent_tangent")


class FemcyError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises FemcyError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FemcyError(f"{LIB_PATH} not found: build it with `python -m femcy_b200.build` "
                         "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_d(a):
    return a.ctypes.data_as(P_d)


def as_i32(a):
    return a.ctypes.data_as(P_i32)


def as_i64(a):
    return a.ctypes.data_as(P_i64)


class Context:
    """Owner of one femcy_ctx (one GPU / rank)."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_ctx()
        rc = self.lib.femcy_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise FemcyError(f"femcy_create(device={device}) failed with code {rc}: no usable CUDA device "
                             "(the femcy_b200 hot path has no CPU fallback)")
        self.h = h
        self.device = device
        # A/B hook for the tools: FEMCY_OPT_<NAME>=<int> in the environment of the PROCESS is applied once, here;
        # the library itself never reads the environment
        for name in OPTIONS:
            v = os.environ.get("FEMCY_OPT_" + name.upper())
            if v is not None:
                self.set_option(name, int(v))

    def set_option(self, name, value):
        self.call("femcy_set_option", name.encode(), int(value))

    def close(self):
        if getattr(self, "h", None):
            self.lib.femcy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            msg = self.lib.femcy_last_error(self.h)
            raise FemcyError(f"{name} failed: {msg.decode() if msg else rc}")
        return rc

    # ---- convenience wrappers ------------------------------------------------------------
    def vec_get(self, which, n):
        out = np.empty(int(n), dtype=np.float64)
        self.call("femcy_vec_get", VEC[which] if isinstance(which, str) else which, as_d(out), out.size)
        return out

    def vec_set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        self.call("femcy_vec_set", VEC[which] if isinstance(which, str) else which, as_d(a), a.size)

    def gp_get(self, which, shape):
        out = np.empty(shape, dtype=np.float64)
        self.call("femcy_gp_get", GP[which], as_d(out), out.size)
        return out

    def gp_set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self.call("femcy_gp_set", GP[which], as_d(a), a.size)

    def norms(self, which):
        out = np.zeros(3)
        self.call("femcy_vec_norms", VEC[which] if isinstance(which, str) else which, as_d(out))
        return out

    def time_ms(self, kind):
        v = C.c_double(0.0)
        self.call("femcy_last_time_ms", int(kind), C.byref(v))
        return v.value

    def cg_breakdown(self):
        """True when the last femcy_cg_solve stopped on a NaN / inf residual"""
        return bool(self.lib.femcy_cg_breakdown(self.h))

    def cg_phase_ns(self):
        out = np.zeros(7)
        self.call("femcy_cg_phase_ns", as_d(out))
        return out

    def launches(self):
        return int(self.lib.femcy_launch_count(self.h))

    def sync(self):
        self.call("femcy_sync")
