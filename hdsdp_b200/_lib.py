"""ctypes binding of libhdsdp_cuda.so (C ABI declared in include/hdsdpcu.h).

The product path has no CPU fallback: loading fails loudly when the shared library is missing, and
every compute entry point fails when no CUDA device is usable.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_long, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhdsdp_cuda.so")

c_double_p = POINTER(c_double)
c_int_p = POINTER(c_int)


class HdsdpCudaError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); mirrors include/hdsdpcu.h one for one
_SIGNATURES = {
    "hdsdpcu_init": (c_int, [c_int]),
    "hdsdpcu_device_count": (c_int, []),
    "hdsdpcu_stream": (c_void_p, []),
    "hdsdpcu_sync": (c_int, []),
    "hdsdpcu_version": (c_char_p, []),
    "hdsdpcu_launch_count": (c_long, [c_int]),
    "hdsdpcu_copy_dev": (c_int, [c_void_p, c_void_p, c_long]),
    "hdsdpcu_set_option": (c_int, [c_char_p, c_int]),
    # B1
    "hdsdpcu_linsys_create": (c_int, [POINTER(c_void_p), c_int]),
    "hdsdpcu_linsys_setparam": (None, [c_void_p, c_void_p]),
    "hdsdpcu_linsys_symbolic": (c_int, [c_void_p, c_int_p, c_int_p]),
    "hdsdpcu_linsys_numeric": (c_int, [c_void_p, c_int_p, c_int_p, c_double_p]),
    "hdsdpcu_linsys_psdcheck": (c_int, [c_void_p, c_int_p, c_int_p, c_double_p, c_int_p]),
    "hdsdpcu_linsys_fsolve": (None, [c_void_p, c_int, c_double_p, c_double_p]),
    "hdsdpcu_linsys_bsolve": (None, [c_void_p, c_int, c_double_p, c_double_p]),
    "hdsdpcu_linsys_solve": (c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    "hdsdpcu_linsys_getdiag": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_linsys_invert": (None, [c_void_p, c_double_p, c_double_p]),
    "hdsdpcu_linsys_destroy": (None, [POINTER(c_void_p)]),
    "hdsdpcu_linsys_padded_dim": (c_int, [c_void_p]),
    "hdsdpcu_linsys_numeric_dev": (c_int, [c_void_p, c_void_p, c_long, c_int_p]),
    "hdsdpcu_linsys_solve_dev": (c_int, [c_void_p, c_int, c_void_p, c_long]),
    "hdsdpcu_linsys_invert_dev": (c_int, [c_void_p, c_void_p]),
    "hdsdpcu_linsys_factor_dev": (c_void_p, [c_void_p]),
    # B2 cone
    "hdsdpcu_cone_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int_p, c_int_p, c_double_p]),
    "hdsdpcu_cone_destroy": (None, [POINTER(c_void_p)]),
    "hdsdpcu_cone_getdim": (c_int, [c_void_p]),
    "hdsdpcu_cone_gettypes": (c_int, [c_void_p, c_int_p]),
    "hdsdpcu_cone_setstart": (None, [c_void_p, c_double]),
    "hdsdpcu_cone_reduceresi": (None, [c_void_p, c_double]),
    "hdsdpcu_cone_setperturb": (None, [c_void_p, c_double]),
    "hdsdpcu_cone_scal": (c_int, [c_void_p, c_double]),
    "hdsdpcu_cone_update": (c_int, [c_void_p, c_double, c_double_p]),
    "hdsdpcu_cone_update_dev": (c_int, [c_void_p, c_double, c_void_p]),
    "hdsdpcu_cone_updatebuffer": (c_int, [c_void_p, c_double, c_double, c_double_p, c_double, c_int]),
    "hdsdpcu_cone_interiorcheck": (c_int, [c_void_p, c_double, c_double_p, c_int_p]),
    "hdsdpcu_cone_interiorcheckexpert": (c_int, [c_void_p, c_double, c_double, c_double_p, c_double, c_int, c_int_p]),
    "hdsdpcu_cone_factorize": (c_int, [c_void_p, c_int, c_int_p]),
    "hdsdpcu_cone_getbarrier": (c_int, [c_void_p, c_double, c_double_p, c_int, c_double_p]),
    "hdsdpcu_cone_addstepandcheck": (c_int, [c_void_p, c_double, c_int, c_int_p]),
    "hdsdpcu_cone_buildschur": (c_int, [c_void_p, c_int, c_void_p, c_int]),
    "hdsdpcu_cone_ratiotest": (c_int, [c_void_p, c_double, c_double_p, c_double, c_int, c_double_p]),
    "hdsdpcu_cone_lanczosmultiply": (c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    "hdsdpcu_cone_lanczossteps": (c_int, [c_void_p]),
    "hdsdpcu_cone_buildprimalxsx": (c_int, [c_void_p, c_double_p, c_double_p, c_int]),
    "hdsdpcu_sym_extreme_eig": (c_int, [c_int, c_double_p, c_int, c_double_p, c_int_p]),
    "hdsdpcu_cone_getprimal": (c_int, [c_void_p, c_double, c_double_p, c_double_p, c_double_p, c_int_p]),
    "hdsdpcu_cone_xdots": (c_int, [c_void_p, c_double_p, c_double_p]),
    "hdsdpcu_cone_getdual": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_timer_start": (c_int, []),
    "hdsdpcu_timer_stop": (c_int, [c_double_p]),
    "hdsdpcu_cone_setsinv": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_cone_setsinv_linsys": (c_int, [c_void_p, c_void_p]),
    "hdsdpcu_cone_getbuffer": (c_int, [c_void_p, c_int, c_double_p]),
    "hdsdpcu_cone_getsinv": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_cone_getfactordiag": (c_int, [c_void_p, c_int, c_double_p]),
    # B2 kkt
    "hdsdpcu_kkt_create": (c_int, [POINTER(c_void_p), c_int]),
    "hdsdpcu_kkt_addcone": (c_int, [c_void_p, c_void_p]),
    "hdsdpcu_kkt_destroy": (None, [POINTER(c_void_p)]),
    "hdsdpcu_kkt_buildup": (c_int, [c_void_p, c_int]),
    "hdsdpcu_kkt_clean": (c_int, [c_void_p, c_int]),
    "hdsdpcu_kkt_buildupextra_bound": (c_int, [c_void_p, c_double_p, c_double_p, c_double_p, c_int]),
    "hdsdpcu_lp_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int_p, c_int_p, c_double_p]),
    "hdsdpcu_lp_setobjective": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_lp_destroy": (None, [POINTER(c_void_p)]),
    "hdsdpcu_kkt_buildupextra_lp": (c_int, [c_void_p, c_void_p, c_double_p, c_double, c_int]),
    "hdsdpcu_kkt_regularize": (c_int, [c_void_p, c_double]),
    "hdsdpcu_kkt_export": (c_int, [c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "hdsdpcu_kkt_factorize": (c_int, [c_void_p]),
    "hdsdpcu_kkt_solve": (c_int, [c_void_p, c_double_p, c_double_p]),
    "hdsdpcu_kkt_solve_status": (c_int, [c_void_p, c_double_p, c_int_p]),
    "hdsdpcu_kkt_set_solver": (c_int, [c_void_p, c_int]),
    "hdsdpcu_kkt_pcg_status": (c_int, [c_void_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "hdsdpcu_kkt_symv": (c_int, [c_void_p, c_double_p, c_double_p]),
    "hdsdpcu_kkt_solve_many": (c_int, [c_void_p, c_int, c_double_p, c_double_p]),
    "hdsdpcu_kkt_registerpsdp": (None, [c_void_p, c_int, POINTER(c_double_p)]),
    "hdsdpcu_kkt_addhost": (c_int, [c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "hdsdpcu_kkt_getmatrix": (c_int, [c_void_p, c_double_p]),
    "hdsdpcu_kkt_padded_dim": (c_int, [c_void_p]),
    "hdsdpcu_kkt_matrix_dev": (c_void_p, [c_void_p]),
    "hdsdpcu_kkt_asinv_dev": (c_void_p, [c_void_p]),
    "hdsdpcu_kkt_solve_dev": (c_int, [c_void_p, c_int, c_void_p]),
    "hdsdpcu_kkt_setshard": (c_int, [c_void_p, c_int, c_int]),
    "hdsdpcu_debug_leafclk": (c_int, [c_void_p]),
    "hdsdpcu_linsys_set_indefinite": (c_int, [c_void_p, c_int]),
    "hdsdpcu_linsys_inertia": (c_int, [c_void_p, c_int_p, c_int_p]),
    "hdsdpcu_kkt_ldl_status": (c_int, [c_void_p, c_int_p, c_int_p, c_int_p]),
    # multi-GPU Schur matrix
    "hdsdpcu_dist_blob_bytes": (c_int, []),
    "hdsdpcu_dist_owner": (c_int, [c_int, c_int, c_int]),
    "hdsdpcu_kkt_dist_init": (c_int, [c_void_p, c_int, c_int, c_int]),
    "hdsdpcu_kkt_dist_export": (c_int, [c_void_p, c_void_p]),
    "hdsdpcu_kkt_dist_connect": (c_int, [c_void_p, c_void_p]),
    "hdsdpcu_distchol_selftest": (c_int, [c_int, c_int, c_int, c_double_p, c_double_p, c_double_p, c_int_p, c_int, c_double_p]),
    "hdsdpcu_dgemm_nt_dev": (c_int, [c_int, c_int, c_int, c_double, c_void_p, c_long, c_void_p, c_long, c_double, c_void_p, c_long, c_int]),
}


def declared_symbols():
    """Every entry point include/hdsdpcu.h declares (used by the CPU symbol-export test)."""
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HdsdpCudaError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def require_gpu(device=-1):
    l = lib()
    if l.hdsdpcu_device_count() <= 0:
        raise HdsdpCudaError("no CUDA device visible: the hdsdp_b200 hot path has no CPU fallback")
    rc = l.hdsdpcu_init(device)
    if rc != 0:
        raise HdsdpCudaError(f"hdsdpcu_init failed with retcode {rc}")
    return l


def check(rc, what):
    if rc != 0:
        raise HdsdpCudaError(f"{what} failed with hdsdp_retcode {rc} ({'FAILED' if rc == 1 else 'MEMORY' if rc == 2 else '?'})")
