"""ctypes binding of include/whitomo_b200.h -- the stub a maintainer of the reference would add to
reach the sm_100a engine instead of ASTRA.  There is no CPU fallback: if the
library is missing or no CUDA device is present the import of any operator raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WT_LIB", os.path.join(_HERE, "libwhitomo_b200.so"))   # WT_LIB: tuning builds only

c_int, c_float, c_double, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_void_p

# every symbol include/whitomo_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "wt_last_error": (ctypes.c_char_p, []),
    "wt_version": (c_int, []),
    "wt_debug_oob_count": (ctypes.c_longlong, []),
    "wt_geom_plan": (c_int, [c_int, c_int, c_int, c_float, ctypes.POINTER(c_float), c_int, c_int] + [ctypes.POINTER(c_int)] * 7),
    "wt_geom_create": (c_int, [c_int, c_int, c_int, c_float, ctypes.POINTER(c_float), c_int, ctypes.POINTER(c_void_p)]),
    "wt_geom_create_ex": (c_int, [c_int, c_int, c_int, c_float, ctypes.POINTER(c_float), c_int, c_int, ctypes.POINTER(c_void_p)]),
    "wt_geom_positions": (c_int, [c_void_p]),
    "wt_geom_destroy": (None, [c_void_p]),
    "wt_geom_dims": (c_int, [c_void_p] + [ctypes.POINTER(c_int)] * 4),
    "wt_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "wt_fp": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "wt_mar_fp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                          c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "wt_bp": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "wt_bp_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t,
                           c_void_p]),
    "wt_bp_row_info": (c_int, [c_void_p, c_int] + [ctypes.POINTER(c_int)] * 3),
    "wt_sirt": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "wt_sirt_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "wt_spectrum_create": (c_int, [c_int, c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_double),
                                   ctypes.POINTER(c_double), c_double, ctypes.POINTER(c_void_p)]),
    "wt_spectrum_destroy": (None, [c_void_p]),
    "wt_white_fp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                            c_void_p, c_size_t, c_void_p]),
    "wt_step_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "wt_white_barrier_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p, c_size_t, c_void_p]),
    "wt_split_plan": (c_int, [c_int] * 5 + [ctypes.POINTER(c_int)] * 3),
    "wt_step_times": (c_int, [c_void_p, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float)]),
    "wt_barrier_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "wt_pair_norm2": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
}
PAIR_NORM_BLOCKS = 32      # WT_PAIR_NORM_BLOCKS of include/whitomo_b200.h



class UpdateParams(ctypes.Structure):
    """wt_update_params of include/whitomo_b200.h"""
    _fields_ = [("alpha", c_float), ("beta_reg", c_float), ("t", c_float), ("reg_w", c_float), ("triv_w", c_float),
                ("lower_only", c_int), ("mode", c_int), ("elem_mask", ctypes.c_uint)]


UPD_STEP_CLIP, UPD_STEP_ONLY, UPD_CLIP_ONLY = 0, 1, 2
POS_ASTRA, POS_EXACT = 0, 1
STEP_INIT, STEP_TIMED, STEP_RING = 1, 2, 64

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load the shared library and type every entry point (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError("whitomo_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise EngineError("whitomo_b200 error %d: %s" % (rc, load().wt_last_error().decode()))
