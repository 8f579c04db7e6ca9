"""ctypes binding of libgsl_scattered_cuda.so (C ABI in include/gsl_scattered_cuda.h).

The library is built in-tree (``gsl_scattered_interpolation_b200/lib/``) by ``make`` in
``csrc/`` (what ``__graft_entry__.build()`` runs).  There is deliberately no fallback: if the
shared library is missing, loading the binding raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_size_t, c_ulong, c_ulonglong,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgsl_scattered_cuda.so")
# A/B experiments only (scripts/): load another build of the same library
LIB_PATH = os.environ.get("GSI_LIB_PATH", LIB_PATH)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gsl_scattered_cuda.h")

SUCCESS, EDOM, EINVAL, EFAILED, ENOMEM, EBADLEN = 0, 1, 4, 5, 8, 19
DEFAULT, NOSTANDARDIZE, ISOSCALE = 0, 1, 2


class Info(ctypes.Structure):
    _fields_ = [
        ("n_nodes", c_size_t), ("n_fields", c_size_t), ("n_simplexes", c_size_t),
        ("n_leaves", c_size_t), ("grid_g", c_size_t), ("slab_bytes", c_size_t),
        ("n_devices", c_size_t), ("build_seconds", c_double), ("flatten_seconds", c_double),
        ("upload_seconds", c_double), ("n_deferred", c_ulonglong), ("n_outside", c_ulonglong),
        ("n_steps", c_ulonglong), ("n_kernel_launches", c_ulonglong),
        ("locate_seconds", c_double), ("exact_seconds", c_double),
        ("n_reordered", c_ulonglong), ("n_contexts", c_size_t),
    ]


ERROR_HANDLER = ctypes.CFUNCTYPE(None, c_char_p, c_char_p, c_int, c_int)

_BATCH_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)


class TypeDescriptor(ctypes.Structure):
    """gsl_scattered_cuda_type (include/gsl_scattered_cuda.h), cf. gsl_interp2d_type"""
    _fields_ = [
        ("name", c_char_p), ("min_size", ctypes.c_uint),
        ("alloc", ctypes.CFUNCTYPE(c_void_p, c_size_t, c_size_t)),
        ("init", ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p)),
        ("eval_e", ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p)),
        ("eval_deriv_x_e", _BATCH_FN), ("eval_deriv_y_e", _BATCH_FN), ("eval_deriv_xx_e", _BATCH_FN),
        ("eval_deriv_xy_e", _BATCH_FN), ("eval_deriv_yy_e", _BATCH_FN),
        ("free", ctypes.CFUNCTYPE(None, c_void_p)),
    ]

_PROTOS = {
    "gsl_scattered_cuda_version": (c_char_p, []),
    "gsl_scattered_cuda_set_error_handler": (c_void_p, [c_void_p]),
    "gsl_scattered_cuda_device_count": (c_int, []),
    "gsl_scattered_cuda_alloc": (c_void_p, [c_size_t, c_size_t]),
    "gsl_scattered_cuda_free": (None, [c_void_p]),
    "gsl_scattered_cuda_set_devices": (c_int, [c_void_p, POINTER(c_int), c_size_t]),
    "gsl_scattered_cuda_set_param": (c_int, [c_void_p, c_char_p, c_double]),
    "gsl_scattered_cuda_shuffle_mt19937": (c_int, [c_ulong, c_size_t, c_void_p]),
    "gsl_scattered_cuda_init": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                                        c_void_p]),
    "gsl_scattered_cuda_build": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "gsl_scattered_cuda_upload": (c_int, [c_void_p, c_void_p, c_size_t]),
    "gsl_scattered_cuda_init_from_tree": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_size_t, c_size_t, c_void_p, c_size_t]),
    "gsl_scattered_cuda_eval_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                          c_void_p]),
    "gsl_scattered_cuda_eval": (c_double, [c_void_p, c_double, c_double]),
    "gsl_scattered_cuda_eval_raster_e": (c_int, [c_void_p, c_double, c_double, c_size_t, c_double,
                                                 c_double, c_size_t, c_void_p, c_void_p]),
    "gsl_scattered_cuda_eval_device_e": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                                 c_void_p, c_size_t, c_void_p, c_void_p]),
    "gsl_scattered_cuda_eval_raster_device_e": (c_int, [c_void_p, c_int, c_double, c_double,
                                                        c_size_t, c_double, c_double, c_size_t,
                                                        c_void_p, c_void_p, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_x_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_y_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_xx_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_yy_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_xy_e": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_eval_deriv_x": (c_double, [c_void_p, c_double, c_double]),
    "gsl_scattered_cuda_eval_deriv_y": (c_double, [c_void_p, c_double, c_double]),
    "gsl_scattered_cuda_eval_deriv_device_e": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p,
                                                       c_void_p, c_size_t, c_void_p, c_void_p]),
    "gsl_scattered_cuda_type_linear_simplex": (c_void_p, []),
    "gsl_scattered_cuda_name": (c_char_p, [c_void_p]),
    "gsl_scattered_cuda_min_size": (ctypes.c_uint, [c_void_p]),
    "gsl_scattered_cuda_type_min_size": (ctypes.c_uint, [c_void_p]),
    "gsl_scattered_cuda_get_info": (c_int, [c_void_p, POINTER(Info)]),
    "gsl_scattered_cuda_reset_counters": (c_int, [c_void_p]),
    "gsl_scattered_cuda_get_tree": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p]),
    "gsl_scattered_cuda_slab_bytes": (c_size_t, [c_void_p]),
    "gsl_scattered_cuda_slab_export_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "gsl_scattered_cuda_init_from_slab_device": (c_int, [c_void_p, c_void_p, c_size_t]),
    "gsl_scattered_cuda_save": (c_int, [c_void_p, c_char_p]),
    "gsl_scattered_cuda_load": (c_int, [c_void_p, c_char_p]),
    "gsl_scattered_cuda_fwrite": (c_int, [c_void_p, c_void_p]),
    "gsl_scattered_cuda_fread": (c_int, [c_void_p, c_void_p]),
    "gsl_scattered_cuda_host_alloc": (c_void_p, [c_size_t]),
    "gsl_scattered_cuda_host_free": (None, [c_void_p]),
    "gsl_scattered_cuda_generate_uniform_device": (c_int, [c_void_p, c_int, c_ulonglong,
                                                           c_ulonglong, c_size_t, c_double,
                                                           c_double, c_double, c_double, c_void_p,
                                                           c_void_p, c_void_p]),
    "gsl_scattered_cuda_generate_uniform_host": (None, [c_ulonglong, c_ulonglong, c_size_t,
                                                        c_double, c_double, c_double, c_double,
                                                        c_void_p, c_void_p]),
    "gsl_scattered_cuda_checksum_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                                   c_void_p, c_void_p]),
    "gsl_scattered_cuda_selftest_division": (c_int, [c_int, c_ulonglong, c_size_t, c_void_p]),
    "gsl_scattered_cuda_build_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "gsl_scattered_cuda_debug_walk_trace": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "gsl_scattered_cuda_compare_device": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_size_t, c_size_t, c_void_p,
                                                  c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load():
    """Load libgsl_scattered_cuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  This package has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
