"""Loads libgsm_b200.so (the sm_100a implementation behind include/gsm_b200.h) through ctypes.

There is no CPU fallback and no alternative backend: if the library has not been built
(`python -c "import __graft_entry__ as g; g.build()"`) or cannot be loaded, importing it raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsm_b200.so")

_lib = None

_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32

# name -> (restype, argtypes); every symbol include/gsm_b200.h declares
SIGNATURES = {
    "gsm_abi_version": (C.c_int, []),
    "gsm_device_count": (C.c_int, []),
    "gsm_create": (C.c_int, [C.POINTER(_vp), _i32, _i64, _i32]),
    "gsm_destroy": (C.c_int, [_vp]),
    "gsm_last_error": (C.c_char_p, [_vp]),
    "gsm_set_mode": (C.c_int, [_vp, _i32, _u32]),
    "gsm_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "gsm_upload_trees": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "gsm_upload_branch_params": (C.c_int, [_vp, _vp, _vp, _vp]),
    "gsm_upload_tree_params": (C.c_int, [_vp, _vp, _vp]),
    "gsm_upload_stubs": (C.c_int, [_vp, _vp, _vp, _vp]),
    "gsm_bind_device_stubs": (C.c_int, [_vp, _vp, _vp, _vp]),
    "gsm_eval": (C.c_int, [_vp, _vp, _vp, _vp]),
    "gsm_eval_host": (C.c_int, [_vp, _i64, _i32] + [_vp] * 11),
    "gsm_eval_dirty": (C.c_int, [_vp, _i64] + [_vp] * 11),
    "gsm_last_dirty_full_count": (_i64, [_vp]),
    "gsm_stub_cdf": (C.c_int, [_vp, _i64, _i32, _vp, _i32, C.POINTER(_i32)]),
    "gsm_stub_cdf_trees": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "gsm_stub_choose": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "gsm_format_newick": (_i64, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, C.c_char_p, _vp, _vp, C.c_char_p, _i64]),
    "gsm_parse_newick": (_i64, [C.c_char_p, _i32, _vp, _vp, _vp, _vp, _vp, _i32, C.c_char_p, _vp]),
    "gsm_effective_leaf_rates": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "gsm_java_double_to_string": (_i64, [C.c_double, C.c_char_p, _i64]),
    "gsm_host_alloc": (C.c_int, [C.POINTER(_vp), _i64]),
    "gsm_host_free": (C.c_int, [_vp]),
    "gsm_device_stride": (_i64, [_i32]),
    "gsm_bind_device": (C.c_int, [_vp, _i64, _i32, _i64] + [_vp] * 8),
    "gsm_eval_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "gsm_sync": (C.c_int, [_vp]),
    "gsm_device_alloc": (C.c_int, [C.POINTER(_vp), _i32, _i64]),
    "gsm_device_free": (C.c_int, [_vp, _i32]),
    "gsm_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "gsm_ipc_open": (C.c_int, [_i32, C.c_char_p, C.POINTER(_vp)]),
    "gsm_ipc_close": (C.c_int, [_vp]),
    "gsm_copy_async": (C.c_int, [_vp, _vp, _i64, _vp]),
    "gsm_measure_fp64_peak": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "gsm_launch_count": (_i64, [_vp]),
}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gammaspikemodel_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
