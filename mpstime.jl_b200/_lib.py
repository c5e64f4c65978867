"""ctypes binding of libmpstime_b200.so (include/mpstime_b200.h).  The library is the product:
there is no CPU fallback -- a missing library or a missing GPU raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpstime_b200.so")

c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_u8_p = C.POINTER(C.c_uint8)


class TrainOpts(C.Structure):
    """mpst_train_opts"""
    _fields_ = [("loss_kind", C.c_int32), ("opt_kind", C.c_int32), ("train_sep", C.c_int32),
                ("update_iters", C.c_int32), ("rescale_before", C.c_int32), ("rescale_after", C.c_int32),
                ("chi_max", C.c_int32), ("reserved", C.c_int32), ("eta", C.c_double), ("cutoff", C.c_double)]


class ImputeOpts(C.Structure):
    """mpst_impute_opts"""
    _fields_ = [("backwards", C.c_int32), ("get_err", C.c_int32), ("max_trials", C.c_int32), ("reserved", C.c_int32),
                ("rejection_threshold", C.c_double), ("max_jump", C.c_double)]


# name -> (restype, argtypes): every symbol include/mpstime_b200.h declares
SIGNATURES = {
    "mpst_version": (C.c_int, []),
    "mpst_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mpst_model_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mpst_destroy": (C.c_int, [C.c_void_p]),
    "mpst_last_error": (C.c_char_p, [C.c_void_p]),
    "mpst_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mpst_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "mpst_encode": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int64, c_double_p]),
    "mpst_set_encoding_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_i32_p, C.c_int64, c_double_p, C.c_int64]),
    "mpst_encode_site": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int64, c_double_p]),
    "mpst_train_load_x": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int, c_i64_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int64, c_i64_p]),
    "mpst_train_load_phi": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int, c_i64_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int64, c_i64_p]),
    "mpst_set_core": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int]),
    "mpst_get_core_dims": (C.c_int, [C.c_void_p, C.c_int, c_i32_p, c_i32_p, c_i32_p]),
    "mpst_get_core": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mpst_build_env": (C.c_int, [C.c_void_p, C.c_int]),
    "mpst_bond_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(TrainOpts), c_double_p, c_double_p,
                                 c_i32_p]),
    "mpst_sweep": (C.c_int, [C.c_void_p, C.POINTER(TrainOpts), C.c_int, c_double_p, c_double_p, c_i32_p]),
    "mpst_sweep_bonds": (C.c_int, [C.c_void_p, C.POINTER(TrainOpts), C.c_int, C.c_int, c_double_p, c_double_p, c_i32_p]),
    "mpst_overlaps": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_i64_p]),
    "mpst_eval_metrics": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_i64_p, c_double_p, c_i64_p]),
    "mpst_impute_batch": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_u8_p, C.c_int64, C.c_int, c_double_p,
                                    C.c_int, c_double_p, C.c_int, C.c_double, c_double_p]),
    "mpst_impute_batch_ex": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_u8_p, C.c_int64, C.c_int, c_double_p,
                                       C.c_int, c_double_p, C.c_int64, C.c_int, C.POINTER(ImputeOpts), c_double_p,
                                       c_double_p]),
    "mpst_bond_loss_grad": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      C.c_int64, C.c_int, C.c_int, C.c_int, c_i64_p, C.c_int, C.c_int, C.c_int,
                                      c_double_p, c_double_p, c_double_p]),
    "mpst_bond_split": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, c_i32_p, c_double_p, c_double_p, c_double_p]),
    "mpst_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mpst_profile_get": (C.c_int, [C.c_void_p, c_double_p, c_i64_p, c_double_p]),
    "mpst_timer_start": (C.c_int, [C.c_void_p]),
    "mpst_timer_stop": (C.c_int, [C.c_void_p, c_double_p]),
    "mpst_profile_reset": (C.c_int, [C.c_void_p]),
    "mpst_launch_count": (C.c_int64, [C.c_void_p]),
    "mpst_debug_set": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "mpst_debug_get": (C.c_int64, [C.c_void_p, C.c_char_p]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every exported symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(mpstime.jl_b200/csrc/build.sh).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
