"""ctypes binding of libmwrt.so (include/mwrt.h).  There is no CPU fallback: if the library is missing
or a CUDA device is absent the product path raises -- it never routes through `oracle/`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MWRT_LIB") or os.path.join(HERE, "libmwrt.so")  # override: experiments only

H2O_COLS = ("fl", "s1", "b2", "w0", "x", "w0s", "xs", "sr", "sh", "xh", "shs", "xhs", "aair", "aself",
            "w2", "w2s", "d2", "d2s")
O2_COLS = ("f", "s300", "be", "w300", "y0", "y1", "g0", "g1", "dnu0", "dnu1")

ST_NAN_INPUT, ST_NEG_ABS, ST_DUCTING, ST_RANGE = 1, 2, 4, 8
F_RAY_TRACING, F_FROM_SAT, F_CLOUDY, F_TAU_LAYERS = 1, 2, 4, 8
E_ARG, E_MODEL, E_NOMEM, E_STATE = -1, -2, -3, -4
PREC_F64, PREC_F32_LINESHAPE = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("n_h2o", "n_o2", "h2o_sd", "o2_mixing", "o2_dens118",
                                          "o2_sd118", "n2_fdep", "liq_model")]
                + [(n, C.c_double) for n in ("h2o_reftline", "h2o_reftcon", "h2o_cf", "h2o_xcf",
                                             "h2o_cs", "h2o_xcs", "h2o_pvap_div", "h2o_den_const",
                                             "h2o_line_const", "o2_x", "o2_wb300", "o2_pvap_div",
                                             "o2_h2o_broad", "o2_nonres", "o2_abs_num", "o2_abs_den",
                                             "o2_iso", "n2_l", "n2_m", "n2_n")]
                + [("h2o_lines", c_dp), ("o2_lines", c_dp)])


class Profiles(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("z", "p", "tk", "rh", "denliq", "denice", "cld_ibase", "cld_itop")]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq",
                                          "tauice", "tau_layer", "status", "tmrcld")]


# every symbol include/mwrt.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = {
    "mwrt_version": (C.c_int, []),
    "mwrt_last_error": (C.c_char_p, []),
    "mwrt_set_device": (C.c_int, [C.c_int]),
    "mwrt_get_device": (C.c_int, [C.POINTER(C.c_int)]),
    "mwrt_builtin_model_count": (C.c_int, []),
    "mwrt_builtin_model_name": (C.c_char_p, [C.c_int]),
    "mwrt_builtin_model_desc": (C.c_int, [C.c_char_p, C.POINTER(ModelDesc)]),
    "mwrt_model_create": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mwrt_model_create_custom": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "mwrt_model_destroy": (C.c_int, [C.c_void_p]),
    "mwrt_vapor_f64": (C.c_int, [C.c_int64] + [C.c_void_p] * 5),
    "mwrt_absorption_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 12),
    "mwrt_absorption_f32ls": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 12),
    "mwrt_slant_path_f64": (C.c_int, [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
                            + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mwrt_rte_f64": (C.c_int, [C.c_int] * 4 + [C.c_void_p] * 7 + [C.c_int, C.c_double]
                     + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p]),
    "mwrt_bandpass_f64": (C.c_int, [C.c_int] * 4 + [C.c_void_p] * 4),
    "mwrt_plan_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p)]),
    "mwrt_plan_destroy": (C.c_int, [C.c_void_p]),
    "mwrt_plan_set_emissivity": (C.c_int, [C.c_void_p, C.c_double]),
    "mwrt_plan_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "mwrt_plan_absorption": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 7),
    "mwrt_plan_set_split": (C.c_int, [C.c_void_p, C.c_int]),
    "mwrt_plan_set_bandpass": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp]),
    "mwrt_plan_launch_count": (C.c_int, [C.c_void_p, C.c_int]),
    "mwrt_plan_forward_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Profiles),
                                           C.POINTER(Outputs), C.c_void_p]),
    "mwrt_plan_forward_host": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Profiles),
                                         C.POINTER(Outputs)]),
    "mwrt_push_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mwrt_push_kernel": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int,
                                   C.c_void_p]),
    "mwrt_plan_set_output_fanout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "mwrt_stats_f64": (C.c_int, [C.c_int, C.c_int64] + [C.c_void_p] * 7),
    "mwrt_thin_mean_f64": (C.c_int, [C.c_int] * 3 + [C.c_void_p] * 12),
    "mwrt_assemble_levels_f64": (C.c_int, [C.c_int] * 4 + [C.c_double] + [C.c_void_p] * 18),
    "mwrt_cloud_features_f64": (C.c_int, [C.c_int, C.c_int] + [C.c_void_p] * 11),
    "mwrt_math_selftest_f64": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mwrt_measure_dfma_peak": (C.c_int, [c_dp, c_dp]),
}

_lib = None


def load():
    """dlopen libmwrt.so and bind every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmwrt.so is not built (%s).  Run `python -c 'import __graft_entry__ as g; g.build()'`"
            " -- this package has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def set_device(index):
    """Make libmwrt's (statically linked) CUDA runtime use the same device as the caller's tensors."""
    check(load().mwrt_set_device(int(index)), "set_device")


def check(rc, what=""):
    """Map the C return convention to Python exceptions (ValueError for bad arguments / models --
    what PyRTlib raises -- RuntimeError for CUDA failures)."""
    if rc == 0:
        return
    msg = load().mwrt_last_error().decode(errors="replace")
    if rc < 0:
        raise ValueError("%s: %s (code %d)" % (what, msg, rc))
    raise RuntimeError("%s: %s (cudaError %d)" % (what, msg, rc))
