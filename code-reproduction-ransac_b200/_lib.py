"""ctypes binding of libransac_b200.so (C ABI in include/ransac_b200.h).

The library is the product: if it is missing, or no sm_100 device is usable, calls fail loudly — there is no
CPU fallback and nothing here imports the oracle."""
import ctypes as C
import os

from . import _build

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_u8_p = C.POINTER(C.c_uint8)
c_i32_p = C.POINTER(C.c_int32)
c_u64_p = C.POINTER(C.c_uint64)


class HParams(C.Structure):
    """b2r_h_params"""
    _fields_ = [("thr", C.c_double), ("max_iters", C.c_int32), ("confidence", C.c_double), ("sampler", C.c_int32),
                ("seed", C.c_uint64), ("arith", C.c_int32), ("mask_semantics", C.c_int32), ("refine", C.c_int32),
                ("hyp_begin", C.c_int64), ("solver", C.c_int32), ("reserved", C.c_int32)]


class HInfo(C.Structure):
    """b2r_h_info"""
    _fields_ = [("status", C.c_int32), ("iters_run", C.c_int32), ("best_iter", C.c_int32), ("best_count", C.c_int32),
                ("sample", C.c_int32 * 4), ("n_inliers", C.c_int32), ("lm_iters", C.c_int32), ("reserved", C.c_int32 * 2)]


class PParams(C.Structure):
    """b2r_p_params"""
    _fields_ = [("thr", C.c_double), ("max_iters", C.c_int32), ("confidence", C.c_double), ("sampler", C.c_int32),
                ("seed", C.c_uint64), ("arith", C.c_int32), ("refine", C.c_int32), ("hyp_begin", C.c_int64),
                ("solver", C.c_int32), ("reserved", C.c_int32)]


class PInfo(C.Structure):
    """b2r_p_info"""
    _fields_ = [("status", C.c_int32), ("iters_run", C.c_int32), ("best_iter", C.c_int32), ("best_count", C.c_int32),
                ("sample", C.c_int32 * 5), ("n_inliers", C.c_int32), ("lm_iters", C.c_int32), ("reserved", C.c_int32),
                ("ransac_rvec", C.c_double * 3), ("ransac_tvec", C.c_double * 3), ("mean_inlier_err", C.c_double),
                ("sum_sq_err", C.c_double)]


# name -> (restype, argtypes); every symbol include/ransac_b200.h declares
SIGNATURES = {
    "b2r_version": (C.c_int, []),
    "b2r_last_error": (C.c_char_p, []),
    "b2r_device_count": (C.c_int, []),
    "b2r_ctx_create": (C.c_void_p, [C.c_int]),
    "b2r_ctx_destroy": (None, [C.c_void_p]),
    "b2r_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "b2r_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "b2r_default_h_params": (None, [C.POINTER(HParams)]),
    "b2r_find_homography": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.POINTER(HParams), c_double_p,
                                      c_u8_p, C.POINTER(HInfo)]),
    "b2r_find_homography_batch": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32,
                                            C.POINTER(HParams), c_double_p, c_u8_p, C.POINTER(HInfo)]),
    "b2r_camera_sweep": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, c_double_p, C.c_int32, C.POINTER(HParams),
                                   c_double_p, c_double_p, c_double_p, c_u8_p, C.POINTER(HInfo), c_i32_p]),
    "b2r_h_problem_upload": (C.c_void_p, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32]),
    "b2r_h_problem_reupload": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32]),
    "b2r_h_problem_free": (None, [C.c_void_p, C.c_void_p]),
    "b2r_h_problem_run": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HParams)]),
    "b2r_h_problem_fetch": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_u8_p, C.POINTER(HInfo)]),
    "b2r_h_problem_score_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HParams), c_u64_p]),
    "b2r_h_problem_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HParams), c_u64_p]),
    "b2r_h_problem_score_shard_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HParams), C.c_void_p]),
    "b2r_h_problem_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(HParams), C.c_void_p]),
    "b2r_h_problem_stage_ms": (C.c_int, [C.c_void_p, C.c_void_p, c_float_p]),
    "b2r_h_problem_peek_hyps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_i32_p, c_float_p, c_i32_p]),
    "b2r_ctx_launch_count": (C.c_int, [C.c_void_p]),
    "b2r_score_h": (C.c_int, [C.c_void_p, c_float_p, C.c_int32, c_float_p, c_float_p, C.c_int32, C.c_float, C.c_int32,
                              c_i32_p]),
    "b2r_solve_h4": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int32, c_i32_p, C.c_int32, C.c_int32, c_double_p,
                               c_u8_p, c_u8_p]),
    "b2r_sample_cv": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int32, C.c_int32, c_i32_p, c_i32_p]),
    "b2r_sample_philox": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int32, C.c_uint64, C.c_int32, C.c_int64,
                                    C.c_int32, c_i32_p]),
    "b2r_refine_h": (C.c_int, [C.c_void_p, c_float_p, c_float_p, C.c_int32, c_u8_p, c_double_p, c_i32_p]),
    "b2r_jacobi_eig": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "b2r_selftest_rcp": (C.c_int, [C.c_void_p, c_u64_p, c_u64_p]),
    "b2r_probe_fp32_peak": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    # ---- PnP path
    "b2r_default_p_params": (None, [C.POINTER(PParams)]),
    "b2r_solve_pnp_ransac": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, c_double_p, C.POINTER(PParams),
                                       c_double_p, c_double_p, c_i32_p, c_i32_p, C.POINTER(PInfo)]),
    "b2r_solve_pnp_ransac_batch": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32, c_double_p,
                                             C.POINTER(PParams), c_double_p, c_double_p, c_i32_p, c_i32_p, C.POINTER(PInfo)]),
    "b2r_solve_pnp_refine_lm": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p,
                                          C.c_int32, c_i32_p]),
    "b2r_p_problem_upload": (C.c_void_p, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32, c_double_p]),
    "b2r_p_problem_reupload": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32,
                                         c_double_p]),
    "b2r_p_problem_free": (None, [C.c_void_p, C.c_void_p]),
    "b2r_p_problem_run": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PParams)]),
    "b2r_p_problem_fetch": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_i32_p, c_i32_p, C.POINTER(PInfo)]),
    "b2r_p_problem_score_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PParams), c_u64_p]),
    "b2r_p_problem_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PParams), c_u64_p]),
    "b2r_p_problem_score_shard_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PParams), C.c_void_p]),
    "b2r_p_problem_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PParams), C.c_void_p]),
    "b2r_p_problem_stage_ms": (C.c_int, [C.c_void_p, C.c_void_p, c_float_p]),
    "b2r_score_p": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, c_double_p, c_double_p, C.c_int32, c_double_p, C.c_float,
                              C.c_int32, c_i32_p]),
    "b2r_pnp_minimal_models": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, c_double_p, c_i32_p, C.c_int32,
                                         C.c_int32, c_double_p, c_double_p, c_double_p, c_u8_p]),
    "b2r_sample_cv_p": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i32_p]),
    # ---- DEM ray-march (row f4)
    "b2r_dem_upload": (C.c_void_p, [C.c_void_p, c_double_p, C.c_int32, c_double_p, C.c_int32, c_double_p]),
    "b2r_dem_free": (None, [C.c_void_p, C.c_void_p]),
    "b2r_ray_march_dem": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int32, c_double_p, C.c_int32, c_double_p, C.c_double,
                                    C.c_double, C.c_int32, c_double_p, c_i32_p, c_i32_p]),
    "b2r_pixels_to_geo": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p,
                                    c_double_p, C.c_int32, c_double_p, C.c_double, C.c_double, C.c_int32, c_double_p, c_i32_p,
                                    c_i32_p, c_double_p]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load the shared library and bind every exported symbol (raises if the library or a symbol is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  ransac_b200 has no CPU fallback.")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def last_error():
    msg = load().b2r_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
