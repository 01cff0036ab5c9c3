"""ctypes binding of oracle/_build/liboracle.so (built by oracle/Makefile) + small NumPy mirrors.

TEST INFRASTRUCTURE ONLY.  Every function cites the reference call site whose arithmetic it checks;
the arithmetic itself lives in OpenCV calib3d 4.13.0 (un-vendored dependency of the reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile the C restatement (gcc, seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "cv_ransac_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, fp, u8p, i32p, ip = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int))
        L.orc_jacobi.argtypes = [dp, C.c_int, dp, dp]
        L.orc_h_run_kernel.argtypes = [fp, fp, C.c_int, dp]
        L.orc_h_run_kernel.restype = C.c_int
        L.orc_h_check_subset.argtypes = [fp, fp, C.c_int]
        L.orc_h_check_subset.restype = C.c_int
        L.orc_h_compute_error.argtypes = [dp, fp, fp, C.c_int, fp]
        L.orc_h_count_inliers_f32.argtypes = [fp, C.c_int, fp, fp, C.c_int, C.c_float, i32p]
        L.orc_update_num_iters.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_update_num_iters.restype = C.c_int
        L.orc_h_ransac_stage.argtypes = [fp, fp, C.c_int, C.c_double, C.c_int, C.c_double, dp, u8p, ip, i32p, i32p,
                                         C.POINTER(C.c_uint64)]
        L.orc_h_ransac_stage.restype = C.c_int
        L.orc_h_lm_refine.argtypes = [fp, fp, C.c_int, dp, C.c_int]
        L.orc_h_lm_refine.restype = C.c_int
        L.orc_find_homography.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, dp, u8p, ip, u8p, dp]
        L.orc_find_homography.restype = C.c_int
        L.orc_rodrigues.argtypes = [dp, dp]
        L.orc_pnp_project_f32.argtypes = [dp, dp, dp, fp, C.c_int, fp]
        L.orc_pnp_count_inliers.argtypes = [dp, dp, dp, fp, fp, C.c_int, C.c_double, u8p]
        L.orc_pnp_count_inliers.restype = C.c_int
        L.orc_pnp_sample_stream.argtypes = [C.c_int, C.c_int, C.c_int, i32p]
        L.orc_rng_next.argtypes = [C.POINTER(C.c_uint64)]
        L.orc_rng_next.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _f32(a, cols):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols))


def rng_stream(n, seed=0xFFFFFFFFFFFFFFFF):
    """First n outputs of cv::RNG(seed).next()."""
    st = C.c_uint64(seed)
    return [int(lib().orc_rng_next(C.byref(st))) for _ in range(n)]


def jacobi(A):
    """cv::eigen on a symmetric matrix: (eigenvalues descending, rows = eigenvectors)."""
    A = np.array(A, dtype=np.float64, order="C")
    n = A.shape[0]
    W = np.zeros(n)
    V = np.zeros((n, n))
    lib().orc_jacobi(_p(A, C.c_double), n, _p(W, C.c_double), _p(V, C.c_double))
    return W, V


def h_run_kernel(src, dst):
    """HomographyEstimatorCallback::runKernel on fp32 points.  Returns H (3,3) or None."""
    s, d = _f32(src, 2), _f32(dst, 2)
    H = np.zeros(9)
    ok = lib().orc_h_run_kernel(_p(s, C.c_float), _p(d, C.c_float), len(s), _p(H, C.c_double))
    return H.reshape(3, 3) if ok else None


def h_check_subset(src4, dst4):
    s, d = _f32(src4, 2), _f32(dst4, 2)
    return bool(lib().orc_h_check_subset(_p(s, C.c_float), _p(d, C.c_float), len(s)))


def h_compute_error(H, src, dst):
    """fp32 un-fused squared reprojection error of every point."""
    s, d = _f32(src, 2), _f32(dst, 2)
    Hd = np.ascontiguousarray(np.asarray(H, dtype=np.float64).reshape(9))
    err = np.zeros(len(s), dtype=np.float32)
    lib().orc_h_compute_error(_p(Hd, C.c_double), _p(s, C.c_float), _p(d, C.c_float), len(s), _p(err, C.c_float))
    return err


def h_count_inliers_f32(models8, src, dst, thr_sq):
    """Inlier counts of fp32 models [H,8] over all points: the K3 kernel's CPU counterpart."""
    m = np.ascontiguousarray(np.asarray(models8, dtype=np.float32).reshape(-1, 8))
    s, d = _f32(src, 2), _f32(dst, 2)
    counts = np.zeros(len(m), dtype=np.int32)
    lib().orc_h_count_inliers_f32(_p(m, C.c_float), len(m), _p(s, C.c_float), _p(d, C.c_float), len(s),
                                  C.c_float(np.float32(thr_sq)), _p(counts, C.c_int32))
    return counts


def update_num_iters(p, ep, model_points, max_iters):
    return int(lib().orc_update_num_iters(p, ep, model_points, max_iters))


def h_ransac_stage(src, dst, thr, max_iters=2000, confidence=0.995, trace=True):
    """The RANSAC stage of cv2.findHomography on fp32-quantised points.

    Returns dict(ok, H, mask, iters, samples [iters,4], counts [iters], draws)."""
    s, d = _f32(src, 2), _f32(dst, 2)
    n = len(s)
    H = np.zeros(9)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    iters = C.c_int(0)
    draws = C.c_uint64(0)
    nit = max(max_iters, 1)
    tidx = np.full((nit, 4), -1, dtype=np.int32)
    tcnt = np.full(nit, -2, dtype=np.int32)
    ok = lib().orc_h_ransac_stage(_p(s, C.c_float), _p(d, C.c_float), n, thr, max_iters, confidence, _p(H, C.c_double),
                                  _p(mask, C.c_uint8), C.byref(iters), _p(tidx, C.c_int32) if trace else None,
                                  _p(tcnt, C.c_int32) if trace else None, C.byref(draws))
    k = iters.value
    return dict(ok=bool(ok), H=H.reshape(3, 3) if ok else None, mask=mask[:n].copy(), iters=k, samples=tidx[:k].copy(),
                counts=tcnt[:k].copy(), draws=int(draws.value))


def h_lm_refine(src_inl, dst_inl, H0, max_iters=10):
    s, d = _f32(src_inl, 2), _f32(dst_inl, 2)
    H = np.ascontiguousarray(np.asarray(H0, dtype=np.float64).reshape(9)).copy()
    it = lib().orc_h_lm_refine(_p(s, C.c_float), _p(d, C.c_float), len(s), _p(H, C.c_double), max_iters)
    return H.reshape(3, 3), it


def find_homography(src, dst, thr, max_iters=2000, confidence=0.995, mask_semantics=0, details=False):
    """cv2.findHomography(src, dst, cv2.RANSAC, thr, maxIters=..., confidence=...) restated.

    Reference call site: /root/reference/main_v1.py:312.  Returns (H or None, mask uint8 (n,1)) like cv2;
    with details=True also a dict with the RANSAC-stage mask/model and iteration count."""
    s = np.ascontiguousarray(np.asarray(src, dtype=np.float64).reshape(-1, 2))
    d = np.ascontiguousarray(np.asarray(dst, dtype=np.float64).reshape(-1, 2))
    n = len(s)
    H = np.zeros(9)
    Hr = np.zeros(9)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    rmask = np.zeros(max(n, 1), dtype=np.uint8)
    iters = C.c_int(0)
    ok = lib().orc_find_homography(_p(s, C.c_double), _p(d, C.c_double), n, thr, max_iters, confidence, mask_semantics,
                                   _p(H, C.c_double), _p(mask, C.c_uint8), C.byref(iters), _p(rmask, C.c_uint8),
                                   _p(Hr, C.c_double))
    out = (H.reshape(3, 3) if ok else None, mask[:n].reshape(n, 1).copy())
    if details:
        return out + (dict(iters=iters.value, ransac_mask=rmask[:n].copy(), ransac_H=Hr.reshape(3, 3)),)
    return out


def rodrigues(rvec):
    r = np.ascontiguousarray(np.asarray(rvec, dtype=np.float64).reshape(3))
    R = np.zeros(9)
    lib().orc_rodrigues(_p(r, C.c_double), _p(R, C.c_double))
    return R.reshape(3, 3)


def pnp_project_f32(R, t, K, obj_f32):
    """cv::projectPoints for zero distortion: fp64 arithmetic, fp32 output."""
    R = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(9))
    t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(3))
    K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    o = _f32(obj_f32, 3)
    out = np.zeros((len(o), 2), dtype=np.float32)
    lib().orc_pnp_project_f32(_p(R, C.c_double), _p(t, C.c_double), _p(K, C.c_double), _p(o, C.c_float), len(o),
                              _p(out, C.c_float))
    return out


def pnp_count_inliers(R, t, K, obj_f32, img_f32, thr):
    """PnPRansacCallback::computeError + findInliers.  Returns (count, mask)."""
    R = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(9))
    t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(3))
    K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    o, im = _f32(obj_f32, 3), _f32(img_f32, 2)
    mask = np.zeros(len(o), dtype=np.uint8)
    c = lib().orc_pnp_count_inliers(_p(R, C.c_double), _p(t, C.c_double), _p(K, C.c_double), _p(o, C.c_float),
                                    _p(im, C.c_float), len(o), thr, _p(mask, C.c_uint8))
    return int(c), mask


def pnp_sample_stream(n_points, iters, model_points=5):
    """The subsets RANSACPointSetRegistrator draws for a callback without checkSubset."""
    idx = np.zeros((iters, model_points), dtype=np.int32)
    lib().orc_pnp_sample_stream(n_points, iters, model_points, _p(idx, C.c_int32))
    return idx
