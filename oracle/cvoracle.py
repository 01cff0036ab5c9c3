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
    srcs = [os.path.join(_HERE, f) for f in ("cv_ransac_oracle.c", "cv_pnp_oracle.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
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
        L.orc_cv_norm_l2sqr.argtypes = [dp, C.c_int]
        L.orc_cv_norm_l2sqr.restype = C.c_double
        L.orc_cv_dot.argtypes = [dp, dp, C.c_int]
        L.orc_cv_dot.restype = C.c_double
        L.orc_cv_gemm_atb.argtypes = [dp, C.c_int, C.c_int, dp, dp]
        L.orc_cv_gemm_axpby.argtypes = [dp, C.c_int, dp, C.c_double, dp, C.c_double, dp]
        L.orc_cv_solve_eig.argtypes = [dp, dp, C.c_int, dp]
        L.orc_cv_invert_eig_diag.argtypes = [dp, C.c_int, dp]
        L.orc_h_refine_eval.argtypes = [dp, fp, fp, C.c_int, dp, dp, dp]
        L.orc_find_homography.argtypes = [dp, dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int, dp, u8p, ip, u8p, dp]
        L.orc_find_homography.restype = C.c_int
        L.orc_rodrigues.argtypes = [dp, dp]
        L.orc_pnp_project_f32.argtypes = [dp, dp, dp, fp, C.c_int, fp]
        L.orc_pnp_count_inliers.argtypes = [dp, dp, dp, fp, fp, C.c_int, C.c_double, u8p]
        L.orc_pnp_count_inliers.restype = C.c_int
        L.orc_pnp_sample_stream.argtypes = [C.c_int, C.c_int, C.c_int, i32p]
        L.orc_epnp.argtypes = [dp, dp, C.c_int, dp, dp, dp]
        L.orc_epnp.restype = C.c_int
        L.orc_rodrigues_inv.argtypes = [dp, dp]
        L.orc_svd.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp]
        L.orc_pnp_minimal_model.argtypes = [fp, fp, C.c_int, dp, dp, dp]
        L.orc_pnp_minimal_model.restype = C.c_int
        L.orc_pnp_ransac_stage.argtypes = [fp, fp, C.c_int, dp, C.c_int, C.c_double, C.c_double, dp, u8p, ip, ip, i32p]
        L.orc_pnp_ransac_stage.restype = C.c_int
        L.orc_pnp_refine_lm.argtypes = [dp, dp, C.c_int, dp, dp, dp, C.c_int]
        L.orc_pnp_refine_lm.restype = C.c_int
        L.orc_pnp_refine_cvlevmarq.argtypes = [dp, dp, C.c_int, dp, dp, dp]
        L.orc_pnp_refine_cvlevmarq.restype = C.c_int
        L.orc_solve_pnp_ransac.argtypes = [dp, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, dp, dp, i32p, ip, ip, dp]
        L.orc_solve_pnp_ransac.restype = C.c_int
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


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def cv_norm_l2sqr(r):
    """cv2.norm(r, cv2.NORM_L2SQR) for float64 data, in the binary's summation order."""
    r = _f64(r).ravel()
    return lib().orc_cv_norm_l2sqr(_p(r, C.c_double), r.size)


def cv_dot(a, b):
    """cv::Mat::dot of two float64 vectors (not exposed by the Python binding; pinned through the LM)."""
    a, b = _f64(a).ravel(), _f64(b).ravel()
    return lib().orc_cv_dot(_p(a, C.c_double), _p(b, C.c_double), a.size)


def cv_gemm_atb(J, r):
    """cv2.gemm(J, r, 1, None, 0, flags=cv2.GEMM_1_T) for a float64 (rows x cols) J and (rows x 1) r."""
    J, r = _f64(J), _f64(r).ravel()
    v = np.zeros(J.shape[1])
    lib().orc_cv_gemm_atb(_p(J, C.c_double), J.shape[0], J.shape[1], _p(r, C.c_double), _p(v, C.c_double))
    return v


def cv_gemm_axpby(A, d, alpha, c, beta):
    """cv2.gemm(A, d, alpha, c, beta) for square float64 A and column vectors d, c."""
    A, d, c = _f64(A), _f64(d).ravel(), _f64(c).ravel()
    out = np.zeros(A.shape[0])
    lib().orc_cv_gemm_axpby(_p(A, C.c_double), A.shape[0], _p(d, C.c_double), alpha, _p(c, C.c_double), beta,
                            _p(out, C.c_double))
    return out


def cv_solve_eig(A, b):
    """cv2.solve(A, b, flags=cv2.DECOMP_EIG) for symmetric float64 A (n <= 12)."""
    A, b = _f64(A), _f64(b).ravel()
    x = np.zeros(A.shape[0])
    lib().orc_cv_solve_eig(_p(A, C.c_double), _p(b, C.c_double), A.shape[0], _p(x, C.c_double))
    return x


def cv_invert_eig_diag(A):
    """np.diag(cv2.invert(A, flags=cv2.DECOMP_EIG)[1]) for symmetric float64 A."""
    A = _f64(A)
    d = np.zeros(A.shape[0])
    lib().orc_cv_invert_eig_diag(_p(A, C.c_double), A.shape[0], _p(d, C.c_double))
    return d


def h_refine_eval(h, src_inl, dst_inl):
    """HomographyRefineCallback::compute + the LM's A = J^T J, v = J^T r: returns (r, A, v)."""
    s, d = _f32(src_inl, 2), _f32(dst_inl, 2)
    h = _f64(h).ravel()
    r, A, v = np.zeros(2 * len(s)), np.zeros((9, 9)), np.zeros(9)
    lib().orc_h_refine_eval(_p(h, C.c_double), _p(s, C.c_float), _p(d, C.c_float), len(s), _p(r, C.c_double),
                            _p(A, C.c_double), _p(v, C.c_double))
    return r, A, v


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


def epnp(obj, img, K):
    """EPnP on n (4..16) correspondences as OpenCV runs it for a RANSAC minimal sample: (R, t) or None."""
    o = np.ascontiguousarray(np.asarray(obj, dtype=np.float64).reshape(-1, 3))
    im = np.ascontiguousarray(np.asarray(img, dtype=np.float64).reshape(-1, 2))
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    R, t = np.zeros(9), np.zeros(3)
    ok = lib().orc_epnp(_p(o, C.c_double), _p(im, C.c_double), len(o), _p(Kd, C.c_double), _p(R, C.c_double), _p(t, C.c_double))
    return (R.reshape(3, 3), t) if ok else None


def rodrigues_inv(R):
    Rd = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(9))
    r = np.zeros(3)
    lib().orc_rodrigues_inv(_p(Rd, C.c_double), _p(r, C.c_double))
    return r


def svd(A):
    """cv::SVD::compute for a small m x n (m >= n) matrix: (w (n,), u (m,n), vt (n,n)) — OpenCV's one-sided Jacobi."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
    m, n = A.shape
    w, Ut, Vt = np.zeros(n), np.zeros((n, m)), np.zeros((n, n))
    lib().orc_svd(_p(A, C.c_double), m, n, _p(w, C.c_double), _p(Ut, C.c_double), _p(Vt, C.c_double))
    return w, Ut.T.copy(), Vt


def pnp_minimal_model(obj5_f32, img5_f32, K):
    """[rvec | tvec] of one minimal sample as PnPRansacCallback::runKernel produces it (EPnP + Rodrigues), or None."""
    o, im = _f32(obj5_f32, 3), _f32(img5_f32, 2)
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    r, t = np.zeros(3), np.zeros(3)
    ok = lib().orc_pnp_minimal_model(_p(o, C.c_float), _p(im, C.c_float), len(o), _p(Kd, C.c_double), _p(r, C.c_double), _p(t, C.c_double))
    return (r, t) if ok else None


def pnp_ransac_stage(obj_f32, img_f32, K, max_iters=5000, thr=30.0, confidence=0.99):
    o, im = _f32(obj_f32, 3), _f32(img_f32, 2)
    n = len(o)
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    model = np.zeros(6)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    iters, best = C.c_int(0), C.c_int(-1)
    trace = np.full(max(max_iters, 1), -2, dtype=np.int32)
    ok = lib().orc_pnp_ransac_stage(_p(o, C.c_float), _p(im, C.c_float), n, _p(Kd, C.c_double), max_iters, thr, confidence,
                                    _p(model, C.c_double), _p(mask, C.c_uint8), C.byref(iters), C.byref(best), _p(trace, C.c_int32))
    return dict(ok=bool(ok), rvec=model[:3].copy(), tvec=model[3:].copy(), mask=mask[:n].copy(), iters=iters.value,
                best_iter=best.value, counts=trace[:iters.value].copy())


def pnp_refine_lm(obj, img, K, rvec, tvec, max_iters=20):
    """cv2.solvePnPRefineLM(obj, img, K, 0, rvec, tvec) restated (main_v1.py:508): classic LMSolver, 20 iterations."""
    o = np.ascontiguousarray(np.asarray(obj, dtype=np.float64).reshape(-1, 3))
    im = np.ascontiguousarray(np.asarray(img, dtype=np.float64).reshape(-1, 2))
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    r = np.ascontiguousarray(np.asarray(rvec, dtype=np.float64).reshape(3)).copy()
    t = np.ascontiguousarray(np.asarray(tvec, dtype=np.float64).reshape(3)).copy()
    lib().orc_pnp_refine_lm(_p(o, C.c_double), _p(im, C.c_double), len(o), _p(Kd, C.c_double), _p(r, C.c_double), _p(t, C.c_double), max_iters)
    return r, t


def solve_pnp_ransac(obj, img, K, iterations_count=100, reprojection_error=8.0, confidence=0.99, details=False):
    """cv2.solvePnPRansac(obj, img, K, zeros, iterationsCount=..., reprojectionError=..., confidence=...) restated.

    Reference call site /root/reference/main_v1.py:497-502.  Returns (ok, rvec (3,1), tvec (3,1), inliers int32 (k,1))."""
    o = np.ascontiguousarray(np.asarray(obj, dtype=np.float64).reshape(-1, 3))
    im = np.ascontiguousarray(np.asarray(img, dtype=np.float64).reshape(-1, 2))
    n = len(o)
    Kd = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(9))
    r, t, model = np.zeros(3), np.zeros(3), np.zeros(6)
    inl = np.zeros(max(n, 1), dtype=np.int32)
    k, iters = C.c_int(0), C.c_int(0)
    ok = lib().orc_solve_pnp_ransac(_p(o, C.c_double), _p(im, C.c_double), n, _p(Kd, C.c_double), iterations_count,
                                    reprojection_error, confidence, _p(r, C.c_double), _p(t, C.c_double), _p(inl, C.c_int32),
                                    C.byref(k), C.byref(iters), _p(model, C.c_double))
    out = (bool(ok), r.reshape(3, 1), t.reshape(3, 1), inl[:k.value].reshape(-1, 1).copy() if ok else None)
    if details:
        return out + (dict(iters=iters.value, ransac_rvec=model[:3].copy(), ransac_tvec=model[3:].copy()),)
    return out
