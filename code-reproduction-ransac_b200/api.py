"""Host-side API over the C ABI: one `Context` per GPU, NumPy in / NumPy out.

Mirrors the call the reference makes, `cv2.findHomography(src, dst, cv2.RANSAC, thr)` (main_v1.py:312), plus the
batched form of the loop around it (`find_homographies`, main_v1.py:254-297) and the individual kernels for the
parity tests.  All arithmetic runs in libransac_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import HInfo, HParams, PInfo, PParams

SAMPLER_CV_REPLAY, SAMPLER_PHILOX = 0, 1
ARITH_EXACT, ARITH_FAST, ARITH_EXACT_UNFILTERED = 0, 1, 2
MASK_CV413, MASK_LEGACY = 0, 1
SOLVER_EXACT, SOLVER_FAST, SOLVER_EXACT_WARP = 0, 1, 2
REFINE_NONE, REFINE_CV, REFINE_PARALLEL = 0, 1, 2
OK, NO_MODEL = 0, 1


class RansacB200Error(RuntimeError):
    pass


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _f64(a, cols):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, cols))


def _f32(a, cols):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols))


def make_params(thr, max_iters=2000, confidence=0.995, sampler=SAMPLER_CV_REPLAY, seed=0, arith=ARITH_EXACT,
                mask_semantics=MASK_CV413, refine=True, hyp_begin=0, solver=SOLVER_EXACT):
    p = HParams()
    p.thr = float(thr)
    p.max_iters = int(max_iters)
    p.confidence = float(confidence)
    p.sampler = int(sampler)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    p.arith = int(arith)
    p.mask_semantics = int(mask_semantics)
    p.refine = int(refine)   # True/1: as cv2 (REFINE_CV); 2: REFINE_PARALLEL; False/0: none
    p.hyp_begin = int(hyp_begin)
    p.solver = int(solver)
    return p


def _info_dict(i):
    return dict(status=i.status, iters_run=i.iters_run, best_iter=i.best_iter, best_count=i.best_count,
                sample=[int(x) for x in i.sample], n_inliers=i.n_inliers, lm_iters=i.lm_iters,
                winner_id=int(i.reserved[0]))   # PHILOX: global hypothesis id of the winner (low 31 bits), on every rank


class _InfoSeq:
    """The per-problem info records of a batched call (b2r_h_info / b2r_p_info array), read-only.  Behaves like the list of
    dicts it replaces (len, indexing, iteration) but builds a dict only when one is asked for: a sweep of hundreds of
    candidates spent more host time building dicts than the GPU spent on RANSAC.  `.status` is the vector of all statuses."""

    def __init__(self, arr, to_dict):
        self._arr, self._to_dict = arr, to_dict
        self.status = np.frombuffer(arr, dtype=np.int32).reshape(len(arr), -1)[:, 0].copy() if len(arr) else np.zeros(0, np.int32)

    def __len__(self):
        return len(self._arr)

    def __getitem__(self, q):
        if isinstance(q, slice):
            return [self._to_dict(self._arr[i]) for i in range(*q.indices(len(self._arr)))]
        return self._to_dict(self._arr[q])

    def __iter__(self):
        return (self._to_dict(i) for i in self._arr)


class HomographyProblem:
    """Q homography-RANSAC problems of n correspondences each, resident in HBM (b2r_h_problem)."""

    def __init__(self, ctx, src, dst, dst_shared=None):
        src = np.asarray(src, dtype=np.float64)
        dst = np.asarray(dst, dtype=np.float64)
        if src.ndim == 2:
            src = src[None]
        if dst_shared is None:
            dst_shared = dst.ndim == 2
        if not dst_shared and dst.ndim == 2:
            dst = dst[None]
        self.Q, self.n = int(src.shape[0]), int(src.shape[1])
        src = np.ascontiguousarray(src)
        dst = np.ascontiguousarray(dst)
        if dst.shape[-2] != self.n or src.shape[-1] != 2 or dst.shape[-1] != 2:
            raise ValueError("src must be (Q,n,2) and dst (n,2) or (Q,n,2)")
        self.ctx = ctx
        self.h2d_bytes = src.nbytes + dst.nbytes
        self._h = ctx._L.b2r_h_problem_upload(ctx._c, _ptr(src, C.c_double), _ptr(dst, C.c_double), 1 if dst_shared else 0,
                                              self.Q, self.n)
        if not self._h:
            raise RansacB200Error(_lib.last_error())

    def reupload(self, src, dst):
        """Replace the points (host arrays, same layout rules as the constructor), reusing the device buffers."""
        src = np.ascontiguousarray(np.asarray(src, dtype=np.float64))
        dst = np.ascontiguousarray(np.asarray(dst, dtype=np.float64))
        if src.ndim == 2:
            src = src[None]
        shared = dst.ndim == 2
        self.Q, self.n = int(src.shape[0]), int(src.shape[1])
        self.h2d_bytes = src.nbytes + dst.nbytes
        self.ctx._check(self.ctx._L.b2r_h_problem_reupload(self.ctx._c, self._h, _ptr(src, C.c_double), _ptr(dst, C.c_double),
                                                           1 if shared else 0, self.Q, self.n))

    def run(self, params):
        self.ctx._check(self.ctx._L.b2r_h_problem_run(self.ctx._c, self._h, C.byref(params)))

    def score_shard(self, params):
        """Stage 1 on this rank's hypothesis-id shard; returns the packed (count, id) keys, uint64 (Q,)."""
        keys = np.zeros(self.Q, dtype=np.uint64)
        self.ctx._check(self.ctx._L.b2r_h_problem_score_shard(self.ctx._c, self._h, C.byref(params), _ptr(keys, C.c_uint64)))
        return keys

    def finish(self, params, keys):
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(self.Q))
        self.ctx._check(self.ctx._L.b2r_h_problem_finish(self.ctx._c, self._h, C.byref(params), _ptr(keys, C.c_uint64)))

    def score_shard_dev(self, params, keys_dev_ptr):
        """Stage 1 with the packed keys written to DEVICE memory (Q uint64 at `keys_dev_ptr`), asynchronous on the context's stream."""
        self.ctx._check(self.ctx._L.b2r_h_problem_score_shard_dev(self.ctx._c, self._h, C.byref(params), C.c_void_p(int(keys_dev_ptr))))

    def finish_dev(self, params, keys_dev_ptr):
        self.ctx._check(self.ctx._L.b2r_h_problem_finish_dev(self.ctx._c, self._h, C.byref(params), C.c_void_p(int(keys_dev_ptr))))

    def fetch(self, want_mask=True):
        H = np.zeros((self.Q, 3, 3))
        mask = np.zeros((self.Q, self.n), dtype=np.uint8) if want_mask else None
        info = (HInfo * self.Q)()
        self.ctx._check(self.ctx._L.b2r_h_problem_fetch(self.ctx._c, self._h, _ptr(H, C.c_double),
                                                        _ptr(mask, C.c_uint8) if want_mask else None, info))
        return H, mask, _InfoSeq(info, _info_dict)

    def peek_hyps(self, first, count, q=0):
        """Samples (count,4), fp32 models (count,8) and inlier counts (count) of hypothesis slots [first, first+count) of
        problem q as the last run left them on the device (parity tests / diagnostics)."""
        smp = np.zeros((count, 4), dtype=np.int32)
        mdl = np.zeros((count, 8), dtype=np.float32)
        cnt = np.zeros(count, dtype=np.int32)
        self.ctx._check(self.ctx._L.b2r_h_problem_peek_hyps(self.ctx._c, self._h, int(q), int(first), int(count),
                                                            _ptr(smp, C.c_int32), _ptr(mdl, C.c_float), _ptr(cnt, C.c_int32)))
        return smp, mdl, cnt

    def stage_ms(self):
        ms = (C.c_float * 5)()
        self.ctx._check(self.ctx._L.b2r_h_problem_stage_ms(self.ctx._c, self._h, ms))
        return dict(sample_solve=ms[0], score=ms[1], select=ms[2], finalize=ms[3], total=ms[4])

    def free(self):
        if self._h:
            self.ctx._L.b2r_h_problem_free(self.ctx._c, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_p_params(thr=8.0, max_iters=100, confidence=0.99, sampler=SAMPLER_CV_REPLAY, seed=0, arith=ARITH_EXACT,
                  refine=True, hyp_begin=0, solver=SOLVER_EXACT):
    """b2r_p_params with cv2.solvePnPRansac's defaults (iterationsCount=100, reprojectionError=8.0, confidence=0.99)."""
    p = PParams()
    p.thr = float(thr)
    p.max_iters = int(max_iters)
    p.confidence = float(confidence)
    p.sampler = int(sampler)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    p.arith = int(arith)
    p.refine = 1 if refine else 0
    p.hyp_begin = int(hyp_begin)
    p.solver = int(solver)
    return p


def _p_info_dict(i):
    return dict(status=i.status, iters_run=i.iters_run, best_iter=i.best_iter, best_count=i.best_count,
                sample=[int(x) for x in i.sample], n_inliers=i.n_inliers, lm_iters=i.lm_iters,
                ransac_rvec=np.array(i.ransac_rvec[:]), ransac_tvec=np.array(i.ransac_tvec[:]),
                mean_inlier_err=float(i.mean_inlier_err), sum_sq_err=float(i.sum_sq_err))


def _K9(K, Q=None):
    K = np.ascontiguousarray(np.asarray(K, dtype=np.float64))
    if Q is None:
        return K.reshape(3, 3)
    return np.ascontiguousarray(K.reshape(Q, 3, 3))


class PnPProblem:
    """Q PnP-RANSAC problems (one camera matrix each) over n correspondences resident in HBM (b2r_p_problem).
    obj (n,3) / img (n,2): the Q problems share the points (testpro-K.py's intrinsics grid); (Q,n,3)/(Q,n,2): own points."""

    def __init__(self, ctx, obj, img, K):
        obj = np.ascontiguousarray(np.asarray(obj, dtype=np.float64))
        img = np.ascontiguousarray(np.asarray(img, dtype=np.float64))
        K = np.asarray(K, dtype=np.float64)
        self.Q = 1 if K.ndim == 2 else int(K.shape[0])
        self.shared = obj.ndim == 2
        self.n = int(obj.shape[-2])
        if obj.shape[-1] != 3 or img.shape[-1] != 2 or img.shape[-2] != self.n or (not self.shared and obj.shape[0] != self.Q):
            raise ValueError("obj must be (n,3) or (Q,n,3), img (n,2) or (Q,n,2), K (3,3) or (Q,3,3)")
        self.ctx = ctx
        self.h2d_bytes = obj.nbytes + img.nbytes + 72 * self.Q
        Kq = _K9(K, self.Q)
        self._h = ctx._L.b2r_p_problem_upload(ctx._c, _ptr(obj, C.c_double), _ptr(img, C.c_double), 1 if self.shared else 0,
                                              self.Q, self.n, _ptr(Kq, C.c_double))
        if not self._h:
            raise RansacB200Error(_lib.last_error())

    def reupload(self, obj, img, K):
        obj = np.ascontiguousarray(np.asarray(obj, dtype=np.float64))
        img = np.ascontiguousarray(np.asarray(img, dtype=np.float64))
        K = np.asarray(K, dtype=np.float64)
        self.Q = 1 if K.ndim == 2 else int(K.shape[0])
        self.shared = obj.ndim == 2
        self.n = int(obj.shape[-2])
        Kq = _K9(K, self.Q)
        self.h2d_bytes = obj.nbytes + img.nbytes + 72 * self.Q
        self.ctx._check(self.ctx._L.b2r_p_problem_reupload(self.ctx._c, self._h, _ptr(obj, C.c_double), _ptr(img, C.c_double),
                                                           1 if self.shared else 0, self.Q, self.n, _ptr(Kq, C.c_double)))

    def run(self, params):
        self.ctx._check(self.ctx._L.b2r_p_problem_run(self.ctx._c, self._h, C.byref(params)))

    def score_shard(self, params):
        keys = np.zeros(self.Q, dtype=np.uint64)
        self.ctx._check(self.ctx._L.b2r_p_problem_score_shard(self.ctx._c, self._h, C.byref(params), _ptr(keys, C.c_uint64)))
        return keys

    def finish(self, params, keys):
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint64).reshape(self.Q))
        self.ctx._check(self.ctx._L.b2r_p_problem_finish(self.ctx._c, self._h, C.byref(params), _ptr(keys, C.c_uint64)))

    def score_shard_dev(self, params, keys_dev_ptr):
        self.ctx._check(self.ctx._L.b2r_p_problem_score_shard_dev(self.ctx._c, self._h, C.byref(params), C.c_void_p(int(keys_dev_ptr))))

    def finish_dev(self, params, keys_dev_ptr):
        self.ctx._check(self.ctx._L.b2r_p_problem_finish_dev(self.ctx._c, self._h, C.byref(params), C.c_void_p(int(keys_dev_ptr))))

    def fetch(self, want_inliers=True):
        """(rvec (Q,3), tvec (Q,3), inliers list of int32 arrays or None, infos)"""
        rvec, tvec = np.zeros((self.Q, 3)), np.zeros((self.Q, 3))
        inl = np.zeros((self.Q, self.n), dtype=np.int32) if want_inliers else None
        ninl = np.zeros(self.Q, dtype=np.int32)
        info = (PInfo * self.Q)()
        self.ctx._check(self.ctx._L.b2r_p_problem_fetch(self.ctx._c, self._h, _ptr(rvec, C.c_double), _ptr(tvec, C.c_double),
                                                        _ptr(inl, C.c_int32) if want_inliers else None, _ptr(ninl, C.c_int32), info))
        lists = [inl[q, :ninl[q]].copy() for q in range(self.Q)] if want_inliers else None
        return rvec, tvec, lists, _InfoSeq(info, _p_info_dict)

    def stage_ms(self):
        ms = (C.c_float * 5)()
        self.ctx._check(self.ctx._L.b2r_p_problem_stage_ms(self.ctx._c, self._h, ms))
        return dict(sample_solve=ms[0], score=ms[1], select=ms[2], finalize=ms[3], total=ms[4])

    def free(self):
        if self._h:
            self.ctx._L.b2r_p_problem_free(self.ctx._c, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Dem:
    """A DEM grid resident in HBM (b2r_dem)."""

    def __init__(self, ctx, grid_y, grid_x, values):
        gy = np.asarray(grid_y, dtype=np.float64).ravel()
        gx = np.asarray(grid_x, dtype=np.float64).ravel()
        v = np.asarray(values, dtype=np.float64)
        if v.shape != (len(gy), len(gx)):
            raise ValueError("values must be (len(grid_y), len(grid_x))")
        if len(gy) > 1 and gy[1] < gy[0]:      # scipy's RegularGridInterpolator flips descending axes (and the values)
            gy, v = gy[::-1], v[::-1, :]
        if len(gx) > 1 and gx[1] < gx[0]:
            gx, v = gx[::-1], v[:, ::-1]
        gy, gx, v = np.ascontiguousarray(gy), np.ascontiguousarray(gx), np.ascontiguousarray(v)
        self.ctx, self.shape = ctx, v.shape
        self._h = ctx._L.b2r_dem_upload(ctx._c, _ptr(gy, C.c_double), len(gy), _ptr(gx, C.c_double), len(gx), _ptr(v, C.c_double))
        if not self._h:
            raise RansacB200Error(_lib.last_error())

    def free(self):
        if self._h:
            self.ctx._L.b2r_dem_free(self.ctx._c, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU context (own CUDA stream and workspaces).  Not thread-safe."""

    def __init__(self, device=0):
        self._L = _lib.load()
        self._c = self._L.b2r_ctx_create(int(device))
        if not self._c:
            raise RansacB200Error(_lib.last_error())
        self.device = int(device)

    def close(self):
        if self._c:
            self._L.b2r_ctx_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RansacB200Error(f"ransac_b200 error {rc}: {_lib.last_error()}")
        return rc

    @property
    def stream(self):
        """cudaStream_t of this context as an int (usable with torch.cuda.ExternalStream)."""
        return int(self._L.b2r_ctx_stream(self._c) or 0)

    def synchronize(self):
        self._check(self._L.b2r_ctx_synchronize(self._c))

    def launch_count(self):
        return int(self._L.b2r_ctx_launch_count(self._c))

    # ---- the reference's call ------------------------------------------------------------------------
    def find_homography(self, src, dst, thr, max_iters=2000, confidence=0.995, **kw):
        """cv2.findHomography(src, dst, cv2.RANSAC, thr, maxIters=..., confidence=...) on the GPU.

        Returns (H (3,3) float64 or None, mask (n,1) uint8, info dict) — cv2's convention plus diagnostics."""
        s, d = _f64(src, 2), _f64(dst, 2)
        if len(s) != len(d):
            raise ValueError("src and dst must have the same number of points")
        n = len(s)
        p = make_params(thr, max_iters, confidence, **kw)
        H = np.zeros(9)
        mask = np.zeros(max(n, 1), dtype=np.uint8)
        info = HInfo()
        rc = self._check(self._L.b2r_find_homography(self._c, _ptr(s, C.c_double), _ptr(d, C.c_double), n, C.byref(p),
                                                     _ptr(H, C.c_double), _ptr(mask, C.c_uint8), C.byref(info)))
        return (H.reshape(3, 3) if rc == OK else None), mask[:n].reshape(n, 1), _info_dict(info)

    def find_homography_batch(self, src, dst, thr, max_iters=2000, confidence=0.995, **kw):
        """Q independent problems in one call: src (Q,n,2); dst (n,2) shared or (Q,n,2).

        Returns (H (Q,3,3), ok (Q,) bool, mask (Q,n) uint8, infos list)."""
        src = np.ascontiguousarray(np.asarray(src, dtype=np.float64))
        dst = np.ascontiguousarray(np.asarray(dst, dtype=np.float64))
        Q, n = src.shape[0], src.shape[1]
        shared = dst.ndim == 2
        p = make_params(thr, max_iters, confidence, **kw)
        H = np.zeros((Q, 3, 3))
        mask = np.zeros((Q, n), dtype=np.uint8)
        info = (HInfo * Q)()
        self._check(self._L.b2r_find_homography_batch(self._c, _ptr(src, C.c_double), _ptr(dst, C.c_double), 1 if shared else 0,
                                                      Q, n, C.byref(p), _ptr(H, C.c_double), _ptr(mask, C.c_uint8), info))
        infos = _InfoSeq(info, _info_dict)
        return H, infos.status == OK, mask, infos

    def upload(self, src, dst, dst_shared=None):
        return HomographyProblem(self, src, dst, dst_shared)

    def camera_sweep(self, pos3d, pixels, cams, thr, max_iters=2000, confidence=0.995, **kw):
        """find_homographies + arg-min (main_v1.py:254-297, :863-866) fused on the device: projection of the landmarks per
        candidate camera, RANSAC, err1/err2, arg-min.  Returns dict(scores (Q,2), M (Q,3,3), H (Q,3,3), mask (Q,n), infos, best)."""
        pos3d, pixels, cams = _f64(pos3d, 3), _f64(pixels, 2), _f64(cams, 3)
        n, Q = len(pos3d), len(cams)
        p = make_params(thr, max_iters, confidence, **kw)
        scores, M, H = np.zeros((Q, 2)), np.zeros((Q, 3, 3)), np.zeros((Q, 3, 3))
        mask = np.zeros((Q, n), dtype=np.uint8)
        info = (HInfo * Q)()
        best = C.c_int32(0)
        self._check(self._L.b2r_camera_sweep(self._c, _ptr(pos3d, C.c_double), _ptr(pixels, C.c_double), n, _ptr(cams, C.c_double), Q,
                                             C.byref(p), _ptr(scores, C.c_double), _ptr(M, C.c_double), _ptr(H, C.c_double),
                                             _ptr(mask, C.c_uint8), info, C.byref(best)))
        return dict(scores=scores, M=M, H=H, mask=mask, infos=_InfoSeq(info, _info_dict), best=int(best.value))

    # ---- cv2.solvePnPRansac / solvePnPRefineLM ----------------------------------------------------------
    def solve_pnp_ransac(self, obj, img, K, iterations_count=100, reprojection_error=8.0, confidence=0.99, **kw):
        """cv2.solvePnPRansac(obj, img, K, zeros, iterationsCount=..., reprojectionError=..., confidence=...) on the GPU
        (main_v1.py:497-502).  Returns (ok, rvec (3,1), tvec (3,1), inliers int32 (k,1) or None, info dict)."""
        o, im = _f64(obj, 3), _f64(img, 2)
        if len(o) != len(im):
            raise ValueError("obj and img must have the same number of points")
        n = len(o)
        Kd = _K9(K)
        p = make_p_params(reprojection_error, iterations_count, confidence, **kw)
        rvec, tvec = np.zeros(3), np.zeros(3)
        inl = np.zeros(max(n, 1), dtype=np.int32)
        k = C.c_int32(0)
        info = PInfo()
        rc = self._check(self._L.b2r_solve_pnp_ransac(self._c, _ptr(o, C.c_double), _ptr(im, C.c_double), n, _ptr(Kd, C.c_double),
                                                      C.byref(p), _ptr(rvec, C.c_double), _ptr(tvec, C.c_double),
                                                      _ptr(inl, C.c_int32), C.byref(k), C.byref(info)))
        ok = rc == OK
        return ok, rvec.reshape(3, 1), tvec.reshape(3, 1), (inl[:k.value].reshape(-1, 1).copy() if ok else None), _p_info_dict(info)

    def solve_pnp_ransac_batch(self, obj, img, Ks, iterations_count=100, reprojection_error=8.0, confidence=0.99, **kw):
        """Q problems in one call (the intrinsics grid of testpro-K.py:58-97 when obj/img are 2-D and shared).
        Returns (ok (Q,), rvec (Q,3), tvec (Q,3), inliers list, infos)."""
        obj = np.ascontiguousarray(np.asarray(obj, dtype=np.float64))
        img = np.ascontiguousarray(np.asarray(img, dtype=np.float64))
        Ks = np.asarray(Ks, dtype=np.float64)
        Q, n, shared = int(Ks.shape[0]), int(obj.shape[-2]), obj.ndim == 2
        Kq = _K9(Ks, Q)
        p = make_p_params(reprojection_error, iterations_count, confidence, **kw)
        rvec, tvec = np.zeros((Q, 3)), np.zeros((Q, 3))
        inl = np.zeros((Q, n), dtype=np.int32)
        ninl = np.zeros(Q, dtype=np.int32)
        info = (PInfo * Q)()
        self._check(self._L.b2r_solve_pnp_ransac_batch(self._c, _ptr(obj, C.c_double), _ptr(img, C.c_double), 1 if shared else 0,
                                                       Q, n, _ptr(Kq, C.c_double), C.byref(p), _ptr(rvec, C.c_double),
                                                       _ptr(tvec, C.c_double), _ptr(inl, C.c_int32), _ptr(ninl, C.c_int32), info))
        infos = _InfoSeq(info, _p_info_dict)
        return infos.status == OK, rvec, tvec, [inl[q, :ninl[q]].copy() for q in range(Q)], infos

    def solve_pnp_refine_lm(self, obj, img, K, rvec, tvec, max_iters=20):
        """cv2.solvePnPRefineLM(obj, img, K, zeros, rvec, tvec) on the GPU (main_v1.py:508).  Returns (rvec (3,1), tvec (3,1), iters)."""
        o, im = _f64(obj, 3), _f64(img, 2)
        Kd = _K9(K)
        r = np.ascontiguousarray(np.asarray(rvec, dtype=np.float64).reshape(3)).copy()
        t = np.ascontiguousarray(np.asarray(tvec, dtype=np.float64).reshape(3)).copy()
        it = C.c_int32(0)
        self._check(self._L.b2r_solve_pnp_refine_lm(self._c, _ptr(o, C.c_double), _ptr(im, C.c_double), len(o), _ptr(Kd, C.c_double),
                                                    _ptr(r, C.c_double), _ptr(t, C.c_double), int(max_iters), C.byref(it)))
        return r.reshape(3, 1), t.reshape(3, 1), it.value

    def upload_pnp(self, obj, img, K):
        return PnPProblem(self, obj, img, K)

    # ---- single kernels (parity tests) -------------------------------------------------------------------
    def score_p(self, models_Rt, obj, img, K, thr_sq, arith=ARITH_EXACT):
        """K3 (PnP): inlier counts of poses models_Rt (m,12) = R row-major | t over the points (quantised to fp32 inside)."""
        m = np.ascontiguousarray(np.asarray(models_Rt, dtype=np.float64).reshape(-1, 12))
        o, im = _f64(obj, 3), _f64(img, 2)
        Kd = _K9(K)
        counts = np.zeros(len(m), dtype=np.int32)
        self._check(self._L.b2r_score_p(self._c, _ptr(m, C.c_double), len(m), _ptr(o, C.c_double), _ptr(im, C.c_double), len(o),
                                        _ptr(Kd, C.c_double), C.c_float(np.float32(thr_sq)), int(arith), _ptr(counts, C.c_int32)))
        return counts

    def pnp_minimal_models(self, obj, img, K, idx, solver=SOLVER_EXACT):
        """K2 (PnP): minimal models of the 5-point samples idx (m,5): (rvec (m,3), tvec (m,3), R (m,3,3), ok (m,)).
        solver: SOLVER_EXACT = OpenCV's EPnP restated, SOLVER_FAST = the throughput path's 5-point solver."""
        o, im = _f64(obj, 3), _f64(img, 2)
        Kd = _K9(K)
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32).reshape(-1, 5))
        m = len(idx)
        rvec, tvec, R = np.zeros((m, 3)), np.zeros((m, 3)), np.zeros((m, 3, 3))
        ok = np.zeros(m, dtype=np.uint8)
        self._check(self._L.b2r_pnp_minimal_models(self._c, _ptr(o, C.c_double), _ptr(im, C.c_double), len(o), _ptr(Kd, C.c_double),
                                                   _ptr(idx, C.c_int32), m, int(solver), _ptr(rvec, C.c_double), _ptr(tvec, C.c_double),
                                                   _ptr(R, C.c_double), _ptr(ok, C.c_uint8)))
        return rvec, tvec, R, ok.astype(bool)

    def sample_cv_p(self, n, n_iters):
        idx = np.zeros((n_iters, 5), dtype=np.int32)
        self._check(self._L.b2r_sample_cv_p(self._c, int(n), int(n_iters), _ptr(idx, C.c_int32)))
        return idx

    def score_h(self, models8, src_f32, dst_f32, thr_sq, arith=ARITH_EXACT):
        m = _f32(models8, 8)
        s, d = _f32(src_f32, 2), _f32(dst_f32, 2)
        counts = np.zeros(len(m), dtype=np.int32)
        self._check(self._L.b2r_score_h(self._c, _ptr(m, C.c_float), len(m), _ptr(s, C.c_float), _ptr(d, C.c_float), len(s),
                                        C.c_float(np.float32(thr_sq)), int(arith), _ptr(counts, C.c_int32)))
        return counts

    def solve_h4(self, src_f32, dst_f32, idx, solver=SOLVER_EXACT):
        s, d = _f32(src_f32, 2), _f32(dst_f32, 2)
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32).reshape(-1, 4))
        k = len(idx)
        H = np.zeros((k, 3, 3))
        ok = np.zeros(k, dtype=np.uint8)
        sub = np.zeros(k, dtype=np.uint8)
        self._check(self._L.b2r_solve_h4(self._c, _ptr(s, C.c_float), _ptr(d, C.c_float), len(s), _ptr(idx, C.c_int32), k,
                                         int(solver), _ptr(H, C.c_double), _ptr(ok, C.c_uint8), _ptr(sub, C.c_uint8)))
        return H, ok.astype(bool), sub.astype(bool)

    def sample_cv(self, src_f32, dst_f32, n_iters):
        s, d = _f32(src_f32, 2), _f32(dst_f32, 2)
        idx = np.full((n_iters, 4), -1, dtype=np.int32)
        gen = C.c_int32(0)
        self._check(self._L.b2r_sample_cv(self._c, _ptr(s, C.c_float), _ptr(d, C.c_float), len(s), n_iters,
                                          _ptr(idx, C.c_int32), C.byref(gen)))
        return idx[:gen.value]

    def sample_philox(self, src_f32, dst_f32, seed, hyp_begin, n_hyp):
        s, d = _f32(src_f32, 2), _f32(dst_f32, 2)
        idx = np.full((n_hyp, 4), -1, dtype=np.int32)
        self._check(self._L.b2r_sample_philox(self._c, _ptr(s, C.c_float), _ptr(d, C.c_float), len(s),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, 0, int(hyp_begin), n_hyp, _ptr(idx, C.c_int32)))
        return idx

    def refine_h(self, src_f32, dst_f32, mask, H0):
        s, d = _f32(src_f32, 2), _f32(dst_f32, 2)
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8).reshape(-1))
        H = np.ascontiguousarray(np.asarray(H0, dtype=np.float64).reshape(9)).copy()
        it = C.c_int32(0)
        self._check(self._L.b2r_refine_h(self._c, _ptr(s, C.c_float), _ptr(d, C.c_float), len(s), _ptr(m, C.c_uint8),
                                         _ptr(H, C.c_double), C.byref(it)))
        return H.reshape(3, 3), it.value

    # ---- row f4: the DEM ray-march (main_v1.py:635-684, :765-785) --------------------------------------------------
    def upload_dem(self, grid_y, grid_x, values):
        """DEM resident in HBM: the grid of RegularGridInterpolator((dem_y, dem_x), dem_array), main_v1.py:454.
        Descending axes are flipped (with the values), as scipy does at construction."""
        return Dem(self, grid_y, grid_x, values)

    def ray_march_dem(self, dem, origins, dirs, max_search_dist=10000, step=1, min_steps=150):
        """ray_intersect_dem (main_v1.py:635-658) for m rays in one launch.  origins (3,) or (m,3); dirs (m,3).
        Returns (geo (m,3), hit_step (m,), status (m,): 0 hit, 1 no intersection, 2 left the DEM)."""
        from . import geo
        o = np.ascontiguousarray(np.asarray(origins, dtype=np.float64))
        d = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).reshape(-1, 3))
        m = len(d)
        shared = o.ndim == 1
        if not shared and o.shape != (m, 3):
            raise ValueError("origins must be (3,) or (m,3)")
        out, hit, st = np.zeros((m, 3)), np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
        u = geo.utm_series_constants()
        self._check(self._L.b2r_ray_march_dem(self._c, dem._h, _ptr(o, C.c_double), 1 if shared else 0, _ptr(d, C.c_double), m,
                                              _ptr(u, C.c_double), float(max_search_dist), float(step), int(min_steps),
                                              _ptr(out, C.c_double), _ptr(hit, C.c_int32), _ptr(st, C.c_int32)))
        return out, hit, st

    def pixels_to_geo(self, dem, pixels, Kinv, R, ray_origin, ctrl_pixels, ctrl_factors, max_search_dist=10000, step=1,
                      min_steps=150):
        """pixel_to_geo (main_v1.py:661-684) for m pixels in one launch sequence.  Returns (geo, hit_step, status, dirs)."""
        from . import geo
        px = np.ascontiguousarray(np.asarray(pixels, dtype=np.float64).reshape(-1, 2))
        cp = np.ascontiguousarray(np.asarray(ctrl_pixels, dtype=np.float64).reshape(-1, 2))
        cf = np.ascontiguousarray(np.asarray(ctrl_factors, dtype=np.float64).reshape(-1, 3))
        if len(cp) != len(cf):
            raise ValueError("every control point needs its optimisation factors (np.average raises in the reference otherwise)")
        Ki = np.ascontiguousarray(np.asarray(Kinv, dtype=np.float64).reshape(3, 3))
        Rm = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(3, 3))
        o = np.ascontiguousarray(np.asarray(ray_origin, dtype=np.float64).reshape(3))
        m = len(px)
        out, hit, st, dirs = np.zeros((m, 3)), np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32), np.zeros((m, 3))
        u = geo.utm_series_constants()
        self._check(self._L.b2r_pixels_to_geo(self._c, dem._h, _ptr(px, C.c_double), m, _ptr(Ki, C.c_double), _ptr(Rm, C.c_double),
                                              _ptr(o, C.c_double), _ptr(cp, C.c_double), _ptr(cf, C.c_double), len(cp),
                                              _ptr(u, C.c_double), float(max_search_dist), float(step), int(min_steps),
                                              _ptr(out, C.c_double), _ptr(hit, C.c_int32), _ptr(st, C.c_int32), _ptr(dirs, C.c_double)))
        return out, hit, st, dirs

    def jacobi_eig(self, A, form=1):
        """Eigenvalues (descending) and eigenvectors (rows) of symmetric matrices A (n_mat, n, n), 2 <= n <= 9, by the
        eigen-solver of the exact-mode kernels (form 0 thread / 1 warp / 2 packed shared memory, n = 9)."""
        A = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
        if A.ndim == 2:
            A = A[None]
        n_mat, n = A.shape[0], A.shape[1]
        W, V = np.zeros((n_mat, n)), np.zeros((n_mat, n, n))
        self._check(self._L.b2r_jacobi_eig(self._c, _ptr(A, C.c_double), n_mat, n, int(form), _ptr(W, C.c_double), _ptr(V, C.c_double)))
        return W, V

    def selftest_rcp(self):
        bad, tested = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.b2r_selftest_rcp(self._c, C.byref(bad), C.byref(tested)))
        return int(bad.value), int(tested.value)

    def probe_fp32_peak(self):
        """(scalar FFMA, packed FFMA2) fp32 FMA lane-operations per second, register resident."""
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._L.b2r_probe_fp32_peak(self._c, C.byref(a), C.byref(b)))
        return a.value, b.value


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
