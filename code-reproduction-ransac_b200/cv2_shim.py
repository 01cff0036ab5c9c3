"""A `cv2`-shaped module for the reference scripts: the hot-path entry points run on the B200 through
libransac_b200.so, every other attribute is forwarded to the real OpenCV (imread, Rodrigues, projectPoints, ...).

    import ransac_b200.cv2_shim as shim
    main_v1.cv2 = shim.module()            # or: sys.modules["cv2"] = shim.module() before importing the script

Replaces exactly the lookups SURVEY.md §8(b) lists: cv2.findHomography with method == cv2.RANSAC
(main_v1.py:312, process.py:200, testpro.py:350, test_pro.py:351, test02.py:263), cv2.solvePnPRansac (main_v1.py:497,
testpro.py:536, test_pro.py:515, testpro-K.py:72), cv2.solvePnPRefineLM (main_v1.py:508, testpro-K.py:122) and the
constant cv2.RANSAC.  Return conventions are cv2's: (H float64 (3,3) or None, mask uint8 (n,1)), fewer than 4 points
raise; (retval bool, rvec (3,1), tvec (3,1), inliers int32 (k,1))."""
import types

import numpy as np

from . import api

RANSAC = 8


class error(Exception):
    """Stands in for cv2.error when real OpenCV is not importable."""


def _real_cv2():
    try:
        import cv2
        return cv2
    except ImportError:
        return None


def findHomography(srcPoints, dstPoints, method=0, ransacReprojThreshold=3.0, mask=None, maxIters=2000,
                   confidence=0.995, _ctx=None, **b2r_kw):
    """cv2.findHomography.  method == cv2.RANSAC runs on the GPU; other methods are not on the reference's hot
    path and are forwarded to OpenCV."""
    if method != RANSAC:
        cv2 = _real_cv2()
        if cv2 is None:
            raise error("findHomography: only method=cv2.RANSAC is implemented by ransac_b200")
        return cv2.findHomography(srcPoints, dstPoints, method, ransacReprojThreshold, mask, maxIters, confidence)
    src = np.asarray(srcPoints, dtype=np.float64).reshape(-1, 2)
    dst = np.asarray(dstPoints, dtype=np.float64).reshape(-1, 2)
    if len(src) != len(dst) or len(src) < 4:
        cv2 = _real_cv2()
        exc = cv2.error if cv2 is not None else error
        raise exc("findHomography: need >= 4 corresponding points (OpenCV asserts in fundam.cpp)")
    ctx = _ctx or api.default_context()
    thr = float(ransacReprojThreshold)
    if not thr > 0:      # OpenCV substitutes its default (fundam.cpp: `if (ransacReprojThreshold <= 0) ... = 3`);
        thr = 3.0        # probed on the 4.13.0 binary: thr = 0, -1, -5.5 all return the thr = 3 result
    H, m, _ = ctx.find_homography(src, dst, thr, int(maxIters), float(confidence), **b2r_kw)
    return H, m


def _zero_distortion(distCoeffs):
    return distCoeffs is None or not np.any(np.asarray(distCoeffs, dtype=np.float64))


def solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec=None, tvec=None, useExtrinsicGuess=False,
                   iterationsCount=100, reprojectionError=8.0, confidence=0.99, inliers=None, flags=0, _ctx=None, **b2r_kw):
    """cv2.solvePnPRansac as the reference calls it (zero distortion, default flags = SOLVEPNP_ITERATIVE whose RANSAC
    kernel is EPnP, no extrinsic guess): runs on the GPU.  Anything else is not on the reference's hot path and is
    forwarded to OpenCV."""
    obj = np.asarray(objectPoints, dtype=np.float64).reshape(-1, 3)
    img = np.asarray(imagePoints, dtype=np.float64).reshape(-1, 2)
    if useExtrinsicGuess or flags != 0 or not _zero_distortion(distCoeffs) or len(obj) == 4:
        cv2 = _real_cv2()
        if cv2 is None:
            raise error("solvePnPRansac: only the reference's configuration is implemented by ransac_b200")
        return cv2.solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec, tvec, useExtrinsicGuess,
                                  iterationsCount, reprojectionError, confidence, inliers, flags)
    if len(obj) != len(img) or len(obj) < 4:
        cv2 = _real_cv2()
        exc = cv2.error if cv2 is not None else error
        raise exc("solvePnPRansac: need >= 4 corresponding points (OpenCV asserts in solvepnp.cpp)")
    ctx = _ctx or api.default_context()
    ok, r, t, inl, _ = ctx.solve_pnp_ransac(obj, img, cameraMatrix, int(iterationsCount), float(reprojectionError),
                                            float(confidence), **b2r_kw)
    if not ok:
        return False, np.zeros((3, 1)), np.zeros((3, 1)), None
    return True, r, t, inl


def solvePnPRefineLM(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec, tvec, criteria=None, _ctx=None):
    """cv2.solvePnPRefineLM (default criteria: 20 iterations, eps FLT_EPSILON) on the GPU; returns (rvec, tvec) (3,1)."""
    if not _zero_distortion(distCoeffs) or criteria is not None:
        cv2 = _real_cv2()
        if cv2 is None:
            raise error("solvePnPRefineLM: only zero distortion / default criteria are implemented by ransac_b200")
        args = (objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec, tvec) + ((criteria,) if criteria is not None else ())
        return cv2.solvePnPRefineLM(*args)
    ctx = _ctx or api.default_context()
    r, t, _ = ctx.solve_pnp_refine_lm(np.asarray(objectPoints, dtype=np.float64).reshape(-1, 3),
                                      np.asarray(imagePoints, dtype=np.float64).reshape(-1, 2), cameraMatrix, rvec, tvec)
    return r, t


def module():
    """A module object usable wherever the scripts use `cv2`."""
    m = types.ModuleType("cv2")
    real = _real_cv2()
    if real is not None:
        for name in dir(real):
            if not name.startswith("__"):
                try:
                    setattr(m, name, getattr(real, name))
                except Exception:
                    pass
    m.RANSAC = RANSAC
    m.findHomography = findHomography
    m.solvePnPRansac = solvePnPRansac
    m.solvePnPRefineLM = solvePnPRefineLM
    m.__ransac_b200__ = True
    return m
