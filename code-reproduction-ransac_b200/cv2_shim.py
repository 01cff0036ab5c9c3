"""A `cv2`-shaped module for the reference scripts: the hot-path entry points run on the B200 through
libransac_b200.so, every other attribute is forwarded to the real OpenCV (imread, Rodrigues, projectPoints, ...).

    import ransac_b200.cv2_shim as shim
    main_v1.cv2 = shim.module()            # or: sys.modules["cv2"] = shim.module() before importing the script

Replaces exactly the lookups SURVEY.md §8(b) lists: cv2.findHomography with method == cv2.RANSAC
(main_v1.py:312, process.py:200, testpro.py:350, test_pro.py:351, test02.py:263) and the constant cv2.RANSAC.
Return conventions are cv2's: (H float64 (3,3) or None, mask uint8 (n,1)); fewer than 4 points raise."""
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
    H, m, _ = ctx.find_homography(src, dst, float(ransacReprojThreshold), int(maxIters), float(confidence), **b2r_kw)
    return H, m


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
    m.__ransac_b200__ = True
    return m
