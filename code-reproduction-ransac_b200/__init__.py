"""ransac_b200 — B200-native (sm_100a) RANSAC camera-location hot path.

Drop-in for the two OpenCV calls the reference makes (cv2.findHomography(..., cv2.RANSAC, thr) at
main_v1.py:312 and cv2.solvePnPRansac at main_v1.py:497) and for the Python loops around them.  The directory
name contains hyphens, so it is imported through the alias module `ransac_b200.py` at the repo root.
"""
from . import _build  # noqa: F401
from .api import (ARITH_EXACT, ARITH_EXACT_UNFILTERED, ARITH_FAST, MASK_CV413, MASK_LEGACY, REFINE_CV, REFINE_NONE, REFINE_PARALLEL, SAMPLER_CV_REPLAY, SAMPLER_PHILOX, SOLVER_EXACT, SOLVER_FAST, Context,
                  HomographyProblem, PnPProblem, RansacB200Error, default_context, make_p_params, make_params)

__all__ = ["Context", "HomographyProblem", "RansacB200Error", "default_context", "make_params", "make_p_params", "PnPProblem", "ARITH_EXACT", "ARITH_EXACT_UNFILTERED", "ARITH_FAST",
           "MASK_CV413", "MASK_LEGACY", "SAMPLER_CV_REPLAY", "SAMPLER_PHILOX", "SOLVER_EXACT", "SOLVER_FAST", "REFINE_NONE", "REFINE_CV", "REFINE_PARALLEL"]
__version__ = "0.1.0"
