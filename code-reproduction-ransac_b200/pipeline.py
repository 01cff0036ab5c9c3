"""Host-side mirror of the reference's callers of the hot path — same names, argument meaning and error behaviour:

    find_homography(recs, pixels, pos3ds, symbols, camera_location, im, show, ransacbound, outputfile)
                                                         main_v1.py:302-422 (process.py:188-291, testpro.py:340-460)
    find_homographies(recs, camera_locations, im, show, ransacbound, output)          main_v1.py:254-297
    best_location(num_matches)                                                        main_v1.py:863-866

The reference loops over the candidate camera locations in Python and calls cv2.findHomography once per candidate;
here the whole sweep is ONE batched GPU call (Q independent RANSAC problems), with the reference's per-candidate
projection (main_v1.py:304-311) and score (main_v1.py:327-348, :419) evaluated vectorised around it.  Plotting,
DEBUG logging and the show=True report files are outside the hot path and not reproduced."""
import csv

import numpy as np

from . import api


def candidate_pos2(pos3ds, camera_location):
    """main_v1.py:306-311: p = pos3d - camera; p = [p2, p1, p0]; p /= p[2]; pos2 = p[:2]   (float64)."""
    p = np.asarray(pos3ds, dtype=np.float64) - np.asarray(camera_location, dtype=np.float64)
    return np.stack([p[..., 2] / p[..., 0], p[..., 1] / p[..., 0]], axis=-1)


def _score(H, mask, pos2, pixels, ransacbound):
    """err1, err2 of main_v1.py:327-348 and :419 for one candidate (H = what cv2.findHomography returned)."""
    M = np.linalg.inv(H)                       # main_v1.py:314
    Minv = np.linalg.inv(M)                    # main_v1.py:335 (recomputed per point there; same value)
    n = len(pos2)
    pp = np.concatenate([pos2, np.ones((n, 1))], axis=1)
    pp2 = pp @ Minv.T
    pp2 = pp2 / pp2[:, 2:3]
    P1 = np.concatenate([pixels, np.ones((n, 1))], axis=1)
    PP2 = P1 @ M.T
    PP2 = PP2 / PP2[:, 2:3]
    inl = mask.reshape(-1) == 1
    err1 = float(np.sum(np.linalg.norm(pixels[inl] - pp2[inl, :2], axis=1)))
    err2 = float(np.sum(np.linalg.norm(pos2[inl] - PP2[inl, :2], axis=1)))
    err2 += float(np.sum(1 - mask.reshape(-1).astype(np.int64))) * ransacbound   # main_v1.py:419
    return M, err1, err2


def find_homography(recs, pixels, pos3ds, symbols, camera_location, im=None, show=False, ransacbound=75.0,
                    outputfile=None, ctx=None, **ransac_kw):
    """One candidate camera location: returns (M, err1, err2) like main_v1.py:422 (M = inv(H)).

    Raises numpy.linalg.LinAlgError when RANSAC finds no model — the reference then fails in np.linalg.inv(None)
    (main_v1.py:314)."""
    ctx = ctx or api.default_context()
    pixels = np.asarray(pixels, dtype=np.float64)
    pos3ds = np.asarray(pos3ds, dtype=np.float64)
    good = (pixels[:, 0] != 0) | (pixels[:, 1] != 0)            # main_v1.py:308
    pos2 = candidate_pos2(pos3ds, camera_location)
    H, mask, _ = ctx.find_homography(pos2[good], pixels[good], ransacbound, **ransac_kw)
    if H is None:
        raise np.linalg.LinAlgError("findHomography returned no model (the reference fails at main_v1.py:314)")
    return _score(H, mask, pos2[good], pixels[good], ransacbound)


def find_homographies(recs, camera_locations, im=None, show=False, ransacbound=75.0, output=None, ctx=None,
                      return_details=False, **ransac_kw):
    """The camera-location sweep, main_v1.py:254-297: returns num_matches (Q, 2) = [err1, err2] per candidate and,
    when `output` is given and show is False, writes the reference's `*_location.csv` (main_v1.py:286-292)."""
    ctx = ctx or api.default_context()
    pixels = np.array([r["pixel"] for r in recs], dtype=np.float64)
    pos3ds = np.array([r["pos3d"] for r in recs], dtype=np.float64)
    grids = np.array([cl["grid_code"] for cl in camera_locations])
    loc3ds = np.array([cl["pos3d"] for cl in camera_locations], dtype=np.float64)
    Q = loc3ds.shape[0]
    good = (pixels[:, 0] != 0) | (pixels[:, 1] != 0)
    pos2 = candidate_pos2(pos3ds[None, good, :], loc3ds[:, None, :])          # (Q, n, 2)
    H, ok, mask, infos = ctx.find_homography_batch(pos2, pixels[good], ransacbound, **ransac_kw)
    num_matches = np.zeros((Q, 2))
    Ms = np.zeros((Q, 3, 3))
    grid_code_min = 0                                                           # main_v1.py:275
    for i in range(Q):
        if grids[i] >= grid_code_min:
            if not ok[i]:
                raise np.linalg.LinAlgError(f"findHomography returned no model for candidate {i} "
                                            "(the reference fails at main_v1.py:314)")
            Ms[i], num_matches[i, 0], num_matches[i, 1] = _score(H[i], mask[i], pos2[i], pixels[good], ransacbound)
    if show is False and output:
        scores = [[i + 1, num_matches[i, 0], num_matches[i, 1], grids[i], loc3ds[i][0], loc3ds[i][1], loc3ds[i][2]]
                  for i in range(Q)]
        with open(output.replace(".jpg", "_location.csv"), "w", newline="", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(["location_id", "min_score", "max_score", "grid_code", "Z", "X", "Y"])
            w.writerows(scores)
    if return_details:
        return num_matches, dict(H=H, M=Ms, mask=mask, infos=infos, pos2=pos2)
    return num_matches


def best_location(num_matches):
    """do_it, main_v1.py:863-866: zeros -> 1e6, argmin of err2."""
    err2 = np.array(num_matches[:, 1], dtype=np.float64)
    err2[err2 == 0] = 1000000
    return int(np.argmin(err2))
