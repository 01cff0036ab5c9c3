"""Host-side mirror of the reference's callers of the hot path — same names, argument meaning and error behaviour:

    find_homography(recs, pixels, pos3ds, symbols, camera_location, im, show, ransacbound, outputfile)
                                                         main_v1.py:302-422 (process.py:188-291, testpro.py:340-460)
    find_homographies(recs, camera_locations, im, show, ransacbound, output)          main_v1.py:254-297
    best_location(num_matches)                                                        main_v1.py:863-866
    camera_matrix_from_image(width, height)                                           main_v1.py:870-883
    estimate_camera_pose(pos3d, pixels, K)                                            main_v1.py:468-512
    estimate_camera_orientation(pos3d, pixels, focal_lengths, sensor_sizes, image_size, known_camera_origin)
                                                                                      testpro-K.py:39-162

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
                      return_details=False, fused=True, **ransac_kw):
    """The camera-location sweep, main_v1.py:254-297: returns num_matches (Q, 2) = [err1, err2] per candidate and,
    when `output` is given and show is False, writes the reference's `*_location.csv` (main_v1.py:286-292).

    fused=True (default): ONE device call (b2r_camera_sweep) does the per-candidate projection, the RANSAC, the scores
    and the arg-min.  fused=False: projection and scores in NumPy around the batched RANSAC call (same results)."""
    ctx = ctx or api.default_context()
    pixels = np.array([r["pixel"] for r in recs], dtype=np.float64)
    pos3ds = np.array([r["pos3d"] for r in recs], dtype=np.float64)
    grids = np.array([cl["grid_code"] for cl in camera_locations])
    loc3ds = np.array([cl["pos3d"] for cl in camera_locations], dtype=np.float64)
    Q = loc3ds.shape[0]
    good = (pixels[:, 0] != 0) | (pixels[:, 1] != 0)
    grid_code_min = 0                                                           # main_v1.py:275
    num_matches = np.zeros((Q, 2))
    Ms = np.zeros((Q, 3, 3))
    if fused:
        res = ctx.camera_sweep(pos3ds[good], pixels[good], loc3ds, ransacbound, **ransac_kw)
        H, mask, infos = res["H"], res["mask"], res["infos"]
        ok = infos.status == api.OK
        pos2 = None
        for i in range(Q):
            if grids[i] >= grid_code_min:
                if not ok[i]:
                    raise np.linalg.LinAlgError(f"findHomography returned no model for candidate {i} "
                                                "(the reference fails at main_v1.py:314)")
                Ms[i], num_matches[i] = res["M"][i], res["scores"][i]
        # The device arg-min (res["best"]) runs over all Q candidates; the reference's rule is applied to the scores the
        # host loop kept (candidates below grid_code_min stay 0 -> 1e6; np.argmin returns the first NaN): Q doubles.
        best = best_location(num_matches)
    else:
        pos2 = candidate_pos2(pos3ds[None, good, :], loc3ds[:, None, :])          # (Q, n, 2)
        H, ok, mask, infos = ctx.find_homography_batch(pos2, pixels[good], ransacbound, **ransac_kw)
        for i in range(Q):
            if grids[i] >= grid_code_min:
                if not ok[i]:
                    raise np.linalg.LinAlgError(f"findHomography returned no model for candidate {i} "
                                                "(the reference fails at main_v1.py:314)")
                Ms[i], num_matches[i, 0], num_matches[i, 1] = _score(H[i], mask[i], pos2[i], pixels[good], ransacbound)
        best = best_location(num_matches)
    if show is False and output:
        scores = [[i + 1, num_matches[i, 0], num_matches[i, 1], grids[i], loc3ds[i][0], loc3ds[i][1], loc3ds[i][2]]
                  for i in range(Q)]
        with open(output.replace(".jpg", "_location.csv"), "w", newline="", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(["location_id", "min_score", "max_score", "grid_code", "Z", "X", "Y"])
            w.writerows(scores)
    if return_details:
        if pos2 is None:
            pos2 = candidate_pos2(pos3ds[None, good, :], loc3ds[:, None, :])
        return num_matches, dict(H=H, M=Ms, mask=mask, infos=infos, pos2=pos2, best=best)
    return num_matches


def best_location(num_matches):
    """do_it, main_v1.py:863-866: zeros -> 1e6, argmin of err2."""
    err2 = np.array(num_matches[:, 1], dtype=np.float64)
    err2[err2 == 0] = 1000000
    return int(np.argmin(err2))


# ---- path B: PnP pose ---------------------------------------------------------------------------------------------
def camera_matrix_from_image(width, height):
    """K of do_it, main_v1.py:870-883: fx = 240/127 * W, fy = 240/178 * H, principal point (982.666819, 697.950868)."""
    return np.array([[240.0 / 127.0 * width, 0.0, 982.666819], [0.0, 240.0 / 178.0 * height, 697.950868], [0.0, 0.0, 1.0]])


def rodrigues(rvec):
    """Rotation matrix of a rotation vector (what the reference gets from cv2.Rodrigues, main_v1.py:895)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(r))
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def estimate_camera_pose(pos3d, pixels, K, ctx=None, **ransac_kw):
    """main_v1.py:468-512 without the plots: solvePnPRansac(5000, 30.0, 0.99), the `< 6 inliers` gate (:504), then
    solvePnPRefineLM on pos3d[inliers] (:508).  Returns (rvec (3,1), tvec (3,1), inliers (k,1) int32) or (None, None, None)."""
    ctx = ctx or api.default_context()
    pos3d = np.asarray(pos3d, dtype=np.float64).reshape(-1, 3)
    pixels = np.asarray(pixels, dtype=np.float64).reshape(-1, 2)
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    ok, rvec, tvec, inliers, _ = ctx.solve_pnp_ransac(pos3d, pixels, K, 5000, 30.0, 0.99, **ransac_kw)
    if not ok or inliers is None or len(inliers) < 6:
        return None, None, None
    idx = inliers.ravel()
    rvec, tvec, _ = ctx.solve_pnp_refine_lm(pos3d[idx], pixels[idx], K, rvec, tvec)
    return rvec, tvec, inliers


def intrinsics_grid(focal_lengths, sensor_sizes, image_size):
    """The K matrices of testpro-K.py:58-70, in loop order, with their (focal, sensor) labels."""
    Ks, labels = [], []
    for f in focal_lengths:
        for (sw, sh) in sensor_sizes:
            fx = f / (sw / image_size[0])
            fy = f / (sh / image_size[1])
            Ks.append([[fx, 0, image_size[0] / 2], [0, fy, image_size[1] / 2], [0, 0, 1]])
            labels.append((f, (sw, sh)))
    return np.array(Ks, dtype=np.float64), labels


def estimate_camera_orientation(pos3d, pixels, focal_lengths, sensor_sizes, image_size, known_camera_origin=None, ctx=None,
                                return_details=False, **ransac_kw):
    """testpro-K.py:39-162 without the prints: one solvePnPRansac per K of the grid — here ONE batched GPU call over
    the shared points — skip K with < 6 inliers (:77), keep the K with the smallest mean inlier reprojection error
    (:80-97), refine that pose with solvePnPRefineLM (:122-125).  Returns (rvec, tvec) like the reference, or
    (None, None) when every K fails (:99-101)."""
    ctx = ctx or api.default_context()
    pos3d = np.asarray(pos3d, dtype=np.float64).reshape(-1, 3)
    pixels = np.asarray(pixels, dtype=np.float64).reshape(-1, 2)
    Ks, labels = intrinsics_grid(focal_lengths, sensor_sizes, image_size)
    ok, rvecs, tvecs, inliers, infos = ctx.solve_pnp_ransac_batch(pos3d, pixels, Ks, 5000, 30.0, 0.99, **ransac_kw)
    best, best_err, results = None, float("inf"), []
    for q in range(len(Ks)):
        if not ok[q] or len(inliers[q]) < 6:
            continue
        err = infos[q]["mean_inlier_err"]
        origin = -rodrigues(rvecs[q]).T @ tvecs[q]
        dist = None if known_camera_origin is None else float(np.linalg.norm(origin - np.asarray(known_camera_origin, dtype=np.float64)))
        results.append(dict(index=q, label=labels[q], mean_error=err, camera_origin=origin, distance=dist, inliers=inliers[q]))
        if err < best_err:
            best, best_err = q, err
    if best is None:
        return (None, None, dict(results=[])) if return_details else (None, None)
    idx = inliers[best]
    rvec, tvec, _ = ctx.solve_pnp_refine_lm(pos3d[idx], pixels[idx], Ks[best], rvecs[best], tvecs[best])
    if return_details:
        return rvec, tvec, dict(results=results, best=best, best_K=Ks[best], best_label=labels[best], best_error=best_err,
                                initial_rvec=rvecs[best].reshape(3, 1), initial_tvec=tvecs[best].reshape(3, 1),
                                inliers=idx.reshape(-1, 1), ok=ok, infos=infos)
    return rvec, tvec
