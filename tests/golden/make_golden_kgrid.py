#!/usr/bin/env python
"""Generates tests/golden/cv2_kgrid.json — run in the BUILD CONTAINER only (needs cv2 4.13.0 and /root/reference).

The intrinsics grid search of the reference, testpro-K.py:39-162, executed with the cv2 binary exactly as the script
does (its data literals and grid are parsed from the file; the script itself runs plots at import): per K the
cv2.solvePnPRansac result (:72-75), the mean inlier reprojection error (:32-36, :80-82), the selected K (:90-97) and
the cv2.solvePnPRefineLM pose of the winner (:122-125).  Also the full estimate_camera_pose chain of
main_v1.py:468-512 on the same points with the K of main_v1.py:870-883."""
import ast
import json
import os
import re
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def tolist(a):
    return np.asarray(a, dtype=np.float64).tolist()


def main():
    text = open(os.path.join(REF, "testpro-K.py"), encoding="utf-8").read()

    def grab(name):
        m = re.search(name + r"\s*=\s*np\.array\((\[.*?\])\s*\)", text, re.S)
        return np.array(ast.literal_eval(m.group(1)), dtype=np.float64)

    def grab_list(name):
        m = re.search(name + r"\s*=\s*(\[.*?\])\s*\n", text, re.S)
        return ast.literal_eval(m.group(1))
    pos3d, pixels = grab("pos3d"), grab("pixels")
    focal_lengths, sensor_sizes = grab_list("focal_lengths"), grab_list("sensor_sizes")
    image_size = ast.literal_eval(re.search(r"image_size\s*=\s*(\(.*?\))", text).group(1))
    known = grab("known_camera_origin") if re.search(r"known_camera_origin\s*=\s*np\.array", text) else \
        np.array(ast.literal_eval(re.search(r"known_camera_origin\s*=\s*(\[.*?\])", text).group(1)), dtype=np.float64)
    dist = np.zeros((4, 1))
    grid, best, best_err = [], None, float("inf")
    for f in focal_lengths:
        for (sw, sh) in sensor_sizes:
            fx, fy = f / (sw / image_size[0]), f / (sh / image_size[1])
            K = np.array([[fx, 0, image_size[0] / 2], [0, fy, image_size[1] / 2], [0, 0, 1]])
            ok, rv, tv, inl = cv2.solvePnPRansac(pos3d, pixels, K, dist, useExtrinsicGuess=False, iterationsCount=5000,
                                                 reprojectionError=30.0, confidence=0.99)
            rec = dict(focal=f, sensor=[sw, sh], K=tolist(K), ok=bool(ok), inliers=None if inl is None else tolist(inl.ravel()),
                       rvec=tolist(rv.ravel()), tvec=tolist(tv.ravel()), used=False, mean_error=None)
            if ok and inl is not None and len(inl) >= 6:
                idx = inl.flatten()
                proj, _ = cv2.projectPoints(pos3d[idx], rv, tv, K, dist)
                err = float(np.mean(np.linalg.norm(pixels[idx] - proj.squeeze(), axis=1)))
                rec["used"], rec["mean_error"] = True, err
                if err < best_err:
                    best, best_err = len(grid), err
            grid.append(rec)
    b = grid[best]
    idx = np.array(b["inliers"], dtype=np.int64)
    r2, t2 = cv2.solvePnPRefineLM(pos3d[idx], pixels[idx], np.array(b["K"]), dist, np.array(b["rvec"]).reshape(3, 1),
                                  np.array(b["tvec"]).reshape(3, 1))
    out = dict(cv2_version=cv2.__version__, pos3d=tolist(pos3d), pixels=tolist(pixels), focal_lengths=focal_lengths,
               sensor_sizes=[list(s) for s in sensor_sizes], image_size=list(image_size), known_camera_origin=tolist(known),
               grid=grid, best=best, best_error=best_err, refined_rvec=tolist(r2.ravel()), refined_tvec=tolist(t2.ravel()))
    # main_v1.py:468-512 with K of main_v1.py:870-883 (image 2142 x 1620)
    K1 = np.array([[240.0 / 127.0 * 2142, 0, 982.666819], [0, 240.0 / 178.0 * 1620, 697.950868], [0, 0, 1]])
    ok, rv, tv, inl = cv2.solvePnPRansac(pos3d, pixels, K1, dist, iterationsCount=5000, reprojectionError=30.0, confidence=0.99)
    r3, t3 = cv2.solvePnPRefineLM(pos3d[inl], pixels[inl], K1, dist, rv.copy(), tv.copy())
    out["estimate_camera_pose"] = dict(K=tolist(K1), ok=bool(ok), inliers=tolist(inl.ravel()), rvec=tolist(r3.ravel()), tvec=tolist(t3.ravel()))
    path = os.path.join(HERE, "cv2_kgrid.json")
    with open(path, "w") as fjs:
        json.dump(out, fjs)
    used = [g for g in grid if g["used"]]
    print("wrote", path, len(grid), "K matrices,", len(used), "used, best", best, b["focal"], b["sensor"], "mean error", best_err)


if __name__ == "__main__":
    main()
