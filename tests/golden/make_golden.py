#!/usr/bin/env python
"""Generates tests/golden/*.json — run in the BUILD CONTAINER only (needs cv2 4.13.0 and /root/reference).

The reference pins nothing for its hot path (no tests, no vectors), so the golden vectors are outputs of the
reference's own implementation of the path run here: the cv2 4.13.0 binary called exactly as the reference calls
it (cv2.findHomography(src, dst, cv2.RANSAC, thr), main_v1.py:312; cv2.solvePnPRansac(...), main_v1.py:497), on
  * the repo's only complete data set (testpro-K.py:198-225, parsed from the file) swept over the 458 candidate
    camera locations of potential_camera_locations.csv (thr 75.0, main_v1.py:862),
  * the reference's recorded run /root/reference/debug.log (inputs recovered from the log, thr 120.0),
  * seeded random problems (sizes 4 ... 300),
plus bit-level vectors for the 4-point solver (cv2.findHomography(s4, d4, 0)), cv2.eigen and cv2.projectPoints.
tests/test_oracle_golden.py replays them against the oracle WITHOUT cv2 or /root/reference; the GPU tests use
the same files.  Floats are written with repr(), which round-trips float64 exactly."""
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

from ransac_b200 import geo, pipeline, synth  # noqa: E402  (host-side helpers only; no GPU needed)


def fixture_a():
    """pos3d (12,3), pixels (12,2) literals of testpro-K.py:198-225 (the script runs its pipeline at import)."""
    text = open(os.path.join(REF, "testpro-K.py"), encoding="utf-8").read()

    def grab(name):
        m = re.search(name + r"\s*=\s*np\.array\((\[.*?\])\s*\)", text, re.S)
        return np.array(ast.literal_eval(m.group(1)), dtype=np.float64)
    return grab("pos3d"), grab("pixels")


def parse_debug_log():
    """The complete candidate blocks of debug.log: logged M (= inv(H)), legacy mask, p1, pp2 per feature."""
    text = open(os.path.join(REF, "debug.log"), encoding="utf-8").read()
    blocks = text.split("Homography Matrix M:")[1:]
    out = []
    for b in blocks:
        head = b.split("- DEBUG - Mask:")[0]
        nums = [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", head.split(" - DEBUG")[0].replace("\n", " "))]
        mtxt = re.search(r"\[\[.*?\]\]", b, re.S).group(0)
        M = np.array([float(x) for x in re.findall(r"[-+]?\d+\.\d*(?:[eE][-+]?\d+)?", mtxt)]).reshape(3, 3)
        feats = re.findall(r"Feature (\d+): mask=\[(\d)\], p1=\[\s*([-\d.]+)\s+([-\d.]+)\s*\], pp2=\[\s*([-\d.eE+]+)\s+([-\d.eE+]+)\s*\]", b)
        feats = [f for f in feats]
        if len(feats) < 12:
            continue
        feats = feats[:12]
        mask = [int(f[1]) for f in feats]
        p1 = np.array([[float(f[2]), float(f[3])] for f in feats])
        pp2 = np.array([[float(f[4]), float(f[5])] for f in feats])
        out.append(dict(M=M, mask=mask, p1=p1, pp2=pp2))
        del nums
    return out


def tolist(a):
    return np.asarray(a).tolist()


def main():
    g = {"cv2_version": cv2.__version__}
    rng = np.random.default_rng(20261018)

    # ---- RNG known answers --------------------------------------------------------------------
    g["rng_first8"] = [130063605, 3133359004, 2578348940, 925327173, 1080261831, 2946015512, 94037301, 2298661280]

    # ---- cv2.eigen, bit level -------------------------------------------------------------------------------
    eig = []
    for _ in range(20):
        B = rng.standard_normal((12, 9))
        A = B.T @ B
        ok, w, v = cv2.eigen(A)
        eig.append(dict(A=tolist(A), w=tolist(w.ravel()), v=tolist(v)))
    g["eigen9"] = eig

    # ---- 4-point solver, bit level --------------------------------------------------------------------------
    k4 = []
    for _ in range(60):
        s = rng.uniform(-3, 3, (4, 2)).astype(np.float32)
        d = rng.uniform(0, 2000, (4, 2)).astype(np.float32)
        H, _ = cv2.findHomography(s.astype(np.float64), d.astype(np.float64), 0)
        if H is not None:
            k4.append(dict(src=tolist(s.astype(np.float64)), dst=tolist(d.astype(np.float64)), H=tolist(H)))
    g["kernel4"] = k4

    # ---- random RANSAC problems --------------------------------------------------------------------------------
    probs = []
    for t in range(40):
        n = int(rng.choice([4, 5, 6, 8, 12, 20, 50, 100, 300]))
        thr = float(rng.choice([3.0, 10.0, 30.0, 75.0]))
        s, d, _ = synth.homography_set(n, float(rng.uniform(0.0, 0.6)), rng, noise_px=float(rng.uniform(0.3, 3.0)))
        H, m = cv2.findHomography(s, d, cv2.RANSAC, thr)
        probs.append(dict(src=tolist(s), dst=tolist(d), thr=thr, H=None if H is None else tolist(H), mask=tolist(m.ravel())))
    # degenerate: all points equal -> no model
    z = np.zeros((10, 2))
    H, m = cv2.findHomography(z, z, cv2.RANSAC, 3.0)
    probs.append(dict(src=tolist(z), dst=tolist(z), thr=3.0, H=None if H is None else tolist(H), mask=tolist(m.ravel())))
    g["ransac_random"] = probs

    # ---- Fixture A sweep: 458 candidate camera locations, thr 75.0 (main_v1.py:862) ------------------------------
    pos3d, pixels = fixture_a()
    locs = geo.read_camera_locations(os.path.join(REF, "potential_camera_locations.csv"))
    loc3ds = np.array([c["pos3d"] for c in locs])
    sweep = dict(pos3d=tolist(pos3d), pixels=tolist(pixels), loc3ds=tolist(loc3ds), grids=[int(c["grid_code"]) for c in locs],
                 thr=75.0, H=[], mask=[], err1=[], err2=[])
    for i in range(len(locs)):
        pos2 = pipeline.candidate_pos2(pos3d, loc3ds[i])
        H, m = cv2.findHomography(pos2, pixels, cv2.RANSAC, 75.0)
        M, e1, e2 = pipeline._score(H, m, pos2, pixels, 75.0)
        sweep["H"].append(tolist(H))
        sweep["mask"].append(tolist(m.ravel()))
        sweep["err1"].append(e1)
        sweep["err2"].append(e2)
    nm = np.stack([sweep["err1"], sweep["err2"]], axis=1)
    sweep["best_index"] = pipeline.best_location(nm)
    g["fixture_a_sweep"] = sweep

    # ---- debug.log known answers (thr 120, legacy mask semantics) ------------------------------------------------
    dbg = []
    for b in parse_debug_log():
        q = np.concatenate([b["pp2"], np.ones((12, 1))], axis=1) @ b["M"].T
        pos2 = q[:, :2] / q[:, 2:3]                       # pos2 = normalize(M [pp2, 1])
        H, m = cv2.findHomography(pos2, b["p1"], cv2.RANSAC, 120.0)
        dbg.append(dict(pos2=tolist(pos2), p1=tolist(b["p1"]), logged_M=tolist(b["M"]), logged_mask=b["mask"],
                        cv413_H=tolist(H), cv413_mask=tolist(m.ravel())))
    g["debug_log"] = dbg

    # ---- PnP: projectPoints bit level + solvePnPRansac on Fixture A ------------------------------------------------
    K = synth.K_1898
    dist = np.zeros((4, 1))
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pos3d, pixels, K, dist, iterationsCount=5000, reprojectionError=30.0, confidence=0.99)
    r2, t2 = cv2.solvePnPRefineLM(pos3d[inl], pixels[inl], K, dist, rvec.copy(), tvec.copy())
    obj32 = pos3d.astype(np.float32)
    proj, _ = cv2.projectPoints(obj32, rvec, tvec, K, dist)
    Rm, _ = cv2.Rodrigues(rvec)
    g["pnp_fixture_a"] = dict(K=tolist(K), ok=bool(ok), rvec=tolist(rvec.ravel()), tvec=tolist(tvec.ravel()),
                              inliers=tolist(inl.ravel()), refined_rvec=tolist(r2.ravel()), refined_tvec=tolist(t2.ravel()),
                              R=tolist(Rm), proj_f32=tolist(proj.reshape(-1, 2).astype(np.float64)))
    pv = []
    for _ in range(10):
        P, px, _ = synth.pnp_set(50, 0.0, rng)
        R, t = synth.look_at_pose()
        rv, _ = cv2.Rodrigues(R)
        rv = rv.ravel() + rng.normal(0, 1e-3, 3)
        tv = t + rng.normal(0, 0.5, 3)
        o32 = P.astype(np.float32)
        pr, _ = cv2.projectPoints(o32, rv, tv, K, dist)
        Rr, _ = cv2.Rodrigues(rv)
        pv.append(dict(obj=tolist(o32.astype(np.float64)), rvec=tolist(rv), tvec=tolist(tv), R=tolist(Rr),
                       proj_f32=tolist(pr.reshape(-1, 2).astype(np.float64))))
    g["project_points"] = pv

    # ---- EPnP minimal solver on 5-point samples (PnPRansacCallback::runKernel) + full solvePnPRansac on synthetic sets ------
    ep = []
    for _ in range(40):
        P, px, _ = synth.pnp_set(5, 0.0, rng, noise_px=float(rng.choice([0.0, 1.0])))
        o32, i32 = P.astype(np.float32), px.astype(np.float32)
        ok, rv, tv = cv2.solvePnP(o32, i32, K, dist, flags=cv2.SOLVEPNP_EPNP)
        ep.append(dict(obj=tolist(o32.astype(np.float64)), img=tolist(i32.astype(np.float64)), ok=bool(ok), rvec=tolist(rv.ravel()), tvec=tolist(tv.ravel())))
    g["epnp5"] = ep
    pr = []
    for _ in range(24):
        n = int(rng.choice([5, 6, 12, 30, 100, 500, 2000]))
        thr = float(rng.choice([8.0, 30.0]))
        P, px, _ = synth.pnp_set(n, float(rng.uniform(0.0, 0.5)), rng)
        ok, rv, tv, inl = cv2.solvePnPRansac(P, px, K, dist, iterationsCount=5000, reprojectionError=thr, confidence=0.99)
        rec = dict(obj=tolist(P), img=tolist(px), thr=thr, ok=bool(ok), rvec=tolist(rv.ravel()), tvec=tolist(tv.ravel()),
                   inliers=None if inl is None else tolist(inl.ravel()))
        if ok and inl is not None and len(inl) >= 6:
            r3, t3 = cv2.solvePnPRefineLM(P[inl.ravel()], px[inl.ravel()], K, dist, rv.copy(), tv.copy())
            rec["refined_rvec"], rec["refined_tvec"] = tolist(r3.ravel()), tolist(t3.ravel())
        pr.append(rec)
    g["pnp_ransac_random"] = pr

    path = os.path.join(HERE, "cv2_golden.json")
    with open(path, "w") as f:
        json.dump(g, f)
    print("wrote", path, os.path.getsize(path), "bytes;", len(dbg), "debug.log blocks;", "best candidate index", sweep["best_index"],
          "err2", sweep["err2"][sweep["best_index"]], "pnp inliers", g["pnp_fixture_a"]["inliers"])


if __name__ == "__main__":
    main()
