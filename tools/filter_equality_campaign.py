"""Equality campaign for the filtered exact predicate: many random shapes, thresholds, noise levels and hypothesis mixes,
B2R_ARITH_EXACT (filtered) against B2R_ARITH_EXACT_UNFILTERED (OpenCV's sequence on every evaluation), both model families.
    python tools/filter_equality_campaign.py [seed] [rounds]          one JSON line per round + a summary line
Every count of every hypothesis is compared; the script exits non-zero on the first difference."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rng = np.random.default_rng(seed)
ctx = ransac_b200.Context(0)
K = synth.K_1898
tot_evals = 0
t0 = time.time()


def rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


for r in range(rounds):
    n = int(rng.choice([64, 100, 333, 1000, 1024, 1025, 4097, 20000, 100000]))
    H = int(rng.choice([1, 100, 1023, 1024, 1025, 5000, 40000]))
    thr = float(rng.choice([0.5, 1.0, 3.0, 8.0, 30.0, 75.0]))
    noise = float(rng.choice([0.0, 0.3, 1.0, 3.0]))
    outl = float(rng.uniform(0, 0.8))
    thr_sq = np.float32(thr * thr)
    # ---- 3x3 ----
    s, d, _ = synth.homography_set(n, outl, rng, noise_px=noise)
    sq, dq = s.astype(np.float32), d.astype(np.float32)
    idx = np.stack([rng.choice(n, 4, replace=False) for _ in range(H)]).astype(np.int32)
    Hm, ok, _ = ctx.solve_h4(sq, dq, idx)
    models = Hm.reshape(-1, 9)[:, :8].astype(np.float32)
    models[~ok] = np.nan
    good = np.flatnonzero(ok)
    if len(good):   # a share of hypotheses within 1e-5 ... 1e-2 of a sampled one: dense near the threshold when it is a good one
        pick = rng.choice(good, size=max(1, H // 3))
        jitter = 10.0 ** rng.uniform(-5, -2, (len(pick), 1))
        models[rng.choice(H, size=len(pick))] = (models[pick] * (1 + jitter * rng.standard_normal((len(pick), 8)))).astype(np.float32)
    a = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_EXACT)
    b = ctx.score_h(models, sq, dq, thr_sq, ransac_b200.ARITH_EXACT_UNFILTERED)
    bad_h = int((a != b).sum())
    # ---- 3x4 ----
    P, px, _ = synth.pnp_set(n, outl, rng, noise_px=noise)
    R0, _ = synth.look_at_pose()
    Hp = min(H, 20000)
    poses = np.zeros((Hp, 12))
    centre = 0.5 * (synth.BOX_LO + synth.BOX_HI)
    for k in range(Hp):
        kind = k % 3
        if kind == 0:
            R = rodrigues(rng.normal(0, 10.0 ** rng.uniform(-5, -2), 3)) @ R0
            cam = synth.CAMERA_ORIGIN + rng.normal(0, 10.0 ** rng.uniform(-2, 1), 3)
        elif kind == 1:
            R = rodrigues(rng.normal(0, 0.2, 3)) @ R0
            cam = synth.CAMERA_ORIGIN + rng.normal(0, 100.0, 3)
        else:
            R = rodrigues(rng.normal(0, 2.0, 3))
            cam = centre + rng.normal(0, 300.0, 3)
        poses[k, :9] = R.ravel()
        poses[k, 9:] = -R @ cam
    pa = ctx.score_p(poses, P, px, K, thr_sq, ransac_b200.ARITH_EXACT)
    pb = ctx.score_p(poses, P, px, K, thr_sq, ransac_b200.ARITH_EXACT_UNFILTERED)
    bad_p = int((pa != pb).sum())
    tot_evals += n * H + n * Hp
    print(json.dumps({"round": r, "n": n, "H": H, "thr_px": thr, "noise_px": noise, "outliers": round(outl, 3),
                      "h_counts_different": bad_h, "h_max_count": int(a.max()), "p_counts_different": bad_p, "p_max_count": int(pa.max())}), flush=True)
    if bad_h or bad_p:
        sys.exit(1)
print(json.dumps({"summary": "filtered == unfiltered on every count", "seed": seed, "rounds": rounds, "evaluations_compared": tot_evals,
                  "seconds": round(time.time() - t0, 1)}))
