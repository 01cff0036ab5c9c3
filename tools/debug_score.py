import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
import oracle as O
ctx = ransac_b200.Context(0)
n, n_models, seed = 1000, 3000, 9
rng = np.random.default_rng(seed)
s, d, _ = synth.homography_set(n, 0.5, rng)
sq, dq = s.astype(np.float32), d.astype(np.float32)
rng = np.random.default_rng(seed)
idx = np.stack([rng.choice(n, 4, replace=False) for _ in range(n_models)]).astype(np.int32)
H, ok, _ = ctx.solve_h4(sq, dq, idx)
models = H.reshape(-1, 9)[:, :8].astype(np.float32)
models[~ok] = np.nan
for thr in (3.0, 75.0):
    thr_sq = np.float32(thr * thr)
    got = ctx.score_h(models, sq, dq, thr_sq, 0)
    ref = O.h_count_inliers_f32(models, sq, dq, thr_sq)
    bad = np.nonzero(got != ref)[0]
    print("thr", thr, "bad", bad, got[bad], ref[bad])
    for b in bad:
        m = models[b:b+1]
        for i in range(n):
            g = ctx.score_h(m, sq[i:i+1], dq[i:i+1], thr_sq, 0)[0]
            r = O.h_count_inliers_f32(m, sq[i:i+1], dq[i:i+1], thr_sq)[0]
            if g != r:
                print("model", b, m.view(np.uint32).tolist(), "point", i, sq[i].view(np.uint32).tolist(), dq[i].view(np.uint32).tolist(), "gpu", g, "cpu", r,
                      "err_cpu", O.h_compute_error(np.r_[m[0].astype(np.float64), 1.0], sq[i:i+1], dq[i:i+1]), "thr_sq", thr_sq)
