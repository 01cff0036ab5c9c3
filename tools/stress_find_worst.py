"""Re-run the homography half of tools/stress_parity.py for one seed and save the problem with the largest deviation of the
refined H from the oracle (development probe): python tools/stress_find_worst.py <seed> <out.npz>"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import ransac_b200
from ransac_b200 import synth

seed = int(sys.argv[1]); out = sys.argv[2]
rng = np.random.default_rng(seed)
ctx = ransac_b200.Context(0)
worst = (0.0, None)
rows = []
for batch in range(40):
    n = int(rng.choice([5, 6, 7, 8, 10, 12, 16, 24, 40, 80, 150, 400]))
    Q = 60
    thr = float(rng.choice([1.0, 3.0, 10.0, 75.0]))
    src, dst = np.zeros((Q, n, 2)), np.zeros((Q, n, 2))
    for q in range(Q):
        src[q], dst[q], _ = synth.homography_set(n, float(rng.uniform(0, 0.7)), rng, noise_px=float(rng.choice([0.0, 0.5, 1.0, 4.0])))
        if rng.random() < 0.1:
            dst[q] = rng.uniform(0, 2000, (n, 2))
    H, ok, mask, infos = ctx.find_homography_batch(src, dst, thr)
    for q in range(Q):
        Hr, mr, det = oracle.find_homography(src[q], dst[q], thr, details=True)
        if Hr is None or not ok[q]:
            continue
        rel = float(np.abs(H[q] - Hr).max() / np.abs(Hr).max())
        if rel > 1e-7:
            rows.append(dict(batch=batch, q=q, n=n, thr=thr, rel=rel, inliers=int(mr.sum()), lm_iters=infos[q]["lm_iters"],
                             mask_equal=bool(np.array_equal(mask[q], mr.ravel()))))
        if rel > worst[0]:
            worst = (rel, dict(src=src[q].copy(), dst=dst[q].copy(), thr=thr, H_gpu=H[q].copy(), H_oracle=Hr.copy(), mask_gpu=mask[q].copy(),
                               mask_oracle=mr.ravel().copy(), ransac_mask=det["ransac_mask"].copy(), ransac_H=det["ransac_H"].copy()))
np.savez(out, **worst[1])
print(json.dumps(dict(worst=worst[0], rows=sorted(rows, key=lambda r: -r["rel"])[:12])))
