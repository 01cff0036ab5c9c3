"""One-off parity stress: many random problems, GPU (C ABI) against the CPU oracle.  python tools/stress_parity.py [seed]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import ransac_b200
from ransac_b200 import synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rng = np.random.default_rng(seed)
ctx = ransac_b200.Context(0)
stats = dict(h_problems=0, h_stage_mismatch=0, h_iters_mismatch=0, h_final_mask_mismatch=0, h_H_worst=0.0, h_none=0, h_small=0, h_small_bit_equal=0,
             p_problems=0, p_inlier_mismatch=0, p_iters_mismatch=0, p_ok_mismatch=0, p_pose_worst=0.0, p_none=0)
t0 = time.time()
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
    _, _, ml, _ = ctx.find_homography_batch(src, dst, thr, mask_semantics=ransac_b200.MASK_LEGACY)
    for q in range(Q):
        Hr, mr, det = oracle.find_homography(src[q], dst[q], thr, details=True)
        stats["h_problems"] += 1
        if (Hr is None) != (not ok[q]):
            stats["h_stage_mismatch"] += 1
            continue
        if Hr is None:
            stats["h_none"] += 1
            continue
        stats["h_iters_mismatch"] += infos[q]["iters_run"] != det["iters"]
        stats["h_stage_mismatch"] += not np.array_equal(ml[q], det["ransac_mask"])
        rel = float(np.abs(H[q] - Hr).max() / np.abs(Hr).max())
        if n <= 128:   # refinement summed in OpenCV's order: the refined H must be the oracle's, bit for bit
            stats["h_small"] += 1
            stats["h_small_bit_equal"] += int(np.array_equal(H[q], Hr))
        if rel < 1e-6:
            stats["h_final_mask_mismatch"] += not np.array_equal(mask[q], mr.ravel())
        stats["h_H_worst"] = max(stats["h_H_worst"], rel)
for batch in range(25):
    n = int(rng.choice([5, 6, 8, 12, 20, 50, 200, 1000]))
    Q = 24
    thr = float(rng.choice([4.0, 8.0, 30.0]))
    obj, img, Ks = np.zeros((Q, n, 3)), np.zeros((Q, n, 2)), np.zeros((Q, 3, 3))
    for q in range(Q):
        obj[q], img[q], _ = synth.pnp_set(n, float(rng.uniform(0, 0.6)), rng, noise_px=float(rng.choice([0.0, 0.5, 2.0])))
        Ks[q] = synth.K_1898
        Ks[q, 0, 0] *= rng.uniform(0.7, 1.3); Ks[q, 1, 1] *= rng.uniform(0.7, 1.3)
    okp, rv, tv, inl, infos = ctx.solve_pnp_ransac_batch(obj, img, Ks, 1000, thr, 0.99)
    for q in range(Q):
        o_ok, r_o, t_o, inl_o, det = oracle.solve_pnp_ransac(obj[q], img[q], Ks[q], 1000, thr, 0.99, details=True)
        stats["p_problems"] += 1
        if bool(okp[q]) != o_ok:
            stats["p_ok_mismatch"] += 1
            continue
        if not o_ok:
            stats["p_none"] += 1
            continue
        stats["p_iters_mismatch"] += infos[q]["iters_run"] != det["iters"]
        stats["p_inlier_mismatch"] += not np.array_equal(inl[q], inl_o.ravel())
        rel = max(float(np.abs(rv[q] - r_o.ravel()).max() / np.abs(r_o).max()), float(np.abs(tv[q] - t_o.ravel()).max() / np.abs(t_o).max()))
        stats["p_pose_worst"] = max(stats["p_pose_worst"], rel)
stats["seconds"] = time.time() - t0
stats["seed"] = seed
print(json.dumps(stats))
