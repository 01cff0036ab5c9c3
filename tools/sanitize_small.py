"""Small end-to-end calls of every entry point (input for compute-sanitizer --tool memcheck)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import pipeline, synth
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(root, "tests", "golden", "cv2_golden.json")))["fixture_a_sweep"]
pos3d, pixels, loc = np.array(g["pos3d"]), np.array(g["pixels"]), np.array(g["loc3ds"])[:40]
ctx = ransac_b200.Context(0)
r = ctx.camera_sweep(pos3d, pixels, loc, g["thr"])
rng = np.random.default_rng(0)
for n in (5, 37, 1000, 5003):
    s, d, _ = synth.homography_set(n, 0.4, rng)
    ctx.find_homography(s, d, 3.0)
    ctx.find_homography(s, d, 3.0, max_iters=777, sampler=ransac_b200.SAMPLER_PHILOX, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
    P, px, _ = synth.pnp_set(n, 0.4, rng)
    ok, rv, tv, inl, _ = ctx.solve_pnp_ransac(P, px, synth.K_1898, 600, 8.0, 0.99)
    ctx.solve_pnp_ransac(P, px, synth.K_1898, 333, 8.0, 0.99, sampler=ransac_b200.SAMPLER_PHILOX, arith=ransac_b200.ARITH_FAST)
    if ok and len(inl) >= 6:
        ctx.solve_pnp_refine_lm(P[inl.ravel()], px[inl.ravel()], synth.K_1898, rv, tv)
s, d, _ = synth.homography_set(40000, 0.5, rng)
ctx.find_homography(s, d, 3.0, max_iters=3000, sampler=ransac_b200.SAMPLER_PHILOX, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
P, px, _ = synth.pnp_set(40000, 0.5, rng)
ctx.solve_pnp_ransac(P, px, synth.K_1898, 2000, 8.0, 0.99, sampler=ransac_b200.SAMPLER_PHILOX)
Ks, _ = pipeline.intrinsics_grid([90, 240], [(127, 178), (102, 127)], (2142, 1620))
ctx.solve_pnp_ransac_batch(pos3d, pixels, Ks, 5000, 30.0, 0.99)
idx = np.stack([rng.choice(1000, 4, replace=False) for _ in range(65)]).astype(np.int32)
s, d, _ = synth.homography_set(1000, 0.3, rng)
ctx.solve_h4(s.astype(np.float32), d.astype(np.float32), idx); ctx.solve_h4(s.astype(np.float32), d.astype(np.float32), idx, solver=2)
print("sanitize_small ok", r["best"])
