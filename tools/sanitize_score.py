"""The fast and the filtered exact scoring kernels alone on small ragged shapes (input for compute-sanitizer --tool memcheck / racecheck: each
CTA re-scales its shared-memory copy of the point tile between the TMA arrival and the scoring loop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
ctx = ransac_b200.Context(0)
rng = np.random.default_rng(0)
for n, m in ((5, 3), (1025, 700), (2500, 1100)):
    s, d, _ = synth.homography_set(n, 0.4, rng)
    models = rng.normal(0, 1, (m, 8)).astype(np.float32)
    ctx.score_h(models, s.astype(np.float32), d.astype(np.float32), np.float32(9.0), ransac_b200.ARITH_FAST)
    if n >= 64:   # the filtered predicate: mostly deferred evaluations (random models), pathological rows, a non-finite point
        models[::7] = np.nan
        models[1::97] *= np.float32(1e20)
        ctx.score_h(models, s.astype(np.float32), d.astype(np.float32), np.float32(9.0), ransac_b200.ARITH_EXACT)
        s2 = s.astype(np.float32).copy(); s2[n // 2, 0] = np.inf
        ctx.score_h(models, s2, d.astype(np.float32), np.float32(9.0), ransac_b200.ARITH_EXACT)
    P, px, _ = synth.pnp_set(n, 0.4, rng)
    R0, t0 = synth.look_at_pose()
    poses = np.tile(np.concatenate([R0.ravel(), t0]), (m, 1)) + rng.normal(0, 1e-3, (m, 12))
    ctx.score_p(poses, P, px, synth.K_1898, np.float32(64.0), ransac_b200.ARITH_FAST)
    if n >= 64:
        poses[::5, 9:] += rng.normal(0, 300.0, (len(poses[::5]), 3))
        poses[3::50, :9] *= 1e20
        ctx.score_p(poses, P, px, synth.K_1898, np.float32(64.0), ransac_b200.ARITH_EXACT)
print("ok")
