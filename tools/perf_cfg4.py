"""configs[4] (4096 problems x 2k points x 2000 hypotheses) and the 458-candidate sweep: stage times (development probe)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
ctx = ransac_b200.Context(0)
c = synth.CONFIGS[4]
s1, d1, _ = synth.homography_set(c["n_points"], c["outliers"], np.random.default_rng(1902))
Q = 4096
src = np.broadcast_to(s1, (Q,) + s1.shape).copy()
dst = d1[None] + np.random.default_rng(7).normal(0, 0.3, (Q,) + d1.shape)
prob = ctx.upload(src, dst)
par = ransac_b200.make_params(3.0, c["hypotheses"], sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
best = None
for _ in range(4):
    prob.run(par); prob.fetch(want_mask=False)
    ms = prob.stage_ms()
    best = ms if best is None or ms["total"] < best["total"] else best
print(json.dumps({"cfg4": best, "step_evals_per_s": Q * 2000.0 * 2000 / (best["total"] * 1e-3)}))
