"""Finalize-stage time of single problems for the launch-shape rule of the cooperative-grid form (development probe)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
ctx = ransac_b200.Context(0)
out = {}
for n in (2000, 5000, 10000, 20000, 50000, 100000, 1000000):
    src, dst, _ = synth.homography_set(n, 0.5, np.random.default_rng(1899))
    prob = ctx.upload(src[None], dst)
    par = ransac_b200.make_params(3.0, 2000, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
    ts = []
    for _ in range(8):
        prob.run(par); prob.fetch(); ts.append(prob.stage_ms()["finalize"])
    out[n] = round(float(np.median(ts[2:])), 4)
    prob.free()
print(json.dumps({"ppt": os.environ.get("B2R_FIN_PPT"), "grid_min_n": os.environ.get("B2R_FIN_GRID_MIN_N"), "finalize_ms": out}))
