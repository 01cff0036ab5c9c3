import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import ransac_b200
from ransac_b200 import synth
ctx = ransac_b200.Context(0)
src, dst, _ = synth.homography_set(1000, 0.3, np.random.default_rng(1899))
prob = ctx.upload(src[None], dst)
par = ransac_b200.make_params(3.0, 10000, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=ransac_b200.ARITH_FAST, solver=ransac_b200.SOLVER_FAST)
for _ in range(3):
    prob.run(par); prob.fetch()
print(prob.stage_ms())
