"""One PnP fast-mode step at the bench shape (profiling target for ncu): python tools/prof_pnp.py [N] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
P, px, _ = synth.pnp_set(N, 0.5, np.random.default_rng(1))
ctx = ransac_b200.Context(0)
prob = ctx.upload_pnp(P, px, synth.K_1898)
for arith in (ransac_b200.ARITH_FAST, ransac_b200.ARITH_EXACT):
    p = ransac_b200.make_p_params(8.0, H, 0.99, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=arith)
    for _ in range(2):
        prob.run(p)
        prob.fetch(want_inliers=False)
print("ok", prob.stage_ms())
