"""EPnP kernel variants (development probe): python tools/perf_epnp_variants.py tools/variants/lib_*.so"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
lib = sys.argv[1]
import ransac_b200
from ransac_b200 import _build, synth
_build.LIB_PATH = os.path.abspath(lib)
_build.needs_build = lambda: False
P, px, _ = synth.pnp_set(100000, 0.5, np.random.default_rng(1))
ctx = ransac_b200.Context(0)
prob = ctx.upload_pnp(P, px, synth.K_1898)
p = ransac_b200.make_p_params(8.0, 100000, 0.99, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=ransac_b200.ARITH_FAST)
best = 1e9
for _ in range(4):
    prob.run(p); prob.fetch(want_inliers=False)
    best = min(best, prob.stage_ms()["sample_solve"])
print(json.dumps({"lib": os.path.basename(lib), "epnp_100k_ms": best}))
