"""Stage times of BASELINE configs[1] (1000 points x 10 000 hypotheses) and of mid-size single problems: python tools/perf_cfg1.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth

ctx = ransac_b200.Context(0)
out = {}
for n, H in ((1000, 10_000), (300, 10_000), (600, 10_000), (2000, 10_000), (5000, 10_000), (20000, 10_000)):
    src, dst, _ = synth.homography_set(n, 0.3, np.random.default_rng(1899))
    prob = ctx.upload(src[None], dst)
    for name, arith, solver in (("fast", ransac_b200.ARITH_FAST, ransac_b200.SOLVER_FAST), ("exact", ransac_b200.ARITH_EXACT, ransac_b200.SOLVER_EXACT)):
        par = ransac_b200.make_params(3.0, H, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=arith, solver=solver)
        for _ in range(4):
            prob.run(par); prob.fetch()
        out[f"n{n}_{name}"] = {k: round(v, 4) for k, v in prob.stage_ms().items()}
print(json.dumps(out))
