"""Stage timings of the PnP path on one GPU (development probe, not the bench): python tools/perf_pnp.py [N] [H]"""
import os
import sys
import json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import synth
if os.environ.get("B2R_DEV_LIB"):   # development only: time a library built with other -D flags
    import ransac_b200._build as _b
    _b.LIB_PATH = os.path.abspath(os.environ["B2R_DEV_LIB"])

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
rng = np.random.default_rng(1)
P, px, _ = synth.pnp_set(N, 0.5, rng)
ctx = ransac_b200.Context(0)
prob = ctx.upload_pnp(P, px, synth.K_1898)
for arith, name in ((ransac_b200.ARITH_FAST, "fast"), (ransac_b200.ARITH_EXACT, "exact")):
    p = ransac_b200.make_p_params(8.0, H, 0.99, sampler=ransac_b200.SAMPLER_PHILOX, seed=3, arith=arith,
                                  solver=ransac_b200.SOLVER_FAST if arith == ransac_b200.ARITH_FAST else ransac_b200.SOLVER_EXACT)
    best = None
    for rep in range(4):
        prob.run(p)
        r, t, inl, info = prob.fetch(want_inliers=False)
        ms = prob.stage_ms()
        if best is None or ms["total"] < best["total"]:
            best = ms
    out = dict(arith=name, N=N, H=H, stage_ms=best, score_evals_per_s=N * H / (best["score"] * 1e-3),
               step_evals_per_s=N * H / (best["total"] * 1e-3), best_count=info[0]["best_count"], n_inliers=info[0]["n_inliers"],
               lm_iters=info[0]["lm_iters"])
    print(json.dumps(out))
