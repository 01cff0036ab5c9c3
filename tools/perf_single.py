"""Stage breakdown of ONE reference-sized problem (12 annotated points): findHomography and solvePnPRansac (development probe)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import pipeline, synth

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(root, "tests", "golden", "cv2_golden.json")))
s = g["fixture_a_sweep"]
pos3d, pixels, loc3ds = np.array(s["pos3d"]), np.array(s["pixels"]), np.array(s["loc3ds"])
ctx = ransac_b200.Context(0)
pos2 = pipeline.candidate_pos2(pos3d[None], loc3ds[:, None, :])
K = synth.K_1898


def timeit(fn, reps=30):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))


out = {}
for cand in (180, 10, 300):
    prob = ctx.upload(pos2[cand:cand + 1], pixels)
    p = ransac_b200.make_params(75.0)
    def run(): prob.run(p); prob.fetch()
    ms = timeit(run)
    _, _, info = prob.fetch()
    out[f"h12_cand{cand}"] = {"ms": ms, "stage_ms": prob.stage_ms(), "iters_run": info[0]["iters_run"], "lm_iters": info[0].get("lm_iters")}
    out[f"h12_cand{cand}_call_ms"] = timeit(lambda: ctx.find_homography(pos2[cand], pixels, 75.0))
pp = ctx.upload_pnp(pos3d, pixels, K)
q = ransac_b200.make_p_params(30.0, 5000, 0.99)
def runp(): pp.run(q); pp.fetch()
out["pnp12"] = {"ms": timeit(runp), "stage_ms": pp.stage_ms()}
print(json.dumps(out))
