"""Latency of the reference-sized problems on one GPU (development probe): the 458-candidate sweep of 12 points
(main_v1.py:254-297), one findHomography, one solvePnPRansac + RefineLM, the 27-K grid (testpro-K.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import pipeline, synth

g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_golden.json")))
kg = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_kgrid.json")))
s = g["fixture_a_sweep"]
pos3d, pixels, loc3ds = np.array(s["pos3d"]), np.array(s["pixels"]), np.array(s["loc3ds"])
ctx = ransac_b200.Context(0)
pos2 = pipeline.candidate_pos2(pos3d[None], loc3ds[:, None, :])
K = synth.K_1898


def timeit(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))


out = {}
out["sweep458_ms"] = timeit(lambda: ctx.find_homography_batch(pos2, pixels, 75.0))
prob = ctx.upload(pos2, pixels)
p = ransac_b200.make_params(75.0)
def run(): prob.run(p); prob.fetch()
out["sweep458_resident_ms"] = timeit(run)
out["sweep458_stage_ms"] = prob.stage_ms()
out["find_homography_12pts_ms"] = timeit(lambda: ctx.find_homography(pos2[180], pixels, 75.0))
out["solve_pnp_ransac_12pts_ms"] = timeit(lambda: ctx.solve_pnp_ransac(pos3d, pixels, K, 5000, 30.0, 0.99))
pp = ctx.upload_pnp(pos3d, pixels, K)
q = ransac_b200.make_p_params(30.0, 5000, 0.99)
def runp(): pp.run(q); pp.fetch()
out["pnp_resident_ms"] = timeit(runp)
out["pnp_stage_ms"] = pp.stage_ms()
out["estimate_camera_pose_ms"] = timeit(lambda: pipeline.estimate_camera_pose(pos3d, pixels, K, ctx=ctx))
out["kgrid27_ms"] = timeit(lambda: pipeline.estimate_camera_orientation(pos3d, pixels, kg["focal_lengths"], [tuple(x) for x in kg["sensor_sizes"]], tuple(kg["image_size"]), ctx=ctx), reps=10)
rng = np.random.default_rng(5)
sN, dN, _ = synth.homography_set(100000, 0.5, rng)
out["find_homography_100k_replay_ms"] = timeit(lambda: ctx.find_homography(sN, dN, 3.0), reps=5)
pr = ctx.upload(sN, dN); p3 = ransac_b200.make_params(3.0)
def run3(): pr.run(p3); pr.fetch()
run3(); out["find_homography_100k_replay_stage_ms"] = pr.stage_ms()
print(json.dumps(out))
