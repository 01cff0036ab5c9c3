"""Is the refined H of small problems bit-identical to the CPU restatement?  (development probe; the checks live in tests/)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import ransac_b200
from ransac_b200 import pipeline, synth

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(root, "tests", "golden", "cv2_golden.json")))
ctx = ransac_b200.Context(0)
out = {}
s = g["fixture_a_sweep"]
pos3d, pixels, loc3ds = np.array(s["pos3d"]), np.array(s["pixels"]), np.array(s["loc3ds"])
pos2 = pipeline.candidate_pos2(pos3d[None], loc3ds[:, None, :])
for name, refine in (("cv", 1), ("parallel", 2)):
    H, ok, mask, infos = ctx.find_homography_batch(pos2, pixels, s["thr"], refine=refine)
    eq = 0; worst = 0.0; worst_cv = 0.0; lm_mis = 0
    for q in range(len(loc3ds)):
        Hr, mr, det = oracle.find_homography(pos2[q], pixels, s["thr"], details=True)
        eq += int(np.array_equal(H[q], Hr))
        worst = max(worst, float(np.abs(H[q] - Hr).max() / np.abs(Hr).max()))
        worst_cv = max(worst_cv, float(np.abs(H[q] - np.array(s["H"][q])).max() / np.abs(np.array(s["H"][q])).max()))
    out["sweep_" + name] = dict(bit_equal=eq, of=len(loc3ds), worst_vs_oracle=worst, worst_vs_cv2=worst_cv)
for name, refine in (("cv", 1), ("parallel", 2)):
    eq = 0; worst = 0.0; worst_cv = 0.0
    for b in g["debug_log"]:
        p2, p1 = np.array(b["pos2"]), np.array(b["p1"])
        H, mask, info = ctx.find_homography(p2, p1, 120.0, refine=refine)
        Hr, mr = oracle.find_homography(p2, p1, 120.0)
        eq += int(np.array_equal(H, Hr))
        worst = max(worst, float(np.abs(H - Hr).max() / np.abs(Hr).max()))
        if "cv413_H" in b:
            Hc = np.array(b["cv413_H"]); worst_cv = max(worst_cv, float(np.abs(H - Hc).max() / np.abs(Hc).max()))
    out["debug_log_" + name] = dict(bit_equal=eq, of=len(g["debug_log"]), worst_vs_oracle=worst, worst_vs_cv2=worst_cv)
rng = np.random.default_rng(11)
eq = tot = 0; worst = 0.0
for t in range(30):
    n = int(rng.choice([5, 6, 8, 12, 20, 33, 64, 100, 128])); Q = 20
    thr = float(rng.choice([1.0, 3.0, 10.0, 75.0]))
    src, dst = np.zeros((Q, n, 2)), np.zeros((Q, n, 2))
    for q in range(Q):
        src[q], dst[q], _ = synth.homography_set(n, float(rng.uniform(0, 0.6)), rng, noise_px=float(rng.choice([0.0, 0.5, 2.0])))
    H, ok, mask, infos = ctx.find_homography_batch(src, dst, thr)
    for q in range(Q):
        Hr, mr = oracle.find_homography(src[q], dst[q], thr)
        if Hr is None or not ok[q]:
            continue
        tot += 1; e = np.array_equal(H[q], Hr); eq += int(e)
        worst = max(worst, float(np.abs(H[q] - Hr).max() / np.abs(Hr).max()))
out["random_small_cv"] = dict(bit_equal=eq, of=tot, worst_vs_oracle=worst)
print(json.dumps(out, indent=1))
