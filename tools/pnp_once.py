"""One solvePnPRansac + RefineLM on the reference's 12 correspondences, three times (for ncu launch lists)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
from ransac_b200 import pipeline, synth
g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_golden.json")))["fixture_a_sweep"]
ctx = ransac_b200.Context(0)
for _ in range(3):
    r = pipeline.estimate_camera_pose(np.array(g["pos3d"]), np.array(g["pixels"]), synth.K_1898, ctx=ctx)
print(r[0].ravel(), r[1].ravel())
