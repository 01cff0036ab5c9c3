import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_b200
g = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cv2_golden.json")))["fixture_a_sweep"]
ctx = ransac_b200.Context(0)
for _ in range(3):
    r = ctx.camera_sweep(np.array(g["pos3d"]), np.array(g["pixels"]), np.array(g["loc3ds"]), g["thr"])
print(r["best"])
