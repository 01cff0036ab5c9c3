#!/usr/bin/env python
"""Writes tests/golden/boundary_1898.json — run in the BUILD CONTAINER only (needs /root/reference).

The annotation the reference feeds to convert_boundary_to_geo (main_v1.py:961-965): the objects of
/root/reference/1898.json (one `__background__` polygon, 21 vertices, 1898.json:12-95) reduced to the three fields that
function reads (group, category, segmentation) plus the image size.  Input data only — no expected outputs: the
reference cannot run this stage (dem_data.tif and pyproj are absent), so the GPU tests compare against the CPU oracle
(oracle/raymarch.py) on a synthetic DEM."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
with open("/root/reference/1898.json", encoding="utf-8") as f:
    d = json.load(f)
out = {"source": "/root/reference/1898.json", "width": d["info"]["width"], "height": d["info"]["height"],
       "objects": [{"group": o["group"], "category": o["category"], "segmentation": o["segmentation"]} for o in d["objects"]]}
with open(os.path.join(HERE, "boundary_1898.json"), "w") as f:
    json.dump(out, f)
print(len(out["objects"]), [len(o["segmentation"]) for o in out["objects"]])
