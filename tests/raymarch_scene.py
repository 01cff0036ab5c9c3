"""Synthetic scene for the DEM ray-march tests (row f4): the reference's dem_data.tif is absent, so the DEM is analytic.

Camera = the pose cv2 returns for the repo's 12 correspondences (golden file), set on the ground + 1.5 m as do_it does
(main_v1.py:910-918); DEM = a GeoTIFF-like grid (latitude axis DESCENDING, as gt[5] < 0 makes it at main_v1.py:432): the
camera stands on a hill (ground = the PnP camera height - 1.5 m), the terrain falls ~100 m towards the landmarks 600-800 m
away (their elevations in testpro-K.py:198-211 are 697-726 m against a camera at 820 m), rises to a ridge behind them, and
carries 10 m hills.  All 21 polygon vertices of 1898.json then hit the terrain — the bottom edge of the image at the
150-step minimum, the skyline vertices 500-900 m out — and a pixel far above the skyline leaves the grid."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(oracle, ny=360, nx=400):
    from oracle import raymarch as rm
    from ransac_b200 import synth
    with open(os.path.join(HERE, "golden", "cv2_golden.json")) as f:
        g = json.load(f)
    p, s = g["pnp_fixture_a"], g["fixture_a_sweep"]
    R = oracle.rodrigues(np.array(p["refined_rvec"]).ravel())
    t = np.array(p["refined_tvec"]).ravel()
    cam = -R.T @ t                                                  # main_v1.py:910
    dem_y = np.linspace(26.150, 26.040, ny)                         # descending, like a north-up GeoTIFF
    dem_x = np.linspace(119.330, 119.450, nx)
    LX, LY = np.meshgrid(dem_x, dem_y)
    E, N = rm.wgs84_to_utm(LX, LY)
    dist = np.hypot(E - cam[0], N - cam[1])
    dem = cam[2] - 1.5 - 0.22 * dist + 0.00011 * dist * dist + 10.0 * np.sin(E / 180.0) * np.cos(N / 230.0)
    dem_data = rm.make_dem_data(dem_y, dem_x, dem)
    lon, lat = rm.utm_to_wgs84(cam[0], cam[1])
    ground = float(dem_data["interpolator"]((lat, lon)))
    ray_origin = np.array([cam[0], cam[1], ground + 1.5])           # main_v1.py:914-918
    pos3d, pixels = np.array(s["pos3d"]), np.array(s["pixels"])
    with open(os.path.join(HERE, "golden", "boundary_1898.json")) as f:
        boundary = json.load(f)
    return dict(K=synth.K_1898.copy(), R=R, t=t, ray_origin=ray_origin, dem_data=dem_data, pos3d=pos3d, pixels=pixels,
                boundary=boundary)
