"""Downstream georeferencing of the reference with the DEM ray-march on the GPU.

Host-side mirror — same names, argument meaning and return conventions — of
    pixel_to_ray                          /root/reference/main_v1.py:547-574
    calculate_weights                     /root/reference/main_v1.py:577-596
    compute_optimization_factors          /root/reference/main_v1.py:599-625
    weighted_average_optimization_factors /root/reference/main_v1.py:627-632
    ray_intersect_dem                     /root/reference/main_v1.py:635-658
    pixel_to_geo                          /root/reference/main_v1.py:661-684
    convert_boundary_to_geo               /root/reference/main_v1.py:765-785
The reference walks every ray in a Python loop of up to 10 000 one-metre steps, each step a pyproj transform and a
RegularGridInterpolator call; here ALL rays of a call are marched by one kernel launch (csrc/raymarch.cuh).  `dem_data`
is the reference's dict (main_v1.py:455-462): its 'interpolator' (a RegularGridInterpolator) supplies the grid; the grid
is uploaded once per dem_data and stays resident.  The small per-control-point arithmetic stays in NumPy, as written in
the reference (prints and logging dropped)."""
import re

import numpy as np

from . import api, geo

_DEM_CACHE = {}   # id(dem_data) -> (ctx, api.Dem)


def dem_handle(dem_data, ctx=None):
    """The device-resident DEM of a reference-style dem_data dict (uploaded on first use)."""
    ctx = ctx or api.default_context()
    key = (id(dem_data), id(ctx))
    if key not in _DEM_CACHE:
        interp = dem_data["interpolator"]
        gy, gx = interp.grid                       # scipy keeps ascending axes (descending input is flipped, values too)
        _DEM_CACHE[key] = (dem_data, ctx.upload_dem(gy, gx, np.asarray(interp.values, dtype=np.float64)))
    return _DEM_CACHE[key][1]


def pixel_to_ray(pixel_x, pixel_y, K, R, ray_origin):
    pixel_homogeneous = np.array([pixel_x, pixel_y, 1.0], dtype=np.float64)
    camera_ray = np.linalg.inv(K) @ pixel_homogeneous
    camera_ray /= np.linalg.norm(camera_ray)
    utm_ray = R.T @ camera_ray
    utm_ray /= np.linalg.norm(utm_ray)
    return ray_origin, utm_ray


def calculate_weights(input_pixel, control_points, max_weight=1, knn_weight=10):
    weights = []
    input_pixel = np.array(input_pixel, dtype=np.float64)
    distances = []
    for cp in control_points:
        pixel = np.array(cp["pixel"], dtype=np.float64)
        distance = np.linalg.norm(input_pixel - pixel)
        distances.append(distance)
        weights.append(min(1.0 / distance if distance != 0 else 1.0, max_weight))
    weights[int(np.argmin(distances))] *= knn_weight
    return np.array(weights)


def compute_optimization_factors(control_points, K, R, ray_origin):
    optimization_factors = []
    for cp in control_points:
        true_geo = np.array(cp["pos3d"], dtype=np.float64)
        ideal_direction = true_geo - ray_origin
        norm_ideal = np.linalg.norm(ideal_direction)
        if norm_ideal == 0:
            continue
        ideal_direction /= norm_ideal
        _, computed_ray = pixel_to_ray(cp["pixel"][0], cp["pixel"][1], K, R, ray_origin)
        computed_ray /= np.linalg.norm(computed_ray)
        f = (ideal_direction[0] / computed_ray[0], ideal_direction[1] / computed_ray[1], ideal_direction[2] / computed_ray[2])
        if abs(f[0]) > 2 or abs(f[1]) > 2 or abs(f[2]) > 2:        # outlier filter, main_v1.py:617
            continue
        optimization_factors.append(f)
        cp["factors"] = f
    return optimization_factors


def weighted_average_optimization_factors(factors, weights):
    normalized_weights = weights / np.sum(weights)
    return np.average(factors, axis=0, weights=normalized_weights)


def ray_intersect_dem(ray_origin, ray_direction, dem_data, max_search_dist=10000, step=1, ctx=None):
    """One ray (the reference's signature); returns np.array([E, N, height]) or None."""
    ctx = ctx or api.default_context()
    geo_, hit, status = ctx.ray_march_dem(dem_handle(dem_data, ctx), np.asarray(ray_origin, dtype=np.float64),
                                          np.asarray(ray_direction, dtype=np.float64).reshape(1, 3), max_search_dist, step)
    return geo_[0].copy() if status[0] == 0 else None


def pixels_to_geo(pixel_coords, K, R, ray_origin, dem_data, control_points, optimization_factors, ctx=None, return_details=False):
    """pixel_to_geo for MANY pixels in one launch sequence: list of np.array([E, N, height]) / None, in input order."""
    ctx = ctx or api.default_context()
    if len(optimization_factors) != len(control_points):
        # np.average(factors, weights=...) raises in the reference when a control point was filtered out (main_v1.py:617-619)
        raise ValueError("Length of weights not compatible with specified axis.")
    px = np.asarray(pixel_coords, dtype=np.float64).reshape(-1, 2)
    cp = np.array([c["pixel"] for c in control_points], dtype=np.float64)
    out, hit, status, dirs = ctx.pixels_to_geo(dem_handle(dem_data, ctx), px, np.linalg.inv(K), R, ray_origin, cp,
                                               np.asarray(optimization_factors, dtype=np.float64))
    res = [out[i].copy() if status[i] == 0 else None for i in range(len(px))]
    if return_details:
        return res, dict(hit_step=hit, status=status, dirs=dirs, geo=out)
    return res


def pixel_to_geo(pixel_coord, K, R, ray_origin, dem_data, control_points, optimization_factors, ctx=None):
    return pixels_to_geo([pixel_coord], K, R, ray_origin, dem_data, control_points, optimization_factors, ctx=ctx)[0]


def convert_boundary_to_geo(json_data, K, R, ray_origin, dem_data, control_points, optimization_factors, ctx=None):
    """All polygon vertices of the annotation in ONE batched march; same grouping, filtering and return value as the
    reference (a vertex is kept when `geo_coord.all()`; a missed vertex makes the reference raise AttributeError on None —
    reproduced)."""
    keys, pixels = [], []
    boundary_points, boundary_geo_coords = {}, {}
    for obj in json_data["objects"]:
        key = (obj["group"], re.sub(r"[^a-zA-Z0-9]", "", obj["category"]))
        if key not in boundary_geo_coords:
            boundary_geo_coords[key] = []
            boundary_points[key] = []
        for pixel_x, pixel_y in obj["segmentation"]:
            keys.append(key)
            pixels.append((pixel_x, pixel_y))
    if not pixels:
        return boundary_geo_coords, boundary_points
    coords = pixels_to_geo(pixels, K, R, ray_origin, dem_data, control_points, optimization_factors, ctx=ctx)
    for key, (pixel_x, pixel_y), geo_coord in zip(keys, pixels, coords):
        if geo_coord.all():            # AttributeError on None, as in the reference (main_v1.py:781)
            boundary_geo_coords[key].append(geo_coord)
            boundary_points[key].append((pixel_x, pixel_y))
    return boundary_geo_coords, boundary_points


def utm_to_wgs84(easting, northing):
    """GeoCoordTransformer.utm_to_wgs84 (main_v1.py:50-58) without pyproj: (lon, lat) degrees."""
    lon, lat = geo.utm50n_to_wgs84(easting, northing)
    if np.any(np.isinf(lat)) or np.any(np.isinf(lon)):
        raise ValueError("Invalid WGS84 coordinates")
    return lon, lat
