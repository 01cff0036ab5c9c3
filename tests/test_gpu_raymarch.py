"""GPU: row f4 — the DEM ray-march (b2r_ray_march_dem / b2r_pixels_to_geo through the C ABI, host mirror georef.py) against
the CPU oracle (oracle/raymarch.py = the reference's loop, main_v1.py:635-684, with scipy's RegularGridInterpolator).

Bar: status and hit step identical; for supplied directions the returned point is the sequentially accumulated position,
hence bit-identical; through pixel_to_geo (direction formed on the device) within 1e-6 m."""
import numpy as np
import pytest

from raymarch_scene import scene

pytestmark = pytest.mark.gpu


def _rays(rng, origin, m):
    az, el = rng.uniform(0, 2 * np.pi, m), rng.uniform(-0.4, 0.5, m)
    el[::3] = rng.uniform(0.6, 1.4, len(el[::3]))        # a third of the rays point at the sky: they leave the grid
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    o = origin[None] + np.c_[rng.uniform(-50, 50, (m, 2)), rng.uniform(0.0, 80.0, m)]
    return o, d


def test_ray_march_matches_oracle(ctx, oracle):
    from oracle import raymarch as rm
    from ransac_b200 import georef
    sc = scene(oracle)
    dem = georef.dem_handle(sc["dem_data"], ctx)
    o, d = _rays(np.random.default_rng(11), sc["ray_origin"], 96)
    geo, hit, status = ctx.ray_march_dem(dem, o, d)
    counts = {0: 0, 1: 0, 2: 0}
    for i in range(len(d)):
        p, s, st = rm.ray_intersect_dem(o[i], d[i], sc["dem_data"])
        assert status[i] == st and hit[i] == s, (i, status[i], st, hit[i], s)
        if st == 0:
            np.testing.assert_array_equal(geo[i], p)        # the accumulated position, bit for bit
        counts[st] += 1
    assert counts[0] >= 30 and counts[2] >= 10, counts
    # shorter search, other step lengths: int(max_search_dist / step) steps, `>= 150 steps` rule unchanged
    for msd, step in ((900.0, 1.0), (4000.0, 2.5), (100.0, 1.0), (160.0, 1.0)):
        geo2, hit2, st2 = ctx.ray_march_dem(dem, o[:24], d[:24], max_search_dist=msd, step=step)
        for i in range(24):
            p, s, st = rm.ray_intersect_dem(o[i], d[i], sc["dem_data"], max_search_dist=msd, step=step)
            assert st2[i] == st and hit2[i] == s
            if st == 0:
                np.testing.assert_array_equal(geo2[i], p)
    # one shared origin; the reference's single-ray signature
    geo3, hit3, st3 = ctx.ray_march_dem(dem, sc["ray_origin"], d[:8])
    for i in range(8):
        p, s, st = rm.ray_intersect_dem(sc["ray_origin"], d[i], sc["dem_data"])
        assert st3[i] == st and hit3[i] == s
        r = georef.ray_intersect_dem(sc["ray_origin"], d[i], sc["dem_data"], ctx=ctx)
        assert (r is None) == (p is None)
        if p is not None:
            np.testing.assert_array_equal(r, p)


def test_boundary_polygon_to_geo(ctx, oracle):
    """convert_boundary_to_geo (main_v1.py:765-785) over the 21 polygon vertices of 1898.json, pose and control points of
    the repo's own data, synthetic DEM: one batched call against pixel_to_geo of the oracle vertex by vertex."""
    from oracle import raymarch as rm
    from ransac_b200 import georef
    sc = scene(oracle)
    K, R, origin, dd = sc["K"], sc["R"], sc["ray_origin"], sc["dem_data"]
    cps = [dict(pixel=sc["pixels"][i], pos3d=sc["pos3d"][i], symbol=str(i)) for i in range(len(sc["pixels"]))]
    factors = georef.compute_optimization_factors(cps, K, R, origin)
    kept = [c for c in cps if "factors" in c]           # main_v1.py:617-621 filters |factor| > 2
    assert len(kept) >= 6 and len(factors) == len(kept)
    if len(kept) < len(cps):                             # np.average raises in the reference when a point was filtered
        with pytest.raises(ValueError):
            georef.pixels_to_geo([[100.0, 900.0]], K, R, origin, dd, cps, factors, ctx=ctx)
    cpix = [c["pixel"] for c in kept]
    geo_coords, points = georef.convert_boundary_to_geo(sc["boundary"], K, R, origin, dd, kept, factors, ctx=ctx)
    key = (1, "background")                              # re.sub(r'[^a-zA-Z0-9]', '', '__background__')
    assert list(geo_coords) == [key] and len(geo_coords[key]) == 21 == len(points[key])
    verts = sc["boundary"]["objects"][0]["segmentation"]
    res, det = georef.pixels_to_geo(verts, K, R, origin, dd, kept, factors, ctx=ctx, return_details=True)
    for i, (px, py) in enumerate(verts):
        p, s, st, d = rm.pixel_to_geo([px, py], K, R, origin, dd, cpix, factors)
        assert st == 0 and det["status"][i] == 0
        assert det["hit_step"][i] == s
        assert np.abs(det["dirs"][i] - d).max() < 1e-14
        assert np.abs(res[i] - p).max() < 1e-6           # metres
        np.testing.assert_array_equal(geo_coords[key][i], res[i])
        assert points[key][i] == (px, py)
    # a pixel above the skyline misses the terrain: pixel_to_geo returns None, convert_boundary_to_geo fails like the reference
    sky = [[1000.0, -3000.0]]
    assert georef.pixel_to_geo(sky[0], K, R, origin, dd, kept, factors, ctx=ctx) is None
    assert rm.pixel_to_geo(sky[0], K, R, origin, dd, cpix, factors)[0] is None
    with pytest.raises(AttributeError):
        georef.convert_boundary_to_geo({"objects": [{"group": 1, "category": "x", "segmentation": sky}]}, K, R, origin, dd, kept,
                                       factors, ctx=ctx)


def test_dem_argument_checks(ctx):
    with pytest.raises(Exception):
        ctx.upload_dem([0.0, 1.0], [0.0, 1.0], np.zeros((3, 2)))
    d = ctx.upload_dem([1.0, 0.0], [0.0, 1.0, 2.0], np.arange(6.0).reshape(2, 3))   # descending latitude axis: flipped
    assert d.shape == (2, 3)
    d.free()
