"""CPU: the oracle of the DEM ray-march (row f4, oracle/raymarch.py) — geodesy closes, and the vectorised walk the GPU
tests use equals the reference's literal step loop (main_v1.py:635-658) bit for bit."""
import numpy as np

from raymarch_scene import scene


def test_utm_series_round_trip_and_check_values(oracle):
    from oracle import raymarch as rm
    from ransac_b200 import geo
    rng = np.random.default_rng(0)
    lon, lat = rng.uniform(118.0, 121.0, 5000), rng.uniform(24.0, 28.0, 5000)
    e, n = rm.wgs84_to_utm(lon, lat)
    lo, la = rm.utm_to_wgs84(e, n)
    assert np.abs(lo - lon).max() < 1e-12 and np.abs(la - lat).max() < 1e-12
    # SURVEY.md Appendix B check values (pyproj outputs recorded from the reference's own literals, testpro-K.py:199)
    e, n = rm.wgs84_to_utm(119.390036, 26.098989)
    assert abs(e - 739031.1998) < 5e-3 and abs(n - 2888840.3870) < 5e-3
    # the product's host-side series (geo.py) is the same transform
    lo2, la2 = geo.utm50n_to_wgs84(e, n)
    lo3, la3 = rm.utm_to_wgs84(e, n)
    assert abs(lo2 - lo3) < 1e-13 and abs(la2 - la3) < 1e-13


def test_vectorised_walk_equals_the_literal_loop(oracle):
    from oracle import raymarch as rm
    sc = scene(oracle, ny=120, nx=140)
    rng = np.random.default_rng(4)
    seen = set()
    for i in range(14):
        az, el = rng.uniform(0, 2 * np.pi), (rng.uniform(-0.35, 0.45) if i % 3 else rng.uniform(0.6, 1.4))
        d = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        o = sc["ray_origin"] + np.array([0.0, 0.0, rng.uniform(0.0, 60.0)])
        msd = (10000, 7000, 400)[i % 3] if i < 12 else 10000       # 400 m: too short for most hits -> status 1
        a = rm.ray_intersect_dem(o, d, sc["dem_data"], max_search_dist=msd)
        b = rm.ray_intersect_dem_literal(o, d, sc["dem_data"], max_search_dist=msd)
        assert a[1:] == b[1:]
        if a[2] == 0:
            np.testing.assert_array_equal(a[0], b[0])
            assert a[1] >= 150
        else:
            assert a[0] is None and b[0] is None
        seen.add(a[2])
    assert seen == {0, 1, 2}       # hits, exhausted searches and DEM exits all occur
