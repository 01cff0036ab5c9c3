"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's DEM ray-march:

    pixel_to_ray                          /root/reference/main_v1.py:547-574
    calculate_weights                     /root/reference/main_v1.py:577-596
    weighted_average_optimization_factors /root/reference/main_v1.py:627-632
    ray_intersect_dem                     /root/reference/main_v1.py:635-658
    pixel_to_geo                          /root/reference/main_v1.py:661-684

`ray_intersect_dem_literal` is the reference's loop statement for statement (one interpolator call per step, positions
accumulated in place); `ray_intersect_dem` is the same computation vectorised — np.cumsum adds sequentially, so the
positions are the literal loop's bit for bit — and is what the GPU tests compare against at 10 000 steps.
The DEM lookup is scipy's RegularGridInterpolator, exactly what the reference builds (main_v1.py:454).

PARITY UNPINNED for the geodesy: the reference transforms UTM -> WGS84 with pyproj (main_v1.py:50-58), which is absent
from this image and not pinned by the reference; `utm_to_wgs84` restates EPSG:32650 -> EPSG:4326 as the published Krueger
series (Karney 2011, to n^6), checked by its round trip with the forward series (< 1e-13 degrees) and by the forward
check values of SURVEY.md Appendix B.  dem_data.tif is absent too: tests use a synthetic DEM."""
import math

import numpy as np

A_WGS84, F_WGS84, K0, LON0, FE, FN = 6378137.0, 1.0 / 298.257223563, 0.9996, math.radians(117.0), 500000.0, 0.0


def _coefficients():
    n = F_WGS84 / (2.0 - F_WGS84)
    p = [n ** k for k in range(7)]
    A = A_WGS84 / (1.0 + n) * (1.0 + p[2] / 4.0 + p[4] / 64.0 + p[6] / 256.0)
    alpha = [p[1] / 2 - 2 * p[2] / 3 + 5 * p[3] / 16 + 41 * p[4] / 180 - 127 * p[5] / 288 + 7891 * p[6] / 37800,
             13 * p[2] / 48 - 3 * p[3] / 5 + 557 * p[4] / 1440 + 281 * p[5] / 630 - 1983433 * p[6] / 1935360,
             61 * p[3] / 240 - 103 * p[4] / 140 + 15061 * p[5] / 26880 + 167603 * p[6] / 181440,
             49561 * p[4] / 161280 - 179 * p[5] / 168 + 6601661 * p[6] / 7257600,
             34729 * p[5] / 80640 - 3418889 * p[6] / 1995840,
             212378941 * p[6] / 319334400]
    beta = [p[1] / 2 - 2 * p[2] / 3 + 37 * p[3] / 96 - p[4] / 360 - 81 * p[5] / 512 + 96199 * p[6] / 604800,
            p[2] / 48 + p[3] / 15 - 437 * p[4] / 1440 + 46 * p[5] / 105 - 1118711 * p[6] / 3870720,
            17 * p[3] / 480 - 37 * p[4] / 840 - 209 * p[5] / 4480 + 5569 * p[6] / 90720,
            4397 * p[4] / 161280 - 11 * p[5] / 504 - 830251 * p[6] / 7257600,
            4583 * p[5] / 161280 - 108847 * p[6] / 3991680,
            20648693 * p[6] / 638668800]
    delta = [2 * p[1] - 2 * p[2] / 3 - 2 * p[3] + 116 * p[4] / 45 + 26 * p[5] / 45 - 2854 * p[6] / 675,
             7 * p[2] / 3 - 8 * p[3] / 5 - 227 * p[4] / 45 + 2704 * p[5] / 315 + 2323 * p[6] / 945,
             56 * p[3] / 15 - 136 * p[4] / 35 - 1262 * p[5] / 105 + 73814 * p[6] / 2835,
             4279 * p[4] / 630 - 332 * p[5] / 35 - 399572 * p[6] / 14175,
             4174 * p[5] / 315 - 144838 * p[6] / 6237,
             601676 * p[6] / 22275]
    return A, alpha, beta, delta


_A, _ALPHA, _BETA, _DELTA = _coefficients()


def wgs84_to_utm(lon_deg, lat_deg):
    """EPSG:4326 -> EPSG:32650 (forward Krueger series); used only to close the round trip of utm_to_wgs84."""
    lon, lat = np.radians(np.asarray(lon_deg, dtype=np.float64)), np.radians(np.asarray(lat_deg, dtype=np.float64))
    e = math.sqrt(F_WGS84 * (2.0 - F_WGS84))
    t = np.sinh(np.arctanh(np.sin(lat)) - e * np.arctanh(e * np.sin(lat)))
    xi, eta = np.arctan2(t, np.cos(lon - LON0)), np.arctanh(np.sin(lon - LON0) / np.sqrt(1.0 + t * t))
    x, y = eta.copy(), xi.copy()
    for j, a in enumerate(_ALPHA, start=1):
        x = x + a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        y = y + a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return FE + K0 * _A * x, FN + K0 * _A * y


def utm_to_wgs84(easting, northing):
    """EPSG:32650 -> EPSG:4326, (lon, lat) in degrees (GeoCoordTransformer.utm_to_wgs84, main_v1.py:50-58)."""
    xi = (np.asarray(northing, dtype=np.float64) - FN) / (K0 * _A)
    eta = (np.asarray(easting, dtype=np.float64) - FE) / (K0 * _A)
    xi_p, eta_p = xi.copy(), eta.copy()
    for j, b in enumerate(_BETA, start=1):
        xi_p = xi_p - b * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        eta_p = eta_p - b * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    chi = np.arcsin(np.sin(xi_p) / np.cosh(eta_p))
    lat = chi.copy()
    for j, d in enumerate(_DELTA, start=1):
        lat = lat + d * np.sin(2 * j * chi)
    return np.degrees(LON0 + np.arctan2(np.sinh(eta_p), np.cos(xi_p))), np.degrees(lat)


def make_dem_data(dem_y, dem_x, dem_array):
    """The dict load_dem_data builds (main_v1.py:454-462), from grid axes instead of a GeoTIFF."""
    from scipy.interpolate import RegularGridInterpolator
    dem_y, dem_x = np.asarray(dem_y, dtype=np.float64), np.asarray(dem_x, dtype=np.float64)
    return {"interpolator": RegularGridInterpolator((dem_y, dem_x), np.asarray(dem_array, dtype=np.float64)),
            "x_range": (dem_x.min(), dem_x.max()), "y_range": (dem_y.min(), dem_y.max()), "data": dem_array}


def ray_intersect_dem_literal(ray_origin, ray_direction, dem_data, max_search_dist=10000, step=1):
    """main_v1.py:635-658, statement for statement (logging dropped).  Returns (point or None, step index, status)."""
    current_pos = np.array(ray_origin, dtype=np.float64)
    step_count = 0
    for _ in range(int(max_search_dist / step)):
        lon, lat = utm_to_wgs84(current_pos[0], current_pos[1])
        try:
            dem_elev = dem_data["interpolator"]((lat, lon))
        except Exception:
            return None, step_count, 2
        if step_count >= 150 and current_pos[2] <= dem_elev:
            return np.array([current_pos[0], current_pos[1], current_pos[2]]), step_count, 0
        current_pos[0] += step * ray_direction[0]
        current_pos[1] += step * ray_direction[1]
        current_pos[2] += step * ray_direction[2]
        step_count += 1
    return None, step_count, 1


def ray_intersect_dem(ray_origin, ray_direction, dem_data, max_search_dist=10000, step=1):
    """The same walk vectorised: positions by np.cumsum (sequential additions = the loop's in-place accumulation), one
    interpolator call for all steps.  Returns (point or None, step index, status: 0 hit, 1 none, 2 left the DEM)."""
    n = int(max_search_dist / step)
    if n == 0:
        return None, 0, 1
    o, d = np.array(ray_origin, dtype=np.float64), np.asarray(ray_direction, dtype=np.float64)
    pos = np.empty((n, 3))
    for k in range(3):
        inc = np.full(n, step * d[k])
        inc[0] = o[k]
        pos[:, k] = np.cumsum(inc)
    lon, lat = utm_to_wgs84(pos[:, 0], pos[:, 1])
    interp = dem_data["interpolator"]
    gy, gx = interp.grid
    outside = (lat < gy[0]) | (lat > gy[-1]) | (lon < gx[0]) | (lon > gx[-1])
    inside = ~outside
    elev = np.full(n, np.nan)
    if inside.any():
        elev[inside] = interp(np.stack([lat[inside], lon[inside]], axis=1))
    with np.errstate(invalid="ignore"):
        hit = inside & (np.arange(n) >= 150) & (pos[:, 2] <= elev)
    event = hit | outside
    if not event.any():
        return None, n, 1
    s = int(np.argmax(event))
    if outside[s]:
        return None, s, 2
    return pos[s].copy(), s, 0


def pixel_to_ray(pixel_x, pixel_y, K, R, ray_origin):
    camera_ray = np.linalg.inv(K) @ np.array([pixel_x, pixel_y, 1.0], dtype=np.float64)
    camera_ray /= np.linalg.norm(camera_ray)
    utm_ray = R.T @ camera_ray
    utm_ray /= np.linalg.norm(utm_ray)
    return ray_origin, utm_ray


def calculate_weights(input_pixel, control_pixels, max_weight=1, knn_weight=10):
    input_pixel = np.array(input_pixel, dtype=np.float64)
    distances = [np.linalg.norm(input_pixel - np.array(p, dtype=np.float64)) for p in control_pixels]
    weights = [min(1.0 / d if d != 0 else 1.0, max_weight) for d in distances]
    weights[int(np.argmin(distances))] *= knn_weight
    return np.array(weights)


def final_ray_direction(pixel_coord, K, R, ray_origin, control_pixels, optimization_factors):
    """pixel_to_geo up to the march (main_v1.py:663-681)."""
    weights = calculate_weights(pixel_coord, control_pixels)
    wf = np.average(np.asarray(optimization_factors, dtype=np.float64), axis=0, weights=weights / np.sum(weights))
    _, d = pixel_to_ray(pixel_coord[0], pixel_coord[1], K, R, ray_origin)
    o = np.array([d[0], d[1], d[2] * wf[2]])
    return o / np.linalg.norm(o)


def pixel_to_geo(pixel_coord, K, R, ray_origin, dem_data, control_pixels, optimization_factors, literal=False):
    d = final_ray_direction(pixel_coord, K, R, ray_origin, control_pixels, optimization_factors)
    f = ray_intersect_dem_literal if literal else ray_intersect_dem
    return f(ray_origin, d, dem_data) + (d,)
