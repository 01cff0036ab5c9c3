"""WGS84 lon/lat <-> UTM zone 50N (EPSG:4326 <-> EPSG:32650), the transforms the reference applies through pyproj
(GeoCoordTransformer, main_v1.py:36-59): forward on both CSVs (main_v1.py:717, :752), inverse at every step of the DEM
ray-march (main_v1.py:641).

Transverse Mercator by the Krueger series in the third flattening n, to n^6 (sub-millimetre inside the zone), so the
reference's CSV inputs can be ingested without pyproj.  Check value (testpro-K.py:199): (119.390036, 26.098989) ->
(739031.1998, 2888840.3870)."""
import csv
import math

import numpy as np

_A = 6378137.0
_F = 1.0 / 298.257223563
_K0 = 0.9996
_LON0 = math.radians(117.0)
_FE, _FN = 500000.0, 0.0


def _series():
    n = _F / (2.0 - _F)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    A = _A / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0)
    alpha = [
        n / 2.0 - 2.0 * n2 / 3.0 + 5.0 * n3 / 16.0 + 41.0 * n4 / 180.0 - 127.0 * n5 / 288.0 + 7891.0 * n6 / 37800.0,
        13.0 * n2 / 48.0 - 3.0 * n3 / 5.0 + 557.0 * n4 / 1440.0 + 281.0 * n5 / 630.0 - 1983433.0 * n6 / 1935360.0,
        61.0 * n3 / 240.0 - 103.0 * n4 / 140.0 + 15061.0 * n5 / 26880.0 + 167603.0 * n6 / 181440.0,
        49561.0 * n4 / 161280.0 - 179.0 * n5 / 168.0 + 6601661.0 * n6 / 7257600.0,
        34729.0 * n5 / 80640.0 - 3418889.0 * n6 / 1995840.0,
        212378941.0 * n6 / 319334400.0,
    ]
    return A, alpha


_A_RECT, _ALPHA = _series()


def _inverse_series():
    """beta_j (rectifying -> conformal sphere) and delta_j (conformal -> geodetic latitude), Krueger 1912 / Karney 2011, to n^6."""
    n = _F / (2.0 - _F)
    n2, n3, n4, n5, n6 = n * n, n ** 3, n ** 4, n ** 5, n ** 6
    beta = [
        n / 2.0 - 2.0 * n2 / 3.0 + 37.0 * n3 / 96.0 - n4 / 360.0 - 81.0 * n5 / 512.0 + 96199.0 * n6 / 604800.0,
        n2 / 48.0 + n3 / 15.0 - 437.0 * n4 / 1440.0 + 46.0 * n5 / 105.0 - 1118711.0 * n6 / 3870720.0,
        17.0 * n3 / 480.0 - 37.0 * n4 / 840.0 - 209.0 * n5 / 4480.0 + 5569.0 * n6 / 90720.0,
        4397.0 * n4 / 161280.0 - 11.0 * n5 / 504.0 - 830251.0 * n6 / 7257600.0,
        4583.0 * n5 / 161280.0 - 108847.0 * n6 / 3991680.0,
        20648693.0 * n6 / 638668800.0,
    ]
    delta = [
        2.0 * n - 2.0 * n2 / 3.0 - 2.0 * n3 + 116.0 * n4 / 45.0 + 26.0 * n5 / 45.0 - 2854.0 * n6 / 675.0,
        7.0 * n2 / 3.0 - 8.0 * n3 / 5.0 - 227.0 * n4 / 45.0 + 2704.0 * n5 / 315.0 + 2323.0 * n6 / 945.0,
        56.0 * n3 / 15.0 - 136.0 * n4 / 35.0 - 1262.0 * n5 / 105.0 + 73814.0 * n6 / 2835.0,
        4279.0 * n4 / 630.0 - 332.0 * n5 / 35.0 - 399572.0 * n6 / 14175.0,
        4174.0 * n5 / 315.0 - 144838.0 * n6 / 6237.0,
        601676.0 * n6 / 22275.0,
    ]
    return beta, delta


_BETA, _DELTA = _inverse_series()


def utm50n_to_wgs84(easting, northing):
    """(lon, lat) in degrees from UTM 50N metres (always_xy order, main_v1.py:50-55).  Krueger series to n^6: the round trip
    with wgs84_to_utm50n closes to < 1e-11 degrees inside the zone."""
    xi = (np.asarray(northing, dtype=np.float64) - _FN) / (_K0 * _A_RECT)
    eta = (np.asarray(easting, dtype=np.float64) - _FE) / (_K0 * _A_RECT)
    xi_p, eta_p = xi.copy(), eta.copy()
    for j, b in enumerate(_BETA, start=1):
        xi_p = xi_p - b * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
        eta_p = eta_p - b * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
    chi = np.arcsin(np.sin(xi_p) / np.cosh(eta_p))
    lat = chi.copy()
    for j, d in enumerate(_DELTA, start=1):
        lat = lat + d * np.sin(2 * j * chi)
    lon = _LON0 + np.arctan2(np.sinh(eta_p), np.cos(xi_p))
    return np.degrees(lon), np.degrees(lat)


def utm_series_constants():
    """The constants the device ray-march needs (b2r_ray_march_dem): k0 * A, lon0 (rad), FE, FN, beta[6], delta[6]."""
    return np.array([_K0 * _A_RECT, _LON0, _FE, _FN] + list(_BETA) + list(_DELTA), dtype=np.float64)


def wgs84_to_utm50n(lon_deg, lat_deg):
    """(easting, northing) in metres; accepts scalars or arrays (always_xy order: lon, lat — main_v1.py:38-43)."""
    lon = np.radians(np.asarray(lon_deg, dtype=np.float64))
    lat = np.radians(np.asarray(lat_deg, dtype=np.float64))
    e = math.sqrt(_F * (2.0 - _F))
    s = np.sin(lat)
    t = np.sinh(np.arctanh(s) - e * np.arctanh(e * s))  # tan of the conformal latitude
    dl = lon - _LON0
    xi = np.arctan2(t, np.cos(dl))
    eta = np.arctanh(np.sin(dl) / np.sqrt(1.0 + t * t))
    x, y = eta.copy(), xi.copy()
    for j, a in enumerate(_ALPHA, start=1):
        x = x + a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        y = y + a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return _FE + _K0 * _A_RECT * x, _FN + _K0 * _A_RECT * y


def read_camera_locations(path):
    """potential_camera_locations.csv -> list of dict(grid_code, pos3d=[E, N, elevation+2.0]); main_v1.py:734-762."""
    out = []
    with open(path, encoding="utf-8") as f:
        rows = csv.reader(f)
        next(rows)
        for row in rows:
            if len(row) < 5 or not row[0].strip():
                continue
            grid_code = int(row[1])
            lon, lat = float(row[2]), float(row[3])
            height = float(row[4]) + 2.0  # observer height, main_v1.py:748
            e, n = wgs84_to_utm50n(lon, lat)
            out.append({"grid_code": grid_code, "pos3d": np.array([float(e), float(n), height])})
    return out


def read_points_data(path, pixel_x, pixel_y, scale=1.0, elevations=None):
    """feature_points_with_annotations.csv -> list of dict(symbol, name, pixel, pos3d); main_v1.py:689-729.

    The shipped CSV has empty Height/Elevation columns (the reference's float('') raises there);
    `elevations` (symbol -> metres) fills them, rows without any elevation are skipped.  (0,0) pixels are dropped
    as in main_v1.py:711."""
    recs = []
    with open(path, encoding="utf-8") as f:
        rows = csv.reader(f)
        names = next(rows)
        ix, iy = names.index(pixel_x), names.index(pixel_y)
        for row in rows:
            if len(row) <= max(ix, iy) or not row[1].strip():
                continue
            symbol, name = row[1], row[2]
            pixel = np.array([int(row[ix]), int(row[iy])]) / scale
            if pixel[0] == 0 and pixel[1] == 0:
                continue
            elev = row[6].strip()
            if elev:
                elevation = float(elev)
            elif elevations is not None and symbol in elevations:
                elevation = float(elevations[symbol])
            else:
                continue
            e, n = wgs84_to_utm50n(float(row[4]), float(row[5]))
            recs.append({"symbol": symbol, "name": name, "pixel": pixel, "pos3d": np.array([float(e), float(n), elevation])})
    return recs
