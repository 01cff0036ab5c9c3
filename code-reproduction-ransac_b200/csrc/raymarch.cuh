// Downstream georeferencing: the reference turns every polygon vertex of the annotation file into a
// ray and marches it over the DEM in 1 m steps,
//     pixel_to_ray            /root/reference/main_v1.py:547-574
//     calculate_weights       /root/reference/main_v1.py:577-596   (+ weighted_average_optimization_factors :627-632)
//     ray_intersect_dem       /root/reference/main_v1.py:635-658   (10 000 steps x pyproj UTM->WGS84 + DEM bilinear lookup)
//     pixel_to_geo            /root/reference/main_v1.py:661-684
//     convert_boundary_to_geo /root/reference/main_v1.py:765-785
// Here: k_pixel_rays (one thread per pixel: weights, corrected ray direction) and k_ray_march_dem (one CTA per ray).
//
// The march is sequential only in its positions: p_{s+1} = p_s + step * dir, one rounded addition per coordinate per step
// (the reference accumulates in place, so p_s is NOT origin + s * step * dir).  Three threads walk the three coordinates of
// a chunk of steps into shared memory; every thread of the CTA then evaluates one step of the chunk — inverse transverse
// Mercator (Krueger series, constants from geo.py), bounds test, bilinear DEM lookup as scipy's RegularGridInterpolator
// does it, the reference's `step_count >= 150 and z <= dem` rule — and the first event in step order wins.  The returned
// point is the accumulated position itself, so it is bit-identical to a sequential walk whenever the hit step agrees.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace b2r {

struct UtmInverse {   // geo.utm_series_constants()
    double k0A, lon0, FE, FN, beta[6], delta[6];
};

struct DemGrid {      // ascending axes (scipy flips descending ones at construction), values[ny][nx]
    const double* gy;   // latitude axis
    const double* gx;   // longitude axis
    const double* values;
    int ny, nx;
};

constexpr int RM_THREADS = 256;   // steps per chunk = threads per CTA
constexpr int RM_MAX_CTRL = 256;  // control points held per pixel in k_pixel_rays

// EPSG:32650 -> EPSG:4326 (always_xy: lon, lat in degrees).  The six-term series are summed with the multiple angles
// from the addition theorems: one sincos + one sinh/cosh pair per series instead of six.
__device__ __forceinline__ void utm_to_lonlat(const UtmInverse& u, double E, double N, double& lon_deg, double& lat_deg) {
    const double xi = (N - u.FN) / u.k0A, eta = (E - u.FE) / u.k0A;
    double s2, c2;
    sincos(2.0 * xi, &s2, &c2);
    const double sh2 = sinh(2.0 * eta), ch2 = cosh(2.0 * eta);
    double sj = s2, cj = c2, shj = sh2, chj = ch2, xip = xi, etap = eta;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        xip = xip - u.beta[j] * sj * chj;
        etap = etap - u.beta[j] * cj * shj;
        const double sn = sj * c2 + cj * s2, cn = cj * c2 - sj * s2;
        const double shn = shj * ch2 + chj * sh2, chn = chj * ch2 + shj * sh2;
        sj = sn; cj = cn; shj = shn; chj = chn;
    }
    const double chi = asin(sin(xip) / cosh(etap));
    double t2, d2;
    sincos(2.0 * chi, &t2, &d2);
    double tj = t2, dj = d2, lat = chi;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        lat = lat + u.delta[j] * tj;
        const double tn = tj * d2 + dj * t2, dn = dj * d2 - tj * t2;
        tj = tn; dj = dn;
    }
    const double lon = u.lon0 + atan2(sinh(etap), cos(xip));
    const double r2d = 180.0 / 3.14159265358979323846;
    lon_deg = lon * r2d;
    lat_deg = lat * r2d;
}

// scipy find_indices: i with grid[i] <= x < grid[i+1], clipped to [0, n-2]; norm = (x - grid[i]) / (grid[i+1] - grid[i])
__device__ __forceinline__ int find_interval(const double* __restrict__ g, int n, double x, double& norm) {
    int lo = 0, hi = n - 1;           // invariant: g[lo] <= x (x is inside the bounds here), answer in [lo, hi)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g + mid) <= x) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    if (lo < 0) lo = 0;
    const double g0 = __ldg(g + lo), den = __ldg(g + lo + 1) - g0;
    norm = den == 0.0 ? 0.0 : (x - g0) / den;
    return lo;
}

// event codes of one step
constexpr int RM_NONE = 0, RM_HIT = 1, RM_OUTSIDE = 2;

__device__ __forceinline__ int march_test(const UtmInverse& u, const DemGrid& d, double E, double N, double z, int s, int min_steps) {
    double lon, lat;
    utm_to_lonlat(u, E, N, lon, lat);
    // RegularGridInterpolator(bounds_error=True): a point outside the grid raises -> the reference returns None
    if (lat < __ldg(d.gy) || lat > __ldg(d.gy + d.ny - 1) || lon < __ldg(d.gx) || lon > __ldg(d.gx + d.nx - 1)) return RM_OUTSIDE;
    if (!(lat == lat) || !(lon == lon)) return RM_NONE;   // f(nan) = nan: the comparison below is False
    double y0, y1;
    const int i0 = find_interval(d.gy, d.ny, lat, y0), i1 = find_interval(d.gx, d.nx, lon, y1);
    const double* v = d.values + (size_t)i0 * d.nx + i1;
    const double elev = __ldg(v) * (1 - y0) * (1 - y1) + __ldg(v + 1) * (1 - y0) * y1 + __ldg(v + d.nx) * y0 * (1 - y1) +
                        __ldg(v + d.nx + 1) * y0 * y1;
    return (s >= min_steps && z <= elev) ? RM_HIT : RM_NONE;
}

// origins/dirs: [m][3] (origin_stride 0: one origin shared by all rays).  status: 0 hit (geo_out = the accumulated position),
// 1 no intersection within n_steps, 2 a step left the DEM (the interpolator raises in the reference).  hit_step: the step.
__global__ void __launch_bounds__(RM_THREADS)
k_ray_march_dem(const double* __restrict__ origins, int origin_stride, const double* __restrict__ dirs, UtmInverse u, DemGrid d,
                int n_steps, double step, int min_steps, double* __restrict__ geo_out, int* __restrict__ hit_step_out,
                int* __restrict__ status_out) {
    __shared__ double pos[3][RM_THREADS];
    __shared__ double carry[3], inc[3];
    __shared__ int first_event[RM_THREADS / 32];
    const int ray = blockIdx.x, tid = threadIdx.x;
    if (tid < 3) {
        carry[tid] = origins[(size_t)ray * origin_stride + tid];
        inc[tid] = step * dirs[(size_t)ray * 3 + tid];      // `step * ray_direction[k]`, formed anew at every step: the same value
    }
    __syncthreads();
    for (int base = 0; base < n_steps; base += RM_THREADS) {
        if (tid < 3) {   // positions of steps base .. base + RM_THREADS - 1, one rounded addition per step
            double p = carry[tid];
            const double a = inc[tid];
            for (int j = 0; j < RM_THREADS; ++j) {
                pos[tid][j] = p;
                p += a;
            }
            carry[tid] = p;
        }
        __syncthreads();
        const int s = base + tid;
        int code = RM_NONE;
        if (s < n_steps) code = march_test(u, d, pos[0][tid], pos[1][tid], pos[2][tid], s, min_steps);
        const unsigned ev = __ballot_sync(0xffffffffu, code != RM_NONE);
        if ((tid & 31) == 0) first_event[tid >> 5] = ev ? (tid + __ffs(ev) - 1) : RM_THREADS;
        __syncthreads();
        int first = RM_THREADS;
#pragma unroll
        for (int w = 0; w < RM_THREADS / 32; ++w) first = min(first, first_event[w]);
        if (first < RM_THREADS) {
            if (tid == first) {
                status_out[ray] = code == RM_HIT ? 0 : 2;
                hit_step_out[ray] = s;
                geo_out[(size_t)ray * 3 + 0] = pos[0][tid];
                geo_out[(size_t)ray * 3 + 1] = pos[1][tid];
                geo_out[(size_t)ray * 3 + 2] = pos[2][tid];
            }
            return;
        }
        __syncthreads();   // pos is rewritten by the next chunk
    }
    if (tid == 0) {
        status_out[ray] = 1;
        hit_step_out[ray] = n_steps;
        geo_out[(size_t)ray * 3 + 0] = geo_out[(size_t)ray * 3 + 1] = geo_out[(size_t)ray * 3 + 2] = 0.0;
    }
}

// pixel_to_geo up to the march: the weighted z factor of the control points (calculate_weights: min(1/d, 1), the nearest one
// x 10; normalised; np.average over the factors) and the corrected, re-normalised ray direction (pixel_to_ray with K^-1 from
// the host's np.linalg.inv, as the reference; R^T; z component scaled).  One thread per pixel.
__global__ void k_pixel_rays(const double* __restrict__ pixels, int m, const double* __restrict__ Kinv, const double* __restrict__ R,
                             const double* __restrict__ ctrl_pixels, const double* __restrict__ ctrl_factors, int n_ctrl,
                             double max_weight, double knn_weight, double* __restrict__ dirs_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double px = pixels[2 * i], py = pixels[2 * i + 1];
    // weights
    double wsum = 0, fz = 0, dmin = 0;
    int imin = -1;
    for (int c = 0; c < n_ctrl; ++c) {
        const double dx = px - ctrl_pixels[2 * c], dy = py - ctrl_pixels[2 * c + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        if (imin < 0 || dist < dmin) { dmin = dist; imin = c; }   // np.argmin: the first minimum
    }
    for (int c = 0; c < n_ctrl; ++c) {
        const double dx = px - ctrl_pixels[2 * c], dy = py - ctrl_pixels[2 * c + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        double w = fmin(dist != 0.0 ? 1.0 / dist : 1.0, max_weight);
        if (c == imin) w *= knn_weight;
        wsum += w;
    }
    double nsum = 0;
    for (int c = 0; c < n_ctrl; ++c) {
        const double dx = px - ctrl_pixels[2 * c], dy = py - ctrl_pixels[2 * c + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        double w = fmin(dist != 0.0 ? 1.0 / dist : 1.0, max_weight);
        if (c == imin) w *= knn_weight;
        const double nw = w / wsum;                  // normalized_weights
        nsum += nw;
        fz += ctrl_factors[3 * c + 2] * nw;          // np.average: sum(a * w) / sum(w)
    }
    fz = fz / nsum;
    // pixel_to_ray
    double cr[3], ur[3];
    for (int k = 0; k < 3; ++k) cr[k] = Kinv[3 * k] * px + Kinv[3 * k + 1] * py + Kinv[3 * k + 2] * 1.0;
    double nrm = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    for (int k = 0; k < 3; ++k) cr[k] /= nrm;
    for (int k = 0; k < 3; ++k) ur[k] = R[k] * cr[0] + R[3 + k] * cr[1] + R[6 + k] * cr[2];   // R.T @ camera_ray
    nrm = sqrt(ur[0] * ur[0] + ur[1] * ur[1] + ur[2] * ur[2]);
    for (int k = 0; k < 3; ++k) ur[k] /= nrm;
    ur[2] = ur[2] * fz;
    nrm = sqrt(ur[0] * ur[0] + ur[1] * ur[1] + ur[2] * ur[2]);
    for (int k = 0; k < 3; ++k) dirs_out[3 * i + k] = ur[k] / nrm;
}

}  // namespace b2r
