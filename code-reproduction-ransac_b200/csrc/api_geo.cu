// C ABI of libransac_b200.so, downstream georeferencing: the DEM ray-march of
// ray_intersect_dem / pixel_to_geo / convert_boundary_to_geo (/root/reference/main_v1.py:635-684, :765-785).
// Host-side orchestration only; the arithmetic runs in raymarch.cuh.  No CPU fallback.
#include "host_common.h"
#include "raymarch.cuh"

using namespace b2r;

struct b2r_dem {
    DevBuf gy, gx, values;
    int ny = 0, nx = 0;
    DemGrid grid() const {
        DemGrid g;
        g.gy = gy.as<double>(); g.gx = gx.as<double>(); g.values = values.as<double>();
        g.ny = ny; g.nx = nx;
        return g;
    }
};

static int utm_from(const double* c, UtmInverse* u) {
    if (!c) return fail(B2R_ERR_ARG, "null UTM series constants%s%s");
    u->k0A = c[0]; u->lon0 = c[1]; u->FE = c[2]; u->FN = c[3];
    for (int j = 0; j < 6; ++j) { u->beta[j] = c[4 + j]; u->delta[j] = c[10 + j]; }
    if (!(u->k0A > 0)) return fail(B2R_ERR_ARG, "bad UTM series constants%s%s");
    return B2R_OK;
}

static int march(b2r_ctx* c, const b2r_dem* dem, const double* origins_dev, int origin_stride, const double* dirs_dev, int m,
                 const double* utm16, double max_search_dist, double step, int min_steps, double* geo_out, int32_t* hit_step_out,
                 int32_t* status_out) {
    UtmInverse u;
    int rc = utm_from(utm16, &u);
    if (rc) return rc;
    if (!(step > 0) || !(max_search_dist >= 0)) return fail(B2R_ERR_ARG, "ray march needs step > 0 and max_search_dist >= 0%s%s");
    const double ns = floor(max_search_dist / step);      // int(max_search_dist / step), main_v1.py:638
    if (ns > 100000000.0) return fail(B2R_ERR_ARG, "too many steps%s%s");
    const int n_steps = (int)ns;
    CU(c->scratch2.reserve(sizeof(double) * 3 * (size_t)m + sizeof(int) * 2 * (size_t)m + 64));
    double* geo = c->scratch2.as<double>();
    int* hit = reinterpret_cast<int*>(geo + 3 * (size_t)m);
    int* status = hit + m;
    LAUNCH(c, k_ray_march_dem, (unsigned)m, RM_THREADS, 0, origins_dev, origin_stride, dirs_dev, u, dem->grid(), n_steps, step, min_steps,
           geo, hit, status);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(geo_out, geo, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    if (hit_step_out) CU(cudaMemcpyAsync(hit_step_out, hit, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(status_out, status, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B2R_OK;
}

extern "C" {

b2r_dem* b2r_dem_upload(b2r_ctx* c, const double* grid_y, int32_t ny, const double* grid_x, int32_t nx, const double* values) {
    if (!c || !grid_y || !grid_x || !values || ny < 2 || nx < 2) {
        fail(B2R_ERR_ARG, "b2r_dem_upload: need a grid of at least 2 x 2 nodes%s%s");
        return nullptr;
    }
    for (int i = 1; i < ny; ++i)
        if (!(grid_y[i] > grid_y[i - 1])) { fail(B2R_ERR_ARG, "b2r_dem_upload: grid_y must be strictly ascending%s%s"); return nullptr; }
    for (int i = 1; i < nx; ++i)
        if (!(grid_x[i] > grid_x[i - 1])) { fail(B2R_ERR_ARG, "b2r_dem_upload: grid_x must be strictly ascending%s%s"); return nullptr; }
    if (cudaSetDevice(c->device) != cudaSuccess) { fail(B2R_ERR_CUDA, "cudaSetDevice failed%s%s"); return nullptr; }
    b2r_dem* d = new b2r_dem();
    d->ny = ny; d->nx = nx;
    const size_t vb = sizeof(double) * (size_t)ny * nx;
    if (d->gy.reserve(sizeof(double) * ny) != cudaSuccess || d->gx.reserve(sizeof(double) * nx) != cudaSuccess ||
        d->values.reserve(vb) != cudaSuccess ||
        cudaMemcpyAsync(d->gy.p, grid_y, sizeof(double) * ny, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(d->gx.p, grid_x, sizeof(double) * nx, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaMemcpyAsync(d->values.p, values, vb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fail(B2R_ERR_CUDA, "b2r_dem_upload: device allocation or copy failed%s%s");
        cudaGetLastError();
        d->gy.release(); d->gx.release(); d->values.release();
        delete d;
        return nullptr;
    }
    return d;
}

void b2r_dem_free(b2r_ctx* c, b2r_dem* d) {
    if (!d) return;
    if (c) cudaSetDevice(c->device);
    d->gy.release(); d->gx.release(); d->values.release();
    delete d;
}

int b2r_ray_march_dem(b2r_ctx* c, const b2r_dem* dem, const double* origins, int32_t origin_shared, const double* dirs, int32_t m,
                      const double* utm_series16, double max_search_dist, double step, int32_t min_steps, double* geo_out,
                      int32_t* hit_step_out, int32_t* status_out) {
    if (!c || !dem || !origins || !dirs || !geo_out || !status_out || m < 1) return fail(B2R_ERR_ARG, "bad argument%s%s");
    CU(cudaSetDevice(c->device));
    const size_t ob = sizeof(double) * 3 * (size_t)(origin_shared ? 1 : m), db = sizeof(double) * 3 * (size_t)m;
    CU(c->scratch0.reserve(ob));
    CU(c->scratch1.reserve(db));
    CU(cudaMemcpyAsync(c->scratch0.p, origins, ob, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->scratch1.p, dirs, db, cudaMemcpyHostToDevice, c->stream));
    return march(c, dem, c->scratch0.as<double>(), origin_shared ? 0 : 3, c->scratch1.as<double>(), m, utm_series16, max_search_dist,
                 step, min_steps, geo_out, hit_step_out, status_out);
}

int b2r_pixels_to_geo(b2r_ctx* c, const b2r_dem* dem, const double* pixels, int32_t m, const double* Kinv, const double* R,
                      const double* ray_origin, const double* ctrl_pixels, const double* ctrl_factors, int32_t n_ctrl,
                      const double* utm_series16, double max_search_dist, double step, int32_t min_steps, double* geo_out,
                      int32_t* hit_step_out, int32_t* status_out, double* dirs_out) {
    if (!c || !dem || !pixels || !Kinv || !R || !ray_origin || !ctrl_pixels || !ctrl_factors || !geo_out || !status_out || m < 1 ||
        n_ctrl < 1)
        return fail(B2R_ERR_ARG, "bad argument%s%s");
    CU(cudaSetDevice(c->device));
    // scratch0: [origin 3][Kinv 9][R 9][pixels 2m][ctrl_pixels 2c][ctrl_factors 3c] ; scratch1: dirs [m][3]
    const size_t nd = 21 + 2 * (size_t)m + 5 * (size_t)n_ctrl;
    CU(c->scratch0.reserve(sizeof(double) * nd));
    CU(c->scratch1.reserve(sizeof(double) * 3 * (size_t)m));
    CU(c->pin_in.reserve(sizeof(double) * nd));
    double* h = (double*)c->pin_in.p;
    memcpy(h, ray_origin, sizeof(double) * 3);
    memcpy(h + 3, Kinv, sizeof(double) * 9);
    memcpy(h + 12, R, sizeof(double) * 9);
    memcpy(h + 21, pixels, sizeof(double) * 2 * (size_t)m);
    memcpy(h + 21 + 2 * (size_t)m, ctrl_pixels, sizeof(double) * 2 * (size_t)n_ctrl);
    memcpy(h + 21 + 2 * (size_t)m + 2 * (size_t)n_ctrl, ctrl_factors, sizeof(double) * 3 * (size_t)n_ctrl);
    CU(cudaMemcpyAsync(c->scratch0.p, h, sizeof(double) * nd, cudaMemcpyHostToDevice, c->stream));
    const double* d0 = c->scratch0.as<double>();
    LAUNCH(c, k_pixel_rays, (unsigned)((m + 127) / 128), 128, 0, d0 + 21, m, d0 + 3, d0 + 12, d0 + 21 + 2 * (size_t)m,
           d0 + 21 + 2 * (size_t)m + 2 * (size_t)n_ctrl, n_ctrl, 1.0, 10.0, c->scratch1.as<double>());
    CU(cudaGetLastError());
    if (dirs_out) CU(cudaMemcpyAsync(dirs_out, c->scratch1.p, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    return march(c, dem, d0, 0, c->scratch1.as<double>(), m, utm_series16, max_search_dist, step, min_steps, geo_out, hit_step_out,
                 status_out);
}

}  // extern "C"
