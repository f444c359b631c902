// xdem_b200 -- C ABI (include/xdem_b200.h): argument validation, constant folding, launches.
#include "../../include/xdem_b200.h"

#include <math.h>
#include <stdarg.h>

#include <atomic>

#include "xb_common.cuh"
#include "xb_terrain.cuh"

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void xb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void xb_count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

static std::atomic<int> g_opt_fl_generic{0};
bool xb_option_florinsky_generic() { return g_opt_fl_generic.load() != 0; }
static std::atomic<int> g_opt_fl_packed{1};
bool xb_option_florinsky_packed() { return g_opt_fl_packed.load() != 0; }
static std::atomic<int> g_opt_fl_tstore{0};
bool xb_option_florinsky_tma_store() { return g_opt_fl_tstore.load() != 0; }
static std::atomic<int> g_opt_w3_generic{0};
bool xb_option_window3_generic() { return g_opt_w3_generic.load() != 0; }
static std::atomic<int> g_opt_vg_full{7};
int xb_option_variogram_full_tiles() { return g_opt_vg_full.load(); }

xb_cuTensorMapEncodeTiled_t xb_get_tensormap_encoder() {
    static xb_cuTensorMapEncodeTiled_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<xb_cuTensorMapEncodeTiled_t>(p);
    }
    return fn;
}

int xb_num_sms(int* out) {
    static int cached[64] = {0};
    int dev = 0;
    XB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && cached[dev]) {
        *out = cached[dev];
        return XB_OK;
    }
    int n = 0;
    XB_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) cached[dev] = n;
    *out = n;
    return XB_OK;
}

int xb_build_terrain_params(xbt::TerrainParams& p, int dtype, double resolution, int fit_id, int curv_method_id,
                                uint32_t surf_mask, uint32_t win_mask, int window_size, int tri_method_id, int degrees,
                                int clip_hillshade, double az, double alt, double zf, int* hs_out, int* hw_out) {
    if (dtype != XB_F32 && dtype != XB_F64) {
        xb_set_error("dtype must be XB_F32 (0) or XB_F64 (1), got %d", dtype);
        return XB_ERR_INVALID;
    }
    if (surf_mask >> 10 || win_mask >> 4) {
        xb_set_error("unknown attribute bits (surf_mask=0x%x win_mask=0x%x)", surf_mask, win_mask);
        return XB_ERR_INVALID;
    }
    if (!surf_mask && !win_mask) {
        xb_set_error("no attribute requested");
        return XB_ERR_INVALID;
    }
    int hs = 0, hw = 0;
    if (surf_mask) {
        if (fit_id < 0 || fit_id > 2) {
            xb_set_error("fit_id must be 0 (Horn), 1 (ZevenbergThorne) or 2 (Florinsky), got %d", fit_id);
            return XB_ERR_INVALID;
        }
        if (fit_id == XB_FIT_HORN && (surf_mask & ~7u)) {
            // terrain.py:296-317
            xb_set_error("'Horn' surface fit method cannot be used for to calculate curvatures");
            return XB_ERR_INVALID;
        }
        if (!(resolution > 0.0) || !isfinite(resolution)) {
            xb_set_error("resolution must be a positive finite number, got %g", resolution);
            return XB_ERR_INVALID;
        }
        hs = (fit_id == XB_FIT_FLORINSKY) ? 2 : 1;
    }
    if (win_mask) {
        if (window_size != 3 && window_size != 5) {
            xb_set_error("the fused kernel supports window_size 3 or 5, got %d", window_size);
            return XB_ERR_UNSUPPORTED;
        }
        if ((win_mask & 8u) && window_size != 3) {
            xb_set_error("rugosity is defined on a 3x3 window only (window.py:505-513)");
            return XB_ERR_INVALID;
        }
        if ((win_mask & 8u) && (!(resolution > 0.0) || !isfinite(resolution))) {
            xb_set_error("rugosity needs a positive finite resolution, got %g", resolution);
            return XB_ERR_INVALID;
        }
        hw = window_size / 2;
    }
    p.surf_mask = surf_mask;
    p.win_mask = win_mask;
    p.fit_id = fit_id;
    p.curv_dir = curv_method_id ? 1 : 0;
    p.tri_wilson = tri_method_id ? 1 : 0;
    p.degrees = degrees ? 1 : 0;
    p.clip_hs = clip_hillshade ? 1 : 0;
    const double r = resolution;
    if (fit_id == XB_FIT_HORN) {
        p.inv_d1 = 1.0 / (8 * r), p.inv_d2 = 0, p.inv_d3 = 0;
    } else if (fit_id == XB_FIT_ZEVENBERG_THORNE) {
        p.inv_d1 = 1.0 / (2 * r), p.inv_d2 = 1.0 / (r * r), p.inv_d3 = 1.0 / (4 * r * r);
    } else {
        p.inv_d1 = 1.0 / (420 * r), p.inv_d2 = 1.0 / (35 * r * r), p.inv_d3 = 1.0 / (100 * r * r);
    }
    p.alg_k3 = p.inv_d2 != 0.0 ? p.inv_d3 / p.inv_d2 : 0.0;
    p.alg_c2 = 100.0 * p.inv_d2;
    // np.rad2deg on a float32 array multiplies by the float32 constant 180.0f/pi_f (terrain.py:591)
    p.rad2deg = dtype == XB_F32 ? (double)(180.0f / 3.14159265358979323846f) : 180.0 / M_PI;
    // surfit.py:614-615
    const double az_rad = (360.0 - az) * (M_PI / 180.0);
    const double alt_rad = alt * (M_PI / 180.0);
    p.hs_sin_alt = sin(alt_rad);
    p.hs_kx = cos(alt_rad) * zf * cos(az_rad);
    p.hs_ky = cos(alt_rad) * zf * sin(az_rad);
    p.zf2 = zf * zf;
    if (dtype == XB_F32) {
        // window.py:628-651: constants in the DEM dtype (NumPy weak-scalar promotion)
        const float L = (float)r;
        const float diag = (float)sqrt(2.0) * L;
        p.rug_dl2_diag = (double)(diag * diag);
        p.rug_dl2_straight = (double)(L * L);
        p.rug_ll = (double)(float)(r * r);
        // RN(1 / L^2) in float32: the candidate nearest to the double quotient or one of its neighbours, whichever
        // minimises |L^2 y - 1| (the product of two float32 values is exact in double)
        const float ll = (float)p.rug_ll;
        float y = (float)(1.0 / (double)ll);
        const float cand[3] = {nextafterf(y, 0.0f), y, nextafterf(y, INFINITY)};
        double best = INFINITY;
        for (float cnd : cand) {
            const double e = fabs((double)ll * (double)cnd - 1.0);
            if (e < best) best = e, y = cnd;
        }
        p.rug_rcp_ll = (double)y;
        // Markstein's correctly rounded a/b needs a mantissa of b that is not all ones
        uint32_t llbits;
        memcpy(&llbits, &ll, 4);
        p.rug_fast_ok = (r >= 1e-6 && r <= 1e6 && (llbits & 0x7fffffu) != 0x7fffffu) ? 1 : 0;
    } else {
        const double diag = sqrt(2.0) * r;
        p.rug_dl2_diag = diag * diag;
        p.rug_dl2_straight = r * r;
        p.rug_ll = r * r;
    }
    p.f.inv1 = (float)p.inv_d1;
    p.f.ang = p.degrees ? (float)p.rad2deg : 1.0f;
    p.f.hs_ky = (float)p.hs_ky;
    p.f.hs_nkx = -(float)p.hs_kx;
    p.f.hs_sa = (float)p.hs_sin_alt;
    p.f.zf2 = (float)p.zf2;
    p.f.curv_nf = (float)(200.0 * p.inv_d2);
    p.f.alg_c2 = (float)p.alg_c2;
    p.f.rug_rcp_ll = (float)p.rug_rcp_ll;
    p.f.rug_nll = -(float)p.rug_ll;
    p.f.rug_l2s = (float)p.rug_dl2_straight;
    p.f.rug_l2d = (float)p.rug_dl2_diag;
    p.f.one = 1.0f;
    p.f.rug_y4 = -0.0625f * p.f.rug_rcp_ll;
    p.f.rug_b4 = -16.0f * p.f.rug_nll;
    *hs_out = hs;
    *hw_out = hw;
    return XB_OK;
}


extern "C" {
#pragma GCC visibility push(default)

const char* xb_last_error(void) { return g_err; }
int xb_version(void) { return 100; }
uint64_t xb_launch_count(void) { return g_launches.load(); }

int xb_set_option(const char* name, int value) {
    if (name && strcmp(name, "florinsky_generic") == 0) {
        g_opt_fl_generic.store(value);
        return XB_OK;
    }
    if (name && strcmp(name, "florinsky_packed") == 0) {
        g_opt_fl_packed.store(value);
        return XB_OK;
    }
    if (name && strcmp(name, "florinsky_tma_store") == 0) {
        g_opt_fl_tstore.store(value);
        return XB_OK;
    }
    if (name && strcmp(name, "window3_generic") == 0) {
        g_opt_w3_generic.store(value);
        return XB_OK;
    }
    if (name && strcmp(name, "variogram_full_tiles") == 0) {
        g_opt_vg_full.store(value);
        return XB_OK;
    }
    xb_set_error("unknown option '%s'", name ? name : "(null)");
    return XB_ERR_INVALID;
}

int xb_terrain_fused(const void* dem_dev, int dtype, int64_t rows_buf, int64_t cols, int64_t ld, int64_t row_begin,
                     int64_t row_end, double resolution, int fit_id, int curv_method_id, uint32_t surf_mask,
                     uint32_t win_mask, int window_size, int tri_method_id, int degrees, int clip_hillshade,
                     double hillshade_azimuth, double hillshade_altitude, double hillshade_z_factor,
                     void* const* out_planes_host, int64_t out_ld, void* stream) {
    if (!dem_dev || !out_planes_host) {
        xb_set_error("null pointer argument");
        return XB_ERR_INVALID;
    }
    if (rows_buf <= 0 || cols <= 0 || ld < cols || row_begin < 0 || row_end > rows_buf || row_begin > row_end ||
        out_ld < cols) {
        xb_set_error("bad geometry rows_buf=%lld cols=%lld ld=%lld rows=[%lld,%lld) out_ld=%lld", (long long)rows_buf,
                     (long long)cols, (long long)ld, (long long)row_begin, (long long)row_end, (long long)out_ld);
        return XB_ERR_INVALID;
    }
    xbt::TerrainParams p;
    memset(&p, 0, sizeof(p));
    int hs = 0, hw = 0;
    int rc = xb_build_terrain_params(p, dtype, resolution, fit_id, curv_method_id, surf_mask, win_mask, window_size,
                                  tri_method_id, degrees, clip_hillshade, hillshade_azimuth, hillshade_altitude,
                                  hillshade_z_factor, &hs, &hw);
    if (rc) return rc;
    for (int i = 0; i < 10; ++i) {
        p.out[i] = (surf_mask >> i) & 1u ? out_planes_host[i] : nullptr;
        if (((surf_mask >> i) & 1u) && !p.out[i]) {
            xb_set_error("surface attribute %d requested but out_planes[%d] is NULL", i, i);
            return XB_ERR_INVALID;
        }
    }
    for (int j = 0; j < 4; ++j) {
        p.out[10 + j] = (win_mask >> j) & 1u ? out_planes_host[10 + j] : nullptr;
        if (((win_mask >> j) & 1u) && !p.out[10 + j]) {
            xb_set_error("windowed index %d requested but out_planes[%d] is NULL", j, 10 + j);
            return XB_ERR_INVALID;
        }
    }
    p.dem = dem_dev;
    p.rows_buf = rows_buf;
    p.cols = cols;
    p.ld = ld;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.out_ld = out_ld;
    if (row_end == row_begin) return XB_OK;
    return xbt::launch(p, dtype, hs, hw, reinterpret_cast<cudaStream_t>(stream));  // counts its kernel launches itself
}

#pragma GCC visibility pop
}  // extern "C"
