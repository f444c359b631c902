/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's Numba terrain engine (checker + CPU baseline).
 *
 * Follows xdem/terrain/surfit.py:977-1088 (`_get_surface_attributes_numba`: per pixel, true convolution of the
 * NaN-padded DEM with float64 kernels, surfit.py:948-969, then `_make_attribute_from_coefs`, surfit.py:451-945) and
 * xdem/terrain/window.py:817-870 (`_get_windowed_indexes_numba`), parallelised over rows with OpenMP the way Numba's
 * prange parallelises the reference loop.  Pinned against the committed reference fixtures through
 * tests/test_oracle_terrain.py::test_c_oracle_*.  Never linked into or called by the product (xdem_b200/).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (oracle/build_oracle.py); no -ffast-math, -ffp-contract=off so that the float
 * arithmetic matches NumPy/Numba (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const double K_H1[9] = {1, 2, 1, 0, 0, 0, -1, -2, -1};
static const double K_H2[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
static const double K_ZT_D[9] = {0, 1, 0, 0, -2, 0, 0, 1, 0};
static const double K_ZT_E[9] = {0, 0, 0, 1, -2, 1, 0, 0, 0};
static const double K_ZT_F[9] = {-1, 0, 1, 0, 0, 0, 1, 0, -1};
static const double K_ZT_G[9] = {0, 1, 0, 0, 0, 0, 0, -1, 0};
static const double K_ZT_H[9] = {0, 0, 0, -1, 0, 1, 0, 0, 0};
static const double K_FL_R[25] = {2, -1, -2, -1, 2, 2, -1, -2, -1, 2, 2, -1, -2, -1, 2, 2, -1, -2, -1, 2, 2, -1, -2, -1, 2};
static const double K_FL_T[25] = {2, 2, 2, 2, 2, -1, -1, -1, -1, -1, -2, -2, -2, -2, -2, -1, -1, -1, -1, -1, 2, 2, 2, 2, 2};
static const double K_FL_S[25] = {-4, -2, 0, 2, 4, -2, -1, 0, 1, 2, 0, 0, 0, 0, 0, 2, 1, 0, -1, -2, 4, 2, 0, -2, -4};
static const double K_FL_P[25] = {31, -44, 0, 44, -31, -5, -62, 0, 62, 5, -17, -68, 0, 68, 17, -5, -62, 0, 62, 5, 31, -44, 0, 44, -31};
static const double K_FL_Q[25] = {-31, 5, 17, 5, -31, 44, 62, 68, 62, 44, 0, 0, 0, 0, 0, -44, -62, -68, -62, -44, 31, -5, -17, -5, 31};

int xo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* attrs bit i = surface attribute i in the order of surfit.py:407-418; out planes [10][H*W] (NULL if unused).
 * fit: 0 Horn 1 ZT 2 Florinsky; curv_dir: 0 geometric 1 directional.  dem is float32 or float64 (is_f64). */
int xo_surface(const void* dem_v, int is_f64, int64_t H, int64_t W, double res, int fit, int curv_dir, uint32_t attrs,
               int degrees, int clip_hs, double az_deg, double alt_deg, double zf, void** out, int nthreads) {
    const int w = fit == 2 ? 5 : 3, h = w / 2, n = w * w;
    const double* kx;
    const double* ky;
    const double *kxx = 0, *kyy = 0, *kxy = 0;
    double d1, d2 = 1, d3 = 1;
    if (fit == 0) {
        kx = K_H2, ky = K_H1, d1 = 8 * res;
    } else if (fit == 1) {
        kx = K_ZT_H, ky = K_ZT_G, kxx = K_ZT_E, kyy = K_ZT_D, kxy = K_ZT_F, d1 = 2 * res, d2 = res * res, d3 = 4 * res * res;
    } else {
        kx = K_FL_P, ky = K_FL_Q, kxx = K_FL_R, kyy = K_FL_T, kxy = K_FL_S, d1 = 420 * res, d2 = 35 * res * res,
        d3 = 100 * res * res;
    }
    /* kernels divided by the resolution factor first (surfit.py:373-377), flipped on use (surfit.py:966) */
    double fx[25], fy[25], fxx[25], fyy[25], fxy[25];
    for (int i = 0; i < n; ++i) {
        fx[i] = kx[n - 1 - i] / d1;
        fy[i] = ky[n - 1 - i] / d1;
        fxx[i] = kxx ? kxx[n - 1 - i] / d2 : 0;
        fyy[i] = kyy ? kyy[n - 1 - i] / d2 : 0;
        fxy[i] = kxy ? kxy[n - 1 - i] / d3 : 0;
    }
    const int need2 = (attrs & ~7u) != 0;
    const double az = (360.0 - az_deg) * (M_PI / 180.0), alt = alt_deg * (M_PI / 180.0);
    const float* df = (const float*)dem_v;
    const double* dd = (const double*)dem_v;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < H; ++r) {
        for (int64_t c = 0; c < W; ++c) {
            double zx = 0, zy = 0, zxx = 0, zyy = 0, zxy = 0;
            /* every tap is multiplied (0*NaN = NaN), so any NaN / out-of-raster cell in the window propagates */
            for (int m1 = 0; m1 < w; ++m1)
                for (int m2 = 0; m2 < w; ++m2) {
                    const int64_t rr = r + m1 - h, cc = c + m2 - h;
                    double v = NAN;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W) v = is_f64 ? dd[rr * W + cc] : (double)df[rr * W + cc];
                    const int i = m1 * w + m2;
                    zx += v * fx[i];
                    zy += v * fy[i];
                    if (need2) {
                        zxx += v * fxx[i];
                        zyy += v * fyy[i];
                        zxy += v * fxy[i];
                    }
                }
            const double g2 = zx * zx + zy * zy;
            double o[10];
            double slope = atan(sqrt(g2));
            double aspect = fmod(-atan2(-zx, zy), 2 * M_PI);
            if (aspect < 0) aspect += 2 * M_PI;
            double slopemap = zf != 1.0 ? atan(tan(slope) * zf) : slope;
            o[2] = 1.5 + 254 * (sin(alt) * cos(slopemap) + cos(alt) * sin(slopemap) * sin(az - aspect));
            o[0] = slope, o[1] = aspect;
            if (need2) {
                const int flat0 = g2 == 0.0, flate = g2 < 10e-15;
                const double n1 = zxx * zx * zx + 2 * zxy * zx * zy + zyy * zy * zy;
                const double n2 = zxx * zy * zy - 2 * zxy * zx * zy + zyy * zx * zx;
                const double n3 = zx * zy * (zxx - zyy) - zxy * (zx * zx - zy * zy);
                const double opg = 1 + g2;
                o[3] = -2.0 * (zxx + zyy) * 100;
                o[4] = (flat0 ? 0.0 : -n1 / (curv_dir ? g2 : g2 * sqrt(opg * opg * opg))) * 100;
                o[5] = (flat0 ? 0.0 : -n2 / (curv_dir ? g2 : g2 * sqrt(opg))) * 100;
                o[6] = (flate ? 0.0 : -n2 / sqrt(g2 * g2 * g2)) * 100;
                if (curv_dir)
                    o[7] = (flat0 ? 0.0 : n3 / pow(g2 * g2 * g2, 0.5)) * 100;
                else
                    o[7] = (flate ? 0.0 : n3 / (pow(g2 * g2 * g2, 0.5) * pow(opg, 0.5))) * 100;
                if (curv_dir) {
                    const double half = (zxx + zyy) / 2, rad = pow(((zxx - zyy) / 2) * ((zxx - zyy) / 2) + zxy * zxy, 0.5);
                    o[8] = (flat0 ? 0.0 : -(half - rad)) * 100;
                    o[9] = (flat0 ? 0.0 : -(half + rad)) * 100;
                } else {
                    const double mn = (1 + zy * zy) * zxx - 2 * zxy * zx * zy + (1 + zx * zx) * zyy;
                    const double den = 2 * pow(opg * opg * opg, 0.5);
                    const double mean = -mn / den;
                    const double uns = pow((mn / den) * (mn / den) - (zxx * zyy - zxy * zxy) / (opg * opg), 0.5);
                    o[8] = (flat0 ? 0.0 : mean + uns) * 100;
                    o[9] = (flat0 ? 0.0 : mean - uns) * 100;
                }
            }
            for (int a = 0; a < 10; ++a) {
                if (!((attrs >> a) & 1u)) continue;
                if (is_f64) {
                    double v = o[a];
                    if (degrees && a < 2) v = v * (180.0 / M_PI);
                    if (clip_hs && a == 2 && !isnan(v)) v = v < 0 ? 0 : (v > 255 ? 255 : v);
                    ((double*)out[a])[r * W + c] = v;
                } else {
                    float v = (float)o[a]; /* cast on store, surfit.py:1086 */
                    if (degrees && a < 2) v = v * (180.0f / 3.14159265358979323846f); /* np.rad2deg in float32, terrain.py:591 */
                    if (clip_hs && a == 2 && !isnan(v)) v = v < 0 ? 0 : (v > 255 ? 255 : v);
                    ((float*)out[a])[r * W + c] = v;
                }
            }
        }
    }
    return 0;
}

/* windowed indexes, float32 DEM, sequential row-major float32 accumulation like the Numba engine (window.py:817-870):
 * attrs bit 0 TPI, 1 TRI, 2 roughness.  (rugosity is only restated in NumPy, oracle/terrain_oracle.py.) */
int xo_windowed_f32(const float* dem, int64_t H, int64_t W, int w, uint32_t attrs, int tri_wilson, float** out,
                    int nthreads) {
    const int h = w / 2, n = w * w;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < H; ++r) {
        for (int64_t c = 0; c < W; ++c) {
            float win[49];
            int bad = 0;
            for (int m1 = 0; m1 < w; ++m1)
                for (int m2 = 0; m2 < w; ++m2) {
                    const int64_t rr = r + m1 - h, cc = c + m2 - h;
                    float v = NAN;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W) v = dem[rr * W + cc];
                    if (!isfinite(v)) bad = 1;
                    win[m1 * w + m2] = v;
                }
            const float ctr = win[n / 2];
            float s = 0, sq = 0, sa = 0, mx = ctr, mn = ctr;
            for (int i = 0; i < n; ++i) {
                s += win[i];
                const float d = fabsf(win[i] - ctr);
                sq += d * d;
                sa += d;
                if (win[i] > mx) mx = win[i];
                if (win[i] < mn) mn = win[i];
            }
            const float nm1 = (float)(n - 1);
            /* Numba: float32 sum, then `/ (window_size**2 - 1)` promotes to float64 (window.py:198) */
            if (attrs & 1u) out[0][r * W + c] = bad ? NAN : (float)((double)ctr - (double)(s - ctr) / (double)(n - 1));
            if (attrs & 2u) out[1][r * W + c] = bad ? NAN : (tri_wilson ? sa / nm1 : sqrtf(sq));
            if (attrs & 4u) out[2][r * W + c] = bad ? NAN : mx - mn;
        }
    }
    return 0;
}

/* rugosity (Jenness 2004), restating the per-pixel function the reference's Numba engine runs (window.py:505-595) with
 * its type promotions: float32 segment arrays (dzs, dls, hsl = sqrt(dzs^2 + dls^2) / 2), the triangle half-perimeter
 * `sum(T) / 2` and Heron product in float64, areas stored float32, `sum(A) / L**2` in float64.  Used for the CPU arm of
 * the "all attributes" benchmark and checked against the Numba-engine fixtures (tests/test_oracle_terrain.py). */
int xo_rugosity_f32(const float* dem, int64_t H, int64_t W, double resolution, float* out, int nthreads) {
    static const int TRI[8][3] = {{3, 0, 12}, {0, 1, 8}, {1, 2, 9}, {2, 4, 14}, {4, 7, 15}, {7, 6, 11}, {6, 5, 10},
                                  {5, 3, 13}};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < H; ++r) {
        for (int64_t c = 0; c < W; ++c) {
            float Z[9];
            int bad = 0;
            for (int m1 = 0; m1 < 3; ++m1)
                for (int m2 = 0; m2 < 3; ++m2) {
                    const int64_t rr = r + m1 - 1, cc = c + m2 - 1;
                    float v = NAN;
                    if (rr >= 0 && rr < H && cc >= 0 && cc < W) v = dem[rr * W + cc];
                    if (!isfinite(v)) bad = 1;
                    Z[m1 * 3 + m2] = v;
                }
            if (bad) {
                out[r * W + c] = NAN;
                continue;
            }
            float dzs[16], dls[16], hsl[16];
            int k = 0, all = 0;
            for (int j = -1; j <= 1; ++j)
                for (int i = -1; i <= 1; ++i) {
                    if (j == 0 && i == 0) {
                        ++all;
                        continue;
                    }
                    dzs[k] = Z[4] - Z[all];
                    dls[k] = (float)(sqrt((double)(j * j + i * i)) * resolution);
                    ++all;
                    ++k;
                }
            dzs[8] = Z[0] - Z[1], dzs[9] = Z[1] - Z[2], dzs[10] = Z[6] - Z[7], dzs[11] = Z[7] - Z[8];
            dzs[12] = Z[0] - Z[3], dzs[13] = Z[3] - Z[6], dzs[14] = Z[2] - Z[5], dzs[15] = Z[5] - Z[8];
            for (int i = 8; i < 16; ++i) dls[i] = (float)resolution;
            for (int i = 0; i < 16; ++i) hsl[i] = sqrtf(dzs[i] * dzs[i] + dls[i] * dls[i]) / 2.0f;
            float area = 0.0f;
            for (int t = 0; t < 8; ++t) {
                const float a = hsl[TRI[t][0]], b = hsl[TRI[t][1]], cc3 = hsl[TRI[t][2]];
                const double hs = (double)((a + b) + cc3) / 2.0;
                const float A = (float)sqrt(hs * (hs - a) * (hs - b) * (hs - cc3));
                area += A;
            }
            out[r * W + c] = (float)((double)area / (resolution * resolution));
        }
    }
    return 0;
}
