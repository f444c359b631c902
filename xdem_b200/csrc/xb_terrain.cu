// xdem_b200 -- fused terrain-attribute stencil kernel (K1) for sm_100a.
//
// One pass over the DEM emits every requested surface-fit attribute (slope, aspect, hillshade, curvature x7) and
// windowed index (TPI, TRI, roughness, rugosity).  Replaces the reference's per-kernel scipy.ndimage.convolve passes
// + whole-array NumPy algebra (surfit.py:1091-1194, window.py:873-923) and its Numba prange loops
// (surfit.py:977-1088, window.py:817-870).
//
// Layout / execution model
//   * persistent CTAs (grid = #SMs x occupancy) walk 128 x TH pixel tiles in row-major order;
//   * each tile (+halo) is staged in shared memory by ONE TMA box load (cp.async.bulk.tensor.2d, mbarrier
//     completion, 2-stage ring) whose tensor map uses CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_..., i.e. out-of-raster
//     cells arrive as NaN -- exactly the reference's `mode="constant", cval=nan` / np.pad(constant_values=nan);
//   * a warp covers one 128-pixel tile row per step, each lane owns 4 consecutive pixels (16-byte aligned), reads
//     its (2h+1) x (4+2h) window from shared memory with vector LDS and writes each attribute plane with one
//     streaming 128-bit store per row (fully coalesced 512 B per warp);
//   * rasters whose base/pitch are not 16-byte aligned use the same kernel with a cooperative bounds-checked loader.
//
// Numerics (see DESIGN.md "Numerical design"): derivative stencils are evaluated as integer-weighted sums of
// *differences* (centre-/pair-differences first), which are exact in fp32 for realistic DEMs, so the kernel matches the
// reference's float64 accumulation without FP64 stencil work; the cancellation-prone curvature algebra runs in FP64;
// windowed indexes use un-contracted IEEE fp32 ops in the reference's (row-major, sequential) order so that
// integer-valued DEMs are bit-exact.
#include <stdlib.h>

#include <type_traits>

#include "xb_terrain_dev.cuh"

namespace xbt {

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory window loads.  `row` points at smem column (XOFF + 4*lane - H) of the wanted row.
// ---------------------------------------------------------------------------------------------------------------
template <int H>
__device__ __forceinline__ void load_row(const float* row, float (&w)[4 + 2 * H]) {
    if constexpr (H == 1) {
        w[0] = row[0];
        float4 m = *reinterpret_cast<const float4*>(row + 1);
        w[1] = m.x, w[2] = m.y, w[3] = m.z, w[4] = m.w;
        w[5] = row[5];
    } else {
        float2 a = *reinterpret_cast<const float2*>(row);
        float4 m = *reinterpret_cast<const float4*>(row + 2);
        float2 b = *reinterpret_cast<const float2*>(row + 6);
        w[0] = a.x, w[1] = a.y, w[2] = m.x, w[3] = m.y, w[4] = m.z, w[5] = m.w, w[6] = b.x, w[7] = b.y;
    }
}
template <int H>
__device__ __forceinline__ void load_row(const double* row, double (&w)[4 + 2 * H]) {
    if constexpr (H == 1) {
        w[0] = row[0];
        double2 m0 = *reinterpret_cast<const double2*>(row + 1);
        double2 m1 = *reinterpret_cast<const double2*>(row + 3);
        w[1] = m0.x, w[2] = m0.y, w[3] = m1.x, w[4] = m1.y;
        w[5] = row[5];
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2 m = *reinterpret_cast<const double2*>(row + 2 * j);
            w[2 * j] = m.x, w[2 * j + 1] = m.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Derivative stencils.  W(r,c) = window value at row offset r, column offset c from the pixel (r,c in -HS..HS).
// All return the *unscaled* integer-weighted sum S; the derivative is S * inv_divider (surfit.py:278-304).
// Effective (flipped) weights: SURVEY.md Appendix A.1; tables surfit.py:61-252; flip surfit.py:966.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct Derivs {
    T sx, sy, sxx, syy, sxy, carrier;  // carrier: +0 or NaN (any non-finite cell in the full window)
};

#define WIN(r, c) win[(r) + H][(c) + H + k]

template <typename T, int H, int HS>
__device__ __forceinline__ Derivs<T> derivs_at(const T (&win)[2 * H + 1][4 + 2 * H], int kk, int fit_id, bool need2) {
    Derivs<T> d;
    d.sxx = d.syy = d.sxy = T(0);
    // `kk` is a compile-time constant after unrolling
    const int k = kk;
    if constexpr (HS == 1) {
        const T c = WIN(0, 0);
        if (fit_id == XB_FIT_HORN_ID) {
            // h2 -> z_x (west-positive), h1 -> z_y; divider 8*res (surfit.py:145-157, 291-292)
            d.sx = ((WIN(-1, -1) - WIN(-1, 1)) + (WIN(1, -1) - WIN(1, 1))) + T(2) * (WIN(0, -1) - WIN(0, 1));
            d.sy = ((WIN(1, -1) - WIN(-1, -1)) + (WIN(1, 1) - WIN(-1, 1))) + T(2) * (WIN(1, 0) - WIN(-1, 0));
            d.carrier = (d.sx + d.sy + c) * T(0);
        } else {
            // zt_h -> z_x, zt_g -> z_y (divider 2*res); zt_e -> z_xx, zt_d -> z_yy (res^2); zt_f -> z_xy (4*res^2)
            // (surfit.py:93-129, 285-289, 565-569)
            d.sx = WIN(0, -1) - WIN(0, 1);
            d.sy = WIN(1, 0) - WIN(-1, 0);
            const T cr1 = WIN(-1, 1) - WIN(-1, -1);
            const T cr2 = WIN(1, 1) - WIN(1, -1);
            if (need2) {
                d.sxx = (WIN(0, -1) - c) + (WIN(0, 1) - c);
                d.syy = (WIN(-1, 0) - c) + (WIN(1, 0) - c);
                d.sxy = cr1 - cr2;
            }
            d.carrier = ((d.sx + d.sy) + (cr1 + cr2) + c) * T(0);
        }
    } else {
        // Florinsky (2009) 5x5: fl_p -> z_x, fl_q -> z_y (420*res); fl_r -> z_xx, fl_t -> z_yy (35*res^2);
        // fl_s -> z_xy (100*res^2)  (surfit.py:204-252, 297-301, 580-584)
        // z_y: sum_c a_c (W(-2,c) - W(2,c)) + b_c (W(1,c) - W(-1,c)),  a = [31,-5,-17,-5,31], b = [44,62,68,62,44]
        T v2[5], v1[5];
#pragma unroll
        for (int c = -2; c <= 2; ++c) {
            v2[c + 2] = WIN(-2, c) - WIN(2, c);
            v1[c + 2] = WIN(1, c) - WIN(-1, c);
        }
        // the weights sum to 35 (v2) and 280 (v1) and the two parts cancel on steep, curved surfaces: combine them at the
        // level of the differences first -- 35 (v2c + 8 v1c) + sum_c a_c (v2_c - v2c) + b_c (v1_c - v1c)
        d.sy = T(35) * (v2[2] + T(8) * v1[2]) + T(31) * ((v2[0] + v2[4]) - T(2) * v2[2]) -
               T(5) * ((v2[1] + v2[3]) - T(2) * v2[2]) + T(44) * ((v1[0] + v1[4]) - T(2) * v1[2]) +
               T(62) * ((v1[1] + v1[3]) - T(2) * v1[2]);
        // z_x: sum_r a_r (W(r,2) - W(r,-2)) + b_r (W(r,-1) - W(r,1))
        T dd[5], ee[5];
#pragma unroll
        for (int r = -2; r <= 2; ++r) {
            dd[r + 2] = WIN(r, 2) - WIN(r, -2);
            ee[r + 2] = WIN(r, -1) - WIN(r, 1);
        }
        d.sx = T(35) * (dd[2] + T(8) * ee[2]) + T(31) * ((dd[0] + dd[4]) - T(2) * dd[2]) -
               T(5) * ((dd[1] + dd[3]) - T(2) * dd[2]) + T(44) * ((ee[0] + ee[4]) - T(2) * ee[2]) +
               T(62) * ((ee[1] + ee[3]) - T(2) * ee[2]);
        d.carrier = (d.sx + d.sy + WIN(0, 0)) * T(0);
        if (need2) {
            // z_xx: every row [2,-1,-2,-1,2]  -> second differences first (exact), then the 5-row sum
            T sxx = T(0), syy = T(0);
#pragma unroll
            for (int r = -2; r <= 2; ++r) {
                const T c0 = WIN(r, 0);
                const T p = (WIN(r, -2) - c0) + (WIN(r, 2) - c0);
                const T q = (WIN(r, -1) - c0) + (WIN(r, 1) - c0);
                sxx += T(2) * p - q;
            }
#pragma unroll
            for (int c = -2; c <= 2; ++c) {
                const T c0 = WIN(0, c);
                const T p = (WIN(-2, c) - c0) + (WIN(2, c) - c0);
                const T q = (WIN(-1, c) - c0) + (WIN(1, c) - c0);
                syy += T(2) * p - q;
            }
            d.sxx = sxx;
            d.syy = syy;
            // z_xy = -sum r*c*W(r,c) = -(M11 + 2 M12 + 2 M21 + 4 M22),  M(r,c) mixed second differences
            const T m11 = ee[1] - ee[3];  // (W(1,1)-W(1,-1)) - (W(-1,1)-W(-1,-1)) = -ee[3] + ee[1]
            const T m12 = dd[3] - dd[1];  // c = 2, r = 1
            const T m21 = ee[0] - ee[4];  // c = 1, r = 2
            const T m22 = dd[4] - dd[0];
            d.sxy = -((m11 + T(2) * (m12 + m21)) + T(4) * m22);
        }
    }
    return d;
}
#undef WIN

// ---------------------------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------------------------
// CMASK != 0: compile-time surface-attribute mask for the common requests (straight-line attribute code the scheduler
// can interleave: measured +6 % on the Florinsky kernel); CMASK == 0: runtime mask from the parameters.
template <typename T, int HS, int HW, int RPW, bool USE_TMA, bool ALG, unsigned CMASK>
__global__ void __launch_bounds__(NTHREADS, ALG ? 2 : 3)
terrain_fused_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TerrainParams p) {
    constexpr int H = (HS > HW) ? HS : HW;
    constexpr int TH = NWARPS * RPW;
    constexpr int BOXH = TH + 2 * H;
    constexpr uint32_t STAGE_BYTES = BOXW * BOXH * sizeof(T);         // bytes one TMA box delivers
    constexpr int STAGE_ELEMS = ((STAGE_BYTES + 127) / 128) * 128 / sizeof(T);  // stage stride, 128 B aligned

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const long long tiles_x = p.tiles_x;
    const long long ntiles = p.ntiles;
    const long long W = p.cols;

    if constexpr (USE_TMA) {
        if (tid == 0) {
            xb_prefetch_tensormap(&tmap);
#pragma unroll
            for (int s = 0; s < NSTAGES; ++s) xb_mbar_init(&full_bar[s], 1);
            xb_fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NSTAGES; ++s) {
                long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
                if (t < ntiles) {
                    const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
                    xb_mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    xb_tma_load_2d(smem + (size_t)s * STAGE_ELEMS, &tmap, &full_bar[s], tx * TW - XOFF,
                                   (int)p.row_begin + ty * TH - H);
                }
            }
        }
    }

    const uint32_t smask = CMASK ? CMASK : p.surf_mask;
    const bool need_surf = smask != 0;
    const bool need2 = (smask & ~7u) != 0;
    const bool need_sah = (smask & 7u) != 0;
    const bool need_curv_alg = ALG && (smask & ~15u) != 0;  // ALG=false kernels carry no FP64 algebra
    const bool vec_ok = p.vec_ok != 0;

    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int stage = USE_TMA ? (it % NSTAGES) : 0;
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const long long y_tile = p.row_begin + (long long)ty * TH;
        const long long x_tile = (long long)tx * TW;
        T* tile = smem + (size_t)stage * STAGE_ELEMS;

        if constexpr (USE_TMA) {
            xb_mbar_wait(&full_bar[stage], (uint32_t)((it / NSTAGES) & 1));
        } else {
            const T* dem = reinterpret_cast<const T*>(p.dem);
            for (int idx = tid; idx < BOXW * BOXH; idx += NTHREADS) {
                const int by = idx / BOXW, bx = idx - by * BOXW;
                const long long gy = y_tile - H + by, gx = x_tile - XOFF + bx;
                T v = Num<T>::nan();
                if (gy >= 0 && gy < p.rows_buf && gx >= 0 && gx < W) v = dem[gy * p.ld + gx];
                tile[idx] = v;
            }
            __syncthreads();
        }

        const long long x0 = x_tile + 4 * lane;
        const bool full = vec_ok && (x0 + 3 < W);
        const int nvalid = (int)((W - x0) < 4 ? (W - x0) : 4);
        // One output row of this warp.  FAST (interior strips: every lane owns four existing pixels, all RPW rows exist,
        // rows are vector-aligned) drops the row test and the ragged-edge store paths -> straight-line code.
        auto row_body = [&](auto fast_tag, int rr) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const int ly = warp * RPW + rr;  // row inside the tile
            const long long y = y_tile + ly;
            if (!FAST && (y >= p.row_end || x0 >= W)) return;  // warp-uniform in y; lanes past the raster edge idle
            const long long off = (y - p.row_begin) * p.out_ld + x0;

            T win[2 * H + 1][4 + 2 * H];
#pragma unroll
            for (int r = 0; r < 2 * H + 1; ++r)
                load_row<H>(tile + (size_t)(ly + r) * BOXW + (XOFF + 4 * lane - H), win[r]);

            // ------------------------------ surface-fit attributes ------------------------------
            if constexpr (HS > 0) {
                if (need_surf) {
                    T sx[4], sy[4], sxx[4], syy[4], sxy[4], car[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // the surface window is centred in the (possibly larger) tile window
                        Derivs<T> d = derivs_at<T, H, HS>(win, k, p.fit_id, need2);
                        sx[k] = d.sx, sy[k] = d.sy, sxx[k] = d.sxx, syy[k] = d.syy, sxy[k] = d.sxy, car[k] = d.carrier;
                    }
                    if (need_sah) {
                        T zx[4], zy[4], g2[4];
                        const T inv1 = (T)p.inv_d1;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            zx[k] = sx[k] * inv1, zy[k] = sy[k] * inv1;
                            g2[k] = xb_fma(zx[k], zx[k], zy[k] * zy[k]);
                        }
                        const T ang = p.degrees ? (T)p.rad2deg : T(1);
                        if (smask & 1u) {
                            T o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = slope_rad(g2[k]) * ang + car[k];  // surfit.py:592
                            store4<T, FAST>(p.out[0], off, full, nvalid, o);
                        }
                        if (smask & 2u) {
                            T o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = aspect_rad(zx[k], zy[k]) * ang + car[k];  // surfit.py:600
                            store4<T, FAST>(p.out[1], off, full, nvalid, o);
                        }
                        if (smask & 4u) {
                            // 1.5 + 254*(sin(alt) cos(s') + cos(alt) sin(s') sin(az - aspect)), s' = atan(zf*|grad|),
                            // evaluated algebraically (surfit.py:606-622): cos(s') = 1/sqrt(1+zf^2 g2),
                            // sin(s') sin(az-asp) = zf (sin(az) zy - cos(az) zx) / sqrt(1+zf^2 g2)
                            T o[4];
                            const T ky = (T)p.hs_ky, kx = -(T)p.hs_kx, sa = (T)p.hs_sin_alt, zf2 = (T)p.zf2;
                            const T lo = p.clip_hs ? T(0) : -CUDART_INF_F, hi = p.clip_hs ? T(255) : CUDART_INF_F;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const T r = xb_rsqrt(xb_fma(zf2, g2[k], T(1)));
                                const T inner = xb_fma(ky, zy[k], xb_fma(kx, zx[k], sa));
                                const T h = xb_fma(T(254) * r, inner, T(1.5));
                                o[k] = fmin(fmax(h, lo), hi) + car[k];  // clip (terrain.py:596); NaN via the carrier
                            }
                            store4<T, FAST>(p.out[2], off, full, nvalid, o);
                        }
                    }
                    if (smask & 8u) {
                        // curvature = -2 (z_xx + z_yy) * 100 (surfit.py:636); z_xx, z_yy share their divider
                        T cv[4];
                        const T f = (T)(-200.0 * p.inv_d2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) cv[k] = (sxx[k] + syy[k]) * f + car[k];
                        store4<T, FAST>(p.out[3], off, full, nvalid, cv);
                    }
                    if constexpr (ALG) if (need_curv_alg) {
                        // Cancellation-prone algebra (surfit.py:638-943) from the exact unscaled sums: FP64 throughout
                        // for float64 rasters; for float32 rasters FP64 numerators + fp32 denominators (curv_alg_mixed)
                        T o4[4], o5[4], o6[4], o7[4], o8[4], o9[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            T r6[6];
                            curv_alg<T>(sx[k], sy[k], sxx[k], syy[k], sxy[k], p, smask, r6);
                            o4[k] = r6[0] + car[k], o5[k] = r6[1] + car[k], o6[k] = r6[2] + car[k];
                            o7[k] = r6[3] + car[k], o8[k] = r6[4] + car[k], o9[k] = r6[5] + car[k];
                        }
                        if (smask & (1u << 4)) store4<T, FAST>(p.out[4], off, full, nvalid, o4);
                        if (smask & (1u << 5)) store4<T, FAST>(p.out[5], off, full, nvalid, o5);
                        if (smask & (1u << 6)) store4<T, FAST>(p.out[6], off, full, nvalid, o6);
                        if (smask & (1u << 7)) store4<T, FAST>(p.out[7], off, full, nvalid, o7);
                        if (smask & (1u << 8)) store4<T, FAST>(p.out[8], off, full, nvalid, o8);
                        if (smask & (1u << 9)) store4<T, FAST>(p.out[9], off, full, nvalid, o9);
                    }
                }
            }

            // ------------------------------ windowed indexes ------------------------------
            if constexpr (HW > 0) {
                if (p.win_mask) {
                    constexpr int WS = 2 * HW + 1;
                    constexpr int OFF = H - HW;  // window centred in the tile window
                    T tpi[4], tri[4], rough[4], rug[4], car_w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const T c = win[H][H + k];
                        // sequential row-major accumulation, un-contracted IEEE ops (window.py:851, 198, 73-74, 130-131)
                        T s = T(0);
#pragma unroll
                        for (int r = 0; r < WS; ++r)
#pragma unroll
                            for (int cc = 0; cc < WS; ++cc) s = Num<T>::add(s, win[OFF + r][OFF + cc + k]);
                        const T carr = Num<T>::mul(s, T(0));
                        car_w[k] = carr;
                        const T nm1 = (T)(WS * WS - 1);
                        // TPI = c - (sum - c)/(n-1)  (window.py:216-220)
                        tpi[k] = Num<T>::add(Num<T>::sub(c, Num<T>::div(Num<T>::sub(s, c), nm1)), carr);
                        if (p.win_mask & 2u) {
                            // TRI Riley sqrt(sum diff^2) (window.py:94-95) / Wilson sum|diff|/(n-1) (window.py:150-155)
                            T acc = T(0);
                            if (p.tri_wilson) {
#pragma unroll
                                for (int r = 0; r < WS; ++r)
#pragma unroll
                                    for (int cc = 0; cc < WS; ++cc)
                                        acc = Num<T>::add(acc, fabs(Num<T>::sub(win[OFF + r][OFF + cc + k], c)));
                                acc = Num<T>::div(acc, nm1);
                            } else {
#pragma unroll
                                for (int r = 0; r < WS; ++r)
#pragma unroll
                                    for (int cc = 0; cc < WS; ++cc) {
                                        const T df = Num<T>::sub(win[OFF + r][OFF + cc + k], c);
                                        acc = Num<T>::add(acc, Num<T>::mul(df, df));
                                    }
                                acc = Num<T>::sqrt(acc);
                            }
                            tri[k] = Num<T>::add(acc, carr);
                        }
                        if (p.win_mask & 4u) {
                            T mx = c, mn = c;
#pragma unroll
                            for (int r = 0; r < WS; ++r)
#pragma unroll
                                for (int cc = 0; cc < WS; ++cc) {
                                    mx = fmax(mx, win[OFF + r][OFF + cc + k]);
                                    mn = fmin(mn, win[OFF + r][OFF + cc + k]);
                                }
                            rough[k] = Num<T>::add(Num<T>::sub(mx, mn), carr);  // window.py:281-287
                        }
                    }
                    if constexpr (HW == 1) {
                        if (p.win_mask & 8u) {
                            // own loop over the 4 pixels: one basic block, so segment lengths shared by neighbouring
                            // pixels (identical expressions) are computed once
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const T c = win[H][H + k];
                                const T carr = car_w[k];
                                // Rugosity, Jenness (2004): window.py:598-683 -- 16 half segment lengths, 8 Heron areas
                                const T z0 = win[OFF + 0][OFF + 0 + k], z1 = win[OFF + 0][OFF + 1 + k],
                                        z2 = win[OFF + 0][OFF + 2 + k], z3 = win[OFF + 1][OFF + 0 + k],
                                        z5 = win[OFF + 1][OFF + 2 + k], z6 = win[OFF + 2][OFF + 0 + k],
                                        z7 = win[OFF + 2][OFF + 1 + k], z8 = win[OFF + 2][OFF + 2 + k];
                                // straight centre segments and ring segments have the same planimetric length L
                                // (window.py:628-651); differences are always taken left-right / top-bottom (the
                                // half-length only depends on dz^2) so that segments shared by neighbouring pixels are
                                // the same expression and are computed once per thread row.
                                const T l2d = (T)p.rug_dl2_diag, l2s = (T)p.rug_dl2_straight;
                                auto hsl = [](T dz, T l2) {
                                    return Num<T>::mul(Num<T>::sqrt(Num<T>::add(Num<T>::mul(dz, dz), l2)), T(0.5));
                                };
                                T h[16];
                                h[0] = hsl(Num<T>::sub(c, z0), l2d);
                                h[1] = hsl(Num<T>::sub(z1, c), l2s);
                                h[2] = hsl(Num<T>::sub(c, z2), l2d);
                                h[3] = hsl(Num<T>::sub(z3, c), l2s);
                                h[4] = hsl(Num<T>::sub(c, z5), l2s);
                                h[5] = hsl(Num<T>::sub(c, z6), l2d);
                                h[6] = hsl(Num<T>::sub(c, z7), l2s);
                                h[7] = hsl(Num<T>::sub(c, z8), l2d);
                                h[8] = hsl(Num<T>::sub(z0, z1), l2s);
                                h[9] = hsl(Num<T>::sub(z1, z2), l2s);
                                h[10] = hsl(Num<T>::sub(z6, z7), l2s);
                                h[11] = hsl(Num<T>::sub(z7, z8), l2s);
                                h[12] = hsl(Num<T>::sub(z0, z3), l2s);
                                h[13] = hsl(Num<T>::sub(z3, z6), l2s);
                                h[14] = hsl(Num<T>::sub(z2, z5), l2s);
                                h[15] = hsl(Num<T>::sub(z5, z8), l2s);
                                auto heron = [](T a, T b, T cc3) {
                                    const T s2 = Num<T>::mul(Num<T>::add(Num<T>::add(a, b), cc3), T(0.5));
                                    T pr = Num<T>::mul(s2, Num<T>::sub(s2, a));
                                    pr = Num<T>::mul(pr, Num<T>::sub(s2, b));
                                    pr = Num<T>::mul(pr, Num<T>::sub(s2, cc3));
                                    return Num<T>::sqrt(pr);
                                };
                                const T a0 = heron(h[3], h[0], h[12]), a1 = heron(h[0], h[1], h[8]),
                                        a2 = heron(h[1], h[2], h[9]), a3 = heron(h[2], h[4], h[14]),
                                        a4 = heron(h[4], h[7], h[15]), a5 = heron(h[7], h[6], h[11]),
                                        a6 = heron(h[6], h[5], h[10]), a7 = heron(h[5], h[3], h[13]);
                                // np.sum(A, axis=-1) on the (..., 8) block reduces with the short axis outermost,
                                // i.e. sequentially (pinned by the SciPy-engine fixtures)
                                T area = Num<T>::add(a0, a1);
                                area = Num<T>::add(area, a2);
                                area = Num<T>::add(area, a3);
                                area = Num<T>::add(area, a4);
                                area = Num<T>::add(area, a5);
                                area = Num<T>::add(area, a6);
                                area = Num<T>::add(area, a7);
                                rug[k] = Num<T>::add(Num<T>::div(area, (T)p.rug_ll), carr);
                            }
                        }
                    }
                    if (p.win_mask & 1u) store4<T, FAST>(p.out[10], off, full, nvalid, tpi);
                    if (p.win_mask & 2u) store4<T, FAST>(p.out[11], off, full, nvalid, tri);
                    if (p.win_mask & 4u) store4<T, FAST>(p.out[12], off, full, nvalid, rough);
                    if constexpr (HW == 1) {
                        if (p.win_mask & 8u) store4<T, FAST>(p.out[13], off, full, nvalid, rug);
                    }
                }
            }
        };
        constexpr bool HAS_FAST = USE_TMA && !ALG && sizeof(T) == 4;  // the hot float32 instantiations only (code size)
        bool warp_fast = false;
        // A/B on B200 (16384^2): slope+aspect 0.73 -> 0.66 ms (Horn), 0.75 -> 0.68 ms (ZT); requests without the
        // slope / aspect / hillshade math (curvature only, windowed indexes) sit at the HBM bound either way and measured
        // 2-7 % slower on the straight-line path, so they keep the compact loop.
        if constexpr (HAS_FAST)
            warp_fast = need_sah && __all_sync(0xffffffffu, full) && (y_tile + (long long)(warp + 1) * RPW <= p.row_end);
        if (HAS_FAST && warp_fast) {
#pragma unroll 1
            for (int rr = 0; rr < RPW; ++rr) row_body(std::true_type{}, rr);
        } else {
#pragma unroll 1
            for (int rr = 0; rr < RPW; ++rr) row_body(std::false_type{}, rr);
        }

        __syncthreads();  // every warp is done with this stage
        if constexpr (USE_TMA) {
            if (tid == 0) {
                const long long tn = t + (long long)NSTAGES * gridDim.x;
                if (tn < ntiles) {
                    const int tyn = (int)(tn / tiles_x), txn = (int)(tn % tiles_x);
                    xb_mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    xb_tma_load_2d(tile, &tmap, &full_bar[stage], txn * TW - XOFF, (int)p.row_begin + tyn * TH - H);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Host-side launch
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int HS, int HW, int RPW, bool USE_TMA, bool ALG, unsigned CMASK = 0u>
static int launch_cfg(const CUtensorMap& tmap, const TerrainParams& p, int num_sms, cudaStream_t stream) {
    constexpr int H = (HS > HW) ? HS : HW;
    constexpr int TH = NWARPS * RPW;
    constexpr int BOXH = TH + 2 * H;
    const size_t smem = (size_t)(USE_TMA ? NSTAGES : 1) * (((size_t)BOXW * BOXH * sizeof(T) + 127) / 128 * 128);
    auto kern = terrain_fused_kernel<T, HS, HW, RPW, USE_TMA, ALG, CMASK>;
    XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    XB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)num_sms * occ;
    if (grid > p.ntiles) grid = p.ntiles;
    kern<<<(unsigned)grid, NTHREADS, smem, stream>>>(tmap, p);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

template <typename T, int HS, int HW>
static int launch_tma_sel(bool use_tma, const CUtensorMap& tmap, const TerrainParams& p, int num_sms,
                          cudaStream_t stream) {
    constexpr int RPW = 8;
    const bool alg = HS > 0 && (p.surf_mask & ~15u) != 0;
    if constexpr (HS > 0) {
        if (alg) {
            if (use_tma) return launch_cfg<T, HS, HW, RPW, true, true>(tmap, p, num_sms, stream);
            return launch_cfg<T, HS, HW, RPW, false, true>(tmap, p, num_sms, stream);
        }
    }
    // compile-time mask for the slope-only request (BASELINE config 1) on the fast float32 / TMA / surface-only path:
    // measured 0.53 -> 0.43 ms at 16384^2 (Horn); for the multi-attribute masks of the 3x3 fits the runtime-mask kernel
    // was as fast or faster (A/B on B200), so they keep it.
    if constexpr (HS > 0 && HW == 0 && sizeof(T) == 4) {
        if (use_tma && p.surf_mask == 1u) return launch_cfg<T, HS, HW, RPW, true, false, 1u>(tmap, p, num_sms, stream);
    }
    if (use_tma) return launch_cfg<T, HS, HW, RPW, true, false>(tmap, p, num_sms, stream);
    return launch_cfg<T, HS, HW, RPW, false, false>(tmap, p, num_sms, stream);
}

template <typename T>
static int launch_halo_sel(int hs, int hw, bool use_tma, const CUtensorMap& tmap, const TerrainParams& p, int num_sms,
                           cudaStream_t stream) {
#define XB_CASE(a, b) \
    if (hs == a && hw == b) return launch_tma_sel<T, a, b>(use_tma, tmap, p, num_sms, stream);
    XB_CASE(1, 0) XB_CASE(2, 0) XB_CASE(0, 1) XB_CASE(0, 2) XB_CASE(1, 1) XB_CASE(1, 2) XB_CASE(2, 1) XB_CASE(2, 2)
#undef XB_CASE
    xb_set_error("unsupported halo combination surface=%d window=%d", hs, hw);
    return XB_ERR_UNSUPPORTED;
}

int launch(const TerrainParams& p_in, int dtype, int hs, int hw, cudaStream_t stream) {
    TerrainParams p = p_in;
    const int H = hs > hw ? hs : hw;
    const int TH = NWARPS * 8;
    const size_t es = dtype == 1 ? 8 : 4;
    p.tiles_x = (p.cols + TW - 1) / TW;
    const long long tiles_y = (p.row_end - p.row_begin + TH - 1) / TH;
    p.ntiles = p.tiles_x * tiles_y;
    if (p.ntiles <= 0) return XB_OK;
    int num_sms = 0;
    int rc = xb_num_sms(&num_sms);
    if (rc) return rc;

    // vector stores need 16-byte aligned plane rows
    bool vec_ok = (p.out_ld * es) % 16 == 0;
    for (int i = 0; i < 14; ++i)
        if (p.out[i] && (reinterpret_cast<uintptr_t>(p.out[i]) % 16) != 0) vec_ok = false;
    p.vec_ok = vec_ok ? 1 : 0;

    // TMA needs a 16-byte aligned base and pitch; otherwise the cooperative loader variant runs
    bool use_tma = (reinterpret_cast<uintptr_t>(p.dem) % 16 == 0) && ((p.ld * es) % 16 == 0) &&
                   p.cols < (1ll << 31) && p.rows_buf < (1ll << 31);
    // float32 requests that mix surface attributes with 3x3 windowed indexes (e.g. "all attributes", BASELINE config 4)
    // run as two specialised launches: the second read of the DEM (4 B/px of 60) costs far less than the combined
    // generic kernel loses to register pressure (r02base: all 13 planes 8.37 ms at 16384^2 vs 3.35 + 4.19 ms apart).
    if (use_tma && dtype == 0 && hs > 0 && hw == 1 && xb_get_tensormap_encoder()) {
        TerrainParams ps = p, pw = p;
        ps.win_mask = 0;
        pw.surf_mask = 0;
        for (int i = 10; i < 14; ++i) ps.out[i] = nullptr;
        for (int i = 0; i < 10; ++i) pw.out[i] = nullptr;
        const int rc = launch(ps, dtype, hs, 0, stream);
        if (rc) return rc;
        return launch(pw, dtype, 0, hw, stream);
    }
    // float32 3x3 windowed indexes: the row-feature-reuse kernel (xb_terrain_w3.cu) shares the Jenness segments between
    // neighbouring pixels and uses packed f32x2 arithmetic.  xb_set_option("window3_generic", 1) forces the generic
    // kernel (A/B tests).
    if (use_tma && dtype == 0 && hs == 0 && hw == 1 && !xb_option_window3_generic() && xb_get_tensormap_encoder() &&
        (!(p.win_mask & 8u) || p.rug_fast_ok))
        return launch_window3_sliding(p, stream);
    // float32 Florinsky with second-derivative attributes: the row-feature-reuse kernel (xb_terrain_fl.cu) is ~25 %
    // faster (ncu/A-B: 1.42 vs 1.86 ms at 16384^2 for slope+aspect+hillshade+curvature); first-derivative-only requests
    // stay on the generic kernel, which computes just z_x, z_y at higher occupancy.  xb_set_option("florinsky_generic",1)
    // forces the generic kernel (A/B tests).
    if (use_tma && dtype == 0 && hs == 2 && hw == 0 && p.fit_id == XB_FIT_FLORINSKY_ID &&
        !xb_option_florinsky_generic() && xb_get_tensormap_encoder() &&
        ((p.surf_mask & ~7u) != 0 || (xb_option_florinsky_packed() && florinsky_has_packed_mask(p.surf_mask))))
        return launch_florinsky_sliding(p, stream);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (use_tma) {
        xb_cuTensorMapEncodeTiled_t enc = xb_get_tensormap_encoder();
        if (!enc) {
            use_tma = false;
        } else {
            cuuint64_t gdim[2] = {(cuuint64_t)p.cols, (cuuint64_t)p.rows_buf};
            cuuint64_t gstr[1] = {(cuuint64_t)(p.ld * es)};
            cuuint32_t box[2] = {(cuuint32_t)BOXW, (cuuint32_t)(TH + 2 * H)};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tmap, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                             const_cast<void*>(p.dem), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
            if (r != CUDA_SUCCESS) {
                xb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
                return XB_ERR_CUDA;
            }
        }
    }
    if (dtype == 1) return launch_halo_sel<double>(hs, hw, use_tma, tmap, p, num_sms, stream);
    return launch_halo_sel<float>(hs, hw, use_tma, tmap, p, num_sms, stream);
}

}  // namespace xbt
