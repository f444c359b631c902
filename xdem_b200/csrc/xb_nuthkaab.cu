// xdem_b200 -- Nuth & Kaab (2011) inner-step kernels (K3) for sm_100a.
//
// Replaces, per iteration of xdem.coreg.affine._nuth_kaab_iteration_step (affine.py:477-536):
//   * _nuth_kaab_aux_vars (affine.py:412-474)            -> nk_aux_kernel      (np.gradient, slope_tan, aspect)
//   * sub_dh_interpolator (affine.py:179-184)            -> nk_dh_kernel       (bilinear dh, uniform shift = 2x2 stencil)
//   * np.nanmedian(dh) (affine.py:504) and the 72-bin
//     nanmedian of dh/slope_tan over aspect
//     (_bin_or_and_fit_nd -> nd_binning -> binned_statistic,
//      base.py:1014-1020, spatialstats.py:147-149)       -> nk_hist_kernel / nk_next_kernel: exact medians by MSD radix
//                                                           select on order-preserving float32 keys (3 histogram passes)
// The 72-point curve_fit stays on the host (affine.py:1038-1045 of base.py).
#include "../../include/xdem_b200.h"

#include <math_constants.h>

#include "xb_common.cuh"

void xb_count_launch(int n);

namespace xbn {

constexpr int NT = 256;

// order-preserving map float32 -> uint32 (ascending), NaN excluded by the callers
__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------------------------
// aux variables: np.gradient (unit spacing; one-sided first/last row & column), float32 arithmetic like NumPy
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
nk_aux_kernel(const float* __restrict__ z, long long rows_buf, long long cols, long long ld, int top_is_border,
              int bottom_is_border, long long row_begin, long long row_end, float* __restrict__ slope_tan,
              float* __restrict__ aspect, long long out_ld) {
    for (long long r = row_begin + blockIdx.x; r < row_end; r += gridDim.x)
    for (long long c = threadIdx.x; c < cols; c += NT) {
        const float* zr = z + r * ld;
        float gx, gy;
        if (c == 0)
            gx = __fsub_rn(zr[1], zr[0]);
        else if (c == cols - 1)
            gx = __fsub_rn(zr[c], zr[c - 1]);
        else
            gx = __fmul_rn(__fsub_rn(zr[c + 1], zr[c - 1]), 0.5f);
        const bool top = (r == 0) && top_is_border;
        const bool bot = (r == rows_buf - 1) && bottom_is_border;
        if (top)
            gy = __fsub_rn(zr[ld + c], zr[c]);
        else if (bot)
            gy = __fsub_rn(zr[c], zr[c - ld]);
        else
            gy = __fmul_rn(__fsub_rn(zr[ld + c], zr[c - ld]), 0.5f);
        // slope_tan = np.sqrt(gx**2 + gy**2); aspect = np.arctan2(-gx, gy) + np.pi   (affine.py:435-438)
        float st = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
        const float asp = __fadd_rn(atan2f(-gx, gy), 3.14159274101257324f);
        // slope_tan[np.isclose(slope_tan, 0)] = np.nan   (affine.py:578-579; atol 1e-8)
        if (fabsf(st) <= 1e-8f) st = CUDART_NAN_F;
        const long long o = (r - row_begin) * out_ld + c;
        slope_tan[o] = st;
        aspect[o] = asp;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dh = ref - bilinear(tba, row + dy, col + dx): the shift is uniform, so the interpolation is a 2x2 stencil with four
// fixed float64 weights at integer offset (i0, j0); a cell outside the raster or NaN makes the result NaN (also under a
// zero weight: 0*NaN = NaN, like scipy.ndimage.map_coordinates(order=1, cval=nan) in the oracle restatement).
// Also reduces min/max of aspect and the count over finite dh (bin range of binned_statistic, spatialstats.py:147-149).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
nk_dh_kernel(const float* __restrict__ ref, const float* __restrict__ tba, const unsigned char* __restrict__ sub_mask,
             const float* __restrict__ aspect, long long rows, long long cols, long long ld, long long tba_ld,
             long long tba_row0, long long tba_rows_total, long long i0, long long j0, double w00, double w01,
             double w10, double w11, float* __restrict__ dh, unsigned* __restrict__ asp_minmax,
             unsigned long long* __restrict__ n_finite) {
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned long long cnt = 0;
    // one CTA walks whole rows (no 64-bit div/mod per element)
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
    for (long long c = threadIdx.x; c < cols; c += NT) {
        float out = CUDART_NAN_F;
        if (sub_mask[r * cols + c]) {
            // tba buffer row index of raster row (r + i0): local shard rows start at raster row tba_row0
            const long long rr = r + i0 + tba_row0, cc = c + j0;
            double acc = CUDART_NAN;
            if (rr >= 0 && rr + 1 < tba_rows_total && cc >= 0 && cc + 1 < cols) {
                const float* t = tba + rr * tba_ld + cc;
                acc = w00 * (double)t[0] + w01 * (double)t[1] + w10 * (double)t[tba_ld] + w11 * (double)t[tba_ld + 1];
            } else if (rr >= 0 && rr < tba_rows_total && cc >= 0 && cc < cols) {
                // on the last row / column: the out-of-raster neighbours only matter if their weight is non-zero
                const bool row_ok = (rr + 1 < tba_rows_total), col_ok = (cc + 1 < cols);
                const float* t = tba + rr * tba_ld + cc;
                acc = w00 * (double)t[0];
                acc += col_ok ? w01 * (double)t[1] : (w01 != 0.0 ? CUDART_NAN : 0.0);
                acc += row_ok ? w10 * (double)t[tba_ld] : (w10 != 0.0 ? CUDART_NAN : 0.0);
                acc += (row_ok && col_ok) ? w11 * (double)t[tba_ld + 1] : (w11 != 0.0 ? CUDART_NAN : 0.0);
            }
            out = (float)((double)ref[r * ld + c] - acc);
        }
        dh[r * cols + c] = out;
        if (isfinite(out)) {
            const unsigned a = __float_as_uint(aspect[r * cols + c]);  // aspect >= 0: bit pattern is monotonic
            lmin = min(lmin, a);
            lmax = max(lmax, a);
            ++cnt;
        }
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) {
        if (lmin != 0xffffffffu) atomicMin(&asp_minmax[0], lmin);
        if (lmax != 0u || cnt) atomicMax(&asp_minmax[1], lmax);
        if (cnt) atomicAdd(n_finite, cnt);
    }
}

// dh at a LIST of pixels (the reference's default: a random subsample of 5e5 valid points, affine.py:2405,
// base.py:576-621): one thread per point, same FP64 expressions as nk_dh_kernel, outputs compact (dh[i] belongs to
// point i).  aspect_pts is the aspect gathered at the points (compact, same order).
__global__ void __launch_bounds__(NT)
nk_dh_points_kernel(const float* __restrict__ ref, const float* __restrict__ tba, const long long* __restrict__ idx,
                    long long n_pts, const float* __restrict__ aspect_pts, long long cols, long long ld, long long tba_ld,
                    long long tba_row0, long long tba_rows_total, long long i0, long long j0, double w00, double w01,
                    double w10, double w11, float* __restrict__ dh, unsigned* __restrict__ asp_minmax,
                    unsigned long long* __restrict__ n_finite) {
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned long long cnt = 0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n_pts; i += (long long)gridDim.x * NT) {
        const long long pix = idx[i];
        const long long r = pix / cols, c = pix - r * cols;
        const long long rr = r + i0 + tba_row0, cc = c + j0;
        double acc = CUDART_NAN;
        if (rr >= 0 && rr + 1 < tba_rows_total && cc >= 0 && cc + 1 < cols) {
            const float* t = tba + rr * tba_ld + cc;
            acc = w00 * (double)t[0] + w01 * (double)t[1] + w10 * (double)t[tba_ld] + w11 * (double)t[tba_ld + 1];
        } else if (rr >= 0 && rr < tba_rows_total && cc >= 0 && cc < cols) {
            const bool row_ok = (rr + 1 < tba_rows_total), col_ok = (cc + 1 < cols);
            const float* t = tba + rr * tba_ld + cc;
            acc = w00 * (double)t[0];
            acc += col_ok ? w01 * (double)t[1] : (w01 != 0.0 ? CUDART_NAN : 0.0);
            acc += row_ok ? w10 * (double)t[tba_ld] : (w10 != 0.0 ? CUDART_NAN : 0.0);
            acc += (row_ok && col_ok) ? w11 * (double)t[tba_ld + 1] : (w11 != 0.0 ? CUDART_NAN : 0.0);
        }
        const float out = (float)((double)ref[r * ld + c] - acc);
        dh[i] = out;
        if (isfinite(out)) {
            const unsigned a = __float_as_uint(aspect_pts[i]);
            lmin = min(lmin, a);
            lmax = max(lmax, a);
            ++cnt;
        }
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) {
        if (lmin != 0xffffffffu) atomicMin(&asp_minmax[0], lmin);
        if (lmax != 0u || cnt) atomicMax(&asp_minmax[1], lmax);
        if (cnt) atomicAdd(n_finite, cnt);
    }
}

// Same pass, four pixels per thread (cols % 4 == 0, 16-byte aligned rows): vector loads of ref / mask / aspect, a vector
// store of dh, and the ten tba values of the four 2x2 stencils loaded up front (memory-level parallelism: the scalar
// kernel was latency bound, ncu r01b: 29 % of DRAM peak at 31 % issue).  Identical FP64 expressions, identical results.
__global__ void __launch_bounds__(NT)
nk_dh_vec4_kernel(const float* __restrict__ ref, const float* __restrict__ tba, const unsigned char* __restrict__ sub_mask,
                  const float* __restrict__ aspect, long long rows, long long cols, long long ld, long long tba_ld,
                  long long tba_row0, long long tba_rows_total, long long i0, long long j0, double w00, double w01,
                  double w10, double w11, float* __restrict__ dh, unsigned* __restrict__ asp_minmax,
                  unsigned long long* __restrict__ n_finite) {
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned long long cnt = 0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long rr = r + i0 + tba_row0;
        const bool rows_in = rr >= 0 && rr + 1 < tba_rows_total;
        for (long long c = 4ll * threadIdx.x; c < cols; c += 4ll * NT) {
            const uchar4 m4 = *reinterpret_cast<const uchar4*>(sub_mask + r * cols + c);
            const float4 r4 = *reinterpret_cast<const float4*>(ref + r * ld + c);
            const float4 a4 = *reinterpret_cast<const float4*>(aspect + r * cols + c);
            const unsigned char mk[4] = {m4.x, m4.y, m4.z, m4.w};
            const float rf[4] = {r4.x, r4.y, r4.z, r4.w};
            const float as[4] = {a4.x, a4.y, a4.z, a4.w};
            float out[4];
            const long long cc = c + j0;
            if (rows_in && cc >= 0 && cc + 4 < cols) {
                const float* t = tba + rr * tba_ld + cc;
                float ta[5], tb[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) ta[k] = t[k], tb[k] = t[tba_ld + k];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double acc =
                        w00 * (double)ta[k] + w01 * (double)ta[k + 1] + w10 * (double)tb[k] + w11 * (double)tb[k + 1];
                    out[k] = mk[k] ? (float)((double)rf[k] - acc) : CUDART_NAN_F;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    out[k] = CUDART_NAN_F;
                    if (!mk[k]) continue;
                    const long long ck = cc + k;
                    double acc = CUDART_NAN;
                    if (rr >= 0 && rr + 1 < tba_rows_total && ck >= 0 && ck + 1 < cols) {
                        const float* t = tba + rr * tba_ld + ck;
                        acc = w00 * (double)t[0] + w01 * (double)t[1] + w10 * (double)t[tba_ld] +
                              w11 * (double)t[tba_ld + 1];
                    } else if (rr >= 0 && rr < tba_rows_total && ck >= 0 && ck < cols) {
                        const bool row_ok = (rr + 1 < tba_rows_total), col_ok = (ck + 1 < cols);
                        const float* t = tba + rr * tba_ld + ck;
                        acc = w00 * (double)t[0];
                        acc += col_ok ? w01 * (double)t[1] : (w01 != 0.0 ? CUDART_NAN : 0.0);
                        acc += row_ok ? w10 * (double)t[tba_ld] : (w10 != 0.0 ? CUDART_NAN : 0.0);
                        acc += (row_ok && col_ok) ? w11 * (double)t[tba_ld + 1] : (w11 != 0.0 ? CUDART_NAN : 0.0);
                    }
                    out[k] = (float)((double)rf[k] - acc);
                }
            }
            *reinterpret_cast<float4*>(dh + r * cols + c) = make_float4(out[0], out[1], out[2], out[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (isfinite(out[k])) {
                    const unsigned a = __float_as_uint(as[k]);
                    lmin = min(lmin, a);
                    lmax = max(lmax, a);
                    ++cnt;
                }
        }
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) {
        if (lmin != 0xffffffffu) atomicMin(&asp_minmax[0], lmin);
        if (lmax != 0u || cnt) atomicMax(&asp_minmax[1], lmax);
        if (cnt) atomicAdd(n_finite, cnt);
    }
}

// bin of binned_statistic(range=None, bins=n): SciPy builds the edges as np.linspace(lo, hi, n + 1) IN THE SAMPLE DTYPE
// (scipy/stats/_binned_statistic.py:_bin_edges, "preserve sample floating point precision"; float32 for the aspect), i.e.
// float32(k*step + lo) with the last edge = hi exactly; np.digitize semantics, the right-most edge belongs to the last
// bin.
__device__ __forceinline__ int aspect_bin(float a, double lo, double hi, double step, double inv_step, int n_bins) {
    const double x = (double)a;
    const double t = (x - lo) * inv_step;
    int k = (int)t;  // t >= 0 for in-range data
    k = max(0, min(n_bins - 1, k));
    const double frac = t - (double)k;
    // far from an edge (the float32 rounding of an edge moves it by < 3e-6 bin widths for n_bins <= 254 over [0, 2 pi])
    if (frac > 1e-4 && frac < 1.0 - 1e-4) return k;
    // near an edge: compare with the float32 edges exactly as NumPy builds them
    auto edge = [&](int j) -> float {
        return j >= n_bins ? (float)hi : (float)__dadd_rn(__dmul_rn((double)j, step), lo);
    };
    while (k > 0 && a < edge(k)) --k;
    while (k < n_bins - 1 && a >= edge(k + 1)) ++k;
    return k;
}

// y key of one element: float32((dh - vshift) / slope_tan) (affine.py:381, 505).  The subtraction is done in float64
// like the reference, the division in IEEE float32 (the key is a float32 anyway; <= 1 ulp from rounding the float64
// quotient).  Returns false for non-finite results.
__device__ __forceinline__ bool y_key(float dhv, float st, double vshift, unsigned& key, float& yf) {
    if (!isfinite(dhv)) return false;
    yf = __fdiv_rn((float)((double)dhv - vshift), st);
    if (!isfinite(yf)) return false;
    key = ordered_key(yf);
    return true;
}

// ---- global select on dh (np.nanmedian(dh), affine.py:504): per-CTA shared-memory histogram of one digit ----------
__global__ void __launch_bounds__(NT)
nk_hist_kernel(const float* __restrict__ dh, long long n, unsigned prefix, unsigned prefix_mask, int shift,
               int n_digits, unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned sh_hist[];
    for (int k = threadIdx.x; k < n_digits; k += NT) sh_hist[k] = 0u;
    __syncthreads();
    // 4 independent coalesced loads in flight per thread (memory-level parallelism), then the shared-memory updates
    const long long stride = (long long)gridDim.x * NT;
    for (long long i0 = (long long)blockIdx.x * NT + threadIdx.x; i0 < n; i0 += 4 * stride) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i0 + u * stride < n) ? dh[i0 + u * stride] : CUDART_NAN_F;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!isfinite(v[u])) continue;
            const unsigned key = ordered_key(v[u]);
            if ((key & prefix_mask) != prefix) continue;
            atomicAdd(&sh_hist[(key >> shift) & (unsigned)(n_digits - 1)], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_digits; k += NT)
        if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
}

__global__ void __launch_bounds__(NT)
nk_next_kernel(const float* __restrict__ dh, long long n, unsigned sel, unsigned* __restrict__ next_key) {
    unsigned best = 0xffffffffu;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const float v = dh[i];
        if (!isfinite(v)) continue;
        const unsigned key = ordered_key(v);
        if (key > sel) best = min(best, key);
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best != 0xffffffffu) atomicMin(next_key, best);
}

// ---- grouped select (72-bin nanmedian of y over aspect, base.py:1014-1020) -----------------------------------------
// Pass 0 computes, once per iteration, the compact (key, group) pair of every element (group 255 = excluded), the first
// digit histogram (per-CTA shared memory: n_groups x n_digits counters) and the moments [n, sum y, sum y^2] of y for the
// initial guess p0 (affine.py:384).  Later passes stream only the 5-byte pairs.
__global__ void __launch_bounds__(NT)
nk_make_keys_kernel(const float* __restrict__ dh, const float* __restrict__ slope_tan, const float* __restrict__ aspect,
                    unsigned char* __restrict__ grp_cache,
                    long long n, double vshift, double asp_lo, double asp_hi, int n_groups,
                    unsigned* __restrict__ key_out, unsigned char* __restrict__ grp_out, int reuse_groups, int shift,
                    int n_digits, unsigned long long* __restrict__ hist, double* __restrict__ moments) {
    extern __shared__ unsigned sh_hist[];
    const int n_cnt = n_groups * n_digits;
    for (int k = threadIdx.x; k < n_cnt; k += NT) sh_hist[k] = 0u;
    __syncthreads();
    const double step = (asp_hi - asp_lo) / (double)n_groups;
    const double inv_step = step > 0.0 ? 1.0 / step : 0.0;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    const long long stride = (long long)gridDim.x * NT;
    for (long long i0 = (long long)blockIdx.x * NT + threadIdx.x; i0 < n; i0 += 4 * stride) {
        float dv[4], sv[4], av[4];
        unsigned char bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            const bool in = i < n;
            dv[u] = in ? dh[i] : CUDART_NAN_F;
            sv[u] = in ? slope_tan[i] : CUDART_NAN_F;
            av[u] = (in && !reuse_groups) ? aspect[i] : CUDART_NAN_F;
            bv[u] = (in && reuse_groups) ? grp_cache[i] : (unsigned char)255;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            if (i >= n) break;
            unsigned key = 0u;
            float yf = 0.f;
            unsigned char grp = 255;
            // a pixel's aspect bin only depends on [asp_lo, asp_hi]: it is computed for EVERY pixel with a finite
            // aspect (valid or not this iteration) and cached; later iterations with the same range read it back
            unsigned char bin = bv[u];
            if (!reuse_groups) {
                bin = isfinite(av[u]) ? (unsigned char)aspect_bin(av[u], asp_lo, asp_hi, step, inv_step, n_groups)
                                      : (unsigned char)255;
                grp_cache[i] = bin;
            }
            if (y_key(dv[u], sv[u], vshift, key, yf) && bin != 255) {
                grp = bin;
                m0 += 1.0;
                m1 += (double)yf;
                m2 += (double)yf * (double)yf;
                atomicAdd(&sh_hist[(int)grp * n_digits + (int)((key >> shift) & (unsigned)(n_digits - 1))], 1u);
            }
            key_out[i] = key;
            grp_out[i] = grp;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_cnt; k += NT)
        if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o);
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    if ((threadIdx.x & 31) == 0 && m0 > 0.0) {
        atomicAdd(&moments[0], m0);
        atomicAdd(&moments[1], m1);
        atomicAdd(&moments[2], m2);
    }
}

// NTH threads share one set of n_groups x n_digits counters: with 72 x 256 counters (72 KB) only 2-3 CTAs fit an SM, so
// the CTA is made large (1024 threads) to keep the SM's warp slots full (ncu r01b: 37 % warps active at 256 threads).
template <int NTH>
__global__ void __launch_bounds__(NTH)
nk_hist_keys_kernel(const unsigned* __restrict__ key, const unsigned char* __restrict__ grp, long long n, int n_groups,
                    const unsigned* __restrict__ prefix, unsigned prefix_mask, int shift, int n_digits,
                    unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned sh_hist[];
    const int n_cnt = n_groups * n_digits;
    for (int k = threadIdx.x; k < n_cnt; k += NTH) sh_hist[k] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * NTH;
    for (long long i0 = (long long)blockIdx.x * NTH + threadIdx.x; i0 < n; i0 += 4 * stride) {
        unsigned g[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool in = i0 + u * stride < n;
            g[u] = in ? grp[i0 + u * stride] : 255u;
            k[u] = in ? key[i0 + u * stride] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g[u] >= (unsigned)n_groups) continue;
            if ((k[u] & prefix_mask) != prefix[g[u]]) continue;
            atomicAdd(&sh_hist[g[u] * n_digits + ((k[u] >> shift) & (unsigned)(n_digits - 1))], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_cnt; k += NTH)
        if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
}

__global__ void __launch_bounds__(NT)
nk_next_keys_kernel(const unsigned* __restrict__ key, const unsigned char* __restrict__ grp, long long n, int n_groups,
                    const unsigned* __restrict__ sel, unsigned* __restrict__ next_key) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long i0 = (long long)blockIdx.x * NT + threadIdx.x; i0 < n; i0 += 4 * stride) {
        unsigned g[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool in = i0 + u * stride < n;
            g[u] = in ? grp[i0 + u * stride] : 255u;
            k[u] = in ? key[i0 + u * stride] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g[u] >= (unsigned)n_groups) continue;
            if (k[u] > sel[g[u]] && k[u] < next_key[g[u]]) atomicMin(&next_key[g[u]], k[u]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Translation-only `apply` (base.py:1567-1570, 1755-1760): dst = bilinear(src at (row + dy, col + dx)) + dz, same 2x2
// fixed-weight stencil and NaN rule as nk_dh_kernel (the regrid of `_reproject_horizontal_shift_samecrs`).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
shift_resample_kernel(const float* __restrict__ src, long long rows, long long cols, long long ld, long long i0,
                      long long j0, double w00, double w01, double w10, double w11, double dz,
                      float* __restrict__ dst, long long dst_ld) {
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
    for (long long c = threadIdx.x; c < cols; c += NT) {
        const long long rr = r + i0, cc = c + j0;
        double acc = CUDART_NAN;
        if (rr >= 0 && rr + 1 < rows && cc >= 0 && cc + 1 < cols) {
            const float* t = src + rr * ld + cc;
            acc = w00 * (double)t[0] + w01 * (double)t[1] + w10 * (double)t[ld] + w11 * (double)t[ld + 1];
        } else if (rr >= 0 && rr < rows && cc >= 0 && cc < cols) {
            const bool row_ok = (rr + 1 < rows), col_ok = (cc + 1 < cols);
            const float* t = src + rr * ld + cc;
            acc = w00 * (double)t[0];
            acc += col_ok ? w01 * (double)t[1] : (w01 != 0.0 ? CUDART_NAN : 0.0);
            acc += row_ok ? w10 * (double)t[ld] : (w10 != 0.0 ? CUDART_NAN : 0.0);
            acc += (row_ok && col_ok) ? w11 * (double)t[ld + 1] : (w11 != 0.0 ? CUDART_NAN : 0.0);
        }
        dst[r * dst_ld + c] = (float)(acc + dz);
    }
}

static int grid_for(long long n, int per_sm) {
    int sms = 0;
    if (xb_num_sms(&sms)) sms = 148;
    long long g = (n + NT - 1) / NT;
    const long long cap = (long long)sms * per_sm;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace xbn

extern "C" {
#pragma GCC visibility push(default)

int xb_nk_aux(const float* ref_dev, int64_t rows_buf, int64_t cols, int64_t ld, int top_is_border,
              int bottom_is_border, int64_t row_begin, int64_t row_end, float* slope_tan_dev, float* aspect_dev,
              int64_t out_ld, void* stream) {
    if (!ref_dev || !slope_tan_dev || !aspect_dev || rows_buf < 2 || cols < 2 || ld < cols || row_begin < 0 ||
        row_end > rows_buf || row_begin > row_end || out_ld < cols) {
        xb_set_error("bad arguments to xb_nk_aux (np.gradient needs at least 2 rows and 2 columns)");
        return XB_ERR_INVALID;
    }
    if ((!top_is_border && row_begin < 1) || (!bottom_is_border && row_end > rows_buf - 1)) {
        xb_set_error("xb_nk_aux: interior shards need one halo row above/below the output rows");
        return XB_ERR_INVALID;
    }
    const long long n = (row_end - row_begin) * cols;
    if (n == 0) return XB_OK;
    xbn::nk_aux_kernel<<<xbn::grid_for(n, 16), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        ref_dev, rows_buf, cols, ld, top_is_border, bottom_is_border, row_begin, row_end, slope_tan_dev, aspect_dev,
        out_ld);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_dh(const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev, const float* aspect_dev,
             int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0, int64_t tba_rows_total,
             double dx_px, double dy_px, float* dh_dev, uint32_t* asp_minmax_dev, unsigned long long* n_finite_dev,
             void* stream) {
    if (!ref_dev || !tba_dev || !sub_mask_dev || !aspect_dev || !dh_dev || !asp_minmax_dev || !n_finite_dev ||
        rows <= 0 || cols <= 0 || ld < cols || tba_ld < cols) {
        xb_set_error("bad arguments to xb_nk_dh");
        return XB_ERR_INVALID;
    }
    if (!isfinite(dx_px) || !isfinite(dy_px) || fabs(dx_px) > 1e9 || fabs(dy_px) > 1e9) {
        xb_set_error("xb_nk_dh: non-finite shift");
        return XB_ERR_INVALID;
    }
    const double fi = floor(dy_px), fj = floor(dx_px);
    const double fy = dy_px - fi, fx = dx_px - fj;
    const double w00 = (1.0 - fy) * (1.0 - fx), w01 = (1.0 - fy) * fx, w10 = fy * (1.0 - fx), w11 = fy * fx;
    const long long n = rows * cols;
    const bool vec4 = cols % 4 == 0 && ld % 4 == 0 && cols >= 4 &&
                      ((reinterpret_cast<uintptr_t>(ref_dev) | reinterpret_cast<uintptr_t>(aspect_dev) |
                        reinterpret_cast<uintptr_t>(dh_dev)) % 16 == 0) &&
                      reinterpret_cast<uintptr_t>(sub_mask_dev) % 4 == 0;
    if (vec4) {
        int grid = xbn::grid_for(n / 4, 8);
        if (grid > rows) grid = (int)rows;
        xbn::nk_dh_vec4_kernel<<<grid, xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            ref_dev, tba_dev, sub_mask_dev, aspect_dev, rows, cols, ld, tba_ld, tba_row0, tba_rows_total, (long long)fi,
            (long long)fj, w00, w01, w10, w11, dh_dev, asp_minmax_dev, n_finite_dev);
    } else {
        xbn::nk_dh_kernel<<<xbn::grid_for(n, 16), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
            ref_dev, tba_dev, sub_mask_dev, aspect_dev, rows, cols, ld, tba_ld, tba_row0, tba_rows_total, (long long)fi,
            (long long)fj, w00, w01, w10, w11, dh_dev, asp_minmax_dev, n_finite_dev);
    }
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_dh_points(const float* ref_dev, const float* tba_dev, const int64_t* idx_dev, int64_t n_pts,
                    const float* aspect_pts_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
                    int64_t tba_rows_total, double dx_px, double dy_px, float* dh_dev, uint32_t* asp_minmax_dev,
                    unsigned long long* n_finite_dev, void* stream) {
    if (!ref_dev || !tba_dev || !idx_dev || !aspect_pts_dev || !dh_dev || !asp_minmax_dev || !n_finite_dev || n_pts <= 0 ||
        rows <= 0 || cols <= 0 || ld < cols || tba_ld < cols) {
        xb_set_error("bad arguments to xb_nk_dh_points");
        return XB_ERR_INVALID;
    }
    if (!isfinite(dx_px) || !isfinite(dy_px) || fabs(dx_px) > 1e9 || fabs(dy_px) > 1e9) {
        xb_set_error("xb_nk_dh_points: non-finite shift");
        return XB_ERR_INVALID;
    }
    const double fi = floor(dy_px), fj = floor(dx_px);
    const double fy = dy_px - fi, fx = dx_px - fj;
    xbn::nk_dh_points_kernel<<<xbn::grid_for(n_pts, 8), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        ref_dev, tba_dev, reinterpret_cast<const long long*>(idx_dev), n_pts, aspect_pts_dev, cols, ld, tba_ld, tba_row0,
        tba_rows_total, (long long)fi, (long long)fj, (1.0 - fy) * (1.0 - fx), (1.0 - fy) * fx, fy * (1.0 - fx), fy * fx,
        dh_dev, asp_minmax_dev, n_finite_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_shift_resample(const float* src_dev, int64_t rows, int64_t cols, int64_t ld, double dx_px, double dy_px,
                      double dz, float* dst_dev, int64_t dst_ld, void* stream) {
    if (!src_dev || !dst_dev || rows <= 0 || cols <= 0 || ld < cols || dst_ld < cols || !isfinite(dx_px) ||
        !isfinite(dy_px) || fabs(dx_px) > 1e9 || fabs(dy_px) > 1e9) {
        xb_set_error("bad arguments to xb_shift_resample");
        return XB_ERR_INVALID;
    }
    const double fi = floor(dy_px), fj = floor(dx_px);
    const double fy = dy_px - fi, fx = dx_px - fj;
    xbn::shift_resample_kernel<<<xbn::grid_for(rows * cols, 16), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src_dev, rows, cols, ld, (long long)fi, (long long)fj, (1.0 - fy) * (1.0 - fx), (1.0 - fy) * fx,
        fy * (1.0 - fx), fy * fx, dz, dst_dev, dst_ld);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_hist(const float* dh_dev, int64_t n, uint32_t prefix, uint32_t prefix_mask, int shift, int n_digits,
               unsigned long long* hist_dev, void* stream) {
    if (!dh_dev || !hist_dev || n <= 0 || n_digits < 2 || (n_digits & (n_digits - 1)) || n_digits > 4096) {
        xb_set_error("bad arguments to xb_nk_hist");
        return XB_ERR_INVALID;
    }
    xbn::nk_hist_kernel<<<xbn::grid_for(n, 8), xbn::NT, (size_t)n_digits * sizeof(unsigned),
                          reinterpret_cast<cudaStream_t>(stream)>>>(dh_dev, n, prefix, prefix_mask, shift, n_digits,
                                                                    hist_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_next(const float* dh_dev, int64_t n, uint32_t sel, uint32_t* next_key_dev, void* stream) {
    if (!dh_dev || !next_key_dev || n <= 0) {
        xb_set_error("bad arguments to xb_nk_next");
        return XB_ERR_INVALID;
    }
    xbn::nk_next_kernel<<<xbn::grid_for(n, 8), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dh_dev, n, sel,
                                                                                                      next_key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

static int grouped_smem(int n_groups, int n_digits, size_t* smem) {
    if (n_groups < 1 || n_groups > 254 || n_digits < 2 || (n_digits & (n_digits - 1)) ||
        (size_t)n_groups * n_digits * sizeof(unsigned) > 160 * 1024) {
        xb_set_error("grouped select needs 1..254 groups and n_groups*n_digits*4 <= 160 KiB (got %d x %d)", n_groups,
                     n_digits);
        return XB_ERR_INVALID;
    }
    *smem = (size_t)n_groups * n_digits * sizeof(unsigned);
    return XB_OK;
}

int xb_nk_make_keys(const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, int64_t n, double vshift,
                    double asp_lo, double asp_hi, int n_groups, uint32_t* key_dev, uint8_t* group_dev,
                    uint8_t* bin_cache_dev, int reuse_bins, int shift, int n_digits, unsigned long long* hist_dev,
                    double* moments_dev, void* stream) {
    size_t smem = 0;
    if (!dh_dev || !slope_tan_dev || !aspect_dev || !key_dev || !group_dev || !bin_cache_dev || !hist_dev ||
        !moments_dev || n <= 0) {
        xb_set_error("bad arguments to xb_nk_make_keys");
        return XB_ERR_INVALID;
    }
    int rc = grouped_smem(n_groups, n_digits, &smem);
    if (rc) return rc;
    XB_CUDA_CHECK(cudaFuncSetAttribute(xbn::nk_make_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xbn::nk_make_keys_kernel<<<xbn::grid_for(n, smem > 72 * 1024 ? 1 : (smem > 36 * 1024 ? 3 : 6)), xbn::NT, smem,
                               reinterpret_cast<cudaStream_t>(stream)>>>(
        dh_dev, slope_tan_dev, aspect_dev, bin_cache_dev, n, vshift, asp_lo, asp_hi, n_groups, key_dev, group_dev,
        reuse_bins ? 1 : 0, shift, n_digits, hist_dev, moments_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_hist_keys(const uint32_t* key_dev, const uint8_t* group_dev, int64_t n, int n_groups,
                    const uint32_t* prefix_dev, uint32_t prefix_mask, int shift, int n_digits,
                    unsigned long long* hist_dev, void* stream) {
    size_t smem = 0;
    if (!key_dev || !group_dev || !prefix_dev || !hist_dev || n <= 0) {
        xb_set_error("bad arguments to xb_nk_hist_keys");
        return XB_ERR_INVALID;
    }
    int rc = grouped_smem(n_groups, n_digits, &smem);
    if (rc) return rc;
    constexpr int NTH = 1024;
    auto kern = xbn::nk_hist_keys_kernel<NTH>;
    XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int num_sms = 0;
    rc = xb_num_sms(&num_sms);
    if (rc) return rc;
    long long grid = (long long)num_sms * (smem > 100 * 1024 ? 1 : 2);
    const long long need = (n + NTH - 1) / NTH;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, NTH, smem, reinterpret_cast<cudaStream_t>(stream)>>>(key_dev, group_dev, n, n_groups, prefix_dev,
                                                                                prefix_mask, shift, n_digits, hist_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_next_keys(const uint32_t* key_dev, const uint8_t* group_dev, int64_t n, int n_groups,
                    const uint32_t* sel_dev, uint32_t* next_key_dev, void* stream) {
    if (!key_dev || !group_dev || !sel_dev || !next_key_dev || n <= 0 || n_groups < 1 || n_groups > 254) {
        xb_set_error("bad arguments to xb_nk_next_keys");
        return XB_ERR_INVALID;
    }
    xbn::nk_next_keys_kernel<<<xbn::grid_for(n, 8), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        key_dev, group_dev, n, n_groups, sel_dev, next_key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
