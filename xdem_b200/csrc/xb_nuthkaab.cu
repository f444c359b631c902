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
    const long long n = (row_end - row_begin) * cols;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const long long r = row_begin + i / cols, c = i % cols;
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
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const long long r = i / cols, c = i % cols;
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

// bin of binned_statistic(range=None, bins=n): edges = linspace(lo, hi, n+1) (= k*step + lo, last edge = hi exactly),
// np.digitize semantics, the right-most edge belongs to the last bin.
__device__ __forceinline__ int aspect_bin(float a, double lo, double hi, double step, double inv_step, int n_bins) {
    const double x = (double)a;
    int k = (int)floor((x - lo) * inv_step);
    k = max(0, min(n_bins - 1, k));
    while (k > 0 && x < __dadd_rn(__dmul_rn((double)k, step), lo)) --k;
    while (k < n_bins - 1) {
        const double e = (k + 1 == n_bins) ? hi : __dadd_rn(__dmul_rn((double)(k + 1), step), lo);
        if (x >= e) ++k; else break;
    }
    return k;
}

struct KeyGroup {
    unsigned key;
    int group;
    bool ok;
    double y;
};

// mode 0: key = dh (one group).  mode 1: key = float32((dh - vshift)/slope_tan) (affine.py:381, 505), group = aspect bin.
__device__ __forceinline__ KeyGroup make_key(int mode, float dhv, const float* __restrict__ slope_tan,
                                             const float* __restrict__ aspect, long long i, double vshift,
                                             double asp_lo, double asp_hi, double step, double inv_step, int n_groups) {
    KeyGroup kg;
    kg.ok = isfinite(dhv);
    kg.group = 0;
    kg.key = 0;
    kg.y = 0.0;
    if (!kg.ok) return kg;
    if (mode == 0) {
        kg.key = ordered_key(dhv);
        return kg;
    }
    const double y = ((double)dhv - vshift) / (double)slope_tan[i];
    const float yf = (float)y;
    kg.ok = isfinite(yf);
    if (!kg.ok) return kg;
    kg.y = y;
    kg.key = ordered_key(yf);
    kg.group = aspect_bin(aspect[i], asp_lo, asp_hi, step, inv_step, n_groups);
    return kg;
}

// One MSD radix-select pass: for keys whose bits under prefix_mask equal prefix[group]: hist[group][digit]++ with
// digit = (key >> shift) & (n_digits-1).  moments (mode 1, first pass): [n, sum y, sum y^2] for p0 (affine.py:384).
__global__ void __launch_bounds__(NT)
nk_hist_kernel(const float* __restrict__ dh, const float* __restrict__ slope_tan, const float* __restrict__ aspect,
               long long n, int mode, double vshift, double asp_lo, double asp_hi, int n_groups,
               const unsigned* __restrict__ prefix, unsigned prefix_mask, int shift, int n_digits,
               unsigned long long* __restrict__ hist, double* __restrict__ moments) {
    extern __shared__ unsigned sh_hist[];  // n_digits counters when n_groups == 1
    const bool use_smem = (n_groups == 1);
    if (use_smem) {
        for (int k = threadIdx.x; k < n_digits; k += NT) sh_hist[k] = 0u;
        __syncthreads();
    }
    const double step = (asp_hi - asp_lo) / (double)n_groups;
    const double inv_step = step > 0.0 ? 1.0 / step : 0.0;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const KeyGroup kg = make_key(mode, dh[i], slope_tan, aspect, i, vshift, asp_lo, asp_hi, step, inv_step, n_groups);
        if (!kg.ok) continue;
        if (moments) {
            m0 += 1.0;
            m1 += kg.y;
            m2 += kg.y * kg.y;
        }
        if ((kg.key & prefix_mask) != prefix[kg.group]) continue;
        const unsigned digit = (kg.key >> shift) & (unsigned)(n_digits - 1);
        if (use_smem)
            atomicAdd(&sh_hist[digit], 1u);
        else
            atomicAdd(&hist[(long long)kg.group * n_digits + digit], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_digits; k += NT)
            if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
    }
    if (moments) {
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
}

// smallest ordered key strictly greater than sel[group] (for the upper median of even-sized groups)
__global__ void __launch_bounds__(NT)
nk_next_kernel(const float* __restrict__ dh, const float* __restrict__ slope_tan, const float* __restrict__ aspect,
               long long n, int mode, double vshift, double asp_lo, double asp_hi, int n_groups,
               const unsigned* __restrict__ sel, unsigned* __restrict__ next_key) {
    const double step = (asp_hi - asp_lo) / (double)n_groups;
    const double inv_step = step > 0.0 ? 1.0 / step : 0.0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const KeyGroup kg = make_key(mode, dh[i], slope_tan, aspect, i, vshift, asp_lo, asp_hi, step, inv_step, n_groups);
        if (!kg.ok) continue;
        if (kg.key > sel[kg.group]) atomicMin(&next_key[kg.group], kg.key);
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
    xbn::nk_dh_kernel<<<xbn::grid_for(n, 16), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        ref_dev, tba_dev, sub_mask_dev, aspect_dev, rows, cols, ld, tba_ld, tba_row0, tba_rows_total, (long long)fi,
        (long long)fj, w00, w01, w10, w11, dh_dev, asp_minmax_dev, n_finite_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_hist(const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, int64_t n, int mode,
               double vshift, double asp_lo, double asp_hi, int n_groups, const uint32_t* prefix_dev,
               uint32_t prefix_mask, int shift, int n_digits, unsigned long long* hist_dev, double* moments_dev,
               void* stream) {
    if (!dh_dev || !prefix_dev || !hist_dev || n <= 0 || n_groups < 1 || n_digits < 2 || (n_digits & (n_digits - 1)) ||
        n_digits > 4096 || (mode != 0 && mode != 1) || (mode == 1 && (!slope_tan_dev || !aspect_dev)) ||
        (mode == 0 && n_groups != 1)) {
        xb_set_error("bad arguments to xb_nk_hist");
        return XB_ERR_INVALID;
    }
    const size_t smem = n_groups == 1 ? (size_t)n_digits * sizeof(unsigned) : 0;
    xbn::nk_hist_kernel<<<xbn::grid_for(n, 8), xbn::NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        dh_dev, slope_tan_dev, aspect_dev, n, mode, vshift, asp_lo, asp_hi, n_groups, prefix_dev, prefix_mask, shift,
        n_digits, hist_dev, moments_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nk_next(const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, int64_t n, int mode,
               double vshift, double asp_lo, double asp_hi, int n_groups, const uint32_t* sel_dev,
               uint32_t* next_key_dev, void* stream) {
    if (!dh_dev || !sel_dev || !next_key_dev || n <= 0 || n_groups < 1 || (mode != 0 && mode != 1)) {
        xb_set_error("bad arguments to xb_nk_next");
        return XB_ERR_INVALID;
    }
    xbn::nk_next_kernel<<<xbn::grid_for(n, 8), xbn::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        dh_dev, slope_tan_dev, aspect_dev, n, mode, vshift, asp_lo, asp_hi, n_groups, sel_dev, next_key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
