// xdem_b200 -- texture shading (Brown 2010 fractional Laplacian) device stages for sm_100a.
//
// The reference (xdem/terrain/freq.py:62-148) does: NaN -> nanmean fill, symmetric pad to a 2/3/5/7-smooth size,
// rfft2, multiply by |f|^alpha (DC zeroed when alpha > 0), irfft2, crop, restore NaN.  The two FFTs are plain library
// transforms (cuFFT, called by the host side); the three elementwise stages around them are the kernels below, each a
// single coalesced pass:
//   xb_texture_prepare : one reduction (sum / count of the non-NaN cells, count of finite cells) + fill-and-pad
//   xb_texture_filter  : in-place scale of the half spectrum by |f|^alpha, filter evaluated in float64 per element
//                        (the reference multiplies its complex64 spectrum by a float64 filter, freq.py:124-137)
//   xb_texture_finish  : crop the padded result and put NaN back where the input was not finite
#include "../../include/xdem_b200.h"

#include <math_constants.h>

#include "xb_common.cuh"

void xb_count_launch(int n);

namespace xbx {

constexpr int NT = 256;

template <typename T> __device__ __forceinline__ T nan_of();
template <> __device__ __forceinline__ float nan_of<float>() { return CUDART_NAN_F; }
template <> __device__ __forceinline__ double nan_of<double>() { return CUDART_NAN; }

// stats[0] = sum over non-NaN cells (np.nanmean keeps +-inf), stats[1] = their count, stats[2] = count of finite cells
template <typename T>
__global__ void __launch_bounds__(NT)
stats_kernel(const T* __restrict__ dem, long long rows, long long cols, long long ld, double* __restrict__ stats) {
    double s = 0.0, n = 0.0, nf = 0.0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
        for (long long c = threadIdx.x; c < cols; c += NT) {
            const T v = dem[r * ld + c];
            if (v == v) s += (double)v, n += 1.0;
            if (isfinite(v)) nf += 1.0;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n > 0.0) atomicAdd(&stats[0], s), atomicAdd(&stats[1], n);
        if (nf > 0.0) atomicAdd(&stats[2], nf);
    }
}

// np.pad(mode="symmetric") source index (edge value repeated): ... 1 0 | 0 1 2 ... n-1 | n-1 n-2 ...
__device__ __forceinline__ long long reflect_index(long long i, long long n) {
    if (n <= 1) return 0;
    const long long period = 2 * n;
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - 1 - i;
}

template <typename T>
__global__ void __launch_bounds__(NT)
pad_kernel(const T* __restrict__ dem, long long rows, long long cols, long long ld, T* __restrict__ padded,
           long long fft_rows, long long fft_cols, long long pad_r, long long pad_c, const double* __restrict__ stats,
           int subtract_mean) {
    // freq.py:93-94: non-finite cells take the mean.  With subtract_mean the whole raster is centred on that mean first
    // (the filter zeroes the DC term for alpha > 0, so the result is unchanged) -- elevations ~1e3 m with ~10 m of relief
    // would otherwise spend 2 of float32's 7 digits on the constant; the difference is taken in float64.
    const double mean = stats[0] / stats[1];
    const T fill = subtract_mean ? (T)0 : (T)mean;
    for (long long i = blockIdx.x; i < fft_rows; i += gridDim.x) {
        const long long si = reflect_index(i - pad_r, rows);
        for (long long j = threadIdx.x; j < fft_cols; j += NT) {
            const long long sj = reflect_index(j - pad_c, cols);
            const T v = dem[si * ld + sj];
            padded[i * fft_cols + j] = isfinite(v) ? (subtract_mean ? (T)((double)v - mean) : v) : fill;
        }
    }
}

template <typename C> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

template <typename T>
__global__ void __launch_bounds__(NT)
filter_kernel(typename Cplx<T>::type* __restrict__ spec, long long fft_rows, long long half_cols, long long fft_cols,
              double alpha) {
    const double inv_r = 1.0 / (double)fft_rows, inv_c = 1.0 / (double)fft_cols;
    for (long long i = blockIdx.x; i < fft_rows; i += gridDim.x) {
        // scipy.fft.fftfreq(n): k / n for k = 0 .. (n-1)//2, then (k - n) / n
        const double fy = (double)(i <= (fft_rows - 1) / 2 ? i : i - fft_rows) * inv_r;
        for (long long j = threadIdx.x; j < half_cols; j += NT) {
            const double fx = (double)j * inv_c;  // rfftfreq
            double filt;
            if (i == 0 && j == 0)
                filt = alpha > 0.0 ? 0.0 : 1.0;  // freq.py:122, 130-131
            else
                filt = pow(fx * fx + fy * fy, 0.5 * alpha);
            typename Cplx<T>::type z = spec[i * half_cols + j];
            z.x = (T)((double)z.x * filt);
            z.y = (T)((double)z.y * filt);
            spec[i * half_cols + j] = z;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(NT)
finish_kernel(const T* __restrict__ padded, long long fft_cols, long long pad_r, long long pad_c,
              const T* __restrict__ dem, long long rows, long long cols, long long ld, T* __restrict__ out,
              long long out_ld) {
    for (long long r = blockIdx.x; r < rows; r += gridDim.x)
        for (long long c = threadIdx.x; c < cols; c += NT) {
            const T v = dem[r * ld + c];
            out[r * out_ld + c] = isfinite(v) ? padded[(r + pad_r) * fft_cols + (c + pad_c)] : nan_of<T>();
        }
}

static int row_grid(long long rows) {
    int sms = 0;
    if (xb_num_sms(&sms)) sms = 148;
    const long long cap = (long long)sms * 8;
    return (int)(rows < cap ? (rows < 1 ? 1 : rows) : cap);
}

}  // namespace xbx

extern "C" {
#pragma GCC visibility push(default)

int xb_texture_prepare(const void* dem_dev, int dtype, int64_t rows, int64_t cols, int64_t ld, void* padded_dev,
                       int64_t fft_rows, int64_t fft_cols, int64_t pad_rows, int64_t pad_cols, int subtract_mean,
                       double* stats_dev, void* stream) {
    if (!dem_dev || !padded_dev || !stats_dev || rows <= 0 || cols <= 0 || ld < cols || fft_rows < rows ||
        fft_cols < cols || pad_rows < 0 || pad_cols < 0 || pad_rows + rows > fft_rows || pad_cols + cols > fft_cols ||
        (dtype != 0 && dtype != 1)) {
        xb_set_error("bad arguments to xb_texture_prepare");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    XB_CUDA_CHECK(cudaMemsetAsync(stats_dev, 0, 3 * sizeof(double), st));
    if (dtype == 0) {
        xbx::stats_kernel<float><<<xbx::row_grid(rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const float*>(dem_dev), rows, cols, ld, stats_dev);
        xbx::pad_kernel<float><<<xbx::row_grid(fft_rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const float*>(dem_dev), rows, cols, ld, reinterpret_cast<float*>(padded_dev), fft_rows,
            fft_cols, pad_rows, pad_cols, stats_dev, subtract_mean);
    } else {
        xbx::stats_kernel<double><<<xbx::row_grid(rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const double*>(dem_dev), rows, cols, ld, stats_dev);
        xbx::pad_kernel<double><<<xbx::row_grid(fft_rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const double*>(dem_dev), rows, cols, ld, reinterpret_cast<double*>(padded_dev), fft_rows,
            fft_cols, pad_rows, pad_cols, stats_dev, subtract_mean);
    }
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(2);
    return XB_OK;
}

int xb_texture_filter(void* spectrum_dev, int dtype, int64_t fft_rows, int64_t fft_cols, double alpha, void* stream) {
    if (!spectrum_dev || fft_rows <= 0 || fft_cols <= 0 || !(alpha >= 0.0 && alpha <= 2.0) ||
        (dtype != 0 && dtype != 1)) {
        xb_set_error("bad arguments to xb_texture_filter (alpha must be within [0, 2])");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long half = fft_cols / 2 + 1;
    if (dtype == 0)
        xbx::filter_kernel<float><<<xbx::row_grid(fft_rows), xbx::NT, 0, st>>>(
            reinterpret_cast<float2*>(spectrum_dev), fft_rows, half, fft_cols, alpha);
    else
        xbx::filter_kernel<double><<<xbx::row_grid(fft_rows), xbx::NT, 0, st>>>(
            reinterpret_cast<double2*>(spectrum_dev), fft_rows, half, fft_cols, alpha);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_texture_finish(const void* padded_dev, int dtype, int64_t fft_rows, int64_t fft_cols, int64_t pad_rows,
                      int64_t pad_cols, const void* dem_dev, int64_t rows, int64_t cols, int64_t ld, void* out_dev,
                      int64_t out_ld, void* stream) {
    if (!padded_dev || !dem_dev || !out_dev || rows <= 0 || cols <= 0 || ld < cols || out_ld < cols ||
        pad_rows < 0 || pad_cols < 0 || pad_rows + rows > fft_rows || pad_cols + cols > fft_cols ||
        (dtype != 0 && dtype != 1)) {
        xb_set_error("bad arguments to xb_texture_finish");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == 0)
        xbx::finish_kernel<float><<<xbx::row_grid(rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const float*>(padded_dev), fft_cols, pad_rows, pad_cols,
            reinterpret_cast<const float*>(dem_dev), rows, cols, ld, reinterpret_cast<float*>(out_dev), out_ld);
    else
        xbx::finish_kernel<double><<<xbx::row_grid(rows), xbx::NT, 0, st>>>(
            reinterpret_cast<const double*>(padded_dev), fft_cols, pad_rows, pad_cols,
            reinterpret_cast<const double*>(dem_dev), rows, cols, ld, reinterpret_cast<double*>(out_dev), out_ld);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
