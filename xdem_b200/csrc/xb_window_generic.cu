// xdem_b200 -- generic odd-window indexes (any window_size up to 31) and fractal roughness, sm_100a.
//
// Complements the fused 3x3 / 5x5 kernel (xb_terrain.cu) for the reference's arbitrary `window_size`
// (window.py:926-1002) and for `fractal_roughness` (Taud & Parrot 2005 box counting, window.py:317-379, default
// 13x13 window).  One pixel per thread, the (tile + halo) block is staged in shared memory
// by a cooperative NaN-filling loader; sums run in the reference's row-major sequential order with un-contracted IEEE
// ops so that integer-valued DEMs are bit-exact and TRI / odd-window TPI match the Numba engine bit-for-bit.
#include "../../include/xdem_b200.h"

#include <math_constants.h>

#include "xb_common.cuh"

void xb_count_launch(int n);

namespace xbw {

constexpr int TX = 32, TY = 8, RPT = 4;  // CTA = 32 x 8 threads, each thread RPT rows -> 32 x 32 pixel tile
constexpr int TILE_H = TY * RPT;
constexpr int MAX_DIV = 8;

struct GenParams {
    const void* dem;
    long long rows_buf, cols, ld, row_begin, row_end, out_ld;
    void* out[5];  // 0 TPI, 1 TRI, 2 roughness, 3 unused, 4 fractal roughness
    unsigned mask;
    int w, tri_wilson;
    int n_div;                 // fractal: divisors q of w//2 (window.py:342-349)
    int div_q[MAX_DIV];
    double log_q[MAX_DIV];     // np.log(q)
    double mean_x, ss_xx;      // regression constants (window.py:363-372)
};

template <typename T> struct N_;
template <> struct N_<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
};
template <> struct N_<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double nan() { return CUDART_NAN; }
};

template <typename T>
__global__ void __launch_bounds__(TX * TY)
window_generic_kernel(const __grid_constant__ GenParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int w = p.w, h = w / 2;
    const int bw = TX + 2 * h, bh = TILE_H + 2 * h;
    const long long x_tile = (long long)blockIdx.x * TX;
    const long long y_tile = p.row_begin + (long long)blockIdx.y * TILE_H;
    const T* dem = reinterpret_cast<const T*>(p.dem);
    const int tid = threadIdx.y * TX + threadIdx.x;
    for (int idx = tid; idx < bw * bh; idx += TX * TY) {
        const int by = idx / bw, bx = idx - by * bw;
        const long long gy = y_tile - h + by, gx = x_tile - h + bx;
        T v = N_<T>::nan();
        if (gy >= 0 && gy < p.rows_buf && gx >= 0 && gx < p.cols) v = dem[gy * p.ld + gx];
        tile[idx] = v;
    }
    __syncthreads();
    const long long x = x_tile + threadIdx.x;
    if (x >= p.cols) return;
    const T nm1 = (T)(w * w - 1);
    for (int rr = 0; rr < RPT; ++rr) {
        const int ly = threadIdx.y * RPT + rr;
        const long long y = y_tile + ly;
        if (y >= p.row_end) break;
        const T* win = tile + (size_t)ly * bw + threadIdx.x;  // top-left of this pixel's window
        const T c = win[(size_t)h * bw + h];
        const long long off = (y - p.row_begin) * p.out_ld + x;
        // window sum (row-major, sequential) -- also the NaN / inf carrier
        T s = T(0);
        for (int j = 0; j < w; ++j)
            for (int k = 0; k < w; ++k) s = N_<T>::add(s, win[(size_t)j * bw + k]);
        const T carr = N_<T>::mul(s, T(0));
        if (p.mask & 1u) {  // TPI (window.py:216-220)
            const T v = N_<T>::sub(c, N_<T>::div(N_<T>::sub(s, c), nm1));
            reinterpret_cast<T*>(p.out[0])[off] = N_<T>::add(v, carr);
        }
        if (p.mask & 2u) {  // TRI Riley (window.py:94-95) / Wilson (window.py:150-155)
            T acc = T(0);
            if (p.tri_wilson) {
                for (int j = 0; j < w; ++j)
                    for (int k = 0; k < w; ++k) acc = N_<T>::add(acc, fabs(N_<T>::sub(win[(size_t)j * bw + k], c)));
                acc = N_<T>::div(acc, nm1);
            } else {
                for (int j = 0; j < w; ++j)
                    for (int k = 0; k < w; ++k) {
                        const T d = N_<T>::sub(win[(size_t)j * bw + k], c);
                        acc = N_<T>::add(acc, N_<T>::mul(d, d));
                    }
                acc = N_<T>::sqrt(acc);
            }
            reinterpret_cast<T*>(p.out[1])[off] = N_<T>::add(acc, carr);
        }
        if (p.mask & 4u) {  // roughness (window.py:281-287)
            T mx = c, mn = c;
            for (int j = 0; j < w; ++j)
                for (int k = 0; k < w; ++k) {
                    mx = fmax(mx, win[(size_t)j * bw + k]);
                    mn = fmin(mn, win[(size_t)j * bw + k]);
                }
            reinterpret_cast<T*>(p.out[2])[off] = N_<T>::add(N_<T>::sub(mx, mn), carr);
        }
        if (p.mask & 16u) {
            // Fractal roughness (window.py:317-379): voxel heights V = clip(z - centre, 0, w); for every divisor q of
            // w//2, Ns(q) = sum over the (w-1)//q x (w-1)//q boxes of max(V) / q; D = -slope of log Ns vs log q.
            const T wT = (T)w;
            double sy = 0.0, sxy = 0.0;
            // NaN rule: only the top-left (w-1) x (w-1) cells are read by the box counting (window.py:356-358, 431)
            bool has_nan = false;
            for (int j = 0; j < w - 1; ++j)
                for (int k = 0; k < w - 1; ++k) has_nan |= isnan(win[(size_t)j * bw + k]);
            for (int l = 0; l < p.n_div; ++l) {
                const int q = p.div_q[l];
                const int nq = (w - 1) / q;
                T sum_ns = T(0);
                for (int j = 0; j < nq; ++j)
                    for (int k = 0; k < nq; ++k) {
                        T mxv = T(0);  // V >= 0
                        for (int a = 0; a < q; ++a)
                            for (int b = 0; b < q; ++b) {
                                T v = N_<T>::sub(win[(size_t)(j * q + a) * bw + (k * q + b)], c);
                                v = fmin(fmax(v, T(0)), wT);
                                mxv = fmax(mxv, v);
                            }
                        sum_ns = N_<T>::add(sum_ns, mxv);
                    }
                const double ns = (double)N_<T>::div(sum_ns, (T)q);
                const double ly_ = log(ns);
                sy += ly_;
                sxy += ly_ * p.log_q[l];
            }
            const double n = (double)p.n_div;
            const double my = sy / n;
            const double b1 = (sxy - n * my * p.mean_x) / p.ss_xx;
            reinterpret_cast<T*>(p.out[4])[off] = has_nan ? N_<T>::nan() : (T)(-b1);
        }
    }
}

}  // namespace xbw

extern "C" {
#pragma GCC visibility push(default)

int xb_windowed_generic(const void* dem_dev, int dtype, int64_t rows_buf, int64_t cols, int64_t ld, int64_t row_begin,
                        int64_t row_end, int window_size, uint32_t win_mask, int tri_method_id,
                        void* const* out_planes_host, int64_t out_ld, void* stream) {
    if (!dem_dev || !out_planes_host || rows_buf <= 0 || cols <= 0 || ld < cols || row_begin < 0 ||
        row_end > rows_buf || row_begin > row_end || out_ld < cols) {
        xb_set_error("bad geometry in xb_windowed_generic");
        return XB_ERR_INVALID;
    }
    if (dtype != XB_F32 && dtype != XB_F64) {
        xb_set_error("dtype must be XB_F32 or XB_F64");
        return XB_ERR_INVALID;
    }
    if (window_size < 3 || window_size > 31 || !(window_size & 1)) {
        xb_set_error("window_size must be odd and in [3, 31], got %d", window_size);
        return XB_ERR_UNSUPPORTED;
    }
    if (!(win_mask & 0x17u) || (win_mask & ~0x17u)) {
        xb_set_error("win_mask must select TPI (1), TRI (2), roughness (4) and/or fractal roughness (16)");
        return XB_ERR_INVALID;
    }
    xbw::GenParams p;
    memset(&p, 0, sizeof(p));
    p.dem = dem_dev, p.rows_buf = rows_buf, p.cols = cols, p.ld = ld, p.row_begin = row_begin, p.row_end = row_end;
    p.out_ld = out_ld, p.mask = win_mask, p.w = window_size, p.tri_wilson = tri_method_id ? 1 : 0;
    for (int i = 0; i < 5; ++i) {
        p.out[i] = out_planes_host[i];
        if (((win_mask >> i) & 1u) && !p.out[i]) {
            xb_set_error("plane %d requested but NULL", i);
            return XB_ERR_INVALID;
        }
    }
    if (win_mask & 16u) {
        const int hw = window_size / 2;
        double sx = 0, sxx = 0;
        for (int q = 1; q <= hw && p.n_div < xbw::MAX_DIV; ++q)
            if (hw % q == 0) {
                p.div_q[p.n_div] = q;
                p.log_q[p.n_div] = log((double)q);
                sx += p.log_q[p.n_div];
                sxx += p.log_q[p.n_div] * p.log_q[p.n_div];
                ++p.n_div;
            }
        p.mean_x = sx / p.n_div;
        p.ss_xx = sxx - p.n_div * p.mean_x * p.mean_x;
    }
    if (row_end == row_begin) return XB_OK;
    const int h = window_size / 2;
    const size_t es = dtype == XB_F64 ? 8 : 4;
    const size_t smem = (size_t)(xbw::TX + 2 * h) * (xbw::TILE_H + 2 * h) * es;
    dim3 grid((unsigned)((cols + xbw::TX - 1) / xbw::TX), (unsigned)((row_end - row_begin + xbw::TILE_H - 1) / xbw::TILE_H));
    dim3 block(xbw::TX, xbw::TY);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == XB_F64) {
        XB_CUDA_CHECK(cudaFuncSetAttribute(xbw::window_generic_kernel<double>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xbw::window_generic_kernel<double><<<grid, block, smem, st>>>(p);
    } else {
        XB_CUDA_CHECK(cudaFuncSetAttribute(xbw::window_generic_kernel<float>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xbw::window_generic_kernel<float><<<grid, block, smem, st>>>(p);
    }
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
