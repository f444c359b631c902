// xdem_b200 -- N-D binned robust statistics (count / median / NMAD per bin) for sm_100a.
//
// Device side of `xdem.spatialstats.nd_binning` (spatialstats.py:91-216), i.e. of the three SciPy calls it makes per
// statistic (`binned_statistic`, `binned_statistic_2d`, `binned_statistic_dd`): every sample gets the flattened number
// of its bin (np.digitize against each variable's edges, right-most edge closed, out-of-range samples dropped --
// scipy/stats/_binned_statistic.py:_bin_numbers), then per-bin exact medians by MSD radix select on order-preserving
// float32 keys (np.nanmedian; 4 passes of 8 bits) and, for the NMAD (geoutils.stats.nmad = 1.4826 * median |x - median|),
// a second select on the absolute deviations from the bin's own median.
//   xb_bin_keys         : (values, <= 3 variables, edges) -> 32-bit key + 16-bit bin number per sample
//   xb_bin_hist         : one digit histogram of the keys whose higher digits match their bin's prefix
//   xb_bin_next         : per bin, the smallest key strictly above a selected key (upper median of even counts)
//   xb_bin_absdev_keys  : keys of |value - median(bin)| (float32 arithmetic, like NumPy on float32 data)
// Histograms live in shared memory when n_bins x 256 counters fit (<= 192 bins), else in global memory.
#include "../../include/xdem_b200.h"

#include <math_constants.h>

#include "xb_common.cuh"

void xb_count_launch(int n);

namespace xbb {

constexpr int NT = 256;
constexpr int MAX_DIMS = 3;
constexpr unsigned short NO_BIN = 0xffffu;

struct BinParams {
    const float* values;
    const float* var[MAX_DIMS];
    const double* edges;           // concatenated edge arrays (float64 copies of the sample-dtype edges)
    int edge_off[MAX_DIMS];        // offset of dimension d's edges
    int n_edges[MAX_DIMS];         // number of edges (bins + 1)
    double p10[MAX_DIMS];          // 10^decimal of scipy's "on the last edge" rounding rule
    int n_dims;
    long long n;
};

__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// np.digitize(x, edges) - 1 with scipy's closed last edge; -1 = outside
__device__ __forceinline__ int bin_of(double x, const double* __restrict__ e, int n_edges, double p10) {
    if (!(x >= e[0])) return -1;
    const double last = e[n_edges - 1];
    if (x >= last) {
        // _bin_numbers: samples >= the last edge that round to it belong to the last bin
        return (rint(x * p10) == rint(last * p10)) ? n_edges - 2 : -1;
    }
    int lo = 0, hi = n_edges - 1;  // e[lo] <= x < e[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= e[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(NT)
bin_keys_kernel(const __grid_constant__ BinParams p, unsigned* __restrict__ key, unsigned short* __restrict__ bin) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < p.n; i += stride) {
        const float v = p.values[i];
        bool ok = isfinite(v);
        int flat = 0;
#pragma unroll
        for (int d = 0; d < MAX_DIMS; ++d) {
            if (d >= p.n_dims) break;
            const float x = p.var[d][i];
            ok = ok && isfinite(x);  // spatialstats.py:128-131: rows with a non-finite value or variable are dropped
            const int b = ok ? bin_of((double)x, p.edges + p.edge_off[d], p.n_edges[d], p.p10[d]) : -1;
            ok = ok && b >= 0;
            flat = flat * (p.n_edges[d] - 1) + (b < 0 ? 0 : b);  // C order, like stats.flatten() (spatialstats.py:173, 193)
        }
        key[i] = ok ? ordered_key(v) : 0u;
        bin[i] = ok ? (unsigned short)flat : NO_BIN;
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(1024)
bin_hist_kernel(const unsigned* __restrict__ key, const unsigned short* __restrict__ bin, long long n, int n_bins,
                const unsigned* __restrict__ prefix, unsigned prefix_mask, int shift,
                unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned sh_hist[];
    const int n_cnt = n_bins * 256;
    if (SMEM) {
        for (int k = threadIdx.x; k < n_cnt; k += blockDim.x) sh_hist[k] = 0u;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        unsigned g[4], k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool in = i0 + u * stride < n;
            g[u] = in ? bin[i0 + u * stride] : (unsigned)NO_BIN;
            k[u] = in ? key[i0 + u * stride] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g[u] >= (unsigned)n_bins) continue;
            if ((k[u] & prefix_mask) != prefix[g[u]]) continue;
            const unsigned slot = g[u] * 256u + ((k[u] >> shift) & 255u);
            if (SMEM) atomicAdd(&sh_hist[slot], 1u);
            else atomicAdd(&hist[slot], 1ull);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_cnt; k += blockDim.x)
            if (sh_hist[k]) atomicAdd(&hist[k], (unsigned long long)sh_hist[k]);
    }
}

__global__ void __launch_bounds__(NT)
bin_next_kernel(const unsigned* __restrict__ key, const unsigned short* __restrict__ bin, long long n, int n_bins,
                const unsigned* __restrict__ sel, unsigned* __restrict__ next_key) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
        const unsigned g = bin[i];
        if (g >= (unsigned)n_bins) continue;
        const unsigned k = key[i];
        if (k > sel[g] && k < next_key[g]) atomicMin(&next_key[g], k);
    }
}

__global__ void __launch_bounds__(NT)
bin_absdev_kernel(const float* __restrict__ values, const unsigned short* __restrict__ bin, long long n, int n_bins,
                  const float* __restrict__ center, unsigned* __restrict__ key) {
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
        const unsigned g = bin[i];
        if (g >= (unsigned)n_bins) continue;
        key[i] = ordered_key(fabsf(__fsub_rn(values[i], center[g])));  // np.abs(data - np.nanmedian(data)) in float32
    }
}

static int grid_1d(long long n, int threads, int per_sm) {
    int sms = 0;
    if (xb_num_sms(&sms)) sms = 148;
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)sms * per_sm;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------------------------------------
// Applying a 1-D binned correction (BiasCorr._apply_rst, biascorr.py:259-310): out = elev + corr(var).
//   mode 0 "linear": interp_nd_binning (spatialstats.py:237-423) in one dimension = RegularGridInterpolator(linear) over
//     the mid-points of the valid bins, extended by one duplicated point on each side so that values outside are held
//     constant: corr = v[i] (1 - t) + v[i+1] t between mids i and i+1, v[0] below the first, v[m-1] above the last.
//   mode 1 "per_bin": get_perbin_nd_binning (spatialstats.py:425-530): the statistic of the bin [left, right) that holds
//     var, NaN outside every bin (x = the m + 1 edges, v = the m statistics).
// float64 arithmetic like the reference, result cast to float32 (base.py:491).  HBM-bound: 12 B per pixel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bin_apply_1d_kernel(const float* __restrict__ elev, const float* __restrict__ var, long long n,
                    const double* __restrict__ x, const double* __restrict__ v, int m, int mode, float* __restrict__ out) {
    extern __shared__ double sh_tab[];
    double* sx = sh_tab;
    double* sv = sh_tab + (mode == 1 ? m + 1 : m);
    for (int k = threadIdx.x; k < (mode == 1 ? m + 1 : m); k += blockDim.x) sx[k] = x[k];
    for (int k = threadIdx.x; k < m; k += blockDim.x) sv[k] = v[k];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = (double)var[i];
        double corr = CUDART_NAN;
        if (a == a) {
            if (mode == 0) {
                if (a <= sx[0]) {
                    corr = sv[0];
                } else if (a >= sx[m - 1]) {
                    corr = sv[m - 1];
                } else {
                    int lo = 0, hi = m - 1;  // sx[lo] <= a < sx[hi]
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (a >= sx[mid]) lo = mid; else hi = mid;
                    }
                    const double t = (a - sx[lo]) / (sx[hi] - sx[lo]);
                    corr = sv[lo] * (1.0 - t) + sv[hi] * t;
                }
            } else if (a >= sx[0] && a < sx[m]) {
                int lo = 0, hi = m;  // sx[lo] <= a < sx[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (a >= sx[mid]) lo = mid; else hi = mid;
                }
                corr = sv[lo];
            }
        }
        out[i] = (float)((double)elev[i] + corr);
    }
}


// Per-bin sum, sum of squares, minimum and maximum (the other built-in statistics of scipy.stats.binned_statistic:
// 'mean', 'std', 'sum', 'min', 'max').  Lanes of a warp that hit the same bin are merged first (__match_any_sync: a few
// hundred bins and unsorted samples mean several lanes per bin), then one float64 atomic per distinct bin and warp.
__global__ void __launch_bounds__(NT)
bin_moments_kernel(const float* __restrict__ val, const unsigned short* __restrict__ bin, long long n, int n_bins,
                   double* __restrict__ sum, double* __restrict__ sumsq, unsigned* __restrict__ minkey,
                   unsigned* __restrict__ maxkey) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * NT;
    const long long n_round = (n + stride - 1) / stride * stride;  // uniform trip count: the body holds warp intrinsics
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n_round; i += stride) {
        unsigned b = NO_BIN;
        float v = 0.f;
        if (i < n) b = bin[i], v = val[i];
        const bool ok = b != NO_BIN && b < (unsigned)n_bins;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? b : 0xffffffffu);
        if (!__any_sync(0xffffffffu, ok)) continue;
        // every lane reduces over its peer group; the lowest lane of the group publishes
        double s = 0.0, s2 = 0.0;
        unsigned kmin = 0xffffffffu, kmax = 0u;
        const unsigned key = __float_as_uint(v) & 0x80000000u ? ~__float_as_uint(v) : (__float_as_uint(v) | 0x80000000u);
        unsigned rest = peers;
        while (rest) {  // peer groups differ between lanes: iterate over the union of set bits
            const int src = __ffs(rest) - 1;
            rest &= rest - 1;
            const float pv = __shfl_sync(peers, v, src);
            const unsigned pk = __shfl_sync(peers, key, src);
            s += (double)pv, s2 += (double)pv * (double)pv;
            kmin = min(kmin, pk), kmax = max(kmax, pk);
        }
        if (ok && lane == __ffs(peers) - 1) {
            atomicAdd(&sum[b], s);
            atomicAdd(&sumsq[b], s2);
            atomicMin(&minkey[b], kmin);
            atomicMax(&maxkey[b], kmax);
        }
    }
}

}  // namespace xbb

extern "C" {
#pragma GCC visibility push(default)

int xb_bin_keys(const float* values_dev, const float* const* vars_dev_host, int n_dims, int64_t n,
                const double* edges_dev, const int32_t* n_edges_host, const double* p10_host, uint32_t* key_dev,
                uint16_t* bin_dev, void* stream) {
    if (!values_dev || !vars_dev_host || !edges_dev || !n_edges_host || !p10_host || !key_dev || !bin_dev || n <= 0 ||
        n_dims < 1 || n_dims > xbb::MAX_DIMS) {
        xb_set_error("bad arguments to xb_bin_keys (1..3 variables)");
        return XB_ERR_INVALID;
    }
    xbb::BinParams p{};
    p.values = values_dev;
    p.edges = edges_dev;
    p.n_dims = n_dims;
    p.n = n;
    long long total = 1;
    int off = 0;
    for (int d = 0; d < n_dims; ++d) {
        if (!vars_dev_host[d] || n_edges_host[d] < 2) {
            xb_set_error("xb_bin_keys: variable %d needs a pointer and at least two edges", d);
            return XB_ERR_INVALID;
        }
        p.var[d] = vars_dev_host[d];
        p.edge_off[d] = off;
        p.n_edges[d] = n_edges_host[d];
        p.p10[d] = p10_host[d];
        off += n_edges_host[d];
        total *= (n_edges_host[d] - 1);
    }
    if (total >= 0xffff) {
        xb_set_error("xb_bin_keys: %lld bins exceed the 16-bit bin numbers", total);
        return XB_ERR_UNSUPPORTED;
    }
    xbb::bin_keys_kernel<<<xbb::grid_1d(n, xbb::NT, 16), xbb::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        p, key_dev, bin_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_bin_hist(const uint32_t* key_dev, const uint16_t* bin_dev, int64_t n, int n_bins, const uint32_t* prefix_dev,
                uint32_t prefix_mask, int shift, unsigned long long* hist_dev, void* stream) {
    if (!key_dev || !bin_dev || !prefix_dev || !hist_dev || n <= 0 || n_bins < 1 || n_bins >= 0xffff || shift < 0 ||
        shift > 24) {
        xb_set_error("bad arguments to xb_bin_hist");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)n_bins * 256 * sizeof(unsigned);
    if (smem <= 192 * 1024) {
        auto kern = xbb::bin_hist_kernel<true>;
        XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = smem > 100 * 1024 ? 1 : 2;
        kern<<<xbb::grid_1d(n, 1024, per_sm), 1024, smem, st>>>(key_dev, bin_dev, n, n_bins, prefix_dev, prefix_mask,
                                                                 shift, hist_dev);
    } else {
        xbb::bin_hist_kernel<false><<<xbb::grid_1d(n, 1024, 2), 1024, 0, st>>>(key_dev, bin_dev, n, n_bins, prefix_dev,
                                                                                prefix_mask, shift, hist_dev);
    }
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_bin_next(const uint32_t* key_dev, const uint16_t* bin_dev, int64_t n, int n_bins, const uint32_t* sel_dev,
                uint32_t* next_key_dev, void* stream) {
    if (!key_dev || !bin_dev || !sel_dev || !next_key_dev || n <= 0 || n_bins < 1 || n_bins >= 0xffff) {
        xb_set_error("bad arguments to xb_bin_next");
        return XB_ERR_INVALID;
    }
    xbb::bin_next_kernel<<<xbb::grid_1d(n, xbb::NT, 16), xbb::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        key_dev, bin_dev, n, n_bins, sel_dev, next_key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_bin_absdev_keys(const float* values_dev, const uint16_t* bin_dev, int64_t n, int n_bins,
                       const float* center_dev, uint32_t* key_dev, void* stream) {
    if (!values_dev || !bin_dev || !center_dev || !key_dev || n <= 0 || n_bins < 1 || n_bins >= 0xffff) {
        xb_set_error("bad arguments to xb_bin_absdev_keys");
        return XB_ERR_INVALID;
    }
    xbb::bin_absdev_kernel<<<xbb::grid_1d(n, xbb::NT, 16), xbb::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        values_dev, bin_dev, n, n_bins, center_dev, key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_bin_apply_1d(const float* elev_dev, const float* var_dev, int64_t n, const double* x_dev, const double* v_dev,
                    int m, int mode, float* out_dev, void* stream) {
    if (!elev_dev || !var_dev || !x_dev || !v_dev || !out_dev || n <= 0 || m < 1 || m > 4000 || (mode != 0 && mode != 1)) {
        xb_set_error("bad arguments to xb_bin_apply_1d (1 <= m <= 4000 table entries, mode 0 linear / 1 per_bin)");
        return XB_ERR_INVALID;
    }
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    const long long need = (n + 255) / 256;
    const int grid = (int)(need < (long long)sms * 8 ? need : (long long)sms * 8);
    xbb::bin_apply_1d_kernel<<<grid, 256, (size_t)(2 * m + 1) * sizeof(double), reinterpret_cast<cudaStream_t>(stream)>>>(
        elev_dev, var_dev, n, x_dev, v_dev, m, mode, out_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_bin_moments(const float* values_dev, const uint16_t* bin_dev, int64_t n, int n_bins, double* sum_dev,
                   double* sumsq_dev, uint32_t* minkey_dev, uint32_t* maxkey_dev, void* stream) {
    if (!values_dev || !bin_dev || !sum_dev || !sumsq_dev || !minkey_dev || !maxkey_dev || n <= 0 || n_bins < 1 ||
        n_bins >= 0xFFFF) {
        xb_set_error("bad arguments to xb_bin_moments");
        return XB_ERR_INVALID;
    }
    xbb::bin_moments_kernel<<<xbb::grid_1d(n, xbb::NT, 8), xbb::NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        values_dev, bin_dev, n, n_bins, sum_dev, sumsq_dev, minkey_dev, maxkey_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
