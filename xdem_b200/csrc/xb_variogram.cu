// xdem_b200 -- all-pairs empirical-variogram lag binning (K2) for sm_100a.
//
// Replaces the O(N^2) pairwise work that scikit-gstat's Variogram does for
// xdem.spatialstats._get_pdist_empirical_variogram (spatialstats.py:1064-1101): for every sample pair i<j the
// Euclidean distance, |v_i - v_j|, the lag class (right-edge bins, pairs >= maxlag dropped), per-class pair count and
// sum of squared differences (Matheron numerator).
//
// Design (DESIGN.md "K2"): samples live on the raster grid, so distances are exact integers d2 = dx^2 + dy^2 in pixel
// units and the float64 bin edges become integer thresholds edge2[k] on the host -- counts are bit-exact by
// construction.  Samples are pre-sorted along a Morton curve and cut into groups of 128 with bounding boxes.  A warp owns
// one (i-group, j-group) tile at a time: the i-group sits in registers (4 samples per lane), the j-group is staged in
// shared memory and broadcast one sample per step (one LDS.128 feeds 4 pairs per lane = 128 pairs per warp step).
// The bounding boxes bound the distance range of the tile, i.e. the <=3 consecutive lag classes it can touch, so every
// lane accumulates into 3 register slots selected by two integer compares -- no shared-memory or global atomics in the
// pair loop (B300_MICROARCH: ATOMS costs ~2 cycles per lane, which would cap this kernel at ~6e10 pairs/s).  Tiles
// spanning more classes are re-swept 3 classes at a time.  Slots are warp-reduced (REDUX for counts, fp64 shuffles for
// sums) and flushed with one global atomic per class per tile.
#include "../../include/xdem_b200.h"

#include <algorithm>

#include "xb_common.cuh"

void xb_count_launch(int n);

namespace xbv {

constexpr int GS = 128;      // samples per group
constexpr int IPL = GS / 32; // i-samples per lane
constexpr int NWARPS = 8;
constexpr int NTHREADS = NWARPS * 32;
constexpr int CHUNK_J = 32;  // j-groups per work unit

__device__ __forceinline__ void box_dist2(const int4& a, const int4& b, unsigned long long& dmin2,
                                          unsigned long long& dmax2) {
    const long long dxmin = max(0, max(a.x - b.z, b.x - a.z));
    const long long dymin = max(0, max(a.y - b.w, b.y - a.w));
    const long long dxmax = max(a.z - b.x, b.z - a.x);
    const long long dymax = max(a.w - b.y, b.w - a.y);
    dmin2 = (unsigned long long)(dxmin * dxmin + dymin * dymin);
    dmax2 = (unsigned long long)(dxmax * dxmax + dymax * dymax);
}

// first k in [0,n) with key < edge2[k]  (n if none); warp-cooperative, identical result in every lane
__device__ __forceinline__ int first_bin_above(const unsigned long long* __restrict__ edge2, int n,
                                               unsigned long long key, int lane) {
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool le = (k < n) && (edge2[k] <= key);
        cnt += __popc(__ballot_sync(0xffffffffu, le));
    }
    return cnt;
}

template <typename D2>
__device__ __forceinline__ D2 dist2(int dx, int dy);
template <>
__device__ __forceinline__ unsigned dist2<unsigned>(int dx, int dy) {
    return (unsigned)(dx * dx) + (unsigned)(dy * dy);
}
template <>
__device__ __forceinline__ unsigned long long dist2<unsigned long long>(int dx, int dy) {
    return (unsigned long long)((long long)dx * dx) + (unsigned long long)((long long)dy * dy);
}

template <typename D2>
__device__ __forceinline__ D2 clamp_edge(unsigned long long e);
template <>
__device__ __forceinline__ unsigned clamp_edge<unsigned>(unsigned long long e) {
    return e > 0xffffffffull ? 0xffffffffu : (unsigned)e;
}
template <>
__device__ __forceinline__ unsigned long long clamp_edge<unsigned long long>(unsigned long long e) {
    return e;
}

// One sweep of a 128 x 128 tile over the 3 lag classes [b, b+3).
template <typename D2, bool CHECK, int EST>
__device__ __forceinline__ void sweep_tile(const int4* __restrict__ sj, const int (&xi)[IPL], const int (&yi)[IPL],
                                           const float (&vi)[IPL], const int (&ii)[IPL], bool diag, D2 L, D2 Ta, D2 Tb,
                                           D2 Tc, unsigned (&cnt)[3], float (&sum)[3]) {
#pragma unroll 4
    for (int jj = 0; jj < GS; ++jj) {
        const int4 pj = sj[jj];  // broadcast LDS.128: x, y, value bits, sorted index (-1 = padding)
        const float vj = __int_as_float(pj.z);
#pragma unroll
        for (int m = 0; m < IPL; ++m) {
            const D2 d2 = dist2<D2>(pj.x - xi[m], pj.y - yi[m]);
            const float df = vj - vi[m];
            // Matheron: sum (v_i - v_j)^2;  Cressie-Hawkins: sum |v_i - v_j|^(1/2)   (skgstat.estimators)
            const float q = EST == 0 ? df * df : sqrtf(fabsf(df));
            bool ok = d2 >= L;
            if (CHECK) ok = ok && (pj.w >= 0) && (ii[m] >= 0) && (!diag || ii[m] < pj.w);
            const bool p0 = ok && (d2 < Ta);
            const bool p1 = ok && (d2 >= Ta) && (d2 < Tb);
            const bool p2 = ok && (d2 >= Tb) && (d2 < Tc);
            cnt[0] += p0 ? 1u : 0u;
            cnt[1] += p1 ? 1u : 0u;
            cnt[2] += p2 ? 1u : 0u;
            sum[0] += p0 ? q : 0.0f;
            sum[1] += p1 ? q : 0.0f;
            sum[2] += p2 ? q : 0.0f;
        }
    }
}

// Interior tile (no padding, not on the diagonal) whose bounding boxes put EVERY pair inside the NCLS <= 3 consecutive
// classes [b, b + NCLS): no lower-bound / validity tests, NCLS - 1 thresholds, exclusive float sums, cumulative integer
// counts (the last class follows from the tile's 128 x 128 pairs).  12-19 instructions per pair instead of 24
// (cuobjdump), and a single-class tile needs no distances at all.
template <typename D2, int NCLS, int EST>
__device__ __forceinline__ void sweep_full(const int4* __restrict__ sj, const int (&xi)[IPL], const int (&yi)[IPL],
                                           const float (&vi)[IPL], D2 Ta, D2 Tb, unsigned (&cum)[2], float (&sum)[3]) {
#pragma unroll 4
    for (int jj = 0; jj < GS; ++jj) {
        const int4 pj = sj[jj];
        const float vj = __int_as_float(pj.z);
#pragma unroll
        for (int m = 0; m < IPL; ++m) {
            const float df = vj - vi[m];
            const float q = EST == 0 ? df * df : sqrtf(fabsf(df));
            if (NCLS == 1) {
                sum[0] += q;
            } else {
                const D2 d2 = dist2<D2>(pj.x - xi[m], pj.y - yi[m]);
                const bool c0 = d2 < Ta;
                cum[0] += c0 ? 1u : 0u;
                sum[0] += c0 ? q : 0.0f;
                if (NCLS == 2) {
                    sum[1] += c0 ? 0.0f : q;
                } else {
                    const bool c1 = d2 < Tb;
                    cum[1] += c1 ? 1u : 0u;
                    sum[1] += (c1 && !c0) ? q : 0.0f;
                    sum[2] += c1 ? 0.0f : q;
                }
            }
        }
    }
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename D2, int EST>
__global__ void __launch_bounds__(NTHREADS, 2)
variogram_pairs_kernel(const int4* __restrict__ pts, const int4* __restrict__ gbox, int G,
                       const unsigned long long* __restrict__ edge2, int n_bins,
                       const long long* __restrict__ unit_prefix, long long unit_begin, long long unit_end,
                       unsigned long long* __restrict__ work_counter, unsigned long long* __restrict__ count,
                       double* __restrict__ sumsq, int full_mask) {
    __shared__ int4 sj_all[NWARPS][GS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4* sj = sj_all[warp];
    const unsigned long long edge_last = edge2[n_bins - 1];

    for (;;) {
        long long u = 0;
        if (lane == 0) u = unit_begin + (long long)atomicAdd(work_counter, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= unit_end) break;
        // unit -> (i-group, chunk): unit_prefix[i] = number of units of rows < i (ascending, G+1 entries)
        int lo = 0, hi = G;  // largest i with unit_prefix[i] <= u
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (unit_prefix[mid] <= u) lo = mid; else hi = mid;
        }
        const int gi = lo;
        const int chunk = (int)(u - unit_prefix[gi]);
        const int j_begin = gi + chunk * CHUNK_J;
        const int j_end = min(G, j_begin + CHUNK_J);

        int xi[IPL], yi[IPL], ii[IPL];
        float vi[IPL];
#pragma unroll
        for (int m = 0; m < IPL; ++m) {
            const int4 p = pts[(long long)gi * GS + lane + 32 * m];
            xi[m] = p.x, yi[m] = p.y, vi[m] = __int_as_float(p.z), ii[m] = p.w;
        }
        const int4 ibox = gbox[gi];

        for (int gj = j_begin; gj < j_end; ++gj) {
            const int4 jbox = gbox[gj];
            unsigned long long dmin2, dmax2;
            box_dist2(ibox, jbox, dmin2, dmax2);
            if (dmin2 >= edge_last) continue;  // the whole tile lies beyond maxlag
            const int b_lo = first_bin_above(edge2, n_bins, dmin2, lane);
            int b_hi = first_bin_above(edge2, n_bins, dmax2, lane);
            if (b_hi > n_bins - 1) b_hi = n_bins - 1;
            __syncwarp();
#pragma unroll
            for (int m = 0; m < IPL; ++m) sj[lane + 32 * m] = pts[(long long)gj * GS + lane + 32 * m];
            __syncwarp();
            const bool diag = (gj == gi);
            const bool check = diag || gi == G - 1 || gj == G - 1;
            if (!check && dmax2 < edge_last && b_hi - b_lo <= 2 && ((full_mask >> (b_hi - b_lo)) & 1)) {
                // every one of the 128 x 128 pairs is valid and lies in [b_lo, b_hi]
                const int ncls = b_hi - b_lo + 1;
                const D2 Ta = clamp_edge<D2>(edge2[b_lo]);
                const D2 Tb = clamp_edge<D2>(edge2[min(b_lo + 1, n_bins - 1)]);
                unsigned cum[2] = {0u, 0u};
                float sum[3] = {0.f, 0.f, 0.f};
                if (ncls == 1)
                    sweep_full<D2, 1, EST>(sj, xi, yi, vi, Ta, Tb, cum, sum);
                else if (ncls == 2)
                    sweep_full<D2, 2, EST>(sj, xi, yi, vi, Ta, Tb, cum, sum);
                else
                    sweep_full<D2, 3, EST>(sj, xi, yi, vi, Ta, Tb, cum, sum);
                const unsigned total = (unsigned)(GS * GS);
                const unsigned k0 = ncls == 1 ? total : __reduce_add_sync(0xffffffffu, cum[0]);
                const unsigned k1 = ncls == 3 ? __reduce_add_sync(0xffffffffu, cum[1]) : total;
                const unsigned c3[3] = {k0, k1 - k0, total - k1};
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    if (s >= ncls || c3[s] == 0u) continue;  // warp-uniform
                    const double ssum = warp_sum_f64((double)sum[s]);
                    if (lane == 0) {
                        atomicAdd(&count[b_lo + s], (unsigned long long)c3[s]);
                        atomicAdd(&sumsq[b_lo + s], ssum);
                    }
                }
                continue;
            }
            for (int b = b_lo; b <= b_hi; b += 3) {
                const D2 L = (b == b_lo) ? (D2)0 : clamp_edge<D2>(edge2[b - 1]);  // first sweep: d2 >= dmin2 >= edge2[b_lo-1]
                const D2 Ta = clamp_edge<D2>(edge2[b]);
                const D2 Tb = (b + 1 < n_bins) ? clamp_edge<D2>(edge2[b + 1]) : Ta;
                const D2 Tc = (b + 2 < n_bins) ? clamp_edge<D2>(edge2[b + 2]) : Tb;
                unsigned cnt[3] = {0u, 0u, 0u};
                float sum[3] = {0.f, 0.f, 0.f};
                if (check)
                    sweep_tile<D2, true, EST>(sj, xi, yi, vi, ii, diag, L, Ta, Tb, Tc, cnt, sum);
                else
                    sweep_tile<D2, false, EST>(sj, xi, yi, vi, ii, diag, L, Ta, Tb, Tc, cnt, sum);
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const unsigned c = __reduce_add_sync(0xffffffffu, cnt[s]);
                    if (c == 0u) continue;  // warp-uniform
                    const double ssum = warp_sum_f64((double)sum[s]);
                    if (lane == 0 && b + s < n_bins) {
                        atomicAdd(&count[b + s], (unsigned long long)c);
                        atomicAdd(&sumsq[b + s], ssum);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dowd's estimator needs the per-class MEDIAN of |v_i - v_j| (skgstat.estimators.dowd): exact selection by MSD radix
// select on the float32 bit pattern of |diff| (non-negative floats order like their bits), 4 passes of 8 bits.  One
// pass = one sweep over all pairs; the class of a pair comes from the same tile thresholds as above, the digit
// histogram (n_bins x 256 counters) is shared by the CTA in shared memory and flushed once.  MODE 1 finds, per class,
// the smallest key strictly greater than sel[class] (upper median of even-sized classes).
// ---------------------------------------------------------------------------------------------------------------
template <typename D2, int MODE>
__global__ void __launch_bounds__(NTHREADS, 2)
variogram_median_kernel(const int4* __restrict__ pts, const int4* __restrict__ gbox, int G,
                        const unsigned long long* __restrict__ edge2, int n_bins,
                        const long long* __restrict__ unit_prefix, long long unit_begin, long long unit_end,
                        unsigned long long* __restrict__ work_counter, const unsigned* __restrict__ prefix,
                        unsigned prefix_mask, int shift, unsigned long long* __restrict__ hist,
                        unsigned* __restrict__ next_key) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    int4* sj_all = reinterpret_cast<int4*>(dyn_smem);
    unsigned* sh = reinterpret_cast<unsigned*>(dyn_smem + (size_t)NWARPS * GS * sizeof(int4));
    const int n_cnt = MODE == 0 ? n_bins * 256 : n_bins;
    for (int k = threadIdx.x; k < n_cnt; k += NTHREADS) sh[k] = MODE == 0 ? 0u : 0xffffffffu;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4* sj = sj_all + (size_t)warp * GS;
    const unsigned long long edge_last = edge2[n_bins - 1];
    for (;;) {
        long long u = 0;
        if (lane == 0) u = unit_begin + (long long)atomicAdd(work_counter, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= unit_end) break;
        int lo = 0, hi = G;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (unit_prefix[mid] <= u) lo = mid; else hi = mid;
        }
        const int gi = lo;
        const int j_begin = gi + (int)(u - unit_prefix[gi]) * CHUNK_J;
        const int j_end = min(G, j_begin + CHUNK_J);
        int xi[IPL], yi[IPL], ii[IPL];
        float vi[IPL];
#pragma unroll
        for (int m = 0; m < IPL; ++m) {
            const int4 p = pts[(long long)gi * GS + lane + 32 * m];
            xi[m] = p.x, yi[m] = p.y, vi[m] = __int_as_float(p.z), ii[m] = p.w;
        }
        const int4 ibox = gbox[gi];
        for (int gj = j_begin; gj < j_end; ++gj) {
            unsigned long long dmin2, dmax2;
            box_dist2(ibox, gbox[gj], dmin2, dmax2);
            if (dmin2 >= edge_last) continue;
            const int b_lo = first_bin_above(edge2, n_bins, dmin2, lane);
            int b_hi = first_bin_above(edge2, n_bins, dmax2, lane);
            if (b_hi > n_bins - 1) b_hi = n_bins - 1;
            __syncwarp();
#pragma unroll
            for (int m = 0; m < IPL; ++m) sj[lane + 32 * m] = pts[(long long)gj * GS + lane + 32 * m];
            __syncwarp();
            const bool diag = (gj == gi);
            for (int jj = 0; jj < GS; ++jj) {
                const int4 pj = sj[jj];
                const float vj = __int_as_float(pj.z);
#pragma unroll
                for (int m = 0; m < IPL; ++m) {
                    const unsigned long long d2 = dist2<unsigned long long>(pj.x - xi[m], pj.y - yi[m]);
                    if (pj.w < 0 || ii[m] < 0 || (diag && ii[m] >= pj.w) || d2 >= edge_last) continue;
                    int b = b_lo;  // the tile touches only classes b_lo..b_hi (usually <= 3)
                    while (b < b_hi && d2 >= edge2[b]) ++b;
                    const unsigned key = __float_as_uint(fabsf(vj - vi[m]));
                    if (MODE == 0) {
                        if ((key & prefix_mask) == prefix[b]) atomicAdd(&sh[b * 256 + ((key >> shift) & 255u)], 1u);
                    } else {
                        if (key > prefix[b] && key < sh[b]) atomicMin(&sh[b], key);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_cnt; k += NTHREADS) {
        if (MODE == 0) {
            if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
        } else {
            if (sh[k] != 0xffffffffu) atomicMin(&next_key[k], sh[k]);
        }
    }
}

// Largest squared pair distance; tiles whose bounding boxes cannot beat the current best are skipped.
__global__ void __launch_bounds__(NTHREADS, 2)
variogram_maxd2_kernel(const int4* __restrict__ pts, const int4* __restrict__ gbox, int G,
                       unsigned long long* __restrict__ work_counter, unsigned long long* __restrict__ best) {
    __shared__ int4 sj_all[NWARPS][GS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int4* sj = sj_all[warp];
    const long long n_units = (long long)G;  // one unit = one i-group against all j >= i
    for (;;) {
        long long u = 0;
        if (lane == 0) u = (long long)atomicAdd(work_counter, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const int gi = (int)u;
        const int4 ibox = gbox[gi];
        int xi[IPL], yi[IPL], ii[IPL];
#pragma unroll
        for (int m = 0; m < IPL; ++m) {
            const int4 p = pts[(long long)gi * GS + lane + 32 * m];
            xi[m] = p.x, yi[m] = p.y, ii[m] = p.w;
        }
        unsigned long long local = 0ull;
        for (int gj = gi; gj < G; ++gj) {
            unsigned long long dmin2, dmax2;
            box_dist2(ibox, gbox[gj], dmin2, dmax2);
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(best);
            cur = __shfl_sync(0xffffffffu, cur, 0);
            if (dmax2 <= cur || dmax2 <= local) continue;
            __syncwarp();
#pragma unroll
            for (int m = 0; m < IPL; ++m) sj[lane + 32 * m] = pts[(long long)gj * GS + lane + 32 * m];
            __syncwarp();
            for (int jj = 0; jj < GS; ++jj) {
                const int4 pj = sj[jj];
#pragma unroll
                for (int m = 0; m < IPL; ++m) {
                    const unsigned long long d2 = dist2<unsigned long long>(pj.x - xi[m], pj.y - yi[m]);
                    if (pj.w >= 0 && ii[m] >= 0 && d2 > local) local = d2;
                }
            }
            unsigned long long w = local;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, w, o);
                w = other > w ? other : w;
            }
            local = w;
            if (lane == 0) atomicMax(best, w);
        }
    }
}

// Work counter of one launch: allocated from the stream-ordered pool (cudaMallocAsync) so that concurrent calls on
// different streams / threads never share state.
struct WorkCounter {
    unsigned long long* ptr = nullptr;
    cudaStream_t st = nullptr;
    int init(cudaStream_t s) {
        st = s;
        XB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&ptr), sizeof(unsigned long long), st));
        XB_CUDA_CHECK(cudaMemsetAsync(ptr, 0, sizeof(unsigned long long), st));
        return XB_OK;
    }
    ~WorkCounter() {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

}  // namespace xbv

extern "C" {
#pragma GCC visibility push(default)

int xb_variogram_group_size(void) { return xbv::GS; }
int xb_variogram_chunk(void) { return xbv::CHUNK_J; }

int xb_variogram_pairs(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                       const unsigned long long* edge2_dev, int n_bins, const int64_t* unit_prefix_dev,
                       int64_t unit_begin, int64_t unit_end, int wide, int estimator, unsigned long long* count_dev,
                       double* sumsq_dev, void* stream) {
    if (!pts_dev || !gbox_dev || !edge2_dev || !unit_prefix_dev || !count_dev || !sumsq_dev || n_groups <= 0 ||
        n_bins <= 0 || n_groups > 0x7fffffff) {
        xb_set_error("bad arguments to xb_variogram_pairs");
        return XB_ERR_INVALID;
    }
    if (unit_end <= unit_begin) return XB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    xbv::WorkCounter wc;
    int rc = wc.init(st);
    if (rc) return rc;
    int num_sms = 0;
    rc = xb_num_sms(&num_sms);
    if (rc) return rc;
    const long long units = unit_end - unit_begin;
    long long grid = std::min<long long>((long long)num_sms * 2, (units + xbv::NWARPS - 1) / xbv::NWARPS);
    if (grid < 1) grid = 1;
    const int4* pts = reinterpret_cast<const int4*>(pts_dev);
    const int4* gbox = reinterpret_cast<const int4*>(gbox_dev);
    const long long* pref = reinterpret_cast<const long long*>(unit_prefix_dev);
    if (estimator != 0 && estimator != 1) {
        xb_set_error("estimator must be 0 (matheron: sum of squares) or 1 (cressie: sum of |diff|^0.5)");
        return XB_ERR_INVALID;
    }
#define XB_VG_LAUNCH(D2T, EST)                                                                                     \
    xbv::variogram_pairs_kernel<D2T, EST><<<(unsigned)grid, xbv::NTHREADS, 0, st>>>(                                \
        pts, gbox, (int)n_groups, edge2_dev, n_bins, pref, unit_begin, unit_end, wc.ptr, count_dev, sumsq_dev,     \
        xb_option_variogram_full_tiles())
    if (wide) {
        if (estimator) XB_VG_LAUNCH(unsigned long long, 1); else XB_VG_LAUNCH(unsigned long long, 0);
    } else {
        if (estimator) XB_VG_LAUNCH(unsigned, 1); else XB_VG_LAUNCH(unsigned, 0);
    }
#undef XB_VG_LAUNCH
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_variogram_median_pass(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                              const unsigned long long* edge2_dev, int n_bins, const int64_t* unit_prefix_dev,
                              int64_t unit_begin, int64_t unit_end, int mode, const uint32_t* prefix_dev,
                              uint32_t prefix_mask, int shift, unsigned long long* hist_dev, uint32_t* next_key_dev,
                              void* stream) {
    if (!pts_dev || !gbox_dev || !edge2_dev || !unit_prefix_dev || !prefix_dev || n_groups <= 0 || n_bins <= 0 ||
        n_groups > 0x7fffffff || (mode == 0 && !hist_dev) || (mode == 1 && !next_key_dev) || (mode != 0 && mode != 1)) {
        xb_set_error("bad arguments to xb_variogram_median_pass");
        return XB_ERR_INVALID;
    }
    const size_t smem = (size_t)xbv::NWARPS * xbv::GS * sizeof(int4) + (size_t)(mode == 0 ? n_bins * 256 : n_bins) * 4;
    if (smem > 200 * 1024) {
        xb_set_error("too many lag classes for the shared-memory histogram (%d)", n_bins);
        return XB_ERR_UNSUPPORTED;
    }
    if (unit_end <= unit_begin) return XB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    xbv::WorkCounter wc;
    int rc = wc.init(st);
    if (rc) return rc;
    int num_sms = 0;
    rc = xb_num_sms(&num_sms);
    if (rc) return rc;
    const long long units = unit_end - unit_begin;
    long long grid = std::min<long long>((long long)num_sms * (smem > 100 * 1024 ? 1 : 2),
                                         (units + xbv::NWARPS - 1) / xbv::NWARPS);
    if (grid < 1) grid = 1;
    const int4* pts = reinterpret_cast<const int4*>(pts_dev);
    const int4* gbox = reinterpret_cast<const int4*>(gbox_dev);
    const long long* pref = reinterpret_cast<const long long*>(unit_prefix_dev);
    auto launch = [&](auto kern) -> int {
        XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, xbv::NTHREADS, smem, st>>>(pts, gbox, (int)n_groups, edge2_dev, n_bins, pref, unit_begin,
                                                          unit_end, wc.ptr, prefix_dev, prefix_mask, shift,
                                                          hist_dev, next_key_dev);
        XB_CUDA_CHECK(cudaGetLastError());
        return XB_OK;
    };
    rc = mode == 0 ? launch(xbv::variogram_median_kernel<unsigned long long, 0>)
                   : launch(xbv::variogram_median_kernel<unsigned long long, 1>);
    if (rc) return rc;
    xb_count_launch(1);
    return XB_OK;
}

int xb_variogram_maxd2(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                       unsigned long long* maxd2_dev, void* stream) {
    if (!pts_dev || !gbox_dev || !maxd2_dev || n_groups <= 0 || n_groups > 0x7fffffff) {
        xb_set_error("bad arguments to xb_variogram_maxd2");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    xbv::WorkCounter wc;
    int rc = wc.init(st);
    if (rc) return rc;
    int num_sms = 0;
    rc = xb_num_sms(&num_sms);
    if (rc) return rc;
    long long grid = std::min<long long>((long long)num_sms * 2, (n_groups + xbv::NWARPS - 1) / xbv::NWARPS);
    xbv::variogram_maxd2_kernel<<<(unsigned)grid, xbv::NTHREADS, 0, st>>>(
        reinterpret_cast<const int4*>(pts_dev), reinterpret_cast<const int4*>(gbox_dev), (int)n_groups,
        wc.ptr, maxd2_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
