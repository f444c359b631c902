// xdem_b200 -- empirical variogram, general-coordinate pair binning (float64 coordinates), sm_100a.
//
// The grid kernel (xb_variogram.cu) works on integer pixel coordinates: exact and fast, but only for samples of a
// regular grid whose spacing is exact in binary.  This kernel takes arbitrary float64 coordinates and pairs either all
// samples of one set (A x A, i < j: `_get_pdist_empirical_variogram`, spatialstats.py:1064-1101, incl. 1-D values +
// `coords=` and non-dyadic `gsd`) or two sets (A x B: `_get_cdist_empirical_variogram`, spatialstats.py:1186-1261, the
// centre-disk x ring samples of the reference's default `cdist_equidistant` sampler and the two random subsets of
// `cdist_point`).  What it replaces is the pair work inside skgstat.Variogram (third-party, absent: the restatement is
// UNPINNED).
//
// Arithmetic = what scipy / cKDTree do in float64, in the same order: d2 = RN(RN(dx*dx) + RN(dy*dy)) (un-contracted),
// d = sqrt(d2).  Instead of taking the square root per pair, the host turns every bin edge e into the threshold
// T = min{t : sqrt_fl(t) >= e} (rule "left"; "> e" for rule "right") -- sqrt is monotone, so d < e <=> d2 < T exactly --
// and a pair's class is found by binary search over the thresholds.  diff = |v_i - v_j| in float64 (exact for float32
// rasters), per class: pair count and sum of diff^2 (Matheron) or sqrt(diff) (Cressie-Hawkins); or, for Dowd's median,
// the class number and the float32 key of diff of every pair are written out for the radix select of xb_binning.cu.
//
// Not HBM-bound (inputs are KBs..MBs): one thread per A-sample, B staged through shared memory in chunks, a per-thread
// run-length cache (class, count, partial sum) so that shared-memory atomics are only issued when the class changes
// (samples arrive sorted along a space-filling curve), one global atomic per class per CTA at the end.
#include "../../include/xdem_b200.h"

#include "xb_common.cuh"

namespace xbx {

constexpr int TPB = 256;
constexpr int CHUNK = 512;
constexpr int MAX_BINS = 1024;

struct Params {
    const double *xa, *ya, *va, *xb, *yb, *vb;
    long long na, nb;
    int same;  // A == B: only pairs i < j
    const double* thr;
    int n_bins, estimator;  // 0 sum diff^2, 1 sum sqrt(diff), 2 emit (class, key) per pair
    unsigned long long* count;
    double* sum;
    unsigned long long* maxd2_bits;
    unsigned short* pair_class;  // estimator 2: [na * nb]
    unsigned int* pair_key;
    // optional pair de-duplication for A x B (both NULL: every pair counts): sample identities and "this sample is in
    // BOTH sets" flags; a pair is skipped when it joins a sample with itself, or -- if both samples belong to both sets,
    // so that the pair shows up in both orientations -- in the orientation ida > idb
    const long long *ida, *idb;
    const unsigned char *dupa, *dupb;
};

__global__ void __launch_bounds__(TPB) pairs_xy_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s_thr = reinterpret_cast<double*>(smem_raw);                   // n_bins
    double* s_sum = s_thr + p.n_bins;                                      // n_bins
    unsigned long long* s_cnt = reinterpret_cast<unsigned long long*>(s_sum + p.n_bins);  // n_bins
    double* s_x = reinterpret_cast<double*>(s_cnt + p.n_bins);             // CHUNK each
    double* s_y = s_x + CHUNK;
    double* s_v = s_y + CHUNK;
    long long* s_id = reinterpret_cast<long long*>(s_v + CHUNK);
    unsigned char* s_dup = reinterpret_cast<unsigned char*>(s_id + CHUNK);
    __shared__ unsigned long long s_max;

    const int tid = threadIdx.x;
    for (int k = tid; k < p.n_bins; k += TPB) {
        s_thr[k] = p.thr[k];
        s_sum[k] = 0.0;
        s_cnt[k] = 0ull;
    }
    if (tid == 0) s_max = 0ull;
    __syncthreads();

    const long long a_tiles = (p.na + TPB - 1) / TPB;
    const long long b_chunks = (p.nb + CHUNK - 1) / CHUNK;
    const int nb1 = p.n_bins - 1;
    unsigned long long my_max = 0ull;
    for (long long unit = blockIdx.x; unit < a_tiles * b_chunks; unit += gridDim.x) {
        const long long at = unit / b_chunks, bc = unit % b_chunks;
        const long long i = at * TPB + tid;
        const long long j0 = bc * CHUNK, j1 = min(p.nb, j0 + CHUNK);
        if (p.same && j1 <= at * TPB + 1) continue;  // chunk entirely at or below the tile's first row (uniform)
        __syncthreads();
        for (long long j = j0 + tid; j < j1; j += TPB) {
            s_x[j - j0] = p.xb[j];
            s_y[j - j0] = p.yb[j];
            s_v[j - j0] = p.vb[j];
            if (p.idb) {
                s_id[j - j0] = p.idb[j];
                s_dup[j - j0] = p.dupb ? p.dupb[j] : (unsigned char)0;
            }
        }
        __syncthreads();
        if (i >= p.na) continue;
        const double xi = p.xa[i], yi = p.ya[i], vi = p.va[i];
        const long long idi = p.ida ? p.ida[i] : -1;
        const bool dupi = p.dupa ? p.dupa[i] != 0 : false;
        int cur = -1;
        unsigned long long cnt = 0ull;
        double acc = 0.0;
        long long jb = j0;
        if (p.same && jb <= i) jb = i + 1;
        for (long long j = jb; j < j1; ++j) {
            if (p.ida && p.idb) {
                const long long idj = s_id[j - j0];
                if (idi == idj || (dupi && s_dup[j - j0] && idi > idj)) continue;  // estimator 2: slot stays 0xFFFF
            }
            const double dx = xi - s_x[j - j0], dy = yi - s_y[j - j0];
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const unsigned long long bits = (unsigned long long)__double_as_longlong(d2);
            my_max = bits > my_max && d2 == d2 ? bits : my_max;
            int k = -1;
            if (d2 < s_thr[nb1]) {  // first k with d2 < thr[k]
                int lo = 0, hi = nb1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (d2 < s_thr[mid]) hi = mid; else lo = mid + 1;
                }
                k = lo;
            }
            const double diff = fabs(vi - s_v[j - j0]);
            if (p.estimator == 2) {
                const size_t o = (size_t)i * (size_t)p.nb + (size_t)j;
                const bool ok = k >= 0 && diff == diff;
                p.pair_class[o] = ok ? (unsigned short)k : (unsigned short)0xFFFF;
                p.pair_key[o] = __float_as_uint((float)diff) | 0x80000000u;  // order-preserving key (xb_binning.cu)
                continue;
            }
            if (k != cur) {
                if (cur >= 0 && cnt) {
                    atomicAdd(&s_cnt[cur], cnt);
                    atomicAdd(&s_sum[cur], acc);
                }
                cur = k, cnt = 0ull, acc = 0.0;
            }
            if (k >= 0 && diff == diff) {
                ++cnt;
                acc += p.estimator == 1 ? sqrt(diff) : diff * diff;
            }
        }
        if (cur >= 0 && cnt) {
            atomicAdd(&s_cnt[cur], cnt);
            atomicAdd(&s_sum[cur], acc);
        }
    }
    if (my_max) atomicMax(&s_max, my_max);
    __syncthreads();
    if (p.estimator != 2)
        for (int k = tid; k < p.n_bins; k += TPB)
            if (s_cnt[k]) {
                atomicAdd(&p.count[k], s_cnt[k]);
                atomicAdd(&p.sum[k], s_sum[k]);
            }
    if (tid == 0 && s_max && p.maxd2_bits) atomicMax(p.maxd2_bits, s_max);
}

// estimator 2 needs every slot of the (na x nb) rectangle initialised (pairs the loop skips: j <= i, dropped pairs)
__global__ void fill_class_kernel(unsigned short* cls, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        cls[i] = 0xFFFF;
}

}  // namespace xbx

extern "C" {
#pragma GCC visibility push(default)

int xb_variogram_pairs_xy(const double* xa_dev, const double* ya_dev, const double* va_dev, int64_t na,
                          const double* xb_dev, const double* yb_dev, const double* vb_dev, int64_t nb,
                          const double* thr_d2_dev, int n_bins, int estimator, unsigned long long* count_dev,
                          double* sum_dev, unsigned long long* maxd2_bits_dev, uint16_t* pair_class_dev,
                          uint32_t* pair_key_dev, const int64_t* ida_dev, const int64_t* idb_dev,
                          const uint8_t* dupa_dev, const uint8_t* dupb_dev, void* stream) {
    if (!xa_dev || !ya_dev || !va_dev || na <= 0 || !thr_d2_dev || n_bins < 1 || n_bins > xbx::MAX_BINS ||
        estimator < 0 || estimator > 2) {
        xb_set_error("bad arguments to xb_variogram_pairs_xy (1 <= n_bins <= %d, estimator 0..2)", xbx::MAX_BINS);
        return XB_ERR_INVALID;
    }
    if (estimator != 2 && (!count_dev || !sum_dev)) {
        xb_set_error("count / sum outputs are NULL");
        return XB_ERR_INVALID;
    }
    if (estimator == 2 && (!pair_class_dev || !pair_key_dev)) {
        xb_set_error("estimator 2 (per-pair output) needs pair_class_dev and pair_key_dev");
        return XB_ERR_INVALID;
    }
    xbx::Params p;
    p.xa = xa_dev, p.ya = ya_dev, p.va = va_dev, p.na = na;
    p.same = (xb_dev == nullptr) ? 1 : 0;
    p.xb = p.same ? xa_dev : xb_dev;
    p.yb = p.same ? ya_dev : yb_dev;
    p.vb = p.same ? va_dev : vb_dev;
    p.nb = p.same ? na : nb;
    if (!p.same && (!yb_dev || !vb_dev || nb <= 0)) {
        xb_set_error("second sample set is incomplete");
        return XB_ERR_INVALID;
    }
    p.thr = thr_d2_dev, p.n_bins = n_bins, p.estimator = estimator;
    p.count = count_dev, p.sum = sum_dev, p.maxd2_bits = maxd2_bits_dev;
    p.pair_class = pair_class_dev, p.pair_key = pair_key_dev;
    p.ida = reinterpret_cast<const long long*>(ida_dev), p.idb = reinterpret_cast<const long long*>(idb_dev);
    p.dupa = dupa_dev, p.dupb = dupb_dev;
    if (p.same) p.ida = p.idb = nullptr, p.dupa = p.dupb = nullptr;
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (estimator == 2) {
        xbx::fill_class_kernel<<<sms * 4, 256, 0, st>>>(pair_class_dev, (size_t)na * (size_t)p.nb);
        XB_CUDA_CHECK(cudaGetLastError());
        xb_count_launch(1);
    }
    const size_t smem = (size_t)n_bins * 24 + (size_t)xbx::CHUNK * (24 + 8 + 1);
    const long long units = ((na + xbx::TPB - 1) / xbx::TPB) * ((p.nb + xbx::CHUNK - 1) / xbx::CHUNK);
    const long long grid = units < (long long)sms * 8 ? units : (long long)sms * 8;
    XB_CUDA_CHECK(cudaFuncSetAttribute(xbx::pairs_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xbx::pairs_xy_kernel<<<(unsigned)grid, xbx::TPB, smem, st>>>(p);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

#pragma GCC visibility pop
}
