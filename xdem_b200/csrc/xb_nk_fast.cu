// xdem_b200 -- Nuth & Kaab (2011) iteration with bracketed exact selection ("fast path"), sm_100a.
//
// What one iteration of xdem.coreg.affine._nuth_kaab_iteration_step (affine.py:477-536) needs: np.nanmedian(dh) over the
// whole raster (:504) and, with y = (dh - median)/slope_tan, the np.nanmedian of y in each of the 72 aspect bins
// (_nuth_kaab_bin_fit -> nd_binning -> binned_statistic, affine.py:358-409, base.py:1014-1020) plus count / mean / std of
// y.  The first implementation (xb_nuthkaab.cu) found every median by MSD radix select over ALL elements: 3 + 4 streaming
// passes per iteration with a host round trip after each (ncu r01c / bench r01e: 6.3 ms per 16384^2 iteration, 0.11 of
// the HBM roofline, ~58 B/px).  Here every median is bracketed first:
//
//   1. dh on a row sample (every `stride`-th row, jittered)            -> ~1.6 % of the pixels
//   2. exact order statistics of the sample (radix select on the small buffer, ranks picked ON THE DEVICE) give a
//      key bracket [lo, hi] that contains the population median with overwhelming probability (+-4 sigma of the
//      sample-rank distribution) and ~0.2 % of the elements
//   3. ONE full pass computes dh, writes it, counts the keys below lo and appends the keys inside [lo, hi] to a compact
//      buffer (17 B/px)
//   4. the exact median = the element of rank (N-1)/2 - below (and N/2 - below) inside the compact buffer: radix select on
//      ~0.5 M keys; its mean is the vertical shift, left on the device
//   5.-8. the same for y per aspect bin: sampled rows -> per-bin brackets -> ONE full pass over (dh, slope_tan, cached bin)
//      that counts per bin (total, below) and appends the in-bracket (key, bin) pairs (9 B/px) -> per-bin exact medians.
//
// 26 B/px instead of 58, two streaming passes instead of eight, and a single host synchronisation per iteration (to read
// the 72 medians for the host curve_fit).  Exactness does not depend on the sample: a bracket that misses its median or
// overflows its buffer raises a flag and the caller repeats that iteration on the exhaustive path -- results are the
// same exact medians (np.nanmedian semantics incl. the mean of the two middle values) either way.
// Multi-GPU: samples and compact buffers are all-gathered (a few MB), counters all-reduced between the stages; every rank
// then selects on the union and obtains identical medians.
#include "../../include/xdem_b200.h"

#include <math_constants.h>

#include <type_traits>

#include "xb_common.cuh"

namespace xbf {

constexpr int NT = 256;
constexpr int MAXB = 96;
constexpr int WSTAGE = 640;  // per-warp staging of compacted entries (a warp sees 2048 pixels of a 16384-wide row)
// layout of the small device state arrays (also exported through xb_nkf_layout)
enum { C_NFIN = 0, C_GBELOW = 1, C_GNC = 2, C_BNC = 3, C_FLAGS = 4, C_RFALL = 5, C_ROWQ_DH = 6, C_ROWQ_Y = 7, C_BTOTAL = 8, C_BBELOW = 8 + MAXB, C_SIZE = 8 + 2 * MAXB };
enum { K_ASPMIN = 0, K_ASPMAX = 1, K_GLO = 2, K_GHI = 3, K_BLO = 4, K_BHI = 4 + MAXB, K_SIZE = 4 + 2 * MAXB };
enum { F_VSHIFT = 0, F_ASPLO = 1, F_ASPHI = 2, F_CLO = 3, F_CHI = 4, F_M0 = 5, F_MED = 8, F_SIZE = 8 + MAXB };
enum { FLAG_GMISS = 1, FLAG_GOVER = 2, FLAG_BMISS = 4, FLAG_BOVER = 8 };

__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned key) {
    return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

// jittered sample of 4-pixel chunks: window k of `stride` consecutive chunks (row-major) contributes one chunk.  Chunks --
// not rows -- because neighbouring pixels are correlated (the bilinear resampling of the noise couples adjacent dh values):
// scattered chunks keep the sample close to independent draws, which the rank margin of the brackets assumes
__device__ __forceinline__ long long sample_chunk(long long k, int stride, unsigned seed, long long n_chunks) {
    unsigned h = (unsigned)k * 2654435761u ^ seed;
    h ^= h >> 16, h *= 0x85ebca6bu, h ^= h >> 13, h *= 0xc2b2ae35u, h ^= h >> 16;
    const long long q = k * stride + (long long)(h % (unsigned)stride);
    return q < n_chunks ? q : n_chunks - 1;
}

// dh of four consecutive pixels (same FP64 expressions as xbn::nk_dh_vec4_kernel; NaN where the mask is 0 or any
// contributing cell is NaN / outside the raster)
__device__ __forceinline__ void dh4(const float* __restrict__ ref, const float* __restrict__ tba,
                                    const unsigned char* __restrict__ sub_mask, long long r, long long c, long long cols,
                                    long long ld, long long tba_ld, long long tba_row0, long long tba_rows_total,
                                    long long i0, long long j0, double w00, double w01, double w10, double w11,
                                    float (&out)[4]) {
    const uchar4 m4 = *reinterpret_cast<const uchar4*>(sub_mask + r * cols + c);
    const float4 r4 = *reinterpret_cast<const float4*>(ref + r * ld + c);
    const unsigned char mk[4] = {m4.x, m4.y, m4.z, m4.w};
    const float rf[4] = {r4.x, r4.y, r4.z, r4.w};
    const long long rr = r + i0 + tba_row0, cc = c + j0;
    if (rr >= 0 && rr + 1 < tba_rows_total && cc >= 0 && cc + 4 < cols) {
        const float* t = tba + rr * tba_ld + cc;
        float ta[5], tb[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) ta[k] = t[k], tb[k] = t[tba_ld + k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double acc = w00 * (double)ta[k] + w01 * (double)ta[k + 1] + w10 * (double)tb[k] + w11 * (double)tb[k + 1];
            out[k] = mk[k] ? (float)((double)rf[k] - acc) : CUDART_NAN_F;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        out[k] = CUDART_NAN_F;
        if (!mk[k]) continue;
        const long long ck = cc + k;
        double acc = CUDART_NAN;
        if (rr >= 0 && rr + 1 < tba_rows_total && ck >= 0 && ck + 1 < cols) {
            const float* t = tba + rr * tba_ld + ck;
            acc = w00 * (double)t[0] + w01 * (double)t[1] + w10 * (double)t[tba_ld] + w11 * (double)t[tba_ld + 1];
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

struct DhArgs {
    const float *ref, *tba, *aspect;
    const unsigned char* sub_mask;
    long long rows, cols, ld, tba_ld, tba_row0, tba_rows_total, i0, j0;
    double w00, w01, w10, w11;
};

// SAMPLE: dh of the sampled rows -> ordered keys (0 = not finite).  FULL: dh of every pixel -> dh[], aspect range and
// count over finite dh, count of keys below the bracket, in-bracket keys appended to gcompact.
template <bool SAMPLE>
__global__ void __launch_bounds__(NT)
nkf_dh_kernel(const DhArgs a, float* __restrict__ dh, unsigned* __restrict__ sample, int stride, unsigned seed,
              long long n_schunks, unsigned long long* __restrict__ cnt, unsigned* __restrict__ keys,
              unsigned* __restrict__ gcompact, unsigned long long gcap) {
    if constexpr (SAMPLE) {
        const long long n_chunks = a.rows * (a.cols / 4);
        for (long long k = (long long)blockIdx.x * NT + threadIdx.x; k < n_schunks; k += (long long)gridDim.x * NT) {
            const long long q = sample_chunk(k, stride, seed, n_chunks);
            const long long r = q / (a.cols / 4), c = 4 * (q - r * (a.cols / 4));
            float out[4];
            dh4(a.ref, a.tba, a.sub_mask, r, c, a.cols, a.ld, a.tba_ld, a.tba_row0, a.tba_rows_total, a.i0, a.j0, a.w00,
                a.w01, a.w10, a.w11, out);
            *reinterpret_cast<uint4*>(sample + 4 * k) =
                make_uint4(isfinite(out[0]) ? ordered_key(out[0]) : 0u, isfinite(out[1]) ? ordered_key(out[1]) : 0u,
                           isfinite(out[2]) ? ordered_key(out[2]) : 0u, isfinite(out[3]) ? ordered_key(out[3]) : 0u);
        }
        return;
    }
    // in-bracket keys are staged per WARP in shared memory (no atomics: only the warp touches its region and counter) and
    // flushed once per raster row with one global atomic per warp (a global atomic per ballot serialised at the L2:
    // 1.76 ms instead of 0.95 ms at 16384^2, ncu r02)
    __shared__ unsigned s_stage[NT / 32][WSTAGE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned wn = 0;  // staged entries of this warp (warp-uniform)
    const unsigned glo = keys[K_GLO], ghi = keys[K_GHI];
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned long long nfin = 0, below = 0;
    for (long long r = blockIdx.x; r < a.rows; r += gridDim.x) {
        for (long long c0 = 0; c0 < a.cols; c0 += 4ll * NT) {
            const long long c = c0 + 4ll * threadIdx.x;
            const bool in = c < a.cols;
            float out[4] = {CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F};
            float as[4] = {0.f, 0.f, 0.f, 0.f};
            if (in) {
                dh4(a.ref, a.tba, a.sub_mask, r, c, a.cols, a.ld, a.tba_ld, a.tba_row0, a.tba_rows_total, a.i0, a.j0,
                    a.w00, a.w01, a.w10, a.w11, out);
                *reinterpret_cast<float4*>(dh + r * a.cols + c) = make_float4(out[0], out[1], out[2], out[3]);
                if (a.aspect) {  // NULL: the aspect range comes from nkf_range_check_kernel (single GPU)
                    const float4 a4 = *reinterpret_cast<const float4*>(a.aspect + r * a.cols + c);
                    as[0] = a4.x, as[1] = a4.y, as[2] = a4.z, as[3] = a4.w;
                }
            }
            unsigned key[4];
            unsigned tmask = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool fin = isfinite(out[u]);
                key[u] = fin ? ordered_key(out[u]) : 0u;
                if (fin) {
                    const unsigned ab = __float_as_uint(as[u]);  // aspect >= 0: bit pattern is monotonic
                    lmin = min(lmin, ab), lmax = max(lmax, ab);
                    ++nfin;
                    below += key[u] < glo;
                    tmask |= (key[u] >= glo && key[u] <= ghi) ? (1u << u) : 0u;
                }
            }
            if (__any_sync(0xffffffffu, tmask != 0u)) {
                // exclusive scan of the lanes' take counts
                const unsigned nt = __popc(tmask);
                unsigned incl = nt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                unsigned pos = wn + incl - nt;
                wn += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (tmask & (1u << u)) {
                        if (pos < WSTAGE) {
                            s_stage[warp][pos] = key[u];
                        } else {  // staging full (heavy ties): straight to the global buffer
                            const unsigned long long gi = atomicAdd(&cnt[C_GNC], 1ull);
                            if (gi < gcap) gcompact[gi] = key[u];
                        }
                        ++pos;
                    }
            }
        }
        // flush the row's staged keys of this warp
        const unsigned n_st = min(wn, (unsigned)WSTAGE);
        if (n_st) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&cnt[C_GNC], (unsigned long long)n_st);
            base = __shfl_sync(0xffffffffu, base, 0);
            __syncwarp();
            for (unsigned i = lane; i < n_st; i += 32)
                if (base + i < gcap) gcompact[base + i] = s_stage[warp][i];
            __syncwarp();
        }
        wn = 0;
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nfin += __shfl_xor_sync(0xffffffffu, nfin, o);
        below += __shfl_xor_sync(0xffffffffu, below, o);
    }
    if (lane == 0) {
        if (a.aspect && lmin != 0xffffffffu) atomicMin(&keys[K_ASPMIN], lmin);
        if (a.aspect && nfin) atomicMax(&keys[K_ASPMAX], lmax);
        if (nfin) atomicAdd(&cnt[C_NFIN], nfin);
        if (below) atomicAdd(&cnt[C_GBELOW], below);
    }
}

// The FULL pass again, specialised (ncu r02f on the kernel above: 58 thread-instructions per pixel of which 31 are 64-bit
// index / bounds arithmetic repeated per 4-pixel chunk, and ten unaligned 4-byte loads per chunk that each touch a
// 512-byte span of the warp): the column shift (c + j0) mod 4 is the same for every chunk of the raster, so the kernel
// is instantiated per SHIFT and reads the two to-be-aligned rows as two ALIGNED float4 each, picking the five values of
// the 2x2 stencils at compile-time positions; everything that only depends on the row is hoisted out of the column
// loop, columns are 32-bit, and the aspect-range reduction only exists in the instantiation that needs it.  Chunks /
// rows that touch the raster border take dh4().  Same FP64 expressions, term for term, as dh4 -> bit-identical dh.
template <int SHIFT, bool HAS_ASPECT>
__global__ void __launch_bounds__(NT)
nkf_dh_full_kernel(const DhArgs a, float* __restrict__ dh, unsigned long long* __restrict__ cnt, unsigned* __restrict__ keys,
                   unsigned* __restrict__ gcompact, unsigned long long gcap) {
    __shared__ unsigned s_stage[NT / 32][WSTAGE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned wn = 0;  // staged entries of this warp (warp-uniform)
    const unsigned glo = keys[K_GLO], ghi = keys[K_GHI];
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned nfin = 0, below = 0;  // per thread: < 2^32 pixels
    const int cols = (int)a.cols;
    const int j0 = (int)a.j0;  // |j0| < cols checked by the host
    const double w00 = a.w00, w01 = a.w01, w10 = a.w10, w11 = a.w11;
    // rows are handed out dynamically (one atomic per row and CTA): the grid is exactly the resident CTAs, so there is
    // no second, partially filled wave (1184 CTAs on 740 resident slots ran at 80 % of the machine, ncu r02f)
    __shared__ long long s_row;
    for (;;) {
        if (threadIdx.x == 0) s_row = (long long)atomicAdd(&cnt[C_ROWQ_DH], 1ull);
        __syncthreads();
        const long long r = s_row;
        __syncthreads();
        if (r >= a.rows) break;
        const long long rr = r + a.i0 + a.tba_row0;
        const bool row_fast = rr >= 0 && rr + 1 < a.tba_rows_total;
        const float* __restrict__ t0 = a.tba + rr * a.tba_ld;
        const float* __restrict__ t1 = t0 + a.tba_ld;
        const float* __restrict__ ref_r = a.ref + r * a.ld;
        const unsigned char* __restrict__ mk_r = a.sub_mask + r * a.cols;
        const float* __restrict__ as_r = HAS_ASPECT ? a.aspect + r * a.cols : nullptr;
        float* __restrict__ dh_r = dh + r * a.cols;
        for (int c0 = 0; c0 < cols; c0 += 4 * NT) {
            const int c = c0 + 4 * (int)threadIdx.x;
            const bool in = c < cols;
            float out[4] = {CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F};
            float as[4] = {0.f, 0.f, 0.f, 0.f};
            if (in) {
                const int base = c + j0 - SHIFT;  // aligned (multiple of 4) start of the eight staged columns
                if (row_fast && base >= 0 && base + 8 <= cols) {
                    const uchar4 m4 = *reinterpret_cast<const uchar4*>(mk_r + c);
                    const float4 r4 = *reinterpret_cast<const float4*>(ref_r + c);
                    const float4 a0 = *reinterpret_cast<const float4*>(t0 + base);
                    const float4 b0 = *reinterpret_cast<const float4*>(t0 + base + 4);
                    const float4 a1 = *reinterpret_cast<const float4*>(t1 + base);
                    const float4 b1 = *reinterpret_cast<const float4*>(t1 + base + 4);
                    const float e0[8] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w};
                    const float e1[8] = {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w};
                    const unsigned char mk[4] = {m4.x, m4.y, m4.z, m4.w};
                    const float rf[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double acc = w00 * (double)e0[SHIFT + k] + w01 * (double)e0[SHIFT + k + 1] +
                                           w10 * (double)e1[SHIFT + k] + w11 * (double)e1[SHIFT + k + 1];
                        out[k] = mk[k] ? (float)((double)rf[k] - acc) : CUDART_NAN_F;
                    }
                } else {
                    dh4(a.ref, a.tba, a.sub_mask, r, c, a.cols, a.ld, a.tba_ld, a.tba_row0, a.tba_rows_total, a.i0, a.j0,
                        w00, w01, w10, w11, out);
                }
                *reinterpret_cast<float4*>(dh_r + c) = make_float4(out[0], out[1], out[2], out[3]);
                if constexpr (HAS_ASPECT) {
                    const float4 a4 = *reinterpret_cast<const float4*>(as_r + c);
                    as[0] = a4.x, as[1] = a4.y, as[2] = a4.z, as[3] = a4.w;
                }
            }
            unsigned key[4];
            unsigned tmask = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool fin = fabsf(out[u]) < CUDART_INF_F;
                key[u] = ordered_key(out[u]);
                if constexpr (HAS_ASPECT) {
                    const unsigned ab = __float_as_uint(as[u]);  // aspect >= 0: bit pattern is monotonic
                    lmin = fin ? min(lmin, ab) : lmin, lmax = fin ? max(lmax, ab) : lmax;
                }
                nfin += fin;
                below += fin && key[u] < glo;
                tmask |= (fin && key[u] >= glo && key[u] <= ghi) ? (1u << u) : 0u;
            }
            if (__any_sync(0xffffffffu, tmask != 0u)) {
                // exclusive scan of the lanes' take counts
                const unsigned nt = __popc(tmask);
                unsigned incl = nt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                unsigned pos = wn + incl - nt;
                wn += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (tmask & (1u << u)) {
                        if (pos < WSTAGE) {
                            s_stage[warp][pos] = key[u];
                        } else {  // staging full (heavy ties): straight to the global buffer
                            const unsigned long long gi = atomicAdd(&cnt[C_GNC], 1ull);
                            if (gi < gcap) gcompact[gi] = key[u];
                        }
                        ++pos;
                    }
            }
        }
        // flush the row's staged keys of this warp
        const unsigned n_st = min(wn, (unsigned)WSTAGE);
        if (n_st) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&cnt[C_GNC], (unsigned long long)n_st);
            base = __shfl_sync(0xffffffffu, base, 0);
            __syncwarp();
            for (unsigned i = lane; i < n_st; i += 32)
                if (base + i < gcap) gcompact[base + i] = s_stage[warp][i];
            __syncwarp();
        }
        wn = 0;
    }
    if constexpr (HAS_ASPECT) {
        lmin = __reduce_min_sync(0xffffffffu, lmin);
        lmax = __reduce_max_sync(0xffffffffu, lmax);
    }
    const unsigned wfin = __reduce_add_sync(0xffffffffu, nfin), wbel = __reduce_add_sync(0xffffffffu, below);
    // (a thread sees < 2^32 pixels, a warp's sum could only wrap beyond 2^32 pixels per warp: rows*cols/gridDim/8)
    if (lane == 0) {
        if (HAS_ASPECT && lmin != 0xffffffffu) atomicMin(&keys[K_ASPMIN], lmin);
        if (HAS_ASPECT && wfin) atomicMax(&keys[K_ASPMAX], lmax);
        if (wfin) atomicAdd(&cnt[C_NFIN], (unsigned long long)wfin);
        if (wbel) atomicAdd(&cnt[C_GBELOW], (unsigned long long)wbel);
    }
}

// aspect bin of binned_statistic(bins=n, range=None): see xbn::aspect_bin (xb_nuthkaab.cu) -- same code
__device__ __forceinline__ int aspect_bin(float a, double lo, double hi, double step, double inv_step, int n_bins) {
    const double x = (double)a;
    const double t = (x - lo) * inv_step;
    int k = (int)t;
    k = max(0, min(n_bins - 1, k));
    const double frac = t - (double)k;
    if (frac > 1e-4 && frac < 1.0 - 1e-4) return k;
    auto edge = [&](int j) -> float {
        return j >= n_bins ? (float)hi : (float)__dadd_rn(__dmul_rn((double)j, step), lo);
    };
    while (k > 0 && a < edge(k)) --k;
    while (k < n_bins - 1 && a >= edge(k + 1)) ++k;
    return k;
}

__device__ __forceinline__ bool y_key(float dhv, float st, double vshift, unsigned& key, float& yf) {
    if (!isfinite(dhv)) return false;
    yf = __fdiv_rn((float)((double)dhv - vshift), st);
    if (!isfinite(yf)) return false;
    key = ordered_key(yf);
    return true;
}

// SAMPLE: (key of y, aspect bin) of the sampled rows.  FULL: every pixel -> per-bin totals and below-bracket counts,
// moments of y, in-bracket (key, bin) pairs appended; the aspect bins are cached per pixel while the range is unchanged.
template <bool SAMPLE>
__global__ void __launch_bounds__(NT)
nkf_y_kernel(const float* __restrict__ dh, const float* __restrict__ slope_tan, const float* __restrict__ aspect,
             unsigned char* __restrict__ bin_cache, long long rows, long long cols, int n_bins,
             unsigned* __restrict__ skey, unsigned char* __restrict__ sgrp, int stride, unsigned seed, long long n_schunks,
             unsigned long long* __restrict__ cnt, const unsigned* __restrict__ keys, double* __restrict__ f64,
             unsigned* __restrict__ bkey, unsigned char* __restrict__ bgrp, unsigned long long bcap) {
    const double vshift = f64[F_VSHIFT], lo = f64[F_ASPLO], hi = f64[F_ASPHI];
    const bool reuse = !SAMPLE && f64[F_CLO] == lo && f64[F_CHI] == hi;
    const double step = (hi - lo) / (double)n_bins;
    const double inv_step = step > 0.0 ? 1.0 / step : 0.0;
    if constexpr (SAMPLE) {
        const long long n_chunks = rows * (cols / 4);
        for (long long k = (long long)blockIdx.x * NT + threadIdx.x; k < n_schunks; k += (long long)gridDim.x * NT) {
            const long long o = 4 * sample_chunk(k, stride, seed, n_chunks);
            const float4 d4 = *reinterpret_cast<const float4*>(dh + o);
            const float4 s4 = *reinterpret_cast<const float4*>(slope_tan + o);
            const float4 a4 = *reinterpret_cast<const float4*>(aspect + o);
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w},
                        av[4] = {a4.x, a4.y, a4.z, a4.w};
            unsigned key[4] = {0u, 0u, 0u, 0u};
            unsigned char grp[4] = {255, 255, 255, 255};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float yf = 0.f;
                unsigned kk = 0u;
                if (isfinite(av[u]) && y_key(dv[u], sv[u], vshift, kk, yf)) {
                    key[u] = kk;
                    grp[u] = (unsigned char)aspect_bin(av[u], lo, hi, step, inv_step, n_bins);
                }
            }
            *reinterpret_cast<uint4*>(skey + 4 * k) = make_uint4(key[0], key[1], key[2], key[3]);
            *reinterpret_cast<uchar4*>(sgrp + 4 * k) = make_uchar4(grp[0], grp[1], grp[2], grp[3]);
        }
        return;
    }
    // FULL pass.  Shared state, addressed through 32-bit shared-window addresses (plain LEA + LDS / ATOMS; static
    // __shared__ arrays indexed per pixel cost a chain of uniform-datapath address instructions each time):
    //   s_lohi[b]       bracket (lo, hi) of bin b; slot MAXB = (0xffffffff, 0): nothing is ever inside it
    //   s_cnt[b]        keys of bin b below the bracket;  s_cnt[MAXB + 1 + b]  keys above it;  slot MAXB: sink
    // A pixel does exactly ONE unconditional shared-memory increment -- below, above, or the sink (in-bracket pixels are
    // counted by the selection's first histogram, invalid ones not at all) -- so the per-pixel code is branch-free.
    // The bin totals follow in sel_pick_kernel (mode 2): N = below + above + in-bracket.
    __shared__ uint2 s_lohi[MAXB + 1];
    __shared__ unsigned s_cnt[2 * (MAXB + 1)];
    __shared__ unsigned s_stage[NT / 32][WSTAGE];
    __shared__ unsigned char s_stage_g[NT / 32][WSTAGE];
    __shared__ unsigned s_wn[NT / 32];  // staged entries per warp
    __shared__ long long s_row;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < NT / 32) s_wn[threadIdx.x] = 0u;
    const unsigned wn_addr = (unsigned)__cvta_generic_to_shared(&s_wn[warp]);
    for (int k = threadIdx.x; k < 2 * (MAXB + 1); k += NT) s_cnt[k] = 0u;
    for (int k = threadIdx.x; k <= MAXB; k += NT)
        s_lohi[k] = k < n_bins ? make_uint2(keys[K_BLO + k], keys[K_BHI + k]) : make_uint2(0xffffffffu, 0u);
    __syncthreads();
    const unsigned lohi_addr = (unsigned)__cvta_generic_to_shared(s_lohi);
    const unsigned cnt_addr = (unsigned)__cvta_generic_to_shared(s_cnt);
    const int lane = threadIdx.x & 31;
    // dh - vshift: the reference subtracts in float64 and rounds to float32; when the median is itself a float32 (odd
    // count, or two equal middle values) the float32 subtraction rounds identically (the float64 difference of two
    // floats is exact unless their exponents are > 29 apart, where both forms return the larger operand)
    const float vs_f = (float)vshift;
    const bool vfloat = (double)vs_f == vshift;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    auto rows_loop = [&](auto vf_tag) {
        constexpr bool VF = decltype(vf_tag)::value;
        for (;;) {  // dynamic row queue, see nkf_dh_full_kernel
            if (threadIdx.x == 0) s_row = (long long)atomicAdd(&cnt[C_ROWQ_Y], 1ull);
            __syncthreads();
            const long long r = s_row;
            __syncthreads();
            if (r >= rows) break;
            const float* dh_r = dh + r * cols;
            const float* st_r = slope_tan + r * cols;
            const float* as_r = aspect + r * cols;
            unsigned char* bc_r = bin_cache + r * cols;
            float f1 = 0.f, f2 = 0.f;  // sum y, sum y^2 of this thread over the row (<= 64 pixels): float is plenty for p0
            unsigned fn = 0;
            for (int c0 = 0; c0 < (int)cols; c0 += 4 * NT) {  // uniform trip count: the body holds warp-wide intrinsics
                const int c = c0 + 4 * (int)threadIdx.x;
                const bool in = c < (int)cols;
                float4 d4 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F), s4 = d4;
                unsigned bins4 = 0xffffffffu;
                if (in) {
                    d4 = *reinterpret_cast<const float4*>(dh_r + c);
                    s4 = *reinterpret_cast<const float4*>(st_r + c);
                    if (reuse) {
                        bins4 = *reinterpret_cast<const unsigned*>(bc_r + c);
                    } else {
                        const float4 a4 = *reinterpret_cast<const float4*>(as_r + c);
                        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                        bins4 = 0u;
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            bins4 |= (isfinite(av[u]) ? (unsigned)aspect_bin(av[u], lo, hi, step, inv_step, n_bins) : 255u)
                                     << (8 * u);
                        *reinterpret_cast<unsigned*>(bc_r + c) = bins4;
                    }
                }
                const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
                unsigned key[4];
                unsigned tmask = 0u;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned b = (bins4 >> (8 * u)) & 255u;
                    float num;
                    if constexpr (VF) num = __fsub_rn(dv[u], vs_f);
                    else num = (float)((double)dv[u] - vshift);
                    const float yf = __fdiv_rn(num, sv[u]);
                    // a finite y implies a finite dh; bin 255 = non-finite aspect
                    const bool valid = (b != 255u) & (fabsf(yf) < CUDART_INF_F);
                    key[u] = ordered_key(yf);
                    const unsigned bb = valid ? b : (unsigned)MAXB;
                    uint2 lh;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lh.x), "=r"(lh.y) : "r"(lohi_addr + 8u * bb));
                    const bool below = key[u] < lh.x, above = key[u] > lh.y;
                    const unsigned slot = below ? bb : (above ? bb + (unsigned)(MAXB + 1) : (unsigned)MAXB);
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cnt_addr + 4u * slot) : "memory");
                    tmask |= (!below && !above) ? (1u << u) : 0u;
                    const float yv = valid ? yf : 0.f;
                    f1 += yv, f2 = fmaf(yv, yv, f2), fn += valid;
                }
                if (__any_sync(0xffffffffu, tmask != 0u)) {
                    // in-bracket pixels (~3 %): each takes a slot of the warp's staging area with a predicated
                    // shared-memory atomic (no warp scan: 20 instead of ~60 instructions per trip, and almost every
                    // trip has at least one such pixel among its 128)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const unsigned take = (tmask >> u) & 1u;
                        unsigned pos;
                        asm volatile(
                            "{ .reg .pred p; setp.ne.u32 p, %1, 0; mov.u32 %0, 0xffffffff;\n\t"
                            "@p atom.shared.add.u32 %0, [%2], 1; }"
                            : "=r"(pos)
                            : "r"(take), "r"(wn_addr)
                            : "memory");
                        const unsigned char b = (unsigned char)((bins4 >> (8 * u)) & 255u);
                        if (pos < (unsigned)WSTAGE) {
                            s_stage[warp][pos] = key[u], s_stage_g[warp][pos] = b;
                        } else if (take) {  // staging full (heavy ties): straight to the global buffer
                            const unsigned long long gi = atomicAdd(&cnt[C_BNC], 1ull);
                            if (gi < bcap) bkey[gi] = key[u], bgrp[gi] = b;
                        }
                    }
                }
            }
            m0 += (double)fn, m1 += (double)f1, m2 += (double)f2;
            __syncwarp();
            const unsigned n_st = min(s_wn[warp], (unsigned)WSTAGE);
            if (n_st) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&cnt[C_BNC], (unsigned long long)n_st);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (unsigned i = lane; i < n_st; i += 32)
                    if (base + i < bcap) bkey[base + i] = s_stage[warp][i], bgrp[base + i] = s_stage_g[warp][i];
            }
            __syncwarp();
            if (lane == 0) s_wn[warp] = 0u;
            __syncwarp();
        }
    };
    if (vfloat) rows_loop(std::true_type{});
    else rows_loop(std::false_type{});
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += NT) {
        if (s_cnt[MAXB + 1 + k]) atomicAdd(&cnt[C_BTOTAL + k], (unsigned long long)s_cnt[MAXB + 1 + k]);  // ABOVE counts
        if (s_cnt[k]) atomicAdd(&cnt[C_BBELOW + k], (unsigned long long)s_cnt[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o);
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    if (lane == 0 && m0 > 0.0) {
        atomicAdd(&f64[F_M0], m0);
        atomicAdd(&f64[F_M0 + 1], m1);
        atomicAdd(&f64[F_M0 + 2], m2);
    }
}

// Point-list fits (the reference's default subsample of 5e5 points): the arrays are small, so there is nothing to
// bracket -- the selection runs on ALL keys, entirely on the device.
__global__ void __launch_bounds__(NT) nkf_keys_points_kernel(const float* __restrict__ dh, long long n,
                                                              unsigned* __restrict__ key) {
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const float v = dh[i];
        key[i] = isfinite(v) ? ordered_key(v) : 0u;
    }
}
// (key of y, aspect bin) of every point + count / sum / sum of squares of y
__global__ void __launch_bounds__(NT)
nkf_y_points_kernel(const float* __restrict__ dh, const float* __restrict__ slope_tan, const float* __restrict__ aspect,
                    long long n, int n_bins, unsigned* __restrict__ skey, unsigned char* __restrict__ sgrp,
                    double* __restrict__ f64) {
    const double vshift = f64[F_VSHIFT], lo = f64[F_ASPLO], hi = f64[F_ASPHI];
    const double step = (hi - lo) / (double)n_bins;
    const double inv_step = step > 0.0 ? 1.0 / step : 0.0;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        const float a = aspect[i];
        unsigned kk = 0u;
        unsigned char g = 255;
        float yf = 0.f;
        if (isfinite(a) && y_key(dh[i], slope_tan[i], vshift, kk, yf)) {
            g = (unsigned char)aspect_bin(a, lo, hi, step, inv_step, n_bins);
            m0 += 1.0, m1 += (double)yf, m2 += (double)yf * (double)yf;
        } else {
            kk = 0u;
        }
        skey[i] = kk, sgrp[i] = g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 += __shfl_xor_sync(0xffffffffu, m0, o);
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    if ((threadIdx.x & 31) == 0 && m0 > 0.0) {
        atomicAdd(&f64[F_M0], m0);
        atomicAdd(&f64[F_M0 + 1], m1);
        atomicAdd(&f64[F_M0 + 2], m2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Radix select on a small buffer, two order statistics per group, ranks chosen on the device.
// ---------------------------------------------------------------------------------------------------------------
// One 8-bit digit histogram of the keys that still match their query's prefix.  The buffer is a sequence of `n_seg`
// segments of `seg_cap` slots of which the first min(seg_count[s], seg_cap) are filled (seg_count == NULL: all of them);
// key 0 marks an empty slot.
__global__ void __launch_bounds__(1024)
sel_hist_kernel(const unsigned* __restrict__ key, const unsigned char* __restrict__ grp, long long n_seg, long long seg_cap,
                const unsigned long long* __restrict__ seg_count, long long seg_count_stride, int G,
                const unsigned* __restrict__ prefix, unsigned mask, int shift, unsigned* __restrict__ hist) {
    extern __shared__ unsigned sh[];  // [G*2*256] counters | [G*2] prefixes
    const int n_cnt = G * 2 * 256;
    unsigned* s_pre = sh + n_cnt;
    for (int k = threadIdx.x; k < n_cnt; k += blockDim.x) sh[k] = 0u;
    for (int k = threadIdx.x; k < 2 * G; k += blockDim.x) s_pre[k] = prefix[k] & mask;
    __syncthreads();
    long long total = n_seg * seg_cap;
    if (seg_count && n_seg == 1) total = (long long)min((unsigned long long)seg_cap, seg_count[0]);
    auto count_one = [&](unsigned k, unsigned g) {
        if (!k || g >= (unsigned)G) return;
        const unsigned d = (k >> shift) & 255u, km = k & mask;
        // while both queries of a group still share their prefix only row 2g is counted (the pick kernel reads it for
        // both); pass 0 (mask == 0) matches every key whatever an earlier select left in `prefix`
        const unsigned p0 = s_pre[2 * g], p1 = s_pre[2 * g + 1];
        if (km == p0) atomicAdd(&sh[(2 * g) * 256 + d], 1u);
        if (km == p1 && p1 != p0) atomicAdd(&sh[(2 * g + 1) * 256 + d], 1u);
    };
    // four entries per thread and trip, two trips in flight (the scalar loop was bound by the latency of its dependent
    // loads: 60 us for 8 M entries, ncu r02)
    const bool vec = (reinterpret_cast<uintptr_t>(key) % 16 == 0) && (!grp || reinterpret_cast<uintptr_t>(grp) % 4 == 0) &&
                     (n_seg == 1 || seg_cap % 4 == 0);
    const long long nvec = vec ? total / 4 : 0;
    const uint4* key4 = reinterpret_cast<const uint4*>(key);
    const uchar4* grp4 = reinterpret_cast<const uchar4*>(grp);
#pragma unroll 2
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
        unsigned live = 4;
        if (seg_count && n_seg > 1) {
            const long long s = (4 * v) / seg_cap, j = 4 * v - s * seg_cap;
            const unsigned long long c = seg_count[s * seg_count_stride];
            live = (unsigned long long)j >= c ? 0u : (unsigned)min(4ull, c - (unsigned long long)j);
        }
        if (!live) continue;
        const uint4 k4 = key4[v];
        const uchar4 g4 = grp ? grp4[v] : make_uchar4(0, 0, 0, 0);
        count_one(k4.x, g4.x);
        if (live > 1) count_one(k4.y, g4.y);
        if (live > 2) count_one(k4.z, g4.z);
        if (live > 3) count_one(k4.w, g4.w);
    }
    for (long long i = 4 * nvec + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        if (seg_count && n_seg > 1) {
            const long long s = i / seg_cap, j = i - s * seg_cap;
            if ((unsigned long long)j >= seg_count[s * seg_count_stride]) continue;
        }
        count_one(key[i], grp ? (unsigned)grp[i] : 0u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_cnt; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

// mode 0 (pivots): ranks (m-1)/2 - margin and m/2 + margin of the m sample keys of the group, margin = 4 sqrt(m) + 8:
//   8 sigma of the rank the population median takes in a sample of independent draws, 4 sigma if the four pixels of a
//   sampled chunk were fully correlated; outputs the two keys.
// mode 1 (median): ranks (N-1)/2 - B and N/2 - B inside the compact buffer, N = ext_total[g] elements of the group in
//   the population, B = ext_below[g] of them below the bracket; outputs the mean of the two middle values (float64).
// mode 2: like 1, but ext_total[g] arrives as the count ABOVE the bracket: N = above + B + (entries of the group in the
//   buffer); N is written back to ext_total[g].
__global__ void __launch_bounds__(256)
sel_pick_kernel(int G, int mode, int pass, int last_pass, int shift, unsigned mask, unsigned* __restrict__ hist,
                unsigned* __restrict__ prefix, unsigned long long* __restrict__ below, long long* __restrict__ rank,
                unsigned long long* __restrict__ ext_total, const unsigned long long* __restrict__ ext_below,
                unsigned* __restrict__ out_lo, unsigned* __restrict__ out_hi, double* __restrict__ out_val,
                unsigned long long* __restrict__ flags, unsigned long long miss_bit) {
    // one warp per GROUP: it serves the group's two queries in turn (query 1 reads query 0's histogram while the two
    // prefixes coincide, see sel_hist_kernel)
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= G) return;
    const unsigned pm0 = prefix[2 * g] & mask, pm1 = prefix[2 * g + 1] & mask;
    const bool shared_row = pass == 0 || pm0 == pm1;
    unsigned keyq[2] = {0u, 0u};
    long long rkq[2] = {-1, -1};
    long long n_pop = -1;
    for (int q = 0; q < 2; ++q) {
        const int t = 2 * g + q;
        const unsigned* h = hist + (size_t)(shared_row ? 2 * g : t) * 256;
        unsigned c[8];
        unsigned long long part = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = h[lane * 8 + j], part += c[j];
        // inclusive scan of the lanes' partial sums
        unsigned long long incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned long long tot = __shfl_sync(0xffffffffu, incl, 31);
        long long rk;
        unsigned long long bel;
        unsigned pre;
        if (pass == 0) {
            bel = 0ull, pre = 0u;
            if (mode == 0) {
                if (tot == 0) {
                    rk = -1;
                } else {
                    const long long m = (long long)tot;
                    const long long margin = (long long)ceil(4.0 * sqrt((double)m)) + 8;
                    rk = q ? min(m - 1, m / 2 + margin) : max(0ll, (m - 1) / 2 - margin);
                }
            } else {
                // mode 2: ext_total holds the count ABOVE the bracket; the buffer's own total completes the population
                const long long B = (long long)ext_below[g];
                const long long N = mode == 2 ? (long long)ext_total[g] + B + (long long)tot : (long long)ext_total[g];
                n_pop = N;
                if (N == 0) {
                    rk = -1;
                } else {
                    rk = (q ? N / 2 : (N - 1) / 2) - B;
                    if (rk < 0 || rk >= (long long)tot) {
                        rk = -2;
                        if (lane == 0) atomicOr(flags, miss_bit);
                    }
                }
            }
        } else {
            rk = rank[t], bel = below[t], pre = prefix[t];
        }
        if (rk >= 0) {
            const unsigned long long want = (unsigned long long)rk - bel;
            const unsigned long long excl = incl - part;
            // the lane whose 8 counters contain the wanted rank
            const bool mine = excl <= want && want < incl;
            unsigned long long cum = excl;
            int digit = 0;
            if (mine) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum + c[j] > want) {
                        digit = lane * 8 + j;
                        break;
                    }
                    cum += c[j];
                }
            }
            const unsigned who = __ballot_sync(0xffffffffu, mine);
            const int src = who ? __ffs(who) - 1 : 31;
            digit = __shfl_sync(0xffffffffu, who ? digit : 255, src);
            cum = __shfl_sync(0xffffffffu, who ? cum : tot, src);
            bel += cum;
            pre |= (unsigned)digit << shift;
        }
        if (lane == 0) rank[t] = rk, below[t] = bel, prefix[t] = pre;
        keyq[q] = pre, rkq[q] = rk;
    }
    __syncwarp();
    if (mode == 2 && pass == 0 && lane == 0) ext_total[g] = (unsigned long long)n_pop;  // callers read the bin totals
    // zero both histogram rows for the next pass / the next select
#pragma unroll
    for (int j = 0; j < 16; ++j) hist[(size_t)(2 * g) * 256 + lane * 16 + j] = 0u;
    if (pass != last_pass || lane != 0) return;
    if (mode == 0) {
        // brackets need not be exact order statistics: after `last_pass` + 1 digits the undetermined low bits are set
        // to 0 (lo) / 1 (hi), which only widens the bracket by the digit bucket; an empty sample group gets the
        // all-inclusive bracket
        const unsigned low = shift ? ((1u << shift) - 1u) : 0u;
        out_lo[g] = rkq[0] >= 0 ? keyq[0] : 1u;
        out_hi[g] = rkq[1] >= 0 ? (keyq[1] | low) : 0xffffffffu;
    } else {
        out_val[g] = (rkq[0] >= 0 && rkq[1] >= 0)
                         ? 0.5 * ((double)key_to_float(keyq[0]) + (double)key_to_float(keyq[1]))
                         : CUDART_NAN;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One-time preparation of a fit: `_nuth_kaab_aux_vars` (np.gradient in NumPy's float32 op order, slope_tan, aspect,
// affine.py:412-474, 578-579 -- same expressions as xbn::nk_aux_kernel) fused with the validity mask of
// `_preprocess_rst_pts_subsample` (inlier & finite(ref, tba, slope_tan, aspect), base.py:653-661), its count, and the
// aspect range over the valid pixels.  17 B/px in one pass instead of the aux pass + nine elementwise torch passes.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
nk_prepare_kernel(const float* __restrict__ z, long long rows_buf, long long cols, long long ld, int top_is_border,
                  int bottom_is_border, long long row_begin, long long row_end, const float* __restrict__ tba,
                  long long tba_ld, const unsigned char* __restrict__ inlier, float* __restrict__ slope_tan,
                  float* __restrict__ aspect, unsigned char* __restrict__ sub_mask, unsigned long long* __restrict__ n_valid,
                  unsigned* __restrict__ asp_minmax) {
    unsigned lmin = 0xffffffffu, lmax = 0u;
    unsigned cnt = 0;
    for (long long r = row_begin + blockIdx.x; r < row_end; r += gridDim.x) {
        const float* zr = z + r * ld;
        const bool top = (r == 0) && top_is_border;
        const bool bot = (r == rows_buf - 1) && bottom_is_border;
        for (long long c = threadIdx.x; c < cols; c += NT) {
            float gx, gy;
            if (c == 0)
                gx = __fsub_rn(zr[1], zr[0]);
            else if (c == cols - 1)
                gx = __fsub_rn(zr[c], zr[c - 1]);
            else
                gx = __fmul_rn(__fsub_rn(zr[c + 1], zr[c - 1]), 0.5f);
            if (top)
                gy = __fsub_rn(zr[ld + c], zr[c]);
            else if (bot)
                gy = __fsub_rn(zr[c], zr[c - ld]);
            else
                gy = __fmul_rn(__fsub_rn(zr[ld + c], zr[c - ld]), 0.5f);
            float st = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
            const float asp = __fadd_rn(atan2f(-gx, gy), 3.14159274101257324f);
            if (fabsf(st) <= 1e-8f) st = CUDART_NAN_F;
            const long long o = (r - row_begin) * cols + c;
            slope_tan[o] = st;
            aspect[o] = asp;
            const bool ok = (!inlier || inlier[o]) && isfinite(zr[c]) && isfinite(tba[(r - row_begin) * tba_ld + c]) &&
                            isfinite(st) && isfinite(asp);
            sub_mask[o] = ok ? 1 : 0;
            if (ok) {
                const unsigned ab = __float_as_uint(asp);  // aspect >= 0: the bit pattern is monotonic
                lmin = min(lmin, ab), lmax = max(lmax, ab);
                ++cnt;
            }
        }
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMin(&asp_minmax[0], lmin);
        atomicMax(&asp_minmax[1], lmax);
        atomicAdd(n_valid, (unsigned long long)cnt);
    }
}

// The same, four pixels per thread (cols % 4 == 0, 16-byte aligned rows): vector loads of the three reference rows, the
// to-be-aligned row and the inlier bytes, vector stores of the three outputs.  Every warp also records its smallest /
// largest valid aspect with one pixel that attains it: rec[4 w .. 4 w + 3] = {min bits, position, max bits, position}
// (0xffffffff / 0 when the warp saw no valid pixel), from which nk_candidates_kernel assembles the range candidates
// without another pass over the aspect plane.
__global__ void __launch_bounds__(NT)
nk_prepare_vec4_kernel(const float* __restrict__ z, long long rows_buf, long long cols, long long ld, int top_is_border,
                       int bottom_is_border, long long row_begin, long long row_end, const float* __restrict__ tba,
                       long long tba_ld, const unsigned char* __restrict__ inlier, float* __restrict__ slope_tan,
                       float* __restrict__ aspect, unsigned char* __restrict__ sub_mask,
                       unsigned long long* __restrict__ n_valid, unsigned* __restrict__ rec,
                       unsigned* __restrict__ row_queue) {
    unsigned lmin = 0xffffffffu, lmax = 0u, pmin = 0u, pmax = 0u;
    unsigned cnt = 0;
    __shared__ long long s_row;
    for (;;) {  // dynamic row queue: the grid is exactly the resident CTAs
        if (threadIdx.x == 0) s_row = row_begin + (long long)atomicAdd(row_queue, 1u);
        __syncthreads();
        const long long r = s_row;
        __syncthreads();
        if (r >= row_end) break;
        const float* zr = z + r * ld;
        const bool top = (r == 0) && top_is_border;
        const bool bot = (r == rows_buf - 1) && bottom_is_border;
        for (long long c = 4ll * threadIdx.x; c < cols; c += 4ll * NT) {
            const float4 m4 = *reinterpret_cast<const float4*>(zr + c);
            const float4 u4 = top ? m4 : *reinterpret_cast<const float4*>(zr - ld + c);
            const float4 d4 = bot ? m4 : *reinterpret_cast<const float4*>(zr + ld + c);
            const float4 t4 = *reinterpret_cast<const float4*>(tba + (r - row_begin) * tba_ld + c);
            const long long o = (r - row_begin) * cols + c;
            const uchar4 i4 = inlier ? *reinterpret_cast<const uchar4*>(inlier + o) : make_uchar4(1, 1, 1, 1);
            const float w[6] = {c > 0 ? zr[c - 1] : 0.f, m4.x, m4.y, m4.z, m4.w, c + 4 < cols ? zr[c + 4] : 0.f};
            const float up[4] = {u4.x, u4.y, u4.z, u4.w}, dn[4] = {d4.x, d4.y, d4.z, d4.w};
            const float tb[4] = {t4.x, t4.y, t4.z, t4.w};
            const unsigned char in4[4] = {i4.x, i4.y, i4.z, i4.w};
            float st4[4], as4[4];
            unsigned char mk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long col = c + k;
                float gx, gy;
                if (col == 0)
                    gx = __fsub_rn(w[k + 2], w[k + 1]);
                else if (col == cols - 1)
                    gx = __fsub_rn(w[k + 1], w[k]);
                else
                    gx = __fmul_rn(__fsub_rn(w[k + 2], w[k]), 0.5f);
                if (top)
                    gy = __fsub_rn(dn[k], w[k + 1]);
                else if (bot)
                    gy = __fsub_rn(w[k + 1], up[k]);
                else
                    gy = __fmul_rn(__fsub_rn(dn[k], up[k]), 0.5f);
                float st = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
                const float asp = __fadd_rn(atan2f(-gx, gy), 3.14159274101257324f);
                if (fabsf(st) <= 1e-8f) st = CUDART_NAN_F;
                st4[k] = st, as4[k] = asp;
                const bool ok = in4[k] && isfinite(w[k + 1]) && isfinite(tb[k]) && isfinite(st) && isfinite(asp);
                mk[k] = ok ? 1 : 0;
                if (ok) {
                    const unsigned ab = __float_as_uint(asp);  // aspect >= 0: the bit pattern is monotonic
                    if (ab < lmin) lmin = ab, pmin = (unsigned)(o + k);
                    if (ab >= lmax) lmax = ab, pmax = (unsigned)(o + k);
                    ++cnt;
                }
            }
            *reinterpret_cast<float4*>(slope_tan + o) = make_float4(st4[0], st4[1], st4[2], st4[3]);
            *reinterpret_cast<float4*>(aspect + o) = make_float4(as4[0], as4[1], as4[2], as4[3]);
            *reinterpret_cast<uchar4*>(sub_mask + o) = make_uchar4(mk[0], mk[1], mk[2], mk[3]);
        }
    }
    const unsigned wmin = __reduce_min_sync(0xffffffffu, lmin), wmax = __reduce_max_sync(0xffffffffu, lmax);
    const unsigned total = __reduce_add_sync(0xffffffffu, cnt);
    const unsigned bmin = __ballot_sync(0xffffffffu, cnt && lmin == wmin), bmax = __ballot_sync(0xffffffffu, cnt && lmax == wmax);
    const unsigned qmin = __shfl_sync(0xffffffffu, pmin, bmin ? __ffs(bmin) - 1 : 0);
    const unsigned qmax = __shfl_sync(0xffffffffu, pmax, bmax ? __ffs(bmax) - 1 : 0);
    if ((threadIdx.x & 31) == 0) {
        unsigned* q = rec + 4ll * ((long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5));
        q[0] = total ? wmin : 0xffffffffu, q[1] = qmin, q[2] = total ? wmax : 0u, q[3] = qmax;
        if (total) atomicAdd(n_valid, (unsigned long long)total);
    }
}

// one CTA: global aspect range of the warps' records and up to RC_K positions attaining each end -> rc (layout below)
__global__ void __launch_bounds__(1024) nk_candidates_kernel(const unsigned* __restrict__ rec, int n_rec,
                                                              unsigned* __restrict__ rc) {
    __shared__ unsigned s_min, s_max, s_nmin, s_nmax;
    if (threadIdx.x == 0) s_min = 0xffffffffu, s_max = 0u, s_nmin = 0u, s_nmax = 0u;
    __syncthreads();
    unsigned lmin = 0xffffffffu, lmax = 0u;
    for (int i = threadIdx.x; i < n_rec; i += blockDim.x) lmin = min(lmin, rec[4 * i]), lmax = max(lmax, rec[4 * i + 2]);
    lmin = __reduce_min_sync(0xffffffffu, lmin), lmax = __reduce_max_sync(0xffffffffu, lmax);
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, lmin), atomicMax(&s_max, lmax);
    __syncthreads();
    const unsigned gmin = s_min, gmax = s_max;
    if (gmin != 0xffffffffu)  // at least one valid pixel
        for (int i = threadIdx.x; i < n_rec; i += blockDim.x) {
            if (rec[4 * i] == gmin) {
                const unsigned k = atomicAdd(&s_nmin, 1u);
                if (k < (unsigned)64) rc[4 + k] = rec[4 * i + 1];
            }
            if (rec[4 * i + 2] == gmax && rec[4 * i] != 0xffffffffu) {
                const unsigned k = atomicAdd(&s_nmax, 1u);
                if (k < (unsigned)64) rc[4 + 64 + k] = rec[4 * i + 3];
            }
        }
    __syncthreads();
    if (threadIdx.x == 0) rc[0] = gmin, rc[1] = gmax, rc[2] = min(s_nmin, 64u), rc[3] = min(s_nmax, 64u);
}

// Positions of (up to RC_K) valid pixels that attain the static aspect minimum / maximum.  The aspect range of an
// iteration is taken over the pixels with a finite dh -- a subset of the valid ones that only differs along the shifted
// raster edge -- so it equals the static range whenever one of these pixels still has a finite dh, which
// nkf_range_check_kernel verifies in O(RC_K) instead of re-reading the aspect plane in every dh pass.
// rc layout (uint32): [0] min bits, [1] max bits, [2] n_min, [3] n_max, [4 .. 4+K) min positions, [4+K .. 4+2K) max.
constexpr int RC_K = 64;
__global__ void __launch_bounds__(NT)
nk_extremes_kernel(const float* __restrict__ aspect, const unsigned char* __restrict__ sub_mask, long long n,
                   unsigned* __restrict__ rc) {
    const unsigned amin = rc[0], amax = rc[1];
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
        if (!sub_mask[i]) continue;
        const unsigned ab = __float_as_uint(aspect[i]);
        if (ab == amin && *reinterpret_cast<volatile unsigned*>(rc + 2) < (unsigned)RC_K) {
            const unsigned k = atomicAdd(&rc[2], 1u);
            if (k < (unsigned)RC_K) rc[4 + k] = (unsigned)i;
        }
        if (ab == amax && *reinterpret_cast<volatile unsigned*>(rc + 3) < (unsigned)RC_K) {
            const unsigned k = atomicAdd(&rc[3], 1u);
            if (k < (unsigned)RC_K) rc[4 + RC_K + k] = (unsigned)i;
        }
    }
}

// one warp: is the static aspect range still attained among the finite dh?  Otherwise cnt[C_RFALL] = 1 and
// nkf_range_full_kernel (always launched; returns at once when the flag is clear) reduces the range over the raster.
__global__ void nkf_range_check_kernel(const float* __restrict__ dh, const unsigned* __restrict__ rc,
                                       unsigned* __restrict__ keys, unsigned long long* __restrict__ cnt) {
    const int lane = threadIdx.x;
    const unsigned n_min = min(rc[2], (unsigned)RC_K), n_max = min(rc[3], (unsigned)RC_K);
    bool okmin = false, okmax = false;
    for (unsigned k = lane; k < n_min; k += 32) okmin |= isfinite(dh[rc[4 + k]]);
    for (unsigned k = lane; k < n_max; k += 32) okmax |= isfinite(dh[rc[4 + RC_K + k]]);
    okmin = __any_sync(0xffffffffu, okmin), okmax = __any_sync(0xffffffffu, okmax);
    if (lane == 0) {
        if (okmin && okmax) keys[K_ASPMIN] = rc[0], keys[K_ASPMAX] = rc[1];
        else cnt[C_RFALL] = 1ull;
    }
}
__global__ void __launch_bounds__(NT)
nkf_range_full_kernel(const float* __restrict__ dh, const float* __restrict__ aspect, long long n,
                      unsigned* __restrict__ keys, const unsigned long long* __restrict__ cnt) {
    if (cnt[C_RFALL] == 0ull) return;
    unsigned lmin = 0xffffffffu, lmax = 0u;
    bool any = false;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT)
        if (isfinite(dh[i])) {
            const unsigned ab = __float_as_uint(aspect[i]);
            lmin = min(lmin, ab), lmax = max(lmax, ab), any = true;
        }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    any = __any_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) {
        atomicMin(&keys[K_ASPMIN], lmin);
        atomicMax(&keys[K_ASPMAX], lmax);
    }
}

__global__ void nkf_reset_kernel(unsigned long long* cnt, unsigned* keys, double* f64, unsigned* hist, int n_hist) {
    for (int i = threadIdx.x; i < C_SIZE; i += blockDim.x) cnt[i] = 0ull;
    for (int i = threadIdx.x; i < K_SIZE; i += blockDim.x) keys[i] = (i == K_ASPMIN) ? 0xffffffffu : 0u;
    for (int i = threadIdx.x; i < F_SIZE; i += blockDim.x)
        if (i != F_CLO && i != F_CHI) f64[i] = (i >= F_MED) ? CUDART_NAN : 0.0;  // the bin-cache range survives
    for (int i = threadIdx.x; i < n_hist; i += blockDim.x) hist[i] = 0u;
}

__global__ void nkf_range_kernel(const unsigned* keys, double* f64) {
    f64[F_ASPLO] = (double)__uint_as_float(keys[K_ASPMIN]);
    f64[F_ASPHI] = (double)__uint_as_float(keys[K_ASPMAX]);
}

__global__ void nkf_finalize_kernel(unsigned long long* cnt, double* f64, unsigned long long gcap, unsigned long long bcap) {
    if (cnt[C_GNC] > gcap) atomicOr(&cnt[C_FLAGS], (unsigned long long)FLAG_GOVER);
    if (cnt[C_BNC] > bcap) atomicOr(&cnt[C_FLAGS], (unsigned long long)FLAG_BOVER);
    f64[F_CLO] = f64[F_ASPLO];
    f64[F_CHI] = f64[F_ASPHI];
}

// exactly the CTAs that are resident at once (for kernels that pull their rows from a queue)
template <class K>
static int grid_resident(K kernel, int threads, long long n_units) {
    int sms = 0, occ = 0;
    if (xb_num_sms(&sms)) sms = 148;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ < 1) occ = 1;
    const long long cap = (long long)sms * occ;
    return (int)(n_units < cap ? (n_units < 1 ? 1 : n_units) : cap);
}

static int grid_rows(long long n_rows, int per_sm) {
    int sms = 0;
    if (xb_num_sms(&sms)) sms = 148;
    const long long cap = (long long)sms * per_sm;
    return (int)(n_rows < cap ? (n_rows < 1 ? 1 : n_rows) : cap);
}

}  // namespace xbf

extern "C" {
#pragma GCC visibility push(default)

int xb_nk_prepare(const float* ref_dev, int64_t rows_buf, int64_t cols, int64_t ld, int top_is_border,
                  int bottom_is_border, int64_t row_begin, int64_t row_end, const float* tba_dev, int64_t tba_ld,
                  const uint8_t* inlier_dev, float* slope_tan_dev, float* aspect_dev, uint8_t* sub_mask_dev,
                  unsigned long long* n_valid_dev, uint32_t* range_cand_dev, void* stream) {
    if (!ref_dev || !tba_dev || !slope_tan_dev || !aspect_dev || !sub_mask_dev || !n_valid_dev || !range_cand_dev ||
        rows_buf < 2 || cols < 2 || ld < cols || tba_ld < cols || row_begin < 0 || row_end > rows_buf ||
        row_begin > row_end) {
        xb_set_error("bad arguments to xb_nk_prepare (np.gradient needs at least 2 rows and 2 columns)");
        return XB_ERR_INVALID;
    }
    if ((!top_is_border && row_begin < 1) || (!bottom_is_border && row_end > rows_buf - 1)) {
        xb_set_error("xb_nk_prepare: interior shards need one halo row above/below the output rows");
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const uint32_t init[4] = {0xffffffffu, 0u, 0u, 0u};
    XB_CUDA_CHECK(cudaMemcpyAsync(range_cand_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    XB_CUDA_CHECK(cudaMemsetAsync(n_valid_dev, 0, sizeof(unsigned long long), st));
    const long long rows = row_end - row_begin;
    if (rows == 0) return XB_OK;
    const bool vec4 = cols % 4 == 0 && ld % 4 == 0 && tba_ld % 4 == 0 && rows * cols < (1ll << 32) &&
                      ((reinterpret_cast<uintptr_t>(ref_dev) | reinterpret_cast<uintptr_t>(tba_dev) |
                        reinterpret_cast<uintptr_t>(slope_tan_dev) | reinterpret_cast<uintptr_t>(aspect_dev)) % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(sub_mask_dev) | reinterpret_cast<uintptr_t>(inlier_dev)) % 4 == 0);
    if (vec4) {
        const int grid = xbf::grid_resident(xbf::nk_prepare_vec4_kernel, xbf::NT, rows);
        const int n_rec = grid * (xbf::NT / 32);
        unsigned* rec = nullptr;  // per-warp records + the row queue counter
        XB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&rec), ((size_t)n_rec * 4 + 1) * sizeof(unsigned), st));
        XB_CUDA_CHECK(cudaMemsetAsync(rec + (size_t)n_rec * 4, 0, sizeof(unsigned), st));
        xbf::nk_prepare_vec4_kernel<<<grid, xbf::NT, 0, st>>>(ref_dev, rows_buf, cols, ld, top_is_border, bottom_is_border,
                                                              row_begin, row_end, tba_dev, tba_ld, inlier_dev,
                                                              slope_tan_dev, aspect_dev, sub_mask_dev, n_valid_dev, rec,
                                                              rec + (size_t)n_rec * 4);
        xbf::nk_candidates_kernel<<<1, 1024, 0, st>>>(rec, n_rec, range_cand_dev);
        XB_CUDA_CHECK(cudaGetLastError());
        XB_CUDA_CHECK(cudaFreeAsync(rec, st));
        xb_count_launch(2);
        return XB_OK;
    }
    xbf::nk_prepare_kernel<<<xbf::grid_rows(rows, 8), xbf::NT, 0, st>>>(
        ref_dev, rows_buf, cols, ld, top_is_border, bottom_is_border, row_begin, row_end, tba_dev, tba_ld, inlier_dev,
        slope_tan_dev, aspect_dev, sub_mask_dev, n_valid_dev, range_cand_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    if (rows * cols < (1ll << 32)) {
        xbf::nk_extremes_kernel<<<xbf::grid_rows((rows * cols + xbf::NT - 1) / xbf::NT, 8), xbf::NT, 0, st>>>(
            aspect_dev, sub_mask_dev, rows * cols, range_cand_dev);
        XB_CUDA_CHECK(cudaGetLastError());
        xb_count_launch(1);
    }
    return XB_OK;
}

int xb_nkf_layout(int32_t* out) {
    if (!out) return XB_ERR_INVALID;
    const int v[16] = {xbf::MAXB,    xbf::C_SIZE,  xbf::K_SIZE,  xbf::F_SIZE, xbf::C_NFIN,   xbf::C_GBELOW,
                       xbf::C_GNC,   xbf::C_BNC,   xbf::C_FLAGS, xbf::C_BTOTAL, xbf::C_BBELOW, xbf::K_ASPMIN,
                       xbf::K_GLO,   xbf::K_BLO,   xbf::F_VSHIFT, xbf::F_MED};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return XB_OK;
}

int xb_nkf_reset(unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev, void* stream) {
    if (!cnt_dev || !keys_dev || !f64_dev || !hist_dev) {
        xb_set_error("bad arguments to xb_nkf_reset");
        return XB_ERR_INVALID;
    }
    xbf::nkf_reset_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cnt_dev, keys_dev, f64_dev, hist_dev,
                                                                                 2 * xbf::MAXB * 256);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nkf_dh(int sample, const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev,
              const float* aspect_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
              int64_t tba_rows_total, double dx_px, double dy_px, float* dh_dev, uint32_t* sample_dev, int stride,
              uint32_t seed, unsigned long long* cnt_dev, uint32_t* keys_dev, uint32_t* gcompact_dev, uint64_t gcap,
              void* stream) {
    if (!ref_dev || !tba_dev || !sub_mask_dev || !cnt_dev || !keys_dev || rows <= 0 || cols < 4 ||
        cols % 4 || ld % 4 || ld < cols || tba_ld < cols || stride < 1 || (sample ? !sample_dev : (!dh_dev || !gcompact_dev)) ||
        ((reinterpret_cast<uintptr_t>(ref_dev) | reinterpret_cast<uintptr_t>(aspect_dev) |
          reinterpret_cast<uintptr_t>(dh_dev) | reinterpret_cast<uintptr_t>(sample_dev)) % 16) ||
        reinterpret_cast<uintptr_t>(sub_mask_dev) % 4) {
        xb_set_error("bad arguments to xb_nkf_dh (the fast path needs cols %% 4 == 0 and 16-byte aligned rasters)");
        return XB_ERR_INVALID;
    }
    if (!isfinite(dx_px) || !isfinite(dy_px) || fabs(dx_px) > 1e9 || fabs(dy_px) > 1e9) {
        xb_set_error("xb_nkf_dh: non-finite shift");
        return XB_ERR_INVALID;
    }
    const double fi = floor(dy_px), fj = floor(dx_px);
    const double fy = dy_px - fi, fx = dx_px - fj;
    xbf::DhArgs a;
    a.ref = ref_dev, a.tba = tba_dev, a.aspect = aspect_dev, a.sub_mask = sub_mask_dev;
    a.rows = rows, a.cols = cols, a.ld = ld, a.tba_ld = tba_ld, a.tba_row0 = tba_row0, a.tba_rows_total = tba_rows_total;
    a.i0 = (long long)fi, a.j0 = (long long)fj;
    a.w00 = (1.0 - fy) * (1.0 - fx), a.w01 = (1.0 - fy) * fx, a.w10 = fy * (1.0 - fx), a.w11 = fy * fx;
    const long long n_schunks = (rows * (cols / 4) + stride - 1) / stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (sample)
        xbf::nkf_dh_kernel<true><<<xbf::grid_resident(xbf::nkf_dh_kernel<true>, xbf::NT, (n_schunks + xbf::NT - 1) / xbf::NT), xbf::NT, 0, st>>>(
            a, nullptr, sample_dev, stride, seed, n_schunks, cnt_dev, keys_dev, nullptr, 0);
    else if (llabs(a.j0) < cols && cols < (1ll << 30) && tba_ld % 4 == 0 && reinterpret_cast<uintptr_t>(tba_dev) % 16 == 0 &&
             rows * cols / 8 < (1ll << 32)) {
        const int shift = (int)(((a.j0 % 4) + 4) % 4);
#define XB_DH_CASE(S)                                                                                                 \
    case S:                                                                                                           \
        if (aspect_dev)                                                                                               \
            xbf::nkf_dh_full_kernel<S, true><<<xbf::grid_resident(xbf::nkf_dh_full_kernel<S, true>, xbf::NT, rows),    \
                                               xbf::NT, 0, st>>>(a, dh_dev, cnt_dev, keys_dev, gcompact_dev, gcap);    \
        else                                                                                                          \
            xbf::nkf_dh_full_kernel<S, false><<<xbf::grid_resident(xbf::nkf_dh_full_kernel<S, false>, xbf::NT, rows),  \
                                                xbf::NT, 0, st>>>(a, dh_dev, cnt_dev, keys_dev, gcompact_dev, gcap);   \
        break;
        switch (shift) { XB_DH_CASE(0) XB_DH_CASE(1) XB_DH_CASE(2) XB_DH_CASE(3) }
#undef XB_DH_CASE
    } else
        xbf::nkf_dh_kernel<false><<<xbf::grid_resident(xbf::nkf_dh_kernel<false>, xbf::NT, rows), xbf::NT, 0, st>>>(a, dh_dev, nullptr, stride, seed,
                                                                               n_schunks, cnt_dev, keys_dev, gcompact_dev,
                                                                               gcap);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nkf_range(const uint32_t* keys_dev, double* f64_dev, void* stream) {
    if (!keys_dev || !f64_dev) return XB_ERR_INVALID;
    xbf::nkf_range_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(keys_dev, f64_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nkf_y(int sample, const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, uint8_t* bin_cache_dev,
             int64_t rows, int64_t cols, int n_bins, uint32_t* skey_dev, uint8_t* sgrp_dev, int stride, uint32_t seed,
             unsigned long long* cnt_dev, const uint32_t* keys_dev, double* f64_dev, uint32_t* bkey_dev,
             uint8_t* bgrp_dev, uint64_t bcap, void* stream) {
    if (!dh_dev || !slope_tan_dev || !aspect_dev || !bin_cache_dev || !cnt_dev || !keys_dev || !f64_dev || rows <= 0 ||
        cols < 4 || cols % 4 || n_bins < 1 || n_bins > xbf::MAXB || stride < 1 ||
        (sample ? (!skey_dev || !sgrp_dev) : (!bkey_dev || !bgrp_dev))) {
        xb_set_error("bad arguments to xb_nkf_y (1 <= n_bins <= %d)", xbf::MAXB);
        return XB_ERR_INVALID;
    }
    const long long n_schunks = (rows * (cols / 4) + stride - 1) / stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (sample)
        xbf::nkf_y_kernel<true><<<xbf::grid_resident(xbf::nkf_y_kernel<true>, xbf::NT, (n_schunks + xbf::NT - 1) / xbf::NT), xbf::NT, 0, st>>>(
            dh_dev, slope_tan_dev, aspect_dev, bin_cache_dev, rows, cols, n_bins, skey_dev, sgrp_dev, stride, seed,
            n_schunks, cnt_dev, keys_dev, f64_dev, nullptr, nullptr, 0);
    else
        xbf::nkf_y_kernel<false><<<xbf::grid_resident(xbf::nkf_y_kernel<false>, xbf::NT, rows), xbf::NT, 0, st>>>(
            dh_dev, slope_tan_dev, aspect_dev, bin_cache_dev, rows, cols, n_bins, nullptr, nullptr, stride, seed,
            n_schunks, cnt_dev, keys_dev, f64_dev, bkey_dev, bgrp_dev, bcap);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

int xb_nkf_select(const uint32_t* key_dev, const uint8_t* grp_dev, int64_t n_seg, int64_t seg_cap,
                  const unsigned long long* seg_count_dev, int64_t seg_count_stride, int n_groups, int mode,
                  unsigned long long* ext_total_dev, const unsigned long long* ext_below_dev, uint32_t* out_lo_dev,
                  uint32_t* out_hi_dev, double* out_val_dev, unsigned long long* flags_dev, uint64_t miss_bit,
                  uint32_t* hist_dev, uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev,
                  void* stream) {
    if (!key_dev || n_seg < 1 || seg_cap < 1 || n_groups < 1 || n_groups > xbf::MAXB || (mode < 0 || mode > 2) ||
        !hist_dev || !prefix_dev || !below_dev || !rank_dev || !flags_dev ||
        (mode == 0 ? (!out_lo_dev || !out_hi_dev) : (!out_val_dev || !ext_total_dev || !ext_below_dev))) {
        xb_set_error("bad arguments to xb_nkf_select");
        return XB_ERR_INVALID;
    }
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)n_groups * 2 * 257 * sizeof(unsigned);
    XB_CUDA_CHECK(cudaFuncSetAttribute(xbf::sel_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long total = n_seg * seg_cap;
    long long grid = (total + 4095) / 4096;
    const long long cap = (long long)sms * (smem > 100 * 1024 ? 1 : 2);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    unsigned mask = 0u;
    const int n_pass = mode == 0 ? 3 : 4;  // brackets: 24 key bits are plenty; medians: all 32
    for (int pass = 0; pass < n_pass; ++pass) {
        const int shift = 24 - 8 * pass;
        xbf::sel_hist_kernel<<<(unsigned)grid, 1024, smem, st>>>(key_dev, grp_dev, n_seg, seg_cap, seg_count_dev,
                                                                 seg_count_stride, n_groups, prefix_dev, mask, shift,
                                                                 hist_dev);
        XB_CUDA_CHECK(cudaGetLastError());
        xbf::sel_pick_kernel<<<(n_groups + 7) / 8, 256, 0, st>>>(n_groups, mode, pass, n_pass - 1, shift, mask, hist_dev, prefix_dev,
                                                                 below_dev, rank_dev, ext_total_dev, ext_below_dev,
                                                                 out_lo_dev, out_hi_dev, out_val_dev, flags_dev,
                                                                 (unsigned long long)miss_bit);
        XB_CUDA_CHECK(cudaGetLastError());
        mask |= 255u << shift;
    }
    xb_count_launch(2 * n_pass);
    return XB_OK;
}

int xb_nkf_finalize(unsigned long long* cnt_dev, double* f64_dev, uint64_t gcap, uint64_t bcap, void* stream) {
    if (!cnt_dev || !f64_dev) return XB_ERR_INVALID;
    xbf::nkf_finalize_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cnt_dev, f64_dev, gcap, bcap);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(1);
    return XB_OK;
}

/* One whole iteration on one GPU: the calls above in order (reset, dh sample, global bracket, dh pass, range, median of
 * dh, y sample, per-bin brackets, y pass, per-bin medians, finalize) -- one C call instead of twelve. */
int xb_nkf_iteration(const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev, const float* slope_tan_dev,
                     const float* aspect_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
                     int64_t tba_rows_total, double dx_px, double dy_px, int n_bins, float* dh_dev, uint8_t* bin_cache_dev,
                     uint32_t* sample_dev, uint8_t* sgrp_dev, int64_t ns, int stride, uint32_t seed,
                     uint32_t* gcompact_dev, uint64_t gcap, uint32_t* bkey_dev, uint8_t* bgrp_dev, uint64_t bcap,
                     unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev,
                     uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev,
                     const uint32_t* range_cand_dev, void* stream) {
    using namespace xbf;
    int rc = xb_nkf_reset(cnt_dev, keys_dev, f64_dev, hist_dev, stream);
    if (rc) return rc;
    rc = xb_nkf_dh(1, ref_dev, tba_dev, sub_mask_dev, aspect_dev, rows, cols, ld, tba_ld, tba_row0, tba_rows_total, dx_px,
                   dy_px, dh_dev, sample_dev, stride, seed, cnt_dev, keys_dev, gcompact_dev, gcap, stream);
    if (rc) return rc;
    rc = xb_nkf_select(sample_dev, nullptr, 1, ns, nullptr, 1, 1, 0, nullptr, nullptr, keys_dev + K_GLO, keys_dev + K_GHI,
                       nullptr, cnt_dev + C_FLAGS, FLAG_GMISS, hist_dev, prefix_dev, below_dev, rank_dev, stream);
    if (rc) return rc;
    const bool rcheck = range_cand_dev != nullptr && rows * cols < (1ll << 32);
    rc = xb_nkf_dh(0, ref_dev, tba_dev, sub_mask_dev, rcheck ? nullptr : aspect_dev, rows, cols, ld, tba_ld, tba_row0,
                   tba_rows_total, dx_px, dy_px, dh_dev, sample_dev, stride, seed, cnt_dev, keys_dev, gcompact_dev, gcap,
                   stream);
    if (rc) return rc;
    if (rcheck) {
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        nkf_range_check_kernel<<<1, 32, 0, st>>>(dh_dev, range_cand_dev, keys_dev, cnt_dev);
        nkf_range_full_kernel<<<grid_rows(rows, 4), NT, 0, st>>>(dh_dev, aspect_dev, rows * cols, keys_dev, cnt_dev);
        XB_CUDA_CHECK(cudaGetLastError());
        xb_count_launch(2);
    }
    rc = xb_nkf_range(keys_dev, f64_dev, stream);
    if (rc) return rc;
    rc = xb_nkf_select(gcompact_dev, nullptr, 1, (int64_t)gcap, cnt_dev + C_GNC, 1, 1, 1, cnt_dev + C_NFIN,
                       cnt_dev + C_GBELOW, nullptr, nullptr, f64_dev + F_VSHIFT, cnt_dev + C_FLAGS, FLAG_GMISS, hist_dev,
                       prefix_dev, below_dev, rank_dev, stream);
    if (rc) return rc;
    rc = xb_nkf_y(1, dh_dev, slope_tan_dev, aspect_dev, bin_cache_dev, rows, cols, n_bins, sample_dev, sgrp_dev, stride,
                  seed ^ 0x5BD1E995u, cnt_dev, keys_dev, f64_dev, bkey_dev, bgrp_dev, bcap, stream);
    if (rc) return rc;
    rc = xb_nkf_select(sample_dev, sgrp_dev, 1, ns, nullptr, 1, n_bins, 0, nullptr, nullptr, keys_dev + K_BLO,
                       keys_dev + K_BHI, nullptr, cnt_dev + C_FLAGS, FLAG_BMISS, hist_dev, prefix_dev, below_dev, rank_dev,
                       stream);
    if (rc) return rc;
    rc = xb_nkf_y(0, dh_dev, slope_tan_dev, aspect_dev, bin_cache_dev, rows, cols, n_bins, sample_dev, sgrp_dev, stride,
                  seed ^ 0x5BD1E995u, cnt_dev, keys_dev, f64_dev, bkey_dev, bgrp_dev, bcap, stream);
    if (rc) return rc;
    rc = xb_nkf_select(bkey_dev, bgrp_dev, 1, (int64_t)bcap, cnt_dev + C_BNC, 1, n_bins, 2, cnt_dev + C_BTOTAL,
                       cnt_dev + C_BBELOW, nullptr, nullptr, f64_dev + F_MED, cnt_dev + C_FLAGS, FLAG_BMISS, hist_dev,
                       prefix_dev, below_dev, rank_dev, stream);
    if (rc) return rc;
    return xb_nkf_finalize(cnt_dev, f64_dev, gcap, bcap, stream);
}

/* One whole iteration of a point-list fit (xdem_b200.coreg._NKState.set_points): dh at the points, its exact median, the
 * aspect range, (key, bin) of y at every point with the moments, the exact per-bin medians and counts -- all on the
 * device, same result block as xb_nkf_iteration (cnt / f64), one host read at the end. */
int xb_nkf_iteration_points(const float* ref_dev, const float* tba_dev, const int64_t* idx_dev, int64_t n_pts,
                            const float* slope_tan_pts_dev, const float* aspect_pts_dev, int64_t rows, int64_t cols,
                            int64_t ld, int64_t tba_ld, int64_t tba_row0, int64_t tba_rows_total, double dx_px,
                            double dy_px, int n_bins, float* dh_pts_dev, uint32_t* key_dev, uint8_t* grp_dev,
                            unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev,
                            uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev, void* stream) {
    using namespace xbf;
    if (!slope_tan_pts_dev || !key_dev || !grp_dev || n_pts <= 0 || n_bins < 1 || n_bins > MAXB) {
        xb_set_error("bad arguments to xb_nkf_iteration_points (1 <= n_bins <= %d)", MAXB);
        return XB_ERR_INVALID;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = xb_nkf_reset(cnt_dev, keys_dev, f64_dev, hist_dev, stream);
    if (rc) return rc;
    rc = xb_nk_dh_points(ref_dev, tba_dev, idx_dev, n_pts, aspect_pts_dev, rows, cols, ld, tba_ld, tba_row0, tba_rows_total,
                         dx_px, dy_px, dh_pts_dev, keys_dev + K_ASPMIN, cnt_dev + C_NFIN, stream);
    if (rc) return rc;
    const int grid = grid_rows((n_pts + NT - 1) / NT, 8);
    nkf_keys_points_kernel<<<grid, NT, 0, st>>>(dh_pts_dev, n_pts, key_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    // mode 2 with zero "above" / "below" counters: the population is exactly the buffer
    rc = xb_nkf_select(key_dev, nullptr, 1, n_pts, nullptr, 1, 1, 2, cnt_dev + C_GNC, cnt_dev + C_GBELOW, nullptr, nullptr,
                       f64_dev + F_VSHIFT, cnt_dev + C_FLAGS, FLAG_GMISS, hist_dev, prefix_dev, below_dev, rank_dev, stream);
    if (rc) return rc;
    rc = xb_nkf_range(keys_dev, f64_dev, stream);
    if (rc) return rc;
    nkf_y_points_kernel<<<grid, NT, 0, st>>>(dh_pts_dev, slope_tan_pts_dev, aspect_pts_dev, n_pts, n_bins, key_dev, grp_dev,
                                             f64_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    xb_count_launch(2);
    rc = xb_nkf_select(key_dev, grp_dev, 1, n_pts, nullptr, 1, n_bins, 2, cnt_dev + C_BTOTAL, cnt_dev + C_BBELOW, nullptr,
                       nullptr, f64_dev + F_MED, cnt_dev + C_FLAGS, FLAG_BMISS, hist_dev, prefix_dev, below_dev, rank_dev,
                       stream);
    if (rc) return rc;
    return xb_nkf_finalize(cnt_dev, f64_dev, ~0ull, ~0ull, stream);
}

#pragma GCC visibility pop
}
