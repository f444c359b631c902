/* xdem_b200 -- C ABI of the B200-native hot path of GlacioHack/xdem (libxdem_b200.so).
 *
 * The reference is pure Python and has no FFI: its plugin seam is the `engine=` keyword dispatched at
 * xdem/terrain/surfit.py:1249/1270 and xdem/terrain/window.py:968/980.  Each entry point below replaces the array-level
 * contract of one reference function (cited per function).  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add (an `engine="b200"` branch next to "scipy"/"numba").
 *
 * Conventions: every pointer named `*_dev` / documented "device" is a CUDA device pointer (e.g. tensor.data_ptr());
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); rasters are C-contiguous row-major,
 * row 0 = north, invalid cells = NaN; every function returns 0 on success or a negative XB_ERR_* code and
 * xb_last_error() returns a thread-local message.  Kernels never raise on data: undefined results are NaN.
 */
#ifndef XDEM_B200_H
#define XDEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XB_OK 0
#define XB_ERR_INVALID (-1)
#define XB_ERR_CUDA (-2)
#define XB_ERR_UNSUPPORTED (-3)

/* dtype codes */
#define XB_F32 0
#define XB_F64 1

/* surface fit ids -- surfit.py:1240 */
#define XB_FIT_HORN 0
#define XB_FIT_ZEVENBERG_THORNE 1
#define XB_FIT_FLORINSKY 2

/* surface attribute bits / plane slots 0..9 -- order of surfit.py:407-418 */
#define XB_SURF_SLOPE 0
#define XB_SURF_ASPECT 1
#define XB_SURF_HILLSHADE 2
#define XB_SURF_CURVATURE 3
#define XB_SURF_PROFILE_CURVATURE 4
#define XB_SURF_TANGENTIAL_CURVATURE 5
#define XB_SURF_PLANFORM_CURVATURE 6
#define XB_SURF_FLOWLINE_CURVATURE 7
#define XB_SURF_MAX_CURVATURE 8
#define XB_SURF_MIN_CURVATURE 9
/* windowed index bits (plane slots 10..13) -- order of window.py:752-758 */
#define XB_WIN_TPI 0
#define XB_WIN_TRI 1
#define XB_WIN_ROUGHNESS 2
#define XB_WIN_RUGOSITY 3
#define XB_N_PLANES 14

const char* xb_last_error(void);
int xb_version(void);
/* Number of kernels this library launched since load (all entry points); bench.py reports it as gpu_launches. */
uint64_t xb_launch_count(void);
/* Tuning / test knobs.  "florinsky_generic" = 1 routes Florinsky requests through the generic fused kernel instead of
 * the row-feature-reuse kernel (both are parity-tested; used for A/B checks); "window3_generic" = 1 does the same for the
 * 3x3 windowed indexes (generic fused kernel instead of xb_terrain_w3.cu).  "variogram_full_tiles" = bit mask (default
 * 7): bit k-1 lets interior variogram tiles that span exactly k lag classes use the threshold-light sweep. */
int xb_set_option(const char* name, int value);

/* Diagnostics: the trivial streaming kernel with the terrain engine's traffic mix (reads n_floats float32 from src,
 * writes n_planes copies to dst[n_planes * n_floats] with streaming vector stores).  bench.py times it to report the
 * bandwidth that mix can reach on the box next to the roofline of the real kernel.  Not counted in xb_launch_count. */
int xb_probe_stream(const void* src_dev, void* dst_dev, int64_t n_floats, int n_planes, void* stream);
/* Diagnostics: the same traffic mix with the planes written by TMA bulk tensor stores (cp.async.bulk.tensor, UTMASTG) from
 * shared-memory staged 8 x 128 tiles instead of st.global.cs -- measures whether the store instruction or the read/write
 * mix sets the bandwidth ceiling of the terrain kernels.  dst holds n_planes planes of rows x cols float32. */
int xb_probe_stream_tma(const void* src_dev, void* dst_dev, int64_t rows, int64_t cols, int n_planes, void* stream);
/* Diagnostics: bit-exactness of the branch-free IEEE cores of the 3x3 windowed kernel against the CUDA round-to-nearest
 * intrinsics, over `count` consecutive float32 bit patterns x starting at bits_begin.  kind 0: the fast-path square root
 * vs __fsqrt_rn(x) (valid range [2^-101, FLT_MAX]); kind 1: the reciprocal-multiply division x / b (rcp_b = RN(1/b)) vs
 * __fdiv_rn(x, b).  The number of differing results is added to mismatches_dev[0]. */
int xb_probe_exact_math(int kind, uint32_t bits_begin, uint64_t count, float b, float rcp_b,
                        unsigned long long* mismatches_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Terrain stencil engine.
 * Replaces `_get_surface_attributes` (surfit.py:1197-1305) and `_get_windowed_indexes` (window.py:926-1002) -- and,
 * with degrees/clip set, the post-processing of `_get_terrain_attribute` (terrain.py:586-596) -- in ONE fused pass.
 *
 *  dem_dev      device raster buffer, `rows_buf` x `cols`, leading dimension `ld` elements, dtype XB_F32 / XB_F64.
 *               Rows outside [0,rows_buf) and columns outside [0,cols) are treated as NaN (reference: cval=nan,
 *               spatialstats.py:2524, window.py:111-112; NaN padding surfit.py:1278-1282, window.py:986).
 *  row_begin/row_end  output rows [row_begin,row_end) of the buffer are computed (row-sharding: the buffer carries
 *               `depth` halo rows from the neighbouring shards, terrain.py:417-432); out planes have
 *               (row_end-row_begin) rows x cols, leading dimension out_ld, same dtype as the DEM.
 *  surf_mask    bit i set = compute surface attribute i into out_planes_host[i]
 *  win_mask     bit j set = compute windowed index j into out_planes_host[10+j]; window_size in {3,5} (rugosity: 3)
 *  degrees      !=0: slope/aspect in degrees (np.rad2deg in the array dtype, terrain.py:591)
 *  clip_hillshade !=0: clip hillshade to [0,255] (terrain.py:596)
 *  out_planes_host  HOST array of XB_N_PLANES device pointers (NULL for planes not requested)
 */
int xb_terrain_fused(const void* dem_dev, int dtype, int64_t rows_buf, int64_t cols, int64_t ld, int64_t row_begin,
                     int64_t row_end, double resolution, int fit_id, int curv_method_id, uint32_t surf_mask,
                     uint32_t win_mask, int window_size, int tri_method_id, int degrees, int clip_hillshade,
                     double hillshade_azimuth, double hillshade_altitude, double hillshade_z_factor,
                     void* const* out_planes_host, int64_t out_ld, void* stream);

/* Same computation from/to HOST buffers: the raster is streamed through the GPU in row blocks with `depth` halo rows
 * (the reference's tiling analogue: geoutils.map_overlap_multiproc_save, terrain.py:412-466), H2D / kernel / D2H
 * overlapped on three streams.  This is the call the reference-facing seam and the `e2e` benchmark figure go through.
 * out_planes_host[i] are HOST pointers here; rows_per_block <= 0 picks ~96 MiB blocks.  Every buffer may be pinned
 * (DMA straight from / to it) or pageable (staged block-wise through pinned scratch by XDEM_B200_HOST_THREADS host
 * threads, default 4-8); the library detects which with cudaPointerGetAttributes. */
int xb_terrain_fused_host(const void* dem_host, int dtype, int64_t rows, int64_t cols, double resolution, int fit_id,
                          int curv_method_id, uint32_t surf_mask, uint32_t win_mask, int window_size,
                          int tri_method_id, int degrees, int clip_hillshade, double hillshade_azimuth,
                          double hillshade_altitude, double hillshade_z_factor, void* const* out_planes_host,
                          int64_t rows_per_block);
/* Row-range form: the host raster holds `rows` rows of which only [row_begin, row_end) are computed -- a row shard that
 * carries `depth` halo rows of its neighbours (multi-GPU streaming of one raster, terrain.py:417-432); the planes have
 * (row_end - row_begin) rows.  Rows outside [0, rows) count as NaN (raster border). */
int xb_terrain_fused_host_rows(const void* dem_host, int dtype, int64_t rows, int64_t cols, int64_t row_begin,
                               int64_t row_end, double resolution, int fit_id, int curv_method_id, uint32_t surf_mask,
                               uint32_t win_mask, int window_size, int tri_method_id, int degrees, int clip_hillshade,
                               double hillshade_azimuth, double hillshade_altitude, double hillshade_z_factor,
                               void* const* out_planes_host, int64_t rows_per_block);
/* Frees the device / pinned scratch the host-buffer path keeps between calls (3 slots of ~96 MiB x (1 + planes)). */
int xb_release_scratch(void);

/* ---------------------------------------------------------------------------------------------------------------
 * N-D binned robust statistics -- the device side of `nd_binning` (xdem/spatialstats.py:91-216), i.e. of the
 * `scipy.stats.binned_statistic[_2d|_dd]` calls it makes with `statistic = "count" / np.nanmedian / nmad`.
 *   xb_bin_keys         bin number (np.digitize per variable against its edges, last edge closed with SciPy's rounding
 *                       rule p10 = 10^decimal, out-of-range or non-finite samples -> 0xFFFF; C-order flattening) and the
 *                       order-preserving 32-bit key of every value.  1..3 variables; vars_dev_host / n_edges_host /
 *                       p10_host are HOST arrays of n_dims entries; edges_dev = the edge arrays concatenated (float64
 *                       copies of the sample-dtype edges SciPy would build).
 *   xb_bin_hist         histogram (n_bins x 256 u64) of one 8-bit digit of the keys whose masked bits equal their bin's
 *                       prefix: one pass of the MSD radix select that finds every bin's exact median.
 *   xb_bin_next         per bin the smallest key > sel[bin] (next_key_dev pre-set to 0xFFFFFFFF): upper middle value.
 *   xb_bin_absdev_keys  keys of |value - center[bin]| in float32 arithmetic: the NMAD's second select.
 */
int xb_bin_keys(const float* values_dev, const float* const* vars_dev_host, int n_dims, int64_t n,
                const double* edges_dev, const int32_t* n_edges_host, const double* p10_host, uint32_t* key_dev,
                uint16_t* bin_dev, void* stream);
int xb_bin_hist(const uint32_t* key_dev, const uint16_t* bin_dev, int64_t n, int n_bins, const uint32_t* prefix_dev,
                uint32_t prefix_mask, int shift, unsigned long long* hist_dev, void* stream);
int xb_bin_next(const uint32_t* key_dev, const uint16_t* bin_dev, int64_t n, int n_bins, const uint32_t* sel_dev,
                uint32_t* next_key_dev, void* stream);
int xb_bin_absdev_keys(const float* values_dev, const uint16_t* bin_dev, int64_t n, int n_bins,
                       const float* center_dev, uint32_t* key_dev, void* stream);

/* Applying a 1-D binned correction -- the device side of `BiasCorr._apply_rst` for one bias variable
 * (biascorr.py:259-310; TerrainBias: the variable is a terrain attribute): out = float32(elev + corr(var)).
 * mode 0 "linear": `interp_nd_binning` in one dimension (spatialstats.py:237-423): piecewise-linear between the m
 * mid-points x of the valid bins with statistics v, held constant outside.  mode 1 "per_bin": `get_perbin_nd_binning`
 * (spatialstats.py:425-530): x = the m + 1 bin edges, v = the m statistics, the bin [left, right) holding var; NaN
 * outside.  x_dev / v_dev are small device tables of float64. */
int xb_bin_apply_1d(const float* elev_dev, const float* var_dev, int64_t n, const double* x_dev, const double* v_dev,
                    int m, int mode, float* out_dev, void* stream);

/* Per-bin sum / sum of squares (float64) and minimum / maximum (order-preserving uint32 keys of the float32 values) of
 * the samples xb_bin_keys assigned to bins: the remaining built-in statistics of scipy.stats.binned_statistic ('mean',
 * 'std', 'sum', 'min', 'max') that nd_binning accepts (spatialstats.py:147-149).  The caller zero-fills sum / sumsq and
 * sets minkey to 0xffffffff, maxkey to 0. */
int xb_bin_moments(const float* values_dev, const uint16_t* bin_dev, int64_t n, int n_bins, double* sum_dev,
                   double* sumsq_dev, uint32_t* minkey_dev, uint32_t* maxkey_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Texture shading (fractional Laplacian, Brown 2010) -- the device stages of `_texture_shading_fft`
 * (xdem/terrain/freq.py:62-148; routed at terrain.py:641-643).  The two FFTs in between are plain library transforms
 * issued by the caller (cuFFT through torch / CuPy; scipy.fft in the reference).
 *   xb_texture_prepare  freq.py:84-112: stats_dev[0..2] = {sum, count} over non-NaN cells (np.nanmean) and the count of
 *                       finite cells; padded[fft_rows x fft_cols] = symmetric pad of the DEM with non-finite cells
 *                       replaced by that mean.  subtract_mean = 1 centres the raster on the mean first (same result for
 *                       alpha > 0, whose filter zeroes the DC term; keeps float32 digits for the relief).  The caller
 *                       checks stats[2] == 0 (all-NaN input -> all-NaN output).
 *   xb_texture_filter   freq.py:114-137: in-place scale of the rfft2 half spectrum [fft_rows x (fft_cols/2+1)] complex
 *                       by hypot(rfftfreq, fftfreq)^alpha evaluated in float64; DC term zeroed when alpha > 0.
 *   xb_texture_finish   freq.py:142-146: crop the irfft2 result and restore NaN where the DEM was not finite.
 * dtype 0 = float32 (complex64 spectrum), 1 = float64 (complex128).
 */
int xb_texture_prepare(const void* dem_dev, int dtype, int64_t rows, int64_t cols, int64_t ld, void* padded_dev,
                       int64_t fft_rows, int64_t fft_cols, int64_t pad_rows, int64_t pad_cols, int subtract_mean,
                       double* stats_dev, void* stream);
int xb_texture_filter(void* spectrum_dev, int dtype, int64_t fft_rows, int64_t fft_cols, double alpha, void* stream);
int xb_texture_finish(const void* padded_dev, int dtype, int64_t fft_rows, int64_t fft_cols, int64_t pad_rows,
                      int64_t pad_cols, const void* dem_dev, int64_t rows, int64_t cols, int64_t ld, void* out_dev,
                      int64_t out_ld, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Generic odd window (3..31) indexes and fractal roughness -- the reference's arbitrary `window_size`
 * (window.py:926-1002) and `fractal_roughness` (window.py:317-379, terrain.py:620-633).
 *  win_mask   bit 0 TPI, bit 1 TRI, bit 2 roughness, bit 4 fractal roughness (plane slots 0, 1, 2, 4 of a HOST array of
 *             5 device pointers); same NaN rule and buffer / row-range conventions as xb_terrain_fused.
 */
int xb_windowed_generic(const void* dem_dev, int dtype, int64_t rows_buf, int64_t cols, int64_t ld, int64_t row_begin,
                        int64_t row_end, int window_size, uint32_t win_mask, int tri_method_id,
                        void* const* out_planes_host, int64_t out_ld, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Empirical variogram: all-pairs distance / squared-difference lag binning.
 * Replaces what skgstat.Variogram (third-party; scikit-gstat>=1.0.18, setup.cfg:56) computes for
 * xdem.spatialstats._get_pdist_empirical_variogram (spatialstats.py:1064-1101): for every sample pair i<j the lag class
 * k with edge2[k-1] <= d2 < edge2[k] (d2 = squared pixel distance; edge2 = integer thresholds the host derives from the
 * float64 right bin edges), count[k] += 1, sumsq[k] += (v_i - v_j)^2 (Matheron numerator).  Pairs with
 * d2 >= edge2[n_bins-1] are dropped (skgstat: distances beyond maxlag).
 *
 *  pts_dev      int32 [n_groups*G][4] = (col, row, float32 value bits, sorted index or -1 for padding); samples sorted
 *               along a space-filling curve and cut into groups of G = xb_variogram_group_size() samples
 *  gbox_dev     int32 [n_groups][4] bounding box (xmin, ymin, xmax, ymax) of the valid samples of each group
 *  edge2_dev    uint64 [n_bins] ascending thresholds
 *  unit_prefix_dev  int64 [n_groups+1]: work units are (group i, chunk of xb_variogram_chunk() j-groups >= i);
 *               unit_prefix[i] = number of units of rows < i.  [unit_begin, unit_end) is this call's share of the units
 *               (multi-GPU: ranks take disjoint ranges and all-reduce count/sumsq, 2*n_bins*8 bytes)
 *  wide         0: d2 fits 32 bits (raster diagonal^2 < 2^32), 1: 64-bit distances
 *  estimator    0: sumsq accumulates (v_i - v_j)^2 (Matheron); 1: sumsq accumulates |v_i - v_j|^0.5 (Cressie-Hawkins)
 *  count_dev    uint64 [n_bins] and sumsq_dev double [n_bins], accumulated into (caller zeroes them)
 */
int xb_variogram_group_size(void);
int xb_variogram_chunk(void);
int xb_variogram_pairs(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                       const unsigned long long* edge2_dev, int n_bins, const int64_t* unit_prefix_dev,
                       int64_t unit_begin, int64_t unit_end, int wide, int estimator, unsigned long long* count_dev,
                       double* sumsq_dev, void* stream);
/* One radix-select pass for Dowd's estimator (per-class median of |v_i - v_j|, skgstat.estimators.dowd): keys are the
 * float32 bit patterns of |diff|.  mode 0: for pairs of class k with (key & prefix_mask) == prefix_dev[k]:
 * hist_dev[k*256 + ((key >> shift) & 255)] += 1.  mode 1: next_key_dev[k] = min(next_key_dev[k], smallest key >
 * prefix_dev[k]).  Same sample / unit conventions as xb_variogram_pairs. */
int xb_variogram_median_pass(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                              const unsigned long long* edge2_dev, int n_bins, const int64_t* unit_prefix_dev,
                              int64_t unit_begin, int64_t unit_end, int mode, const uint32_t* prefix_dev,
                              uint32_t prefix_mask, int shift, unsigned long long* hist_dev, uint32_t* next_key_dev,
                              void* stream);
/* Largest squared pixel distance over all sample pairs, accumulated with max into maxd2_dev[0] (skgstat's "even"
 * binning clips maxlag to the largest sampled distance). */
int xb_variogram_maxd2(const int32_t* pts_dev, const int32_t* gbox_dev, int64_t n_groups,
                       unsigned long long* maxd2_dev, void* stream);

/* General-coordinate form of the pair binning (float64 coordinates and values): every pair i < j of one sample set
 * (xb_dev == NULL; `_get_pdist_empirical_variogram`, spatialstats.py:1064-1101, incl. 1-D values + `coords=`) or every
 * pair between two sets A x B (`_get_cdist_empirical_variogram`, spatialstats.py:1186-1261: centre-disk x ring samples
 * of the default `cdist_equidistant` sampler, the two random subsets of `cdist_point`).  Squared distance exactly as
 * scipy / cKDTree form it in float64, d2 = RN(RN(dx dx) + RN(dy dy)); thr_d2_dev[k] = the float64 threshold on d2 the
 * host derives from right edge k (d < edge <=> d2 < thr, sqrt being monotone); class = first k with d2 < thr[k], pairs
 * with d2 >= thr[n_bins-1] are dropped.  estimator 0: sum_dev[k] += (v_i - v_j)^2; 1: += |v_i - v_j|^0.5;
 * 2 (Dowd): nothing is summed, pair_class_dev[i*nb + j] (0xFFFF = no class) and pair_key_dev[i*nb + j] (order-preserving
 * key of float32 |v_i - v_j|, as xb_bin_keys builds it) are written for the radix select of xb_bin_hist / xb_bin_next.  maxd2_bits_dev (optional): bit pattern
 * of the largest d2 over all pairs (atomicMax; skgstat's "even" binning clips maxlag to the largest distance).
 * ida/idb (optional, A x B only): sample identities; a pair of a sample with itself is skipped, and a pair whose two
 * samples both carry dupa/dupb != 0 (they belong to BOTH sets, so the pair occurs in both orientations) is only taken in
 * the orientation ida < idb. */
int xb_variogram_pairs_xy(const double* xa_dev, const double* ya_dev, const double* va_dev, int64_t na,
                          const double* xb_dev, const double* yb_dev, const double* vb_dev, int64_t nb,
                          const double* thr_d2_dev, int n_bins, int estimator, unsigned long long* count_dev,
                          double* sum_dev, unsigned long long* maxd2_bits_dev, uint16_t* pair_class_dev,
                          uint32_t* pair_key_dev, const int64_t* ida_dev, const int64_t* idb_dev,
                          const uint8_t* dupa_dev, const uint8_t* dupb_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Nuth & Kaab (2011) inner step.
 * xb_nk_aux replaces `_nuth_kaab_aux_vars` + the zero-slope mask (affine.py:412-474, 578-579): np.gradient semantics
 * (unit spacing, one-sided at raster borders), slope_tan = sqrt(gx^2+gy^2), aspect = arctan2(-gx,gy)+pi, all float32.
 * The buffer may be a row shard: top_is_border / bottom_is_border say whether buffer row 0 / rows_buf-1 are raster
 * borders (otherwise they are halo rows and must lie outside [row_begin,row_end)).
 */
int xb_nk_aux(const float* ref_dev, int64_t rows_buf, int64_t cols, int64_t ld, int top_is_border,
              int bottom_is_border, int64_t row_begin, int64_t row_end, float* slope_tan_dev, float* aspect_dev,
              int64_t out_ld, void* stream);

/* dh at a list of pixels (row-major linear indices idx_dev[n_pts] into the rows x cols raster): the reference's default
 * fit works on a random subsample of 5e5 valid points (affine.py:2405, base.py:576-621) -- the pass then touches only
 * those points instead of streaming the rasters.  dh_dev[i] belongs to point i; aspect_pts_dev is the aspect gathered at
 * the points; asp_minmax / n_finite as in xb_nk_dh. */
int xb_nk_dh_points(const float* ref_dev, const float* tba_dev, const int64_t* idx_dev, int64_t n_pts,
                    const float* aspect_pts_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
                    int64_t tba_rows_total, double dx_px, double dy_px, float* dh_dev, uint32_t* asp_minmax_dev,
                    unsigned long long* n_finite_dev, void* stream);

/* xb_nk_aux fused with the validity mask of the fit (`_preprocess_rst_pts_subsample`, base.py:653-661):
 * sub_mask = inlier (NULL: all) & finite(ref, tba, slope_tan, aspect), n_valid_dev[0] = its population, and
 * range_cand_dev (uint32[4 + 2*64]) = {aspect min bits, max bits over the valid pixels, n_min, n_max, up to 64 pixel
 * indices attaining the minimum, up to 64 attaining the maximum} -- what xb_nkf_iteration needs to confirm the aspect
 * range of an iteration without re-reading the aspect plane.  tba_dev points at the to-be-aligned row of output row 0;
 * slope_tan / aspect / sub_mask are contiguous (row_end - row_begin) x cols. */
int xb_nk_prepare(const float* ref_dev, int64_t rows_buf, int64_t cols, int64_t ld, int top_is_border,
                  int bottom_is_border, int64_t row_begin, int64_t row_end, const float* tba_dev, int64_t tba_ld,
                  const uint8_t* inlier_dev, float* slope_tan_dev, float* aspect_dev, uint8_t* sub_mask_dev,
                  unsigned long long* n_valid_dev, uint32_t* range_cand_dev, void* stream);

/* dh = ref - bilinear(tba at (row + dy_px, col + dx_px)) where sub_mask != 0, NaN elsewhere (affine.py:179-184;
 * geoutils `_interp_points`, linear, NaN-propagating -- restated).  dh / sub_mask / aspect are
 * contiguous rows x cols; tba_dev is a buffer of tba_rows_total rows whose row `tba_row0` is raster row 0 of this
 * shard (halo rows above/below for row shards).  asp_minmax_dev[0/1] receive min/max (float32 bit patterns) of aspect
 * over finite dh (caller initialises to 0xffffffff / 0), n_finite_dev[0] the finite count (accumulated). */
int xb_nk_dh(const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev, const float* aspect_dev,
             int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0, int64_t tba_rows_total,
             double dx_px, double dy_px, float* dh_dev, uint32_t* asp_minmax_dev, unsigned long long* n_finite_dev,
             void* stream);

/* Translation-only `Coreg.apply` (base.py:1567-1570 shifts the transform; base.py:1755-1760 regrids on the original
 * grid with `_reproject_horizontal_shift_samecrs`): dst[r,c] = bilinear(src at (r + dy_px, c + dx_px)) + dz, NaN rule as
 * xb_nk_dh.  For a fitted NuthKaab: dx_px = -shift_x / transform.a, dy_px = -shift_y / transform.e, dz = shift_z. */
int xb_shift_resample(const float* src_dev, int64_t rows, int64_t cols, int64_t ld, double dx_px, double dy_px,
                      double dz, float* dst_dev, int64_t dst_ld, void* stream);

/* Exact medians by MSD radix select on order-preserving float32 keys (np.nanmedian, affine.py:504 and the per-bin
 * nanmedian of `nd_binning`, base.py:1014-1020 -> spatialstats.py:147-149).
 *
 * Global select on dh:  for finite dh with (key & prefix_mask) == prefix:  hist_dev[(key >> shift) & (n_digits-1)] += 1.
 * xb_nk_next: next_key_dev[0] = min(next_key_dev[0], smallest key > sel)  (upper median of even counts). */
int xb_nk_hist(const float* dh_dev, int64_t n, uint32_t prefix, uint32_t prefix_mask, int shift, int n_digits,
               unsigned long long* hist_dev, void* stream);
int xb_nk_next(const float* dh_dev, int64_t n, uint32_t sel, uint32_t* next_key_dev, void* stream);

/* Grouped select over aspect bins.  xb_nk_make_keys computes once per iteration, for every element, the key of
 * y = float32((dh - vshift)/slope_tan) (affine.py:381, 505) and its aspect bin among n_groups equal-width bins of
 * [asp_lo, asp_hi] (scipy.stats.binned_statistic semantics: edges = linspace, right-most edge closed); elements with
 * non-finite dh or y get group 255.  It also fills the first digit histogram hist_dev[group*n_digits + digit] and
 * moments_dev = [n, sum y, sum y^2] (p0 of affine.py:384).  bin_cache_dev [n] keeps every pixel's aspect bin; with
 * reuse_bins != 0 (same [asp_lo, asp_hi] as the call that filled it) the bins are read back instead of recomputed.
 * xb_nk_hist_keys / xb_nk_next_keys refine on the 5-byte
 * (key, group) pairs with per-group prefixes / selections. */
int xb_nk_make_keys(const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, int64_t n, double vshift,
                    double asp_lo, double asp_hi, int n_groups, uint32_t* key_dev, uint8_t* group_dev,
                    uint8_t* bin_cache_dev, int reuse_bins, int shift, int n_digits, unsigned long long* hist_dev,
                    double* moments_dev, void* stream);
int xb_nk_hist_keys(const uint32_t* key_dev, const uint8_t* group_dev, int64_t n, int n_groups,
                    const uint32_t* prefix_dev, uint32_t prefix_mask, int shift, int n_digits,
                    unsigned long long* hist_dev, void* stream);
int xb_nk_next_keys(const uint32_t* key_dev, const uint8_t* group_dev, int64_t n, int n_groups,
                    const uint32_t* sel_dev, uint32_t* next_key_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Nuth & Kaab iteration with bracketed exact selection (csrc/xb_nk_fast.cu): the same exact medians as the entry points
 * above (np.nanmedian(dh), affine.py:504; per-aspect-bin np.nanmedian of (dh - median)/slope_tan, base.py:1014-1020) from
 * TWO streaming passes per iteration and no host round trip until the end.  Every median is first bracketed from a row
 * sample; one full pass counts the keys below the bracket and compacts the keys inside it; a radix select on the compact
 * buffer (ranks picked on the device) yields the exact middle values.  A bracket that misses or overflows sets a flag in
 * cnt[C_FLAGS] and the caller repeats the iteration with the exhaustive entry points.
 *
 * State lives in three small device arrays whose layout xb_nkf_layout reports: out[0..15] = {MAXB, C_SIZE, K_SIZE, F_SIZE,
 * C_NFIN, C_GBELOW, C_GNC, C_BNC, C_FLAGS, C_BTOTAL, C_BBELOW, K_ASPMIN, K_GLO, K_BLO, F_VSHIFT, F_MED};
 *   cnt  uint64[C_SIZE]: finite-dh count, keys below the global bracket, global / per-bin compact counts, flags,
 *        per-bin totals [C_BTOTAL + b] and below-bracket counts [C_BBELOW + b]
 *   keys uint32[K_SIZE]: aspect min / max bit patterns, global bracket [K_GLO, K_GLO+1], per-bin brackets [K_BLO + b] /
 *        [K_BLO + MAXB + b] (order-preserving float32 keys)
 *   f64  double[F_SIZE]: vertical shift (median of dh), aspect range, cached range of the per-pixel bin cache,
 *        moments {n, sum y, sum y^2} at [F_VSHIFT + 5 ..], per-bin medians [F_MED + b]
 * Multi-GPU callers all-gather the sample / compact buffers and all-reduce the counters between the calls (segments:
 * n_seg buffers of seg_cap slots with seg_count filled slots each).  Rasters need cols % 4 == 0 and 16-byte alignment. */
/* The two full passes (xb_nkf_dh / xb_nkf_y with sample == 0) hand their rows out through queue counters inside
 * cnt_dev: call xb_nkf_reset before every iteration (xb_nkf_iteration does), otherwise a second full pass finds its
 * queue exhausted and processes nothing. */
int xb_nkf_layout(int32_t* out16);
int xb_nkf_reset(unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev, void* stream);
/* sample != 0: dh of the jittered row sample (one row of every `stride`) -> sample_dev[(rows+stride-1)/stride * cols]
 * order-preserving keys, 0 = not finite.  sample == 0: dh of every pixel -> dh_dev, aspect range / finite count / count of
 * keys below keys[K_GLO] into cnt / keys, keys within [K_GLO, K_GLO+1] appended to gcompact_dev (capacity gcap). */
int xb_nkf_dh(int sample, const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev,
              const float* aspect_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
              int64_t tba_rows_total, double dx_px, double dy_px, float* dh_dev, uint32_t* sample_dev, int stride,
              uint32_t seed, unsigned long long* cnt_dev, uint32_t* keys_dev, uint32_t* gcompact_dev, uint64_t gcap,
              void* stream);
/* f64[aspect range] = float32 min / max of the aspect over finite dh (after the counters were all-reduced) */
int xb_nkf_range(const uint32_t* keys_dev, double* f64_dev, void* stream);
/* sample != 0: (key of y, aspect bin) of the sampled rows -> skey_dev / sgrp_dev.  sample == 0: every pixel -> per-bin
 * above-bracket (cnt[C_BTOTAL..]) / below-bracket counts / moments into cnt / f64, in-bracket (key, bin) pairs appended
 * to bkey_dev / bgrp_dev; xb_nkf_select mode 2 turns the above-counts into the bin totals. */
int xb_nkf_y(int sample, const float* dh_dev, const float* slope_tan_dev, const float* aspect_dev, uint8_t* bin_cache_dev,
             int64_t rows, int64_t cols, int n_bins, uint32_t* skey_dev, uint8_t* sgrp_dev, int stride, uint32_t seed,
             unsigned long long* cnt_dev, const uint32_t* keys_dev, double* f64_dev, uint32_t* bkey_dev,
             uint8_t* bgrp_dev, uint64_t bcap, void* stream);
/* Two order statistics per group of a small key buffer by 4 x (digit histogram + device-side digit pick).  mode 0: the
 * bracket keys around the sample median of each group -> out_lo_dev / out_hi_dev; mode 1: the exact median of the
 * population (ext_total elements, ext_below below the bracket) from the compact buffer -> out_val_dev (mean of the two
 * middle float32 values in float64, NaN for an empty group); a rank outside the buffer sets miss_bit in flags_dev[0].
 * mode 2: like 1, but ext_total_dev arrives holding the count ABOVE the bracket (what xb_nkf_y accumulates): the
 * population is above + below + the group's entries in the buffer, and that total is written back to ext_total_dev. */
int xb_nkf_select(const uint32_t* key_dev, const uint8_t* grp_dev, int64_t n_seg, int64_t seg_cap,
                  const unsigned long long* seg_count_dev, int64_t seg_count_stride, int n_groups, int mode,
                  unsigned long long* ext_total_dev, const unsigned long long* ext_below_dev, uint32_t* out_lo_dev,
                  uint32_t* out_hi_dev, double* out_val_dev, unsigned long long* flags_dev, uint64_t miss_bit,
                  uint32_t* hist_dev, uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev,
                  void* stream);
int xb_nkf_finalize(unsigned long long* cnt_dev, double* f64_dev, uint64_t gcap, uint64_t bcap, void* stream);
/* One whole single-GPU iteration: the calls above in order (sample_dev / sgrp_dev hold ns = 4 * ceil(rows*cols/4/stride)
 * entries and are reused for the dh and the y sample).  range_cand_dev (from xb_nk_prepare, or NULL): with it the dh pass
 * does not read the aspect plane -- the aspect range is confirmed on the candidates, with a full reduction only when
 * every candidate lost its dh.  aspect_dev may be NULL in xb_nkf_dh(sample = 0) for the same purpose. */
int xb_nkf_iteration(const float* ref_dev, const float* tba_dev, const uint8_t* sub_mask_dev, const float* slope_tan_dev,
                     const float* aspect_dev, int64_t rows, int64_t cols, int64_t ld, int64_t tba_ld, int64_t tba_row0,
                     int64_t tba_rows_total, double dx_px, double dy_px, int n_bins, float* dh_dev, uint8_t* bin_cache_dev,
                     uint32_t* sample_dev, uint8_t* sgrp_dev, int64_t ns, int stride, uint32_t seed,
                     uint32_t* gcompact_dev, uint64_t gcap, uint32_t* bkey_dev, uint8_t* bgrp_dev, uint64_t bcap,
                     unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev,
                     uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev,
                     const uint32_t* range_cand_dev, void* stream);

/* One whole iteration of a point-list fit (see xb_nk_dh_points): exact median of dh, aspect range, exact per-bin medians
 * / counts / moments of y = (dh - median)/slope_tan at the points, all on the device; results in the same cnt / f64
 * block as xb_nkf_iteration.  key_dev / grp_dev: scratch of n_pts entries. */
int xb_nkf_iteration_points(const float* ref_dev, const float* tba_dev, const int64_t* idx_dev, int64_t n_pts,
                            const float* slope_tan_pts_dev, const float* aspect_pts_dev, int64_t rows, int64_t cols,
                            int64_t ld, int64_t tba_ld, int64_t tba_row0, int64_t tba_rows_total, double dx_px,
                            double dy_px, int n_bins, float* dh_pts_dev, uint32_t* key_dev, uint8_t* grp_dev,
                            unsigned long long* cnt_dev, uint32_t* keys_dev, double* f64_dev, uint32_t* hist_dev,
                            uint32_t* prefix_dev, unsigned long long* below_dev, long long* rank_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XDEM_B200_H */
