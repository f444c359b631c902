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

/* Same computation from/to HOST buffers (pinned memory recommended): the raster is streamed through the GPU in row
 * blocks with `depth` halo rows (the reference's tiling analogue: geoutils.map_overlap_multiproc_save,
 * terrain.py:412-466), H2D / kernel / D2H overlapped on three streams.  This is the call the `e2e` benchmark figure
 * times.  out_planes_host[i] are HOST pointers here; rows_per_block <= 0 picks ~96 MiB blocks. */
int xb_terrain_fused_host(const void* dem_host, int dtype, int64_t rows, int64_t cols, double resolution, int fit_id,
                          int curv_method_id, uint32_t surf_mask, uint32_t win_mask, int window_size,
                          int tri_method_id, int degrees, int clip_hillshade, double hillshade_azimuth,
                          double hillshade_altitude, double hillshade_z_factor, void* const* out_planes_host,
                          int64_t rows_per_block);

#ifdef __cplusplus
}
#endif
#endif /* XDEM_B200_H */
