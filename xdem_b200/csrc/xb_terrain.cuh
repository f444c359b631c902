// xdem_b200 -- parameters of the fused terrain kernel (K1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XB_FIT_HORN_ID 0
#define XB_FIT_ZT_ID 1
#define XB_FIT_FLORINSKY_ID 2

namespace xbt {

struct TerrainParams {
    const void* dem;       // device raster buffer
    long long rows_buf;    // rows of the buffer (rows outside are NaN)
    long long cols;
    long long ld;          // leading dimension (elements)
    long long row_begin;   // output rows [row_begin,row_end)
    long long row_end;
    void* out[14];         // planes: 0..9 surface attributes, 10..13 windowed indexes (NULL if not requested)
    long long out_ld;
    long long tiles_x, ntiles;
    uint32_t surf_mask, win_mask;
    int fit_id;            // 0 Horn, 1 ZevenbergThorne, 2 Florinsky
    int curv_dir;          // 0 geometric, 1 directional
    int tri_wilson;        // 0 Riley, 1 Wilson
    int degrees, clip_hs, vec_ok;
    double inv_d1, inv_d2, inv_d3;  // 1/divider of (z_x,z_y), (z_xx,z_yy), z_xy  (surfit.py:278-304)
    double rad2deg;
    double hs_sin_alt, hs_kx, hs_ky, zf2;  // hillshade constants (surfit.py:606-622)
    double rug_dl2_diag, rug_dl2_straight, rug_ll;  // rugosity constants in the DEM dtype (window.py:628-651)
};

int launch(const TerrainParams& p, int dtype, int hs, int hw, cudaStream_t stream);
// float32 Florinsky surface attributes with row-feature reuse (xb_terrain_fl.cu); needs a TMA-eligible raster
int launch_florinsky_sliding(const TerrainParams& p, cudaStream_t stream);

}  // namespace xbt
