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
    double alg_k3, alg_c2;  // curvature algebra: d2/d3 (ratio of the z_xx and z_xy dividers) and 100/d2 (xb_terrain_dev.cuh)
    double hs_sin_alt, hs_kx, hs_ky, zf2;  // hillshade constants (surfit.py:606-622)
    double rug_dl2_diag, rug_dl2_straight, rug_ll;  // rugosity constants in the DEM dtype (window.py:628-651)
    double rug_rcp_ll;     // float32 RN(1 / rug_ll) for the correctly rounded reciprocal-multiply division (float32 only)
    // float32 copies of the constants above, rounded on the host exactly as `(float)value` would be in the kernel: the
    // float32 sliding kernels read them straight from the constant bank (ncu r02a: the in-loop F2F conversions of these
    // doubles were re-materialised per row and loaded the XU pipe)
    struct F32 {
        float inv1, ang, hs_ky, hs_nkx, hs_sa, zf2, curv_nf, alg_c2, rug_rcp_ll, rug_nll, rug_l2s, rug_l2d;
        float one;  // 1.0f as a run-time value (addp2, xb_terrain_dev.cuh)
        float rug_y4, rug_b4;  // -RN(1/L^2)/16 and +16 L^2: division of the negated, x16-scaled area sum (xb_terrain_w3.cu)
    } f;
    int rug_fast_ok;       // resolution inside the argument range of the fast IEEE sqrt / division (xb_terrain_w3.cu)
};

int launch(const TerrainParams& p, int dtype, int hs, int hw, cudaStream_t stream);
// float32 Florinsky surface attributes with row-feature reuse (xb_terrain_fl.cu); needs a TMA-eligible raster
int launch_florinsky_sliding(const TerrainParams& p, cudaStream_t stream);
// attribute masks with a packed (f32x2) compile-time specialisation of that kernel
inline bool florinsky_has_packed_mask(uint32_t m) {
    return m == 1u || m == 2u || m == 3u || m == 4u || m == 7u || m == 8u || m == 11u || m == 15u || m == 0x3F7u ||
           m == 0x3FFu;
}
// float32 3x3 windowed indexes with row-feature reuse (xb_terrain_w3.cu); needs a TMA-eligible raster
int launch_window3_sliding(const TerrainParams& p, cudaStream_t stream);

}  // namespace xbt
