// xdem_b200 -- host-buffer entry of the terrain engine: streams a host raster through the GPU in row blocks.
//
// This is the path a reference caller holding NumPy arrays takes (DEM.slope() -> ... -> engine seam).  The reference's
// analogue is geoutils.map_overlap_multiproc_save (terrain.py:412-466): overlapping tiles with `depth` halo rows.
// Here each row block (+depth halo rows) is copied H2D, processed by the fused kernel and its planes copied D2H on one
// of NSLOT streams, so H2D(b+1), kernel(b) and D2H(b-1) overlap (PCIe is full duplex).
#include "../../include/xdem_b200.h"

#include <algorithm>
#include <mutex>

#include "xb_common.cuh"
#include "xb_terrain.cuh"

int xb_build_terrain_params(xbt::TerrainParams& p, int dtype, double resolution, int fit_id, int curv_method_id,
                            uint32_t surf_mask, uint32_t win_mask, int window_size, int tri_method_id, int degrees,
                            int clip_hillshade, double az, double alt, double zf, int* hs_out, int* hw_out);
void xb_count_launch(int n);

namespace {
constexpr int NSLOT = 3;
struct Scratch {
    void* buf[NSLOT] = {nullptr, nullptr, nullptr};
    size_t bytes = 0;
    cudaStream_t stream[NSLOT] = {nullptr, nullptr, nullptr};
    int device = -1;
};
Scratch g_scratch;
std::mutex g_mu;

int ensure_scratch(size_t bytes) {
    int dev = 0;
    XB_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_scratch.device != dev || g_scratch.bytes < bytes) {
        for (int i = 0; i < NSLOT; ++i) {
            if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
            g_scratch.buf[i] = nullptr;
            if (g_scratch.device != dev && g_scratch.stream[i]) {
                cudaStreamDestroy(g_scratch.stream[i]);
                g_scratch.stream[i] = nullptr;
            }
        }
        g_scratch.bytes = 0;
        for (int i = 0; i < NSLOT; ++i) {
            XB_CUDA_CHECK(cudaMalloc(&g_scratch.buf[i], bytes));
            if (!g_scratch.stream[i]) XB_CUDA_CHECK(cudaStreamCreateWithFlags(&g_scratch.stream[i], cudaStreamNonBlocking));
        }
        g_scratch.bytes = bytes;
        g_scratch.device = dev;
    }
    return XB_OK;
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int xb_terrain_fused_host(const void* dem_host, int dtype, int64_t rows, int64_t cols, double resolution, int fit_id,
                          int curv_method_id, uint32_t surf_mask, uint32_t win_mask, int window_size,
                          int tri_method_id, int degrees, int clip_hillshade, double hillshade_azimuth,
                          double hillshade_altitude, double hillshade_z_factor, void* const* out_planes_host,
                          int64_t rows_per_block) {
    if (!dem_host || !out_planes_host || rows <= 0 || cols <= 0) {
        xb_set_error("bad arguments to xb_terrain_fused_host");
        return XB_ERR_INVALID;
    }
    xbt::TerrainParams base;
    memset(&base, 0, sizeof(base));
    int hs = 0, hw = 0;
    int rc = xb_build_terrain_params(base, dtype, resolution, fit_id, curv_method_id, surf_mask, win_mask, window_size,
                                     tri_method_id, degrees, clip_hillshade, hillshade_azimuth, hillshade_altitude,
                                     hillshade_z_factor, &hs, &hw);
    if (rc) return rc;
    const int depth = std::max(hs, hw);
    const size_t es = dtype == XB_F64 ? 8 : 4;
    int slots[XB_N_PLANES], n_planes = 0;
    for (int i = 0; i < 10; ++i)
        if ((surf_mask >> i) & 1u) slots[n_planes++] = i;
    for (int j = 0; j < 4; ++j)
        if ((win_mask >> j) & 1u) slots[n_planes++] = 10 + j;
    for (int k = 0; k < n_planes; ++k)
        if (!out_planes_host[slots[k]]) {
            xb_set_error("plane %d requested but its host pointer is NULL", slots[k]);
            return XB_ERR_INVALID;
        }
    if (rows_per_block <= 0) {
        // ~96 MiB of input per block, a multiple of the kernel's tile height
        rows_per_block = std::max<int64_t>(64, (int64_t)((96ull << 20) / (cols * es)) / 64 * 64);
    }
    rows_per_block = std::min(rows_per_block, rows);
    // device leading dimension padded to 16 B so that TMA loads and vector stores apply for any width
    const int64_t ld = (cols * (int64_t)es + 15) / 16 * 16 / (int64_t)es;
    const size_t in_bytes = (size_t)(rows_per_block + 2 * depth) * ld * es;
    const size_t plane_bytes = (size_t)rows_per_block * ld * es;
    const size_t in_bytes_al = (in_bytes + 255) / 256 * 256;
    const size_t slot_bytes = in_bytes_al + plane_bytes * n_planes;

    std::lock_guard<std::mutex> lock(g_mu);
    rc = ensure_scratch(slot_bytes);
    if (rc) return rc;

    int b = 0;
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per_block, ++b) {
        const int s = b % NSLOT;
        cudaStream_t st = g_scratch.stream[s];
        const int64_t r1 = std::min(rows, r0 + rows_per_block);
        const int64_t b0 = std::max<int64_t>(0, r0 - depth), b1 = std::min(rows, r1 + depth);
        char* dbuf = reinterpret_cast<char*>(g_scratch.buf[s]);
        // stream order on `st` guarantees the previous use of this slot (its D2H copies) has finished
        XB_CUDA_CHECK(cudaMemcpy2DAsync(dbuf, ld * es, reinterpret_cast<const char*>(dem_host) + (size_t)b0 * cols * es,
                                        cols * es, cols * es, b1 - b0, cudaMemcpyHostToDevice, st));
        xbt::TerrainParams p = base;
        p.dem = dbuf;
        p.rows_buf = b1 - b0;
        p.cols = cols;
        p.ld = ld;
        p.row_begin = r0 - b0;
        p.row_end = r1 - b0;
        p.out_ld = ld;
        for (int k = 0; k < n_planes; ++k) p.out[slots[k]] = dbuf + in_bytes_al + plane_bytes * k;
        rc = xbt::launch(p, dtype, hs, hw, st);
        if (rc) return rc;
        xb_count_launch(1);
        for (int k = 0; k < n_planes; ++k) {
            char* dst = reinterpret_cast<char*>(out_planes_host[slots[k]]) + (size_t)r0 * cols * es;
            XB_CUDA_CHECK(cudaMemcpy2DAsync(dst, cols * es, dbuf + in_bytes_al + plane_bytes * k, ld * es, cols * es,
                                            r1 - r0, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < NSLOT; ++i) XB_CUDA_CHECK(cudaStreamSynchronize(g_scratch.stream[i]));
    return XB_OK;
}

#pragma GCC visibility pop
}
