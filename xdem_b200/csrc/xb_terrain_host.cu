// xdem_b200 -- host-buffer entry of the terrain engine: streams a host raster through the GPU in row blocks.
//
// This is the path a reference caller holding NumPy arrays takes (DEM.slope() -> ... -> engine seam).  The reference's
// analogue is geoutils.map_overlap_multiproc_save (terrain.py:412-466): overlapping tiles with `depth` halo rows.
// Here each row block (+depth halo rows) is copied H2D, processed by the fused kernel(s) and its planes copied D2H on
// one of NSLOT streams, so H2D(b+1), kernel(b) and D2H(b-1) overlap (PCIe is full duplex).
//
// Host buffers may be pinned or pageable (cudaPointerGetAttributes decides per buffer):
//   * pinned (cudaHostAlloc / cudaHostRegister, e.g. the planes the Python layer allocates): DMA straight from / to it;
//   * pageable input (the ndarray a reference caller holds): each block is first copied into a pinned staging buffer of
//     its slot by a few host threads (a single cudaMemcpy from pageable memory stages through one driver thread at a
//     fraction of the link rate), then DMA'd; the staging copy of block b+1 overlaps the DMA / kernel of block b;
//   * pageable output: planes are DMA'd into the slot's pinned staging buffer and drained to the caller's array by the
//     same threads when the slot comes up for reuse (or at the end).
#include "../../include/xdem_b200.h"

#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "xb_common.cuh"
#include "xb_terrain.cuh"

int xb_build_terrain_params(xbt::TerrainParams& p, int dtype, double resolution, int fit_id, int curv_method_id,
                            uint32_t surf_mask, uint32_t win_mask, int window_size, int tri_method_id, int degrees,
                            int clip_hillshade, double az, double alt, double zf, int* hs_out, int* hw_out);
void xb_count_launch(int n);

namespace {
constexpr int NSLOT = 3;
struct Scratch {
    void* buf[NSLOT] = {nullptr, nullptr, nullptr};          // device: [input block | planes]
    size_t bytes = 0;
    void* stage_in[NSLOT] = {nullptr, nullptr, nullptr};     // pinned host staging of a pageable input block
    size_t stage_in_bytes = 0;
    void* stage_out[NSLOT] = {nullptr, nullptr, nullptr};    // pinned host staging of pageable output planes
    size_t stage_out_bytes = 0;
    cudaStream_t stream[NSLOT] = {nullptr, nullptr, nullptr};
    int device = -1;
};
Scratch g_scratch;
std::mutex g_mu;

void free_scratch_locked() {
    for (int i = 0; i < NSLOT; ++i) {
        if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
        if (g_scratch.stage_in[i]) cudaFreeHost(g_scratch.stage_in[i]);
        if (g_scratch.stage_out[i]) cudaFreeHost(g_scratch.stage_out[i]);
        g_scratch.buf[i] = g_scratch.stage_in[i] = g_scratch.stage_out[i] = nullptr;
        if (g_scratch.stream[i]) cudaStreamDestroy(g_scratch.stream[i]);
        g_scratch.stream[i] = nullptr;
    }
    g_scratch.bytes = g_scratch.stage_in_bytes = g_scratch.stage_out_bytes = 0;
    g_scratch.device = -1;
}

int ensure_scratch(size_t dev_bytes, size_t in_bytes, size_t out_bytes) {
    int dev = 0;
    XB_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_scratch.device != dev && g_scratch.device >= 0) {
        int cur = dev;
        cudaSetDevice(g_scratch.device);
        free_scratch_locked();
        cudaSetDevice(cur);
    }
    g_scratch.device = dev;
    for (int i = 0; i < NSLOT; ++i)
        if (!g_scratch.stream[i]) XB_CUDA_CHECK(cudaStreamCreateWithFlags(&g_scratch.stream[i], cudaStreamNonBlocking));
    if (g_scratch.bytes < dev_bytes) {
        for (int i = 0; i < NSLOT; ++i) {
            if (g_scratch.buf[i]) cudaFree(g_scratch.buf[i]);
            g_scratch.buf[i] = nullptr;
        }
        g_scratch.bytes = 0;
        for (int i = 0; i < NSLOT; ++i) XB_CUDA_CHECK(cudaMalloc(&g_scratch.buf[i], dev_bytes));
        g_scratch.bytes = dev_bytes;
    }
    if (g_scratch.stage_in_bytes < in_bytes) {
        for (int i = 0; i < NSLOT; ++i) {
            if (g_scratch.stage_in[i]) cudaFreeHost(g_scratch.stage_in[i]);
            g_scratch.stage_in[i] = nullptr;
        }
        g_scratch.stage_in_bytes = 0;
        for (int i = 0; i < NSLOT; ++i) XB_CUDA_CHECK(cudaHostAlloc(&g_scratch.stage_in[i], in_bytes, cudaHostAllocDefault));
        g_scratch.stage_in_bytes = in_bytes;
    }
    if (g_scratch.stage_out_bytes < out_bytes) {
        for (int i = 0; i < NSLOT; ++i) {
            if (g_scratch.stage_out[i]) cudaFreeHost(g_scratch.stage_out[i]);
            g_scratch.stage_out[i] = nullptr;
        }
        g_scratch.stage_out_bytes = 0;
        for (int i = 0; i < NSLOT; ++i)
            XB_CUDA_CHECK(cudaHostAlloc(&g_scratch.stage_out[i], out_bytes, cudaHostAllocDefault));
        g_scratch.stage_out_bytes = out_bytes;
    }
    return XB_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();  // clear
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int host_threads() {
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("XDEM_B200_HOST_THREADS");
        int v = e ? atoi(e) : 0;
        if (v <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            v = hw >= 32 ? 8 : (hw >= 8 ? 4 : 1);
        }
        n = std::min(v, 64);
    }
    return n;
}

// dst/src pitched 2-D copy, rows split over a few threads (memcpy is bandwidth-bound per thread)
void parallel_copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, int64_t rows) {
    const int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, (int64_t)(width * rows) >> 22));
    auto work = [=](int64_t r0, int64_t r1) {
        if (dpitch == width && spitch == width) {
            memcpy(dst + (size_t)r0 * width, src + (size_t)r0 * width, (size_t)(r1 - r0) * width);
        } else {
            for (int64_t r = r0; r < r1; ++r) memcpy(dst + (size_t)r * dpitch, src + (size_t)r * spitch, width);
        }
    };
    if (nt <= 1) {
        work(0, rows);
        return;
    }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work, rows * t / nt, rows * (t + 1) / nt);
    work(0, rows / nt);
    for (auto& t : th) t.join();
}

// On every exit path the slot streams are drained before the caller's buffers can go away (the async copies read /
// write them).
struct StreamDrain {
    ~StreamDrain() {
        for (int i = 0; i < NSLOT; ++i)
            if (g_scratch.stream[i]) cudaStreamSynchronize(g_scratch.stream[i]);
    }
};
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int xb_release_scratch(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_scratch.device >= 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(g_scratch.device);
        free_scratch_locked();
        cudaSetDevice(cur);
    }
    return XB_OK;
}

int xb_terrain_fused_host_rows(const void* dem_host, int dtype, int64_t rows, int64_t cols, int64_t row_begin,
                               int64_t row_end, double resolution, int fit_id, int curv_method_id, uint32_t surf_mask,
                               uint32_t win_mask, int window_size, int tri_method_id, int degrees, int clip_hillshade,
                               double hillshade_azimuth, double hillshade_altitude, double hillshade_z_factor,
                               void* const* out_planes_host, int64_t rows_per_block) {
    if (!dem_host || !out_planes_host || rows <= 0 || cols <= 0 || row_begin < 0 || row_end > rows ||
        row_begin > row_end) {
        xb_set_error("bad arguments to xb_terrain_fused_host (rows=%lld cols=%lld output rows [%lld,%lld))",
                     (long long)rows, (long long)cols, (long long)row_begin, (long long)row_end);
        return XB_ERR_INVALID;
    }
    xbt::TerrainParams base;
    memset(&base, 0, sizeof(base));
    int hs = 0, hw = 0;
    int rc = xb_build_terrain_params(base, dtype, resolution, fit_id, curv_method_id, surf_mask, win_mask, window_size,
                                     tri_method_id, degrees, clip_hillshade, hillshade_azimuth, hillshade_altitude,
                                     hillshade_z_factor, &hs, &hw);
    if (rc) return rc;
    if (row_end == row_begin) return XB_OK;
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
    const int64_t out_rows = row_end - row_begin;
    if (rows_per_block <= 0) {
        // at most ~96 MiB of input per block (a multiple of the kernels' tile heights); rasters smaller than eight such
        // blocks are cut into ~8 blocks of at least 8 MiB, so that staging, H2D, kernel and D2H still overlap (one
        // block = a fully serial chain: a 4096^2 raster took 5.7 ms end to end instead of ~3)
        const int64_t cap = std::max<int64_t>(64, (int64_t)((96ull << 20) / (cols * es)) / 64 * 64);
        const int64_t floor_rows = std::max<int64_t>(64, (int64_t)((8ull << 20) / (cols * es)) / 64 * 64);
        const int64_t want = ((out_rows + 7) / 8 + 63) / 64 * 64;
        rows_per_block = std::min(cap, std::max(floor_rows, want));
    }
    rows_per_block = std::min(rows_per_block, out_rows);
    // device leading dimension padded to 16 B so that TMA loads and vector stores apply for any width
    const int64_t ld = (cols * (int64_t)es + 15) / 16 * 16 / (int64_t)es;
    const size_t row_bytes = (size_t)cols * es;
    const size_t in_bytes = (size_t)(rows_per_block + 2 * depth) * ld * es;
    const size_t plane_bytes = (size_t)rows_per_block * ld * es;
    const size_t in_bytes_al = (in_bytes + 255) / 256 * 256;
    const size_t slot_bytes = in_bytes_al + plane_bytes * n_planes;

    const bool in_pinned = is_pinned(dem_host);
    bool out_pinned = true;
    for (int k = 0; k < n_planes; ++k) out_pinned = out_pinned && is_pinned(out_planes_host[slots[k]]);
    const size_t stage_in_bytes = in_pinned ? 0 : (size_t)(rows_per_block + 2 * depth) * row_bytes;
    const size_t stage_plane = (size_t)rows_per_block * row_bytes;
    const size_t stage_out_bytes = out_pinned ? 0 : stage_plane * n_planes;

    std::lock_guard<std::mutex> lock(g_mu);
    rc = ensure_scratch(slot_bytes, stage_in_bytes, stage_out_bytes);
    if (rc) return rc;
    StreamDrain drain;  // runs before `lock` is released

    // pageable outputs: rows [r0,r1) of a finished block sit in stage_out[s] until drained
    struct Pending {
        int64_t r0 = 0, r1 = 0;
    } pending[NSLOT];
    auto drain_slot = [&](int s) {
        if (out_pinned || pending[s].r1 == pending[s].r0) return;
        const int64_t r0 = pending[s].r0, r1 = pending[s].r1;
        for (int k = 0; k < n_planes; ++k)
            parallel_copy2d(reinterpret_cast<char*>(out_planes_host[slots[k]]) + (size_t)(r0 - row_begin) * row_bytes,
                            row_bytes, reinterpret_cast<const char*>(g_scratch.stage_out[s]) + stage_plane * k,
                            row_bytes, row_bytes, r1 - r0);
        pending[s].r0 = pending[s].r1 = 0;
    };

    int b = 0;
    for (int64_t r0 = row_begin; r0 < row_end; r0 += rows_per_block, ++b) {
        const int s = b % NSLOT;
        cudaStream_t st = g_scratch.stream[s];
        const int64_t r1 = std::min(row_end, r0 + rows_per_block);
        const int64_t b0 = std::max<int64_t>(0, r0 - depth), b1 = std::min(rows, r1 + depth);
        char* dbuf = reinterpret_cast<char*>(g_scratch.buf[s]);
        const char* src = reinterpret_cast<const char*>(dem_host) + (size_t)b0 * row_bytes;
        if (!in_pinned || !out_pinned) {
            // the slot's staging buffers are free once its previous block has fully drained
            XB_CUDA_CHECK(cudaStreamSynchronize(st));
            drain_slot(s);
        }
        if (!in_pinned) {
            parallel_copy2d(reinterpret_cast<char*>(g_scratch.stage_in[s]), row_bytes, src, row_bytes, row_bytes,
                            b1 - b0);
            src = reinterpret_cast<const char*>(g_scratch.stage_in[s]);
        }
        // stream order on `st` guarantees the previous use of this slot (its D2H copies) has finished
        XB_CUDA_CHECK(cudaMemcpy2DAsync(dbuf, ld * es, src, row_bytes, row_bytes, b1 - b0, cudaMemcpyHostToDevice, st));
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
        for (int k = 0; k < n_planes; ++k) {
            char* dst = out_pinned
                            ? reinterpret_cast<char*>(out_planes_host[slots[k]]) + (size_t)(r0 - row_begin) * row_bytes
                            : reinterpret_cast<char*>(g_scratch.stage_out[s]) + stage_plane * k;
            XB_CUDA_CHECK(cudaMemcpy2DAsync(dst, row_bytes, dbuf + in_bytes_al + plane_bytes * k, ld * es, row_bytes,
                                            r1 - r0, cudaMemcpyDeviceToHost, st));
        }
        pending[s].r0 = r0;
        pending[s].r1 = r1;
    }
    for (int i = 0; i < NSLOT; ++i) {
        XB_CUDA_CHECK(cudaStreamSynchronize(g_scratch.stream[i]));
        drain_slot(i);
    }
    return XB_OK;
}

int xb_terrain_fused_host(const void* dem_host, int dtype, int64_t rows, int64_t cols, double resolution, int fit_id,
                          int curv_method_id, uint32_t surf_mask, uint32_t win_mask, int window_size,
                          int tri_method_id, int degrees, int clip_hillshade, double hillshade_azimuth,
                          double hillshade_altitude, double hillshade_z_factor, void* const* out_planes_host,
                          int64_t rows_per_block) {
    return xb_terrain_fused_host_rows(dem_host, dtype, rows, cols, 0, rows, resolution, fit_id, curv_method_id,
                                      surf_mask, win_mask, window_size, tri_method_id, degrees, clip_hillshade,
                                      hillshade_azimuth, hillshade_altitude, hillshade_z_factor, out_planes_host,
                                      rows_per_block);
}

#pragma GCC visibility pop
}
