// xdem_b200 -- common device/host helpers (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define XB_OK 0
#define XB_ERR_INVALID (-1)
#define XB_ERR_CUDA (-2)
#define XB_ERR_UNSUPPORTED (-3)

void xb_set_error(const char* fmt, ...);
void xb_count_launch(int n);  // every kernel launch of the library is counted (xb_launch_count)

#define XB_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            xb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return XB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t xb_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void xb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xb_smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void xb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void xb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void xb_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xb_smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool xb_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(xb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void xb_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!xb_mbar_try_wait(bar, parity)) {
    }
}

// 2-D tiled TMA load: box (c0 = innermost/x coordinate, c1 = y coordinate); out-of-bounds elements are filled
// according to the tensor map (NaN for the DEM map).
__device__ __forceinline__ void xb_tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(xb_smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(xb_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void xb_prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Host: cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
typedef CUresult (*xb_cuTensorMapEncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                CUtensorMapFloatOOBfill);
xb_cuTensorMapEncodeTiled_t xb_get_tensormap_encoder();
int xb_num_sms(int* out);
bool xb_option_florinsky_generic();
bool xb_option_florinsky_packed();   // f32x2 (FFMA2) arithmetic in the sliding Florinsky kernel
bool xb_option_florinsky_tma_store();  // headline Florinsky requests: planes written by TMA bulk stores (A/B)
bool xb_option_window3_generic();    // 3x3 windowed indexes through the generic fused kernel (A/B tests)
int xb_option_variogram_full_tiles();  // bit k-1: interior tiles spanning k lag classes take the threshold-light sweep
