// xdem_b200 -- streaming probe: the trivial kernel with the terrain engine's memory traffic (one float32 plane read,
// n_planes float32 planes written with streaming vector stores, no arithmetic).  bench.py times it next to the real
// kernel so that the roofline line can also be read against the bandwidth this access mix can reach at all on the box
// (write-heavy traffic does not reach the copy bandwidth of MEASURED_PEAKS.json).  Diagnostics only.
#include "../../include/xdem_b200.h"

#include "xb_terrain_dev.cuh"

void xb_count_launch(int n);

namespace xbp {

template <int NP>
__global__ void __launch_bounds__(256)
stream_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
#pragma unroll
        for (int p = 0; p < NP; ++p) __stcs(dst + (size_t)p * n4 + i, make_float4(v.x + (float)p, v.y, v.z, v.w));
    }
}

// Same traffic mix with the planes written by TMA: every CTA stages a (8 x 128) tile of each plane in shared memory
// and one thread issues cp.async.bulk.tensor.3d stores (UTMASTG) through one tensor map over dst[plane][row][col].
// Answers whether the ~5.4 TB/s of the st.global.cs probe is a property of the 1-read / 4-write mix or of the store
// instruction.
constexpr int TS_W = 128, TS_H = 8, TS_PLANES = 4;

__global__ void __launch_bounds__(256)
stream_tma_store_kernel(const float* __restrict__ src, const __grid_constant__ CUtensorMap tmap, long long rows,
                        long long cols, int n_planes) {
    __shared__ __align__(128) float tile[2][TS_PLANES][TS_H][TS_W];  // double-buffered: 2 x 16 KB
    const long long tiles_x = cols / TS_W, tiles_y = rows / TS_H;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 lanes x float4 = 128 columns, 8 row groups
    int it = 0;
    for (long long t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x, ++it) {
        const long long tyi = t / tiles_x, txi = t - tyi * tiles_x;
        const int buf = it & 1;
        // the bulk stores that read this buffer two iterations ago must have finished reading shared memory
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int r = ty; r < TS_H; r += 8) {
            const float4 v = *reinterpret_cast<const float4*>(src + (tyi * TS_H + r) * cols + txi * TS_W + 4 * tx);
#pragma unroll
            for (int p = 0; p < TS_PLANES; ++p)
                if (p < n_planes)
                    *reinterpret_cast<float4*>(&tile[buf][p][r][4 * tx]) = make_float4(v.x + (float)p, v.y, v.z, v.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int p = 0; p < n_planes; ++p)
                asm volatile(
                    "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                        reinterpret_cast<uint64_t>(&tmap)),
                    "r"(xb_smem_u32(&tile[buf][p][0][0])), "r"((int)(txi * TS_W)), "r"((int)(tyi * TS_H)), "r"(p)
                    : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Bit-exactness probes of the branch-free IEEE cores used by the 3x3 windowed kernel (xb_terrain_w3.cu).
// kind 0: sqrt2_rn_fast(x) vs __fsqrt_rn(x);  kind 1: div2_rn_const(x, y, -b) vs __fdiv_rn(x, b);
// x runs over the `count` consecutive float32 bit patterns starting at bits_begin.
__global__ void __launch_bounds__(256)
exact_math_kernel(int kind, uint32_t bits_begin, unsigned long long count, float b, float y,
                  unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; 2 * i < count;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t b0 = bits_begin + (uint32_t)(2 * i), b1 = (2 * i + 1 < count) ? b0 + 1 : b0;
        const float2 x = make_float2(__uint_as_float(b0), __uint_as_float(b1));
        float2 fast, ref;
        if (kind == 0) {
            fast = xbt::sqrt2_rn_fast(x);
            ref = make_float2(__fsqrt_rn(x.x), __fsqrt_rn(x.y));
        } else {
            fast = xbt::div2_rn_const(x, y, -b);
            ref = make_float2(__fdiv_rn(x.x, b), __fdiv_rn(x.y, b));
        }
        bad += (__float_as_uint(fast.x) != __float_as_uint(ref.x)) + (__float_as_uint(fast.y) != __float_as_uint(ref.y));
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace xbp

extern "C" {
#pragma GCC visibility push(default)

int xb_probe_stream(const void* src_dev, void* dst_dev, int64_t n_floats, int n_planes, void* stream) {
    if (!src_dev || !dst_dev || n_floats < 4 || n_floats % 4 != 0 || n_planes < 1 || n_planes > 4 ||
        reinterpret_cast<uintptr_t>(src_dev) % 16 != 0 || reinterpret_cast<uintptr_t>(dst_dev) % 16 != 0) {
        xb_set_error("bad arguments to xb_probe_stream (16-byte aligned buffers, n_floats %% 4 == 0, 1..4 planes)");
        return XB_ERR_INVALID;
    }
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n4 = (size_t)n_floats / 4;
    const float4* s = reinterpret_cast<const float4*>(src_dev);
    float4* d = reinterpret_cast<float4*>(dst_dev);
    const int grid = sms * 8;
    switch (n_planes) {
        case 1: xbp::stream_kernel<1><<<grid, 256, 0, st>>>(s, d, n4); break;
        case 2: xbp::stream_kernel<2><<<grid, 256, 0, st>>>(s, d, n4); break;
        case 3: xbp::stream_kernel<3><<<grid, 256, 0, st>>>(s, d, n4); break;
        default: xbp::stream_kernel<4><<<grid, 256, 0, st>>>(s, d, n4); break;
    }
    XB_CUDA_CHECK(cudaGetLastError());
    return XB_OK;
}

int xb_probe_stream_tma(const void* src_dev, void* dst_dev, int64_t rows, int64_t cols, int n_planes, void* stream) {
    if (!src_dev || !dst_dev || rows < xbp::TS_H || cols < xbp::TS_W || rows % xbp::TS_H || cols % xbp::TS_W ||
        n_planes < 1 || n_planes > xbp::TS_PLANES || reinterpret_cast<uintptr_t>(src_dev) % 16 ||
        reinterpret_cast<uintptr_t>(dst_dev) % 16) {
        xb_set_error("bad arguments to xb_probe_stream_tma (rows %% 8 == 0, cols %% 128 == 0, 1..4 planes)");
        return XB_ERR_INVALID;
    }
    xb_cuTensorMapEncodeTiled_t enc = xb_get_tensormap_encoder();
    if (!enc) return XB_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)n_planes};
    cuuint64_t gstr[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * (cuuint64_t)rows * 4};
    cuuint32_t box[3] = {(cuuint32_t)xbp::TS_W, (cuuint32_t)xbp::TS_H, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dst_dev, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        xb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return XB_ERR_CUDA;
    }
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    xbp::stream_tma_store_kernel<<<sms * 6, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float*>(src_dev), tmap, rows, cols, n_planes);
    XB_CUDA_CHECK(cudaGetLastError());
    return XB_OK;
}

int xb_probe_exact_math(int kind, uint32_t bits_begin, uint64_t count, float b, float rcp_b,
                        unsigned long long* mismatches_dev, void* stream) {
    if ((kind != 0 && kind != 1) || !mismatches_dev || count == 0) {
        xb_set_error("bad arguments to xb_probe_exact_math");
        return XB_ERR_INVALID;
    }
    int sms = 0;
    int rc = xb_num_sms(&sms);
    if (rc) return rc;
    xbp::exact_math_kernel<<<sms * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        kind, bits_begin, (unsigned long long)count, b, rcp_b, mismatches_dev);
    XB_CUDA_CHECK(cudaGetLastError());
    return XB_OK;
}

#pragma GCC visibility pop
}
