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
