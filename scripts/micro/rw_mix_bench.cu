// Micro-benchmark: achievable HBM bandwidth for the terrain kernel's traffic mix (4 B read + 16 B written per pixel,
// four output planes, streaming stores) next to a plain copy, a pure read and a pure write -- the practical ceiling the
// K1 roofline fraction should be read against (MEASURED_PEAKS.json holds the COPY bandwidth).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rw_mix_bench rw_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ a, float4* __restrict__ o, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(o + i, a[i]);
}
__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ a, float* __restrict__ o, size_t n4) {
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = a[i];
        s += v.x + v.y + v.z + v.w;
    }
    if (s == 123.456f) o[0] = s;
}
__global__ void __launch_bounds__(256) k_write(float4* __restrict__ o, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(o + i, make_float4(1.f, 2.f, 3.f, (float)i));
}
template <int NP>
__global__ void __launch_bounds__(256) k_mix(const float4* __restrict__ a, float4* __restrict__ o, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = a[i];
#pragma unroll
        for (int p = 0; p < NP; ++p) __stcs(o + (size_t)p * n4 + i, make_float4(v.x + p, v.y, v.z, v.w));
    }
}

template <typename F>
static float time_ms(F launch) {
    launch();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t n = (size_t)32768 * 32768, n4 = n / 4;
    float4 *a, *o;
    cudaMalloc(&a, n * 4);
    cudaMalloc(&o, n * 4 * 4);
    cudaMemset(a, 0, n * 4);
    const int grid = sms * 8;
    float ms;
    ms = time_ms([&] { k_copy<<<grid, 256>>>(a, o, n4); });
    printf("copy  (4 B read + 4 B written / px)   %7.3f ms  %7.1f GB/s\n", ms, 8.0 * n / ms / 1e6);
    ms = time_ms([&] { k_read<<<grid, 256>>>(a, (float*)o, n4); });
    printf("read  (4 B read / px)                 %7.3f ms  %7.1f GB/s\n", ms, 4.0 * n / ms / 1e6);
    ms = time_ms([&] { k_write<<<grid, 256>>>(o, n4 * 4); });
    printf("write (16 B written / px)             %7.3f ms  %7.1f GB/s\n", ms, 16.0 * n / ms / 1e6);
    ms = time_ms([&] { k_mix<4><<<grid, 256>>>(a, o, n4); });
    printf("mix   (4 B read + 16 B written / px)  %7.3f ms  %7.1f GB/s   <- terrain, 4 planes\n", ms, 20.0 * n / ms / 1e6);
    ms = time_ms([&] { k_mix<3><<<grid, 256>>>(a, o, n4); });
    printf("mix   (4 B read + 12 B written / px)  %7.3f ms  %7.1f GB/s   <- terrain, 3 planes\n", ms, 16.0 * n / ms / 1e6);
    ms = time_ms([&] { k_mix<1><<<grid, 256>>>(a, o, n4); });
    printf("mix   (4 B read + 4 B written / px)   %7.3f ms  %7.1f GB/s   <- terrain, 1 plane\n", ms, 8.0 * n / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
