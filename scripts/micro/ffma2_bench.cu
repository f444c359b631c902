// Micro-benchmark: throughput of packed fp32 (FFMA2 / FADD2 / FMUL2, sm_100) vs scalar FFMA, alone and mixed with
// integer work, to decide whether pairing the two pixels of a lane into f32x2 operations relieves an issue-bound kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float seed) {
    float2 a[8];
    unsigned u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = make_float2(seed + k + threadIdx.x, seed - k), u[k] = threadIdx.x * 7 + k;
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0 || MODE == 2) {  // scalar: 2 FFMA per element pair
                a[k].x = fmaf(a[k].x, m.x, c.x);
                a[k].y = fmaf(a[k].y, m.y, c.y);
            } else {  // packed: 1 FFMA2
                a[k] = __ffma2_rn(a[k], m, c);
            }
            if (MODE >= 2) {  // 2 integer ops per element pair
                u[k] = (u[k] ^ (u[k] >> 3)) + 0x9e3779b9u;
            }
        }
    }
    float s = 0.f;
    unsigned su = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y, su += u[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)su;
}

template <int MODE>
void run(const char* name, float* d, int sms) {
    const int iters = 20000, grid = sms * 8;
    kern<MODE><<<grid, 256>>>(d, 100, 1.f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<MODE><<<grid, 256>>>(d, iters, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * (double)iters * grid * 256;  // 16 scalar FMAs per thread-iteration
    printf("%-28s %8.3f ms  %8.2f TFLOP/s fp32  (%.1f scalar-FMA lanes/clk/SM at 1.965 GHz)\n", name, ms,
           flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* d;
    cudaMalloc(&d, sizeof(float) * sms * 8 * 256);
    run<0>("scalar FFMA", d, sms);
    run<1>("packed FFMA2", d, sms);
    run<2>("scalar FFMA + int ops", d, sms);
    run<3>("packed FFMA2 + int ops", d, sms);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
