// xdem_b200 -- branch-free fp32 math for the terrain kernel (sm_100a): MUFU.RCP / MUFU.RSQ seeds + polynomial cores.
// Accuracy (checked on the GPU by tests/test_terrain_gpu.py::test_math_accuracy_sweep): atan / atan2 <= 2 ulp,
// sqrt <= 1 ulp -- the same class as the CUDA math library functions they replace, without their slow-path branches.
#pragma once
#include <cuda_runtime.h>

namespace xbm {

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// sqrt(x) for x >= 0 (x == 0 -> 0), one Newton step on the MUFU.RSQ seed: <= 1 ulp
__device__ __forceinline__ float sqrt_fast(float x) {
    const float xs = fmaxf(x, 1.17549435e-38f);
    const float r = rsqrt_approx(xs);
    float g = x * r;
    const float e = fmaf(-g, g, x);
    return fmaf(0.5f * r, e, g);
}

// atan(t) for t in [0,1]:  t + t*z*q(z), z = t^2, q = degree-7 near-minimax fit (max 1.6 ulp)
__device__ __forceinline__ float atan_unit(float t) {
    const float z = t * t;
    float q = 2.8423242409e-03f;
    q = fmaf(q, z, -1.6053270867e-02f);
    q = fmaf(q, z, 4.2698739575e-02f);
    q = fmaf(q, z, -7.5086833966e-02f);
    q = fmaf(q, z, 1.0645598343e-01f);
    q = fmaf(q, z, -1.4205896306e-01f);
    q = fmaf(q, z, 1.9993145739e-01f);
    q = fmaf(q, z, -3.3333126241e-01f);
    return fmaf(t * z, q, t);
}

constexpr float PI_F = 3.14159265358979323846f;
constexpr float HALF_PI_F = 1.57079632679489661923f;
constexpr float TWO_PI_F = 6.28318530717958647692f;

// atan(x), x >= 0 (inf allowed)
__device__ __forceinline__ float atan_pos(float x) {
    const bool big = x > 1.0f;
    const float t = big ? rcp_approx(x) : x;
    const float r = atan_unit(t);
    return big ? HALF_PI_F - r : r;
}

// (-atan2(-zx, zy)) mod 2*pi  ==  angle of (zy, zx) measured from +zy towards +zx, in [0, 2*pi]  (surfit.py:600)
__device__ __forceinline__ float aspect_angle(float zx, float zy) {
    const float ax = fabsf(zx), ay = fabsf(zy);
    const float mx = fmaxf(fmaxf(ax, ay), 1.17549435e-38f);
    const float mn = fminf(ax, ay);
    float v = atan_unit(mn * rcp_approx(mx));
    v = ax > ay ? HALF_PI_F - v : v;
    v = zy < 0.0f ? PI_F - v : v;
    v = zx < 0.0f ? TWO_PI_F - v : v;
    return v;
}

}  // namespace xbm
