// xdem_b200 -- Florinsky (5x5) surface-fit kernel with row-feature reuse ("sliding window"), float32, sm_100a.
//
// The generic fused kernel (xb_terrain.cu) rebuilds every pixel's 5x5 stencil from scratch: ~93 FP ops per pixel for the
// four derivatives of slope+aspect+hillshade+curvature, which makes it FMA-pipe bound (ncu: 206 thread-instructions per
// pixel, 66 % FMA-pipe utilisation).  Here each lane owns a 2-pixel-wide column strip and marches down it; for every
// input row it computes, ONCE, the per-pixel row features
//     p = (w[-2]-c) + (w[2]-c),  q = (w[-1]-c) + (w[1]-c)   (second differences about the row centre c = w[0]),
//     D = w[2] - w[-2],          E = w[-1] - w[1]           (first differences),
// keeps the features of the last five rows in registers (a statically indexed ring, the row loop is unrolled by 5) and
// combines them vertically with the Florinsky weights (surfit.py:204-252; effective weights SURVEY.md A.1):
//     z_x  ~ sum_r a_r D_r + b_r E_r                     a = [31,-5,-17,-5,31], b = [44,62,68,62,44]
//     z_y  ~ 31 (p-2 - p2) - 5 (q-2 - q2) + 44 (p1 - p-1) + 62 (q1 - q-1) + 35 (c-2 - c2) + 280 (c1 - c-1)
//     z_xx ~ sum_r (2 p_r - q_r)
//     z_yy ~ [2 (P-2 + P2) - (P-1 + P1) - 2 P0] with P = p + q,  + 5 [2((c-2-c0)+(c2-c0)) - ((c-1-c0)+(c1-c0))]
//     z_xy ~ -(M11 + 2 (M12 + M21) + 4 M22),  M from differences of D / E between rows +-1, +-2
// Every term is a difference-first integer combination, exact in fp32 like the generic kernel.  ~54 FP ops
// per pixel for the same four derivatives.  Same tiles, TMA staging, NaN rule, attribute math and outputs as the generic
// kernel; results are identical up to the association order of exact sums (tested bit-for-bit on integer DEMs and to
// <= 1e-6 relative against the generic kernel / reference fixtures).
//
// Three specialisations on top of that (each A/B-measured): compile-time attribute masks for the two
// headline requests, a warp-uniform branch-free path for interior strips, and packed f32x2 arithmetic (FADD2 / FMUL2 /
// FFMA2) for the two pixels of a lane -- 126 -> 78 issue slots per pixel; the kernel now runs at 0.94 of the bandwidth
// a no-arithmetic kernel with the same read/write mix reaches.
#include <stdlib.h>

#include <type_traits>

#include "xb_terrain_dev.cuh"

namespace xbt {

constexpr int FL_RPW = 15;                 // output rows per warp (multiple of 5: ring period)
constexpr int FL_WY = 4;                   // warps along y
constexpr int FL_TH = FL_WY * FL_RPW;      // 60 output rows per tile
constexpr int FL_BOXH = FL_TH + 4;         // 64 staged rows

struct RowFeat {
    float p[2], q[2], D[2], E[2], c[2];
};

__device__ __forceinline__ void make_features(const float* row, RowFeat& f) {
    // row points at shared-memory column (x0 - 2); 6 values cover both pixels' 5-wide windows
    const float2 a = *reinterpret_cast<const float2*>(row);
    const float2 b = *reinterpret_cast<const float2*>(row + 2);
    const float2 d = *reinterpret_cast<const float2*>(row + 4);
    const float w[6] = {a.x, a.y, b.x, b.y, d.x, d.y};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float c = w[k + 2];
        f.c[k] = c;
        f.p[k] = (w[k] - c) + (w[k + 4] - c);
        f.q[k] = (w[k + 1] - c) + (w[k + 3] - c);
        f.D[k] = w[k + 4] - w[k];
        f.E[k] = w[k + 1] - w[k + 3];
    }
}

// rows r = -2..2 of the current output row are ring slots S0..S4
// CMASK != 0: the surface-attribute mask is a compile-time constant (common requests), which turns the attribute blocks
// into one straight-line region the scheduler can interleave; CMASK == 0: runtime mask.
// FAST: the whole warp strip is inside the raster and 8-byte aligned -> unconditional vector stores, no row test.
template <bool ALG, unsigned CMASK, bool FAST>
__device__ __forceinline__ void emit_row(const RowFeat& r0, const RowFeat& r1, const RowFeat& r2, const RowFeat& r3,
                                         const RowFeat& r4, const TerrainParams& p, long long off, bool full, int nvalid) {
    const unsigned mask = CMASK ? CMASK : p.surf_mask;
    const bool need2 = (mask & ~7u) != 0;
    const bool need_sah = (mask & 7u) != 0;
    float sx[2], sy[2], sxx[2], syy[2], sxy[2], car[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        // z_x (fl_p, divider 420 res): the column weights sum to 35 (D) and 280 (E) and on a surface that is steep but
        // curved across x the two parts cancel; they are combined at the level of the differences FIRST --
        // 35 (D2 + 8 E2) + sum_r a_r (D_r - D2) + b_r (E_r - E2) -- so that the large parts cancel exactly and only
        // small second-order terms meet the big weights (see DESIGN.md "Numerical design")
        {
            const float s1 = fmaf(-2.0f, r2.D[k], r0.D[k] + r4.D[k]), s2 = fmaf(-2.0f, r2.D[k], r1.D[k] + r3.D[k]);
            const float s3 = fmaf(-2.0f, r2.E[k], r0.E[k] + r4.E[k]), s4 = fmaf(-2.0f, r2.E[k], r1.E[k] + r3.E[k]);
            const float tx = fmaf(8.0f, r2.E[k], r2.D[k]);
            sx[k] = fmaf(35.0f, tx, fmaf(31.0f, s1, fmaf(-5.0f, s2, fmaf(44.0f, s3, 62.0f * s4))));
        }
        // z_y (fl_q): row sums about the row centres + the centre-column differences (same ordering rule)
        sy[k] = (31.0f * (r0.p[k] - r4.p[k]) - 5.0f * (r0.q[k] - r4.q[k])) +
                (44.0f * (r3.p[k] - r1.p[k]) + 62.0f * (r3.q[k] - r1.q[k])) +
                35.0f * fmaf(8.0f, r3.c[k] - r1.c[k], r0.c[k] - r4.c[k]);
        const float pp2 = r0.p[k] + r4.p[k], pp1 = r1.p[k] + r3.p[k];
        const float qq2 = r0.q[k] + r4.q[k], qq1 = r1.q[k] + r3.q[k];
        // z_xx (fl_r, 35 res^2) = sum_r (2 p_r - q_r): propagates NaN / inf from every cell of the 5x5 window
        const float sp = (pp2 + pp1) + r2.p[k], sq = (qq2 + qq1) + r2.q[k];
        sxx[k] = 2.0f * sp - sq;
        car[k] = sxx[k] * 0.0f;
        sxy[k] = 0.0f;
        syy[k] = 0.0f;
        if (need2) {
            // z_yy (fl_t, 35 res^2)
            const float tp = 2.0f * (pp2 - r2.p[k]) - pp1;
            const float tq = 2.0f * (qq2 - r2.q[k]) - qq1;
            const float a2 = (r0.c[k] - r2.c[k]) + (r4.c[k] - r2.c[k]);
            const float a1 = (r1.c[k] - r2.c[k]) + (r3.c[k] - r2.c[k]);
            syy[k] = (tp + tq) + 5.0f * (2.0f * a2 - a1);
            if (ALG) {
                // z_xy (fl_s, 100 res^2): mixed second differences
                const float m11 = r1.E[k] - r3.E[k], m12 = r3.D[k] - r1.D[k];
                const float m21 = r0.E[k] - r4.E[k], m22 = r4.D[k] - r0.D[k];
                sxy[k] = -((m11 + 2.0f * (m12 + m21)) + 4.0f * m22);
            }
        }
    }
    if (need_sah) {
        float zx[2], zy[2], g2[2];
        const float inv1 = p.f.inv1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            zx[k] = sx[k] * inv1, zy[k] = sy[k] * inv1;
            g2[k] = fmaf(zx[k], zx[k], zy[k] * zy[k]);
        }
        const float ang = p.f.ang;
        if (mask & 1u)
            store2<FAST>(p.out[0], off, full, nvalid, slope_rad(g2[0]) * ang + car[0], slope_rad(g2[1]) * ang + car[1]);
        if (mask & 2u)
            store2<FAST>(p.out[1], off, full, nvalid, aspect_rad(zx[0], zy[0]) * ang + car[0],
                   aspect_rad(zx[1], zy[1]) * ang + car[1]);
        if (mask & 4u) {
            const float ky = p.f.hs_ky, kx = p.f.hs_nkx, sa = p.f.hs_sa, zf2 = p.f.zf2;
            const float lo = p.clip_hs ? 0.0f : -CUDART_INF_F, hi = p.clip_hs ? 255.0f : CUDART_INF_F;
            float o[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float r = xb_rsqrt(fmaf(zf2, g2[k], 1.0f));
                const float inner = fmaf(ky, zy[k], fmaf(kx, zx[k], sa));
                o[k] = fminf(fmaxf(fmaf(254.0f * r, inner, 1.5f), lo), hi) + car[k];
            }
            store2<FAST>(p.out[2], off, full, nvalid, o[0], o[1]);
        }
    }
    if (mask & 8u) {
        const float f = -p.f.curv_nf;
        store2<FAST>(p.out[3], off, full, nvalid, (sxx[0] + syy[0]) * f + car[0], (sxx[1] + syy[1]) * f + car[1]);
    }
    if constexpr (ALG) {
        if (mask & ~15u) {
            float r6a[6], r6b[6];
            curv_alg<float>(sx[0], sy[0], sxx[0], syy[0], sxy[0], p, mask, r6a);
            curv_alg<float>(sx[1], sy[1], sxx[1], syy[1], sxy[1], p, mask, r6b);
#pragma unroll
            for (int a = 0; a < 6; ++a)
                if (mask & (1u << (4 + a)))
                    store2<FAST>(p.out[4 + a], off, full, nvalid, r6a[a] + car[0], r6b[a] + car[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Packed variant: the two pixels of a lane ride in one f32x2 register pair and the arithmetic uses sm_100's packed
// FADD2 / FMUL2 / FFMA2 (crt/sm_100_rt.h).  A packed op occupies the FMA pipe for two cycles (micro-benchmark
// scripts/micro/ffma2_bench.cu: 106 vs 120 scalar-FMA lanes/clk/SM) but takes ONE issue slot -- and this kernel is
// issue-bound (ncu r01c: 85 % issue-active, FMA pipe 67 %).  Same formulas as emit_row; second derivatives are carried
// negated (nsxx = -z_xx sums, nsyy) so that every subtraction is a single FFMA2 with a -1 / -2 constant (packed ops have
// no operand-negate modifier).  MUFU seeds, min/max and the quadrant selects stay scalar.
// ---------------------------------------------------------------------------------------------------------------

struct RowFeat2 {
    f2 p, q, D, E, c;
};

__device__ __forceinline__ void make_features(const float* row, RowFeat2& f) {
    const f2 a = *reinterpret_cast<const f2*>(row);      // w[0], w[1]
    const f2 b = *reinterpret_cast<const f2*>(row + 2);  // w[2], w[3]  = the two centres
    const f2 d = *reinterpret_cast<const f2*>(row + 4);  // w[4], w[5]
    const f2 m1 = make_float2(a.y, b.x);                  // w[k+1]
    const f2 m3 = make_float2(b.y, d.x);                  // w[k+3]
    f.c = b;
    f.p = add2(sub2(a, b), sub2(d, b));
    f.q = add2(sub2(m1, b), sub2(m3, b));
    f.D = sub2(d, a);
    f.E = sub2(m1, m3);
}

// atan(t), t in [0,1], both components (same polynomial as xbm::atan_unit)
__device__ __forceinline__ f2 atan_unit2(f2 t) {
    const f2 z = mul2(t, t);
    f2 q = S2(2.8423242409e-03f);
    q = fma2(q, z, S2(-1.6053270867e-02f));
    q = fma2(q, z, S2(4.2698739575e-02f));
    q = fma2(q, z, S2(-7.5086833966e-02f));
    q = fma2(q, z, S2(1.0645598343e-01f));
    q = fma2(q, z, S2(-1.4205896306e-01f));
    q = fma2(q, z, S2(1.9993145739e-01f));
    q = fma2(q, z, S2(-3.3333126241e-01f));
    return fma2(mul2(t, z), q, t);
}

constexpr int FL_TS_ROWS = 5;                 // TMA-store variant: rows staged per bulk store (one trip of the row ring)
constexpr int FL_TS_PLANE = FL_TS_ROWS * 64;  // floats per plane in a warp's staging area (64-pixel-wide strip)

// TST: the planes of this row go to the warp's shared-memory staging area (row-major [plane][row][64]; `stg` already
// points at this row and lane) for a later TMA bulk store instead of to global memory.
template <unsigned CMASK, bool FAST, bool TST = false, int DIR = -1>
__device__ __forceinline__ void emit_row(const RowFeat2& r0, const RowFeat2& r1, const RowFeat2& r2, const RowFeat2& r3,
                                         const RowFeat2& r4, const TerrainParams& p, long long off, bool full,
                                         int nvalid, float* stg = nullptr) {
    static_assert(CMASK != 0, "packed path: compile-time attribute mask");
    auto put = [&](auto kc, float a, float b) {
        constexpr int k = decltype(kc)::value;
        if constexpr (TST) {
            constexpr int ord = __builtin_popcount(CMASK & ((1u << k) - 1u));
            *reinterpret_cast<float2*>(stg + ord * FL_TS_PLANE) = make_float2(a, b);
        } else {
            store2<FAST>(p.out[k], off, full, nvalid, a, b);
        }
    };
    constexpr bool NEED2 = (CMASK & ~7u) != 0;    // any second-derivative attribute
    constexpr bool NEEDALG = (CMASK & ~15u) != 0;  // curvature algebra (per-pixel FP64 numerators, curv_alg<float>)
    // z_x (fl_p, divider 420 res): 35 (D2 + 8 E2) + sum_r a_r (D_r - D2) + b_r (E_r - E2) -- the large, mutually
    // cancelling D and E parts are combined exactly at the level of the differences before any big weight is applied
    const f2 s1 = fma2(S2(-2.0f), r2.D, add2(r0.D, r4.D)), s2 = fma2(S2(-2.0f), r2.D, add2(r1.D, r3.D));
    const f2 s3 = fma2(S2(-2.0f), r2.E, add2(r0.E, r4.E)), s4 = fma2(S2(-2.0f), r2.E, add2(r1.E, r3.E));
    const f2 tx = fma2(S2(8.0f), r2.E, r2.D);
    f2 sx = mul2(S2(62.0f), s4);
    sx = fma2(S2(44.0f), s3, sx);
    sx = fma2(S2(-5.0f), s2, sx);
    sx = fma2(S2(31.0f), s1, sx);
    sx = fma2(S2(35.0f), tx, sx);
    // z_y (fl_q)
    const f2 t1 = fma2(S2(-5.0f), sub2(r0.q, r4.q), mul2(S2(31.0f), sub2(r0.p, r4.p)));
    const f2 t2 = fma2(S2(62.0f), sub2(r3.q, r1.q), mul2(S2(44.0f), sub2(r3.p, r1.p)));
    const f2 t3 = mul2(S2(35.0f), fma2(S2(8.0f), sub2(r3.c, r1.c), sub2(r0.c, r4.c)));
    const f2 sy = add2(add2(t1, t2), t3);
    f2 pp2, pp1, qq2, qq1, nsxx, car;
    if constexpr (NEED2) {
        pp2 = add2(r0.p, r4.p), pp1 = add2(r1.p, r3.p);
        qq2 = add2(r0.q, r4.q), qq1 = add2(r1.q, r3.q);
        const f2 sp = add2(add2(pp2, pp1), r2.p), sq = add2(add2(qq2, qq1), r2.q);
        nsxx = fma2(S2(-2.0f), sp, sq);  // -(2 sp - sq): touches every cell of the window
        car = mul2(nsxx, S2(0.0f));
    } else {
        // first-derivative requests: z_x touches every column but the centre one, z_y every row but the centre one
        car = mul2(add2(add2(sx, sy), r2.c), S2(0.0f));
    }
    if (CMASK & 7u) {
        const float inv1 = p.f.inv1;
        const f2 zx = mul2(sx, S2(inv1)), zy = mul2(sy, S2(inv1));
        const f2 g2 = fma2(zx, zx, mul2(zy, zy));
        const float ang = p.f.ang;
        if (CMASK & 1u) {
            // sqrt_fast on both components: r = rsqrt(max(x, tiny)); g = x r; g += (0.5 r)(x - g g)
            const f2 r = make_float2(xbm::rsqrt_approx(fmaxf(g2.x, 1.17549435e-38f)),
                                     xbm::rsqrt_approx(fmaxf(g2.y, 1.17549435e-38f)));
            f2 g = mul2(g2, r);
            const f2 e = fma2(mul2(g, S2(-1.0f)), g, g2);
            g = fma2(mul2(r, S2(0.5f)), e, g);
            // atan_pos
            const bool b0 = g.x > 1.0f, b1 = g.y > 1.0f;
            const f2 t = make_float2(b0 ? xbm::rcp_approx(g.x) : g.x, b1 ? xbm::rcp_approx(g.y) : g.y);
            const f2 a = atan_unit2(t);
            const f2 v = make_float2(b0 ? xbm::HALF_PI_F - a.x : a.x, b1 ? xbm::HALF_PI_F - a.y : a.y);
            const f2 o = fma2(v, S2(ang), car);
            put(std::integral_constant<int, 0>{}, o.x, o.y);
        }
        if (CMASK & 2u) {
            const float ax0 = fabsf(zx.x), ay0 = fabsf(zy.x), ax1 = fabsf(zx.y), ay1 = fabsf(zy.y);
            const f2 mn = make_float2(fminf(ax0, ay0), fminf(ax1, ay1));
            const f2 rc = make_float2(xbm::rcp_approx(fmaxf(fmaxf(ax0, ay0), 1.17549435e-38f)),
                                      xbm::rcp_approx(fmaxf(fmaxf(ax1, ay1), 1.17549435e-38f)));
            const f2 a = atan_unit2(mul2(mn, rc));
            float v0 = a.x, v1 = a.y;
            v0 = ax0 > ay0 ? xbm::HALF_PI_F - v0 : v0;
            v1 = ax1 > ay1 ? xbm::HALF_PI_F - v1 : v1;
            v0 = zy.x < 0.0f ? xbm::PI_F - v0 : v0;
            v1 = zy.y < 0.0f ? xbm::PI_F - v1 : v1;
            v0 = zx.x < 0.0f ? xbm::TWO_PI_F - v0 : v0;
            v1 = zx.y < 0.0f ? xbm::TWO_PI_F - v1 : v1;
            const f2 o = fma2(make_float2(v0, v1), S2(ang), car);
            put(std::integral_constant<int, 1>{}, o.x, o.y);
        }
        if (CMASK & 4u) {
            const float ky = p.f.hs_ky, kx = p.f.hs_nkx, sa = p.f.hs_sa, zf2 = p.f.zf2;
            const float lo = p.clip_hs ? 0.0f : -CUDART_INF_F, hi = p.clip_hs ? 255.0f : CUDART_INF_F;
            const f2 den = fma2(S2(zf2), g2, S2(1.0f));
            const f2 r = make_float2(xb_rsqrt(den.x), xb_rsqrt(den.y));
            const f2 inner = fma2(S2(ky), zy, fma2(S2(kx), zx, S2(sa)));
            const f2 h = fma2(mul2(S2(254.0f), r), inner, S2(1.5f));
            const f2 hc = make_float2(fminf(fmaxf(h.x, lo), hi), fminf(fmaxf(h.y, lo), hi));
            const f2 o = add2(hc, car);
            put(std::integral_constant<int, 2>{}, o.x, o.y);
        }
    }
    if constexpr (NEED2) {
        // -z_yy sums: (pp1 - 2 (pp2 - p2)) + (qq1 - 2 (qq2 - q2)) + 5 (a1 - 2 a2)
        const f2 ntp = fma2(S2(-2.0f), sub2(pp2, r2.p), pp1);
        const f2 ntq = fma2(S2(-2.0f), sub2(qq2, r2.q), qq1);
        const f2 a2 = add2(sub2(r0.c, r2.c), sub2(r4.c, r2.c));
        const f2 a1 = add2(sub2(r1.c, r2.c), sub2(r3.c, r2.c));
        const f2 nsyy = fma2(S2(5.0f), fma2(S2(-2.0f), a2, a1), add2(ntp, ntq));
        if (CMASK & 8u) {
            const float nf = p.f.curv_nf;  // curvature = -200 (z_xx + z_yy) / d2 = (nsxx + nsyy) 200 / d2
            const f2 o = fma2(add2(nsxx, nsyy), S2(nf), car);
            put(std::integral_constant<int, 3>{}, o.x, o.y);
        }
        if constexpr (NEEDALG) {
            // -z_xy sums (fl_s, 100 res^2): M11 + 2 (M12 + M21) + 4 M22, mixed second differences of D / E
            const f2 m11 = sub2(r1.E, r3.E), m12 = sub2(r3.D, r1.D), m21 = sub2(r0.E, r4.E), m22 = sub2(r4.D, r0.D);
            const f2 nsxy = fma2(S2(4.0f), m22, fma2(S2(2.0f), add2(m12, m21), m11));
            float r6a[6], r6b[6];
            // g2 in fp32 exactly as the slope block forms it (and as curv_alg<float> recomputes it)
            const f2 zxa = mul2(sx, S2(p.f.inv1)), zya = mul2(sy, S2(p.f.inv1));
            const f2 g2a = fma2(zxa, zxa, mul2(zya, zya));
            curv_alg_f32<DIR>(sx.x, sy.x, -nsxx.x, -nsyy.x, -nsxy.x, g2a.x, p, CMASK, r6a);
            curv_alg_f32<DIR>(sx.y, sy.y, -nsxx.y, -nsyy.y, -nsxy.y, g2a.y, p, CMASK, r6b);
#define XB_FL_PUT_ALG(A)                             \
    if constexpr ((CMASK & (1u << (4 + A))) != 0) \
        put(std::integral_constant<int, 4 + A>{}, r6a[A] + car.x, r6b[A] + car.y);
            XB_FL_PUT_ALG(0) XB_FL_PUT_ALG(1) XB_FL_PUT_ALG(2) XB_FL_PUT_ALG(3) XB_FL_PUT_ALG(4) XB_FL_PUT_ALG(5)
#undef XB_FL_PUT_ALG
        }
    }
}

template <bool ALG, unsigned CMASK, bool PACKED, int DIR = -1>
__global__ void __launch_bounds__(NTHREADS, 2)
florinsky_sliding_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TerrainParams p) {
    constexpr uint32_t STAGE_BYTES = BOXW * FL_BOXH * sizeof(float);
    constexpr int STAGE_ELEMS = ((STAGE_BYTES + 127) / 128) * 128 / sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wx = warp & 1, wy = warp >> 1;  // 2 x 4 warps: 64-pixel-wide strips, 15 rows each
    const long long tiles_x = p.tiles_x, ntiles = p.ntiles, W = p.cols;

    if (tid == 0) {
        xb_prefetch_tensormap(&tmap);
#pragma unroll
        for (int s = 0; s < NSTAGES; ++s) xb_mbar_init(&full_bar[s], 1);
        xb_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGES; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < ntiles) {
                const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
                xb_mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                xb_tma_load_2d(smem + (size_t)s * STAGE_ELEMS, &tmap, &full_bar[s], tx * TW - XOFF,
                               (int)p.row_begin + ty * FL_TH - 2);
            }
        }
    }
    const bool vec_ok = p.vec_ok != 0;

    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int stage = it % NSTAGES;
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const long long y_tile = p.row_begin + (long long)ty * FL_TH;
        const long long x0 = (long long)tx * TW + wx * 64 + 2 * lane;
        float* tile = smem + (size_t)stage * STAGE_ELEMS;
        xb_mbar_wait(&full_bar[stage], (uint32_t)((it / NSTAGES) & 1));

        const bool active = x0 < W;
        const bool full = vec_ok && (x0 + 1 < W);
        const int nvalid = (int)((W - x0) < 2 ? (W - x0) : 2);
        // first staged row of this warp = tile row wy*15 (i.e. output row - 2); column x0 - 2
        const float* base = tile + (size_t)(wy * FL_RPW) * BOXW + (XOFF + wx * 64 + 2 * lane - 2);
        const long long y_first = y_tile + wy * FL_RPW;
        // interior strips (all 32 lanes write two pixels, all 15 rows exist) take the branch-free path
        const bool warp_fast = __all_sync(0xffffffffu, full) && (y_first + FL_RPW <= p.row_end);
        if (active && y_first < p.row_end) {
            typename std::conditional<PACKED, RowFeat2, RowFeat>::type f0, f1, f2, f3, f4;
            make_features(base + 0 * BOXW, f0);
            make_features(base + 1 * BOXW, f1);
            make_features(base + 2 * BOXW, f2);
            make_features(base + 3 * BOXW, f3);
            long long off = (y_first - p.row_begin) * p.out_ld + x0;
            // five output rows per trip: the ring returns to its starting assignment
#define XB_FL_RING(STEP)               \
    STEP(f4, f0, f1, f2, f3, f4)       \
    STEP(f0, f1, f2, f3, f4, f0)       \
    STEP(f1, f2, f3, f4, f0, f1)       \
    STEP(f2, f3, f4, f0, f1, f2)       \
    STEP(f3, f4, f0, f1, f2, f3)
            if (warp_fast) {
#pragma unroll 1
                for (int g = 0; g < FL_RPW / 5; ++g) {
                    const float* rp = base + (size_t)(4 + 5 * g) * BOXW;
#define XB_FL_STEP(NEW, A, B, C, D, E)                                   \
    make_features(rp, NEW);                                              \
    if constexpr (PACKED) emit_row<CMASK, true, false, DIR>(A, B, C, D, E, p, off, true, 2);      \
    else emit_row<ALG, CMASK, true>(A, B, C, D, E, p, off, true, 2);                               \
    rp += BOXW, off += p.out_ld;
                    XB_FL_RING(XB_FL_STEP)
#undef XB_FL_STEP
                }
            } else {
                long long y = y_first;
#pragma unroll 1
                for (int g = 0; g < FL_RPW / 5; ++g) {
                    const float* rp = base + (size_t)(4 + 5 * g) * BOXW;
#define XB_FL_STEP(NEW, A, B, C, D, E)                                                            \
    make_features(rp, NEW);                                                                       \
    if (y < p.row_end) {                                                                          \
        if constexpr (PACKED) emit_row<CMASK, false, false, DIR>(A, B, C, D, E, p, off, full, nvalid); \
        else emit_row<ALG, CMASK, false>(A, B, C, D, E, p, off, full, nvalid);                    \
    }                                                                                             \
    rp += BOXW, off += p.out_ld, ++y;
                    XB_FL_RING(XB_FL_STEP)
#undef XB_FL_STEP
                }
            }
#undef XB_FL_RING
        }
        __syncthreads();
        if (tid == 0) {
            const long long tn = t + (long long)NSTAGES * gridDim.x;
            if (tn < ntiles) {
                const int tyn = (int)(tn / tiles_x), txn = (int)(tn % tiles_x);
                xb_mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                xb_tma_load_2d(tile, &tmap, &full_bar[stage], txn * TW - XOFF, (int)p.row_begin + tyn * FL_TH - 2);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-store variant of the packed kernel.  The r02 streaming probes (profiles/tma_store_probe_r02*.txt) show that for the
// terrain traffic mix (1 plane read, 3-4 written) bulk tensor stores from shared memory reach 5.5-6.0 TB/s where
// st.global.cs vectors stop at 5.35-5.45 TB/s.  Every warp stages FL_TS_ROWS rows of its 64-pixel strip per plane in
// shared memory and lane 0 issues one cp.async.bulk.tensor.2d store per plane; the tensor maps clip at the raster /
// shard edges, so this variant has no edge paths at all.
// ---------------------------------------------------------------------------------------------------------------
struct OutMaps {
    CUtensorMap m[4];
};

template <unsigned CMASK>
__global__ void __launch_bounds__(NTHREADS, 2)
florinsky_sliding_tstore_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ OutMaps omaps,
                                const __grid_constant__ TerrainParams p) {
    constexpr int NP = __builtin_popcount(CMASK);
    constexpr uint32_t STAGE_BYTES = BOXW * FL_BOXH * sizeof(float);
    constexpr int STAGE_ELEMS = ((STAGE_BYTES + 127) / 128) * 128 / sizeof(float);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[NSTAGES];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wx = warp & 1, wy = warp >> 1;
    const long long tiles_x = p.tiles_x, ntiles = p.ntiles, W = p.cols;
    float* wstage = smem + (size_t)NSTAGES * STAGE_ELEMS + (size_t)warp * (NP * FL_TS_PLANE);

    if (tid == 0) {
        xb_prefetch_tensormap(&tmap);
#pragma unroll
        for (int k = 0; k < NP; ++k) xb_prefetch_tensormap(&omaps.m[k]);
#pragma unroll
        for (int s = 0; s < NSTAGES; ++s) xb_mbar_init(&full_bar[s], 1);
        xb_fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGES; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < ntiles) {
                const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
                xb_mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                xb_tma_load_2d(smem + (size_t)s * STAGE_ELEMS, &tmap, &full_bar[s], tx * TW - XOFF,
                               (int)p.row_begin + ty * FL_TH - 2);
            }
        }
    }
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int stage = it % NSTAGES;
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const long long y_tile = p.row_begin + (long long)ty * FL_TH;
        const long long xs = (long long)tx * TW + wx * 64;  // first column of the warp's strip
        float* tile = smem + (size_t)stage * STAGE_ELEMS;
        xb_mbar_wait(&full_bar[stage], (uint32_t)((it / NSTAGES) & 1));
        const float* base = tile + (size_t)(wy * FL_RPW) * BOXW + (XOFF + wx * 64 + 2 * lane - 2);
        const long long y_first = y_tile + wy * FL_RPW;
        if (xs < W && y_first < p.row_end) {
            RowFeat2 f0, f1, f2, f3, f4;
            make_features(base + 0 * BOXW, f0);
            make_features(base + 1 * BOXW, f1);
            make_features(base + 2 * BOXW, f2);
            make_features(base + 3 * BOXW, f3);
#pragma unroll 1
            for (int g = 0; g < FL_RPW / 5; ++g) {
                const float* rp = base + (size_t)(4 + 5 * g) * BOXW;
                float* sp = wstage + 2 * lane;
                // the previous bulk stores of this warp must have finished READING the staging area
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#define XB_FL_STEP(NEW, A, B, C, D, E)                             \
    make_features(rp, NEW);                                        \
    emit_row<CMASK, true, true>(A, B, C, D, E, p, 0, true, 2, sp); \
    rp += BOXW, sp += 64;
                XB_FL_STEP(f4, f0, f1, f2, f3, f4)
                XB_FL_STEP(f0, f1, f2, f3, f4, f0)
                XB_FL_STEP(f1, f2, f3, f4, f0, f1)
                XB_FL_STEP(f2, f3, f4, f0, f1, f2)
                XB_FL_STEP(f3, f4, f0, f1, f2, f3)
#undef XB_FL_STEP
                xb_fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int oy = (int)(y_first - p.row_begin) + 5 * g;
#pragma unroll
                    for (int k = 0; k < NP; ++k)
                        asm volatile(
                            "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                reinterpret_cast<uint64_t>(&omaps.m[k])),
                            "r"(xb_smem_u32(wstage + k * FL_TS_PLANE)), "r"((int)xs), "r"(oy)
                            : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            const long long tn = t + (long long)NSTAGES * gridDim.x;
            if (tn < ntiles) {
                const int tyn = (int)(tn / tiles_x), txn = (int)(tn % tiles_x);
                xb_mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                xb_tma_load_2d(tile, &tmap, &full_bar[stage], txn * TW - XOFF, (int)p.row_begin + tyn * FL_TH - 2);
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA exits
}

// Eligibility: float32, Florinsky surface attributes only, 16-byte aligned raster (TMA), no windowed indexes.
int launch_florinsky_sliding(const TerrainParams& p_in, cudaStream_t stream) {
    TerrainParams p = p_in;
    p.tiles_x = (p.cols + TW - 1) / TW;
    const long long tiles_y = (p.row_end - p.row_begin + FL_TH - 1) / FL_TH;
    p.ntiles = p.tiles_x * tiles_y;
    if (p.ntiles <= 0) return XB_OK;
    int num_sms = 0;
    int rc = xb_num_sms(&num_sms);
    if (rc) return rc;
    bool vec_ok = (p.out_ld * 4) % 8 == 0;  // 8-byte vector stores
    for (int i = 0; i < 14; ++i)
        if (p.out[i] && (reinterpret_cast<uintptr_t>(p.out[i]) % 8) != 0) vec_ok = false;
    p.vec_ok = vec_ok ? 1 : 0;
    xb_cuTensorMapEncodeTiled_t enc = xb_get_tensormap_encoder();
    if (!enc) return XB_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    cuuint64_t gdim[2] = {(cuuint64_t)p.cols, (cuuint64_t)p.rows_buf};
    cuuint64_t gstr[1] = {(cuuint64_t)(p.ld * 4)};
    cuuint32_t box[2] = {(cuuint32_t)BOXW, (cuuint32_t)FL_BOXH};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p.dem), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) {
        xb_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return XB_ERR_CUDA;
    }
    const size_t smem = (size_t)NSTAGES * (((size_t)BOXW * FL_BOXH * 4 + 127) / 128 * 128);
    const bool alg = (p.surf_mask & ~15u) != 0;
    auto launch_one = [&](auto kern) -> int {
        XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        XB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem));
        if (occ < 1) occ = 1;
        long long grid = (long long)num_sms * occ;
        if (grid > p.ntiles) grid = p.ntiles;
        kern<<<(unsigned)grid, NTHREADS, smem, stream>>>(tmap, p);
        XB_CUDA_CHECK(cudaGetLastError());
        xb_count_launch(1);
        return XB_OK;
    };
    // 2 CTAs/SM: the 5-row feature ring needs ~100 registers (a 3-CTA build spills and measured 40 % slower).
    // The two headline requests get compile-time attribute masks.
    // TMA-store variant for the two headline requests when every plane is 16-byte aligned with a 16-byte pitch
    if (xb_option_florinsky_packed() && xb_option_florinsky_tma_store() && (p.surf_mask == 15u || p.surf_mask == 11u)) {
        bool ok = (p.out_ld * 4) % 16 == 0 && p.cols < (1ll << 31) && (p.row_end - p.row_begin) < (1ll << 31);
        for (int i = 0; i < 4; ++i)
            if (((p.surf_mask >> i) & 1u) && reinterpret_cast<uintptr_t>(p.out[i]) % 16 != 0) ok = false;
        OutMaps om;
        memset(&om, 0, sizeof(om));
        int np = 0;
        for (int i = 0; i < 4 && ok; ++i) {
            if (!((p.surf_mask >> i) & 1u)) continue;
            cuuint64_t od[2] = {(cuuint64_t)p.cols, (cuuint64_t)(p.row_end - p.row_begin)};
            cuuint64_t os[1] = {(cuuint64_t)(p.out_ld * 4)};
            cuuint32_t ob[2] = {64u, (cuuint32_t)FL_TS_ROWS};
            cuuint32_t oe[2] = {1, 1};
            if (enc(&om.m[np++], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.out[i], od, os, ob, oe,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                ok = false;
        }
        if (ok) {
            const size_t smem_ts = smem + (size_t)NWARPS * np * FL_TS_PLANE * sizeof(float);
            auto launch_ts = [&](auto kern) -> int {
                XB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts));
                int occ = 0;
                XB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem_ts));
                if (occ < 1) occ = 1;
                long long grid = (long long)num_sms * occ;
                if (grid > p.ntiles) grid = p.ntiles;
                kern<<<(unsigned)grid, NTHREADS, smem_ts, stream>>>(tmap, om, p);
                XB_CUDA_CHECK(cudaGetLastError());
                xb_count_launch(1);
                return XB_OK;
            };
            return p.surf_mask == 15u ? launch_ts(florinsky_sliding_tstore_kernel<15u>)
                                      : launch_ts(florinsky_sliding_tstore_kernel<11u>);
        }
    }
    // Packed (f32x2) kernels with compile-time masks for the common requests: the API default DEM.slope() / aspect /
    // hillshade and their combinations (first derivatives only), the two headline requests, curvature, and the
    // "all attributes" request of BASELINE config 4 (nine planes, with and without the deprecated `curvature`).
    const bool packed = xb_option_florinsky_packed();
    if (packed) {
        switch (p.surf_mask) {
#define XB_FL_CASE(M) \
    case M: return launch_one(florinsky_sliding_kernel<((M) & ~15u) != 0, M, true>);
            XB_FL_CASE(1u) XB_FL_CASE(2u) XB_FL_CASE(3u) XB_FL_CASE(4u) XB_FL_CASE(7u) XB_FL_CASE(8u) XB_FL_CASE(11u)
            XB_FL_CASE(15u)
#undef XB_FL_CASE
            // the "all attributes" requests: curvature method fixed at compile time (the other method's arithmetic
            // drops out of the per-pixel algebra)
            case 0x3F7u:
                return p.curv_dir ? launch_one(florinsky_sliding_kernel<true, 0x3F7u, true, 1>)
                                  : launch_one(florinsky_sliding_kernel<true, 0x3F7u, true, 0>);
            case 0x3FFu:
                return p.curv_dir ? launch_one(florinsky_sliding_kernel<true, 0x3FFu, true, 1>)
                                  : launch_one(florinsky_sliding_kernel<true, 0x3FFu, true, 0>);
            default: break;
        }
    }
    if (alg) return launch_one(florinsky_sliding_kernel<true, 0u, false>);
    if (p.surf_mask == 15u) return launch_one(florinsky_sliding_kernel<false, 15u, false>);
    if (p.surf_mask == 11u) return launch_one(florinsky_sliding_kernel<false, 11u, false>);
    return launch_one(florinsky_sliding_kernel<false, 0u, false>);
}

}  // namespace xbt
