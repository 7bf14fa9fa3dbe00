// render_math.cuh -- the per-(pixel, splat) arithmetic shared by every compositing kernel of libtgs.so
// (render.cu: the B200 kernels; refstructure.cu: the upstream-structured comparison kernels), so that all
// of them take BIT-IDENTICAL skip / stop decisions for every (pixel, splat) pair.
#pragma once
#include "tgs_math.cuh"

// power = -0.5*(A dx^2 + C dy^2) - B dx dy with a PINNED rounding sequence: forward and backward must
// take bit-identical skip decisions (power > 0, alpha < 1/255) for every (pixel, splat) pair, so the
// contraction into FMAs is spelled out instead of being left to the compiler per kernel.
static __device__ __forceinline__ float splat_power(const float4 q, float dx, float dy) {
    const float ax = __fmul_rn(q.x, dx);
    const float cy = __fmul_rn(q.z, dy);
    const float s = __fmaf_rn(cy, dy, __fmul_rn(ax, dx));
    const float bxy = __fmul_rn(__fmul_rn(q.y, dx), dy);
    return __fmaf_rn(-0.5f, s, -bxy);
}
static __device__ __forceinline__ float splat_alpha(float opacity, float G) {
    return fminf(TGS_ALPHA_MAX, __fmul_rn(opacity, G));
}
// exp(power) = ex2(power*log2e) with flush-to-zero: one FMUL + one MUFU.EX2, no denormal fix-up code.
// (power <= 0 here; results below 2^-126 flush to 0, far under the 1/255 alpha threshold.)
static __device__ __forceinline__ float splat_exp(float power) {
    float y;
    const float x = __fmul_rn(power, 1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
static __device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// power threshold of the alpha >= 1/255 test: o*exp(power) >= 1/255  <=>  power >= thr := -ln(255 o).
// lg2.approx is accurate to ~1e-6 relative; the culls compare against thr with an absolute margin of 0.05.
static __device__ __forceinline__ float splat_thr(float opacity) {
    return -0.6931471805599453f * __log2f(255.0f * opacity);
}

// EXACT warp-level cull.  A splat can only contribute to a pixel if alpha = o*exp(power) >= 1/255,
// i.e. power >= thr := -ln(255 o)  (thr is precomputed per splat in record.c.w).  power = -q/2 with
// q(p) = (p-mu)^T Q (p-mu) convex, so over a pixel rectangle R (a warp's patch, or a whole tile at pack time) the maximum power is -q_min/2
// where q_min is 0 if mu lies in R and otherwise the minimum over the four edges (each a clamped 1-D
// quadratic).  If even that maximum is below thr (minus a safety margin that dominates the fp32
// rounding of this bound and of the per-pixel evaluation), NO pixel of the patch would pass the
// alpha test, so skipping the splat for the whole warp is bit-identical to evaluating it.
static __device__ __forceinline__ bool rect_may_touch(const float4 a, const float4 q, float thr, float x0, float x1,
                                                      float y0, float y1) {
    const float ex0 = x0 - a.x, ex1 = x1 - a.x, ey0 = y0 - a.y, ey1 = y1 - a.y;
    if (ex0 <= 0.0f && ex1 >= 0.0f && ey0 <= 0.0f && ey1 >= 0.0f) return thr <= 0.05f;
    const float A = q.x, B = q.y, Cc = q.z;
    const float rA = __fdividef(1.0f, A), rC = __fdividef(1.0f, Cc);
    float qmin, tmax;
    {   // vertical edges: dx fixed, dy* = clamp(-B dx / C)
        float dy = fminf(fmaxf(-B * ex0 * rC, ey0), ey1);
        float t1 = A * ex0 * ex0, t2 = Cc * dy * dy, t3 = 2.0f * B * ex0 * dy;
        qmin = t1 + t2 + t3; tmax = t1 + t2 + fabsf(t3);
        dy = fminf(fmaxf(-B * ex1 * rC, ey0), ey1);
        t1 = A * ex1 * ex1; t2 = Cc * dy * dy; t3 = 2.0f * B * ex1 * dy;
        float qq = t1 + t2 + t3;
        if (qq < qmin) { qmin = qq; tmax = t1 + t2 + fabsf(t3); }
    }
    {   // horizontal edges: dy fixed, dx* = clamp(-B dy / A)
        float dx = fminf(fmaxf(-B * ey0 * rA, ex0), ex1);
        float t1 = A * dx * dx, t2 = Cc * ey0 * ey0, t3 = 2.0f * B * dx * ey0;
        float qq = t1 + t2 + t3;
        if (qq < qmin) { qmin = qq; tmax = t1 + t2 + fabsf(t3); }
        dx = fminf(fmaxf(-B * ey1 * rA, ex0), ex1);
        t1 = A * dx * dx; t2 = Cc * ey1 * ey1; t3 = 2.0f * B * dx * ey1;
        qq = t1 + t2 + t3;
        if (qq < qmin) { qmin = qq; tmax = t1 + t2 + fabsf(t3); }
    }
    const float margin = 0.05f + 1e-5f * tmax;
    return -0.5f * qmin >= thr - margin;
}

