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
static __device__ __forceinline__ float splat_alpha(float opacity, float G, float alpha_max = TGS_ALPHA_MAX) {
    return fminf(alpha_max, __fmul_rn(opacity, G));
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

