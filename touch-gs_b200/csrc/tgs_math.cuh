// tgs_math.cuh -- per-Gaussian math of the rasterizer hot path, usable from device code and
// (for the CPU-side math tests only: tests/host_math_harness.cpp) from host code.
//
// Forward math follows SURVEY.md §8(a) row A1 in EXACTLY the operation order of
// oracle/gs_oracle.py::preprocess, so that with FMA contraction disabled (the CUDA translation
// unit is compiled with --fmad=false, the host harness with -ffp-contract=off) radii, tile
// rectangles and depth keys are bit-identical to the oracle.  Backward math is our own
// derivation of the chain rule of A1, validated against oracle autograd.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TGS_HD __host__ __device__ __forceinline__
#else
#define TGS_HD inline
#endif

#define TGS_TILE 16
#define TGS_NEAR_Z 0.2f
#define TGS_COV_BLUR 0.3f
#define TGS_ALPHA_MAX 0.99f
#define TGS_ALPHA_MIN (1.0f / 255.0f)
#define TGS_T_EPS 0.0001f
#define TGS_NGRAD 10   // screen-space gradient slots per Gaussian

#define TGS_SH_C0 0.28209479177387814f
#define TGS_SH_C1 0.4886025119029199f
#define TGS_SH_C2_0 1.0925484305920792f
#define TGS_SH_C2_1 -1.0925484305920792f
#define TGS_SH_C2_2 0.31539156525252005f
#define TGS_SH_C2_3 -1.0925484305920792f
#define TGS_SH_C2_4 0.5462742152960396f
#define TGS_SH_C3_0 -0.5900435899266435f
#define TGS_SH_C3_1 2.890611442640554f
#define TGS_SH_C3_2 -0.4570457994644658f
#define TGS_SH_C3_3 0.3731763325901154f
#define TGS_SH_C3_4 -0.4570457994644658f
#define TGS_SH_C3_5 1.445305721320277f
#define TGS_SH_C3_6 -0.5900435899266435f

// Host-computed camera constants (fp32, same expressions as oracle camera_scalars()).
struct TgsCam {
    float fx, fy, limx, limy;
    float mod;
    int W, H, Tx, Ty, row0, row1;
    int deg, K;
    float near_z;          // near-plane cull threshold (TGS_NEAR_Z unless the caller overrides it)
    float ppx, ppy;        // principal point offset from the image centre, pixels
    float alpha_max;       // alpha clamp (TGS_ALPHA_MAX unless the caller overrides it)
};

// ((m[0][col]*x + m[1][col]*y) + m[2][col]*z) + m[3][col], m = transposed 4x4 (row-major [k][j])
TGS_HD float tgs_xform(const float* m, float x, float y, float z, int col) {
    return ((m[col] * x + m[4 + col] * y) + m[8 + col] * z) + m[12 + col];
}

TGS_HD float tgs_clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// Sigma = R diag(mod*s)^2 R^T, 6 upper-triangular floats.
TGS_HD void tgs_cov3d(const float* scale, float mod, const float* q, float* cov) {
    float sx = mod * scale[0], sy = mod * scale[1], sz = mod * scale[2];
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R00 = 1.0f - 2.0f * (y * y + z * z), R01 = 2.0f * (x * y - r * z), R02 = 2.0f * (x * z + r * y);
    float R10 = 2.0f * (x * y + r * z), R11 = 1.0f - 2.0f * (x * x + z * z), R12 = 2.0f * (y * z - r * x);
    float R20 = 2.0f * (x * z - r * y), R21 = 2.0f * (y * z + r * x), R22 = 1.0f - 2.0f * (x * x + y * y);
    float M00 = R00 * sx, M01 = R01 * sy, M02 = R02 * sz;
    float M10 = R10 * sx, M11 = R11 * sy, M12 = R12 * sz;
    float M20 = R20 * sx, M21 = R21 * sy, M22 = R22 * sz;
    cov[0] = (M00 * M00 + M01 * M01) + M02 * M02;
    cov[1] = (M00 * M10 + M01 * M11) + M02 * M12;
    cov[2] = (M00 * M20 + M01 * M21) + M02 * M22;
    cov[3] = (M10 * M10 + M11 * M11) + M12 * M12;
    cov[4] = (M10 * M20 + M11 * M21) + M12 * M22;
    cov[5] = (M20 * M20 + M21 * M21) + M22 * M22;
}

// Intermediates of the EWA projection shared by forward and backward.
struct TgsEwa {
    float tx, ty, tz;      // view-space mean
    float cx, cy;          // fov-clamped tx, ty
    bool inx, iny;         // clamp inactive
    float m0[3], m1[3];    // rows of J*W
    float u[3], v[3];      // Sigma*m0, Sigma*m1
    float a, b, c;         // 2D covariance incl. blur
};

TGS_HD void tgs_ewa(const float* vm, const TgsCam& cam, float tx, float ty, float tz,
                    const float* cov, TgsEwa& e) {
    e.tx = tx; e.ty = ty; e.tz = tz;
    float qx = tx / tz, qy = ty / tz;
    e.inx = (qx >= -cam.limx) && (qx <= cam.limx);
    e.iny = (qy >= -cam.limy) && (qy <= cam.limy);
    e.cx = tgs_clampf(qx, -cam.limx, cam.limx) * tz;
    e.cy = tgs_clampf(qy, -cam.limy, cam.limy) * tz;
    float itz = 1.0f / tz;
    float J00 = cam.fx * itz, J11 = cam.fy * itz;
    float tz2 = tz * tz;
    float J02 = -(cam.fx * e.cx) / tz2;
    float J12 = -(cam.fy * e.cy) / tz2;
    for (int k = 0; k < 3; ++k) {           // W[r][k] = vm[k][r]
        e.m0[k] = J00 * vm[4 * k + 0] + J02 * vm[4 * k + 2];
        e.m1[k] = J11 * vm[4 * k + 1] + J12 * vm[4 * k + 2];
    }
    const float S[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
    for (int k = 0; k < 3; ++k) {
        e.u[k] = (S[k][0] * e.m0[0] + S[k][1] * e.m0[1]) + S[k][2] * e.m0[2];
        e.v[k] = (S[k][0] * e.m1[0] + S[k][1] * e.m1[1]) + S[k][2] * e.m1[2];
    }
    e.a = ((e.m0[0] * e.u[0] + e.m0[1] * e.u[1]) + e.m0[2] * e.u[2]) + TGS_COV_BLUR;
    e.b = (e.m0[0] * e.v[0] + e.m0[1] * e.v[1]) + e.m0[2] * e.v[2];
    e.c = ((e.m1[0] * e.v[0] + e.m1[1] * e.v[1]) + e.m1[2] * e.v[2]) + TGS_COV_BLUR;
}

struct TgsProj {
    float px, py, depth;
    float conA, conB, conC;
    int radius;                       // 0 = invisible in the full image
    int rminx, rminy, rmaxx, rmaxy;   // tile rect clipped to the band
    int tiles;                        // tiles touched in the band
};

TGS_HD void tgs_rect1(float p, float r, int n, int& lo, int& hi) {
    float flo = tgs_clampf((p - r) / (float)TGS_TILE, 0.0f, (float)n);
    float fhi = tgs_clampf(((p + r) + (float)(TGS_TILE - 1)) / (float)TGS_TILE, 0.0f, (float)n);
    lo = (flo == flo) ? (int)flo : 0;   // NaN -> 0 (oracle: nan_to_num)
    hi = (fhi == fhi) ? (int)fhi : 0;
}

// A1 geometry: returns false (and radius = 0, tiles = 0) when culled.  `cov` must hold the 3D
// covariance (computed by the caller via tgs_cov3d or taken from cov3D_precomp).
TGS_HD bool tgs_project(const float* vm, const float* pm, const TgsCam& cam,
                        float x, float y, float z, const float* cov, TgsProj& o) {
    o.radius = 0; o.tiles = 0; o.rminx = o.rminy = o.rmaxx = o.rmaxy = 0;
    o.px = o.py = o.depth = 0.0f; o.conA = o.conB = o.conC = 0.0f;
    float tz = tgs_xform(vm, x, y, z, 2);
    if (!(tz > cam.near_z)) return false;
    float tx = tgs_xform(vm, x, y, z, 0), ty = tgs_xform(vm, x, y, z, 1);
    float hx = tgs_xform(pm, x, y, z, 0), hy = tgs_xform(pm, x, y, z, 1), hw = tgs_xform(pm, x, y, z, 3);
    float pw = 1.0f / (hw + 0.0000001f);
    float ndcx = hx * pw, ndcy = hy * pw;
    TgsEwa e;
    tgs_ewa(vm, cam, tx, ty, tz, cov, e);
    float det = e.a * e.c - e.b * e.b;
    if (det == 0.0f) return false;
    float det_inv = 1.0f / det;
    float mid = 0.5f * (e.a + e.c);
    float disc = sqrtf(fmaxf(mid * mid - det, 0.1f));
    float lam = fmaxf(mid + disc, mid - disc);
    float rad = ceilf(3.0f * sqrtf(fmaxf(lam, 0.0f)));
    float px = ((ndcx + 1.0f) * (float)cam.W - 1.0f) * 0.5f + cam.ppx;
    float py = ((ndcy + 1.0f) * (float)cam.H - 1.0f) * 0.5f + cam.ppy;
    int x0, x1, y0, y1;
    tgs_rect1(px, rad, cam.Tx, x0, x1);
    tgs_rect1(py, rad, cam.Ty, y0, y1);
    if ((x1 - x0) * (y1 - y0) <= 0) return false;
    // visible in the full image from here on
    o.radius = (int)rad;
    y0 = y0 < cam.row0 ? cam.row0 : (y0 > cam.row1 ? cam.row1 : y0);
    y1 = y1 < cam.row0 ? cam.row0 : (y1 > cam.row1 ? cam.row1 : y1);
    o.rminx = x0; o.rmaxx = x1; o.rminy = y0; o.rmaxy = y1;
    o.tiles = (x1 - x0) * (y1 - y0);
    o.px = px; o.py = py; o.depth = tz;
    o.conA = e.c * det_inv; o.conB = -e.b * det_inv; o.conC = e.a * det_inv;
    return true;
}

// --------------------------------------------------------------------------------- SH colour
// sh: [K][3] for this Gaussian.  Returns rgb (+0.5, clamped >= 0) and the clamp mask (bit c).
// the 16 real-SH basis values at direction (dx,dy,dz)/|.| (zero above the active degree)
TGS_HD void tgs_sh_bases(int deg, float dx, float dy, float dz, float* b) {
    float ln = sqrtf((dx * dx + dy * dy) + dz * dz);
    float x = dx / ln, y = dy / ln, z = dz / ln;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 16; ++k) b[k] = 0.0f;
    b[0] = TGS_SH_C0;
    if (deg > 0) {
        b[1] = -TGS_SH_C1 * y; b[2] = TGS_SH_C1 * z; b[3] = -TGS_SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = TGS_SH_C2_0 * xy; b[5] = TGS_SH_C2_1 * yz; b[6] = TGS_SH_C2_2 * (2.0f * zz - xx - yy);
            b[7] = TGS_SH_C2_3 * xz; b[8] = TGS_SH_C2_4 * (xx - yy);
            if (deg > 2) {
                b[9] = TGS_SH_C3_0 * y * (3.0f * xx - yy);
                b[10] = TGS_SH_C3_1 * xy * z;
                b[11] = TGS_SH_C3_2 * y * (4.0f * zz - xx - yy);
                b[12] = TGS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = TGS_SH_C3_4 * x * (4.0f * zz - xx - yy);
                b[14] = TGS_SH_C3_5 * z * (xx - yy);
                b[15] = TGS_SH_C3_6 * x * (xx - 3.0f * yy);
            }
        }
    }
}

TGS_HD void tgs_sh_forward(int deg, const float* sh, float dx, float dy, float dz,
                           float* rgb, unsigned& clamped) {
    // NOTE: `sh` must hold 48 floats with zeros above the active degree: all loops below run over
    // the full 16 bases with compile-time indices so that device code keeps everything in registers.
    float b[16];
    tgs_sh_bases(deg, dx, dy, dz, b);
    clamped = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 3; ++c) {
        float acc = 0.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 16; ++k) acc += b[k] * sh[3 * k + c];
        acc += 0.5f;
        if (acc < 0.0f) { clamped |= (1u << c); acc = 0.0f; }
        rgb[c] = acc;
    }
}

// Basis function k of the real SH expansion at unit direction (x,y,z) and its gradient.
// `k` is a compile-time constant at every call site (fully unrolled loops), so the switch folds.
TGS_HD void tgs_sh_basis(int k, float x, float y, float z, float& b, float& bx, float& by, float& bz) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b = bx = by = bz = 0.0f;
    switch (k) {
        case 0: b = TGS_SH_C0; break;
        case 1: b = -TGS_SH_C1 * y; by = -TGS_SH_C1; break;
        case 2: b = TGS_SH_C1 * z; bz = TGS_SH_C1; break;
        case 3: b = -TGS_SH_C1 * x; bx = -TGS_SH_C1; break;
        case 4: b = TGS_SH_C2_0 * xy; bx = TGS_SH_C2_0 * y; by = TGS_SH_C2_0 * x; break;
        case 5: b = TGS_SH_C2_1 * yz; by = TGS_SH_C2_1 * z; bz = TGS_SH_C2_1 * y; break;
        case 6: b = TGS_SH_C2_2 * (2.0f * zz - xx - yy);
                bx = TGS_SH_C2_2 * -2.0f * x; by = TGS_SH_C2_2 * -2.0f * y; bz = TGS_SH_C2_2 * 4.0f * z; break;
        case 7: b = TGS_SH_C2_3 * xz; bx = TGS_SH_C2_3 * z; bz = TGS_SH_C2_3 * x; break;
        case 8: b = TGS_SH_C2_4 * (xx - yy); bx = TGS_SH_C2_4 * 2.0f * x; by = TGS_SH_C2_4 * -2.0f * y; break;
        case 9: b = TGS_SH_C3_0 * y * (3.0f * xx - yy);
                bx = TGS_SH_C3_0 * 6.0f * xy; by = TGS_SH_C3_0 * (3.0f * xx - 3.0f * yy); break;
        case 10: b = TGS_SH_C3_1 * xy * z; bx = TGS_SH_C3_1 * yz; by = TGS_SH_C3_1 * xz; bz = TGS_SH_C3_1 * xy; break;
        case 11: b = TGS_SH_C3_2 * y * (4.0f * zz - xx - yy);
                 bx = TGS_SH_C3_2 * -2.0f * xy; by = TGS_SH_C3_2 * (4.0f * zz - xx - 3.0f * yy);
                 bz = TGS_SH_C3_2 * 8.0f * yz; break;
        case 12: b = TGS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                 bx = TGS_SH_C3_3 * -6.0f * xz; by = TGS_SH_C3_3 * -6.0f * yz;
                 bz = TGS_SH_C3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy); break;
        case 13: b = TGS_SH_C3_4 * x * (4.0f * zz - xx - yy);
                 bx = TGS_SH_C3_4 * (4.0f * zz - 3.0f * xx - yy); by = TGS_SH_C3_4 * -2.0f * xy;
                 bz = TGS_SH_C3_4 * 8.0f * xz; break;
        case 14: b = TGS_SH_C3_5 * z * (xx - yy);
                 bx = TGS_SH_C3_5 * 2.0f * xz; by = TGS_SH_C3_5 * -2.0f * yz; bz = TGS_SH_C3_5 * (xx - yy); break;
        case 15: b = TGS_SH_C3_6 * x * (xx - 3.0f * yy);
                 bx = TGS_SH_C3_6 * (3.0f * xx - 3.0f * yy); by = TGS_SH_C3_6 * -6.0f * xy; break;
        default: break;
    }
}

// Backward of 4 consecutive coefficients (k = 4*grp .. 4*grp+3): sh12 / dsh12 hold [4][3] floats.
// nb = number of ACTIVE coefficients; inactive ones get zero gradient and add nothing to ddir.
// Streaming the 16 coefficients in 4 groups keeps the register footprint small (12 + 12 live floats).
TGS_HD void tgs_sh_backward_group(int grp, int nb, float x, float y, float z, const float* sh12,
                                  const float* g, float* dsh12, float& ddx, float& ddy, float& ddz) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) {
        const int k = 4 * grp + i;
        float b, bx, by, bz;
        tgs_sh_basis(k, x, y, z, b, bx, by, bz);
        const bool act = k < nb;
        const float s = act ? (sh12[3 * i] * g[0] + sh12[3 * i + 1] * g[1] + sh12[3 * i + 2] * g[2]) : 0.0f;
        ddx += bx * s; ddy += by * s; ddz += bz * s;
        dsh12[3 * i] = act ? b * g[0] : 0.0f;
        dsh12[3 * i + 1] = act ? b * g[1] : 0.0f;
        dsh12[3 * i + 2] = act ? b * g[2] : 0.0f;
    }
}

// dL/dsh [K][3] (fully written: zero above the active degree) and dL/d(mean) through the view
// direction.  drgb is zeroed where the forward clamped.  `sh` / `dsh` may be global pointers: they are
// read / written in groups of 4 coefficients.
TGS_HD void tgs_sh_backward(int deg, int K, const float* sh, float dx, float dy, float dz,
                            const float* drgb_in, unsigned clamped, float* dsh, float* dmean,
                            bool sh_in_global = true) {   // false: `sh` points into shared memory (no ld.global.nc)
    float g[3];
    for (int c = 0; c < 3; ++c) g[c] = ((clamped >> c) & 1u) ? 0.0f : drgb_in[c];
    const float ln = sqrtf((dx * dx + dy * dy) + dz * dz);
    const float inv = 1.0f / ln;
    const float x = dx * inv, y = dy * inv, z = dz * inv;
    const int nb = (deg + 1) * (deg + 1);
    float ddx = 0.0f, ddy = 0.0f, ddz = 0.0f;
    const bool vec = ((3 * K) & 3) == 0;           // K*12 bytes is a multiple of 16: float4 path
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int grp = 0; grp < 4; ++grp) {
        if (4 * grp >= K) break;
        float s12[12], d12[12];
        if (vec) {
#if defined(__CUDA_ARCH__)
            if (4 * grp < nb) {
                const float4* p = reinterpret_cast<const float4*>(sh + 12 * grp);
                float4 v0, v1, v2;
                if (sh_in_global) { v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2); }
                else { v0 = p[0]; v1 = p[1]; v2 = p[2]; }
                s12[0] = v0.x; s12[1] = v0.y; s12[2] = v0.z; s12[3] = v0.w; s12[4] = v1.x; s12[5] = v1.y;
                s12[6] = v1.z; s12[7] = v1.w; s12[8] = v2.x; s12[9] = v2.y; s12[10] = v2.z; s12[11] = v2.w;
            } else {
#pragma unroll
                for (int i = 0; i < 12; ++i) s12[i] = 0.0f;
            }
#else
            for (int i = 0; i < 12; ++i) s12[i] = (4 * grp < nb) ? sh[12 * grp + i] : 0.0f;
#endif
        } else {
            for (int i = 0; i < 12; ++i) s12[i] = (12 * grp + i < 3 * K && 4 * grp + i / 3 < nb) ? sh[12 * grp + i] : 0.0f;
        }
        tgs_sh_backward_group(grp, nb, x, y, z, s12, g, d12, ddx, ddy, ddz);
        if (vec) {
#if defined(__CUDA_ARCH__)
            float4* q = reinterpret_cast<float4*>(dsh + 12 * grp);
            q[0] = make_float4(d12[0], d12[1], d12[2], d12[3]);
            q[1] = make_float4(d12[4], d12[5], d12[6], d12[7]);
            q[2] = make_float4(d12[8], d12[9], d12[10], d12[11]);
#else
            for (int i = 0; i < 12; ++i) dsh[12 * grp + i] = d12[i];
#endif
        } else {
            for (int i = 0; i < 12; ++i)
                if (12 * grp + i < 3 * K) dsh[12 * grp + i] = d12[i];
        }
    }
    // d = raw/|raw| :  dL/draw = (dd - d (d.dd)) / |raw|
    const float dot = x * ddx + y * ddy + z * ddz;
    dmean[0] += (ddx - x * dot) * inv;
    dmean[1] += (ddy - y * dot) * inv;
    dmean[2] += (ddz - z * dot) * inv;
}

// ------------------------------------------------------------------------ geometry backward
// sg = screen grads (dx,dy [pixel units], dA,dB,dC, dopacity, dr,dg,db, ddepth).
// Accumulates into dmean[3]; writes dcov[6].
TGS_HD void tgs_project_backward(const float* vm, const float* pm, const TgsCam& cam,
                                 float x, float y, float z, const float* cov,
                                 const float* sg, float* dmean, float* dcov) {
    float tx = tgs_xform(vm, x, y, z, 0), ty = tgs_xform(vm, x, y, z, 1), tz = tgs_xform(vm, x, y, z, 2);
    TgsEwa e;
    tgs_ewa(vm, cam, tx, ty, tz, cov, e);
    float det = e.a * e.c - e.b * e.b;
    float det_inv = 1.0f / det;
    float A = e.c * det_inv, B = -e.b * det_inv, C = e.a * det_inv;
    // dL/dSigma2 = -Q G Q with Q = [[A,B],[B,C]], G = [[gA, gB/2],[gB/2, gC]]
    float gA = sg[2], gB = 0.5f * sg[3], gC = sg[4];
    float q00 = A * gA + B * gB, q01 = A * gB + B * gC;
    float q10 = B * gA + C * gB, q11 = B * gB + C * gC;
    float da = -(q00 * A + q01 * B);
    float db = -2.0f * (q00 * B + q01 * C);
    float dc = -(q10 * B + q11 * C);
    // Sigma2 = [m0;m1] Sigma [m0;m1]^T (+blur)
    dcov[0] = da * e.m0[0] * e.m0[0] + db * e.m0[0] * e.m1[0] + dc * e.m1[0] * e.m1[0];
    dcov[3] = da * e.m0[1] * e.m0[1] + db * e.m0[1] * e.m1[1] + dc * e.m1[1] * e.m1[1];
    dcov[5] = da * e.m0[2] * e.m0[2] + db * e.m0[2] * e.m1[2] + dc * e.m1[2] * e.m1[2];
    dcov[1] = 2.0f * da * e.m0[0] * e.m0[1] + db * (e.m0[0] * e.m1[1] + e.m0[1] * e.m1[0]) + 2.0f * dc * e.m1[0] * e.m1[1];
    dcov[2] = 2.0f * da * e.m0[0] * e.m0[2] + db * (e.m0[0] * e.m1[2] + e.m0[2] * e.m1[0]) + 2.0f * dc * e.m1[0] * e.m1[2];
    dcov[4] = 2.0f * da * e.m0[1] * e.m0[2] + db * (e.m0[1] * e.m1[2] + e.m0[2] * e.m1[1]) + 2.0f * dc * e.m1[1] * e.m1[2];
    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
    for (int k = 0; k < 3; ++k) {
        float dm0 = 2.0f * da * e.u[k] + db * e.v[k];
        float dm1 = db * e.u[k] + 2.0f * dc * e.v[k];
        dJ00 += dm0 * vm[4 * k + 0]; dJ02 += dm0 * vm[4 * k + 2];
        dJ11 += dm1 * vm[4 * k + 1]; dJ12 += dm1 * vm[4 * k + 2];
    }
    float itz = 1.0f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    float dtx = e.inx ? -cam.fx * itz2 * dJ02 : 0.0f;
    float dty = e.iny ? -cam.fy * itz2 * dJ12 : 0.0f;
    float dtz = -cam.fx * itz2 * dJ00 - cam.fy * itz2 * dJ11
              + 2.0f * cam.fx * e.cx * itz3 * dJ02 + 2.0f * cam.fy * e.cy * itz3 * dJ12;
    dtz += sg[9];                                   // expected-depth channel: depth = t.z
    for (int k = 0; k < 3; ++k)
        dmean[k] += vm[4 * k + 0] * dtx + vm[4 * k + 1] * dty + vm[4 * k + 2] * dtz;
    // pixel mean: px = ((hx*pw + 1) W - 1)/2
    float hx = tgs_xform(pm, x, y, z, 0), hy = tgs_xform(pm, x, y, z, 1), hw = tgs_xform(pm, x, y, z, 3);
    float pw = 1.0f / (hw + 0.0000001f);
    float dndx = sg[0] * 0.5f * (float)cam.W, dndy = sg[1] * 0.5f * (float)cam.H;
    float dhx = dndx * pw, dhy = dndy * pw;
    float dhw = -(dndx * hx + dndy * hy) * pw * pw;
    for (int k = 0; k < 3; ++k)
        dmean[k] += pm[4 * k + 0] * dhx + pm[4 * k + 1] * dhy + pm[4 * k + 3] * dhw;
}

// dL/dcov3D -> dL/dscale, dL/dquat (quaternion used as given, not normalised)
TGS_HD void tgs_cov3d_backward(const float* scale, float mod, const float* q, const float* dcov,
                               float* dscale, float* dq) {
    float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[3][3] = {
        {1.0f - 2.0f * (y * y + z * z), 2.0f * (x * y - r * z), 2.0f * (x * z + r * y)},
        {2.0f * (x * y + r * z), 1.0f - 2.0f * (x * x + z * z), 2.0f * (y * z - r * x)},
        {2.0f * (x * z - r * y), 2.0f * (y * z + r * x), 1.0f - 2.0f * (x * x + y * y)}};
    // symmetric gradient matrix (off-diagonals carry half of the unique-entry gradient)
    float G[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                     {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                     {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
    float dR[3][3];
    for (int j = 0; j < 3; ++j) {
        float ds = 0.0f;
        for (int i = 0; i < 3; ++i) {
            // dM[i][j] = 2 * sum_k G[i][k] * M[k][j],  M[k][j] = R[k][j]*s[j]
            float dM = 2.0f * (G[i][0] * R[0][j] + G[i][1] * R[1][j] + G[i][2] * R[2][j]) * s[j];
            ds += dM * R[i][j];
            dR[i][j] = dM * s[j];
        }
        dscale[j] = mod * ds;
    }
    dq[0] = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
    dq[1] = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - r * dR[1][2]
                    + z * dR[2][0] + r * dR[2][1] - 2.0f * x * dR[2][2]);
    dq[2] = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + r * dR[0][2] + x * dR[1][0] + z * dR[1][2]
                    - r * dR[2][0] + z * dR[2][1] - 2.0f * y * dR[2][2]);
    dq[3] = 2.0f * (-2.0f * z * dR[0][0] - r * dR[0][1] + x * dR[0][2] + r * dR[1][0] - 2.0f * z * dR[1][1]
                    + y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
}
