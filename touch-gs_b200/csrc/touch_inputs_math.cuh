// touch_inputs_math.cuh -- per-pixel fp64 math of the touch / vision fusion (see touch_inputs.cu), usable
// from device code and, for the CPU tests only (tests/host_math_harness.cpp), from host code.
#pragma once
#include <math.h>
#include <stdint.h>
#if defined(__CUDACC__)
#define TGS_FHD __host__ __device__ __forceinline__
#else
#define TGS_FHD inline
#endif

struct FuseParams {
    double scale, offset, offset2, unit;   // unit = 1e-3 * scene_scale
    int real_world;
};

struct PixelOut { unsigned short va, ds, fu, fs; float target, weight; };

// numpy's float64 -> uint16 ``astype`` (reference :373-376) is a C cast: on x86-64 it truncates toward zero
// into a signed integer and keeps the low 16 bits, so NEGATIVE values WRAP (the baseline map is negative where
// the vision depth is 0: -464 mm is stored as 65072).  A direct f64->u16 conversion on the GPU saturates to 0,
// so go through a signed int to stay byte-identical with the PNGs the reference writes.
TGS_FHD unsigned short enc_mm(double x) { return (unsigned short)(int)(x * 1000.0); }

TGS_FHD PixelOut fuse_pixel(unsigned short touch_mm, unsigned short vision_mm,
                                               unsigned short tsig_mm, const FuseParams& p) {
    const double t = (double)touch_mm / 1000.0, vv = (double)vision_mm / 1000.0, ts = (double)tsig_mm / 1000.0;
    double v = (p.scale * vv) + p.offset;                     // :288
    const double ds = v;                                      // :291
    double diff = v - t;                                      // :294
    if (diff > 3.0) diff = 0.0;                               // :295
    const double t2a = p.real_world ? t * ((diff > 0.0) ? 1.0 : 0.0) : t;   // :297
    if (t2a > 0.0) v = v + p.offset2;                         // :298,:304
    v = fmax(v, 0.0);                                         // :306
    const double vs = fmin(fmax(v * 0.05, 0.0), 10.0) + 5.0;  // create_uncertainty_from_depth.py:21, :312-313
    // vs is always in [5, 15]: 1/vs and v/vs are finite, the reference's inf / nan patches (:121,:144) never fire.
    const double rv = 1.0 / vs;                               // :116
    const double mu_v = v / vs;                               // :143
    // touch side.  ts == 0 (no touch here: ~90 % of the pixels): mask = 0, 1/0 = inf -> 0 (:117,:120) and
    // (t*0)/0 = nan -> 0 (:136-141), so both divisions can be skipped with identical results.
    double rt = 0.0, mu_t = 0.0;
    if (ts > 0.0) { rt = 1.0 / ts; mu_t = (t * 1.0) / ts; }   // :109,:117,:136,:140 (finite: no patch fires)
    double sigma = 1.0 / (rt + rv);                           // :124  (rt + rv >= 1/15 > 0: never inf, :126)
    double fused = sigma * (mu_t + mu_v);                     // :146
    fused = fmax(fused, 0.0);                                 // :360
    sigma = fmin(fmax(sigma, 0.0), 10.0);                     // :361
    PixelOut o;
    o.va = enc_mm(v); o.ds = enc_mm(ds); o.fu = enc_mm(fused); o.fs = enc_mm(sigma);   // :373-376
    o.target = (float)((double)o.fu * p.unit);                // trainer-side decode
    const double sg = (double)o.fs / 1000.0;
    o.weight = (float)((sg > 0.0) ? 1.0 / sg : 0.0);
    return o;
}

