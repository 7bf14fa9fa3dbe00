// screen_api.cu -- the hot path split at the screen-space boundary, for trainers written against the
// gsplat-0.1-style three-call API: project_gaussians -> (the caller's own
// colour code, e.g. spherical_harmonics) -> rasterize_gaussians.  The nerfstudio splat model of early 2024 that the
// Touch-GS fork builds on (reference .gitmodules:7-9, scripts/train_bunny_real.sh:52) calls exactly these three.
// Everything reuses the kernels of the fused path (preprocess.cu, binning.cu, render.cu); only the glue kernels that
// build per-Gaussian records from caller-provided screen-space tensors and the stand-alone SH evaluation live here.
// Convention differences (alpha clamp 0.999, near plane = clip_thresh, principal point, pixel-centre offset) are
// runtime switches of TgsSettings / arguments, never separate code paths.
#include "tgs_common.cuh"

namespace {

// records + binning inputs from caller-provided screen-space data (what preprocess.cu produces in the fused path)
__global__ void __launch_bounds__(256)
k_screen_records(int N, const float* __restrict__ xys, const float* __restrict__ depths, const int32_t* __restrict__ radii,
                 const float* __restrict__ conics, const float* __restrict__ colors, const float* __restrict__ opac,
                 float pixel_offset, TgsCam cam, TgsRecord* __restrict__ rec, uint32_t* __restrict__ tiles,
                 uint2* __restrict__ rect, uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ ids) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float px = xys[2 * i], py = xys[2 * i + 1], d = depths[i];
    const int r = radii[i];
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, cnt = 0;
    if (r > 0) {
        tgs_rect1(px, (float)r, cam.Tx, x0, x1);
        tgs_rect1(py, (float)r, cam.Ty, y0, y1);
        y0 = y0 < cam.row0 ? cam.row0 : (y0 > cam.row1 ? cam.row1 : y0);
        y1 = y1 < cam.row0 ? cam.row0 : (y1 > cam.row1 ? cam.row1 : y1);
        cnt = (x1 - x0) * (y1 - y0);
        if (cnt < 0) cnt = 0;
    }
    const float o = cnt > 0 ? opac[i] : 0.0f;
    TgsRecord R;
    // the compositing kernels sample at integer pixel coordinates: a sample point (x + off, y + off) is the same as
    // integer sampling of a splat moved by -off
    R.a = make_float4(px - pixel_offset, py - pixel_offset, d, __int_as_float(i));
    R.b = make_float4(conics[3 * i], conics[3 * i + 1], conics[3 * i + 2], o);
    R.c = make_float4(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2], cnt > 0 ? -logf(255.0f * o) : 3.0e38f);
    rec[i] = R;
    tiles[i] = (uint32_t)cnt;
    rect[i] = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)y0 | ((uint32_t)y1 << 16));
    depth_keys[i] = cnt > 0 ? __float_as_uint(d) : 0xFFFFFFFFu;
    ids[i] = (uint32_t)i;
}

// colours[n,c] = sum_k basis_k(dir_n) * coeffs[n,k,c]  (raw: the caller adds 0.5 and clamps)
__global__ void __launch_bounds__(256)
k_sh_eval(int N, int nb, int K, const float* __restrict__ dirs, const float* __restrict__ coeffs, float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    float acc[3] = {0.f, 0.f, 0.f};
    const float* c = coeffs + (size_t)3 * K * i;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < nb) {
            float b, bx, by, bz;
            tgs_sh_basis(k, dx, dy, dz, b, bx, by, bz);
            acc[0] += b * c[3 * k]; acc[1] += b * c[3 * k + 1]; acc[2] += b * c[3 * k + 2];
        }
    }
    out[3 * i] = acc[0]; out[3 * i + 1] = acc[1]; out[3 * i + 2] = acc[2];
}
__global__ void __launch_bounds__(256)
k_sh_eval_bwd(int N, int nb, int K, const float* __restrict__ dirs, const float* __restrict__ v_out,
              float* __restrict__ v_coeffs) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
    dx *= inv; dy *= inv; dz *= inv;
    const float g0 = v_out[3 * i], g1 = v_out[3 * i + 1], g2 = v_out[3 * i + 2];
    float* c = v_coeffs + (size_t)3 * K * i;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (k < K) {
            float b = 0.f, bx, by, bz;
            if (k < nb) tgs_sh_basis(k, dx, dy, dz, b, bx, by, bz);
            c[3 * k] = b * g0; c[3 * k + 1] = b * g1; c[3 * k + 2] = b * g2;
        }
    }
}

int check_geometry(const TgsSettings* s, const TgsGaussians* g, const char* who) {
    if (!s || !g) { tgs_set_error("%s: NULL settings / gaussians", who); return TGS_EINVAL; }
    if (s->image_width <= 0 || s->image_height <= 0 || g->N < 0) { tgs_set_error("%s: bad sizes", who); return TGS_EINVAL; }
    if (!s->viewmatrix || !s->projmatrix) { tgs_set_error("%s: viewmatrix / projmatrix must be non-NULL", who); return TGS_EINVAL; }
    if (g->N > 0) {
        if (!g->means3D) { tgs_set_error("%s: means3D must be non-NULL", who); return TGS_EINVAL; }
        const bool sr = g->scales && g->rotations;
        if (sr == (g->cov3D_precomp != nullptr) || ((g->scales || g->rotations) && !sr)) {
            tgs_set_error("%s: provide exactly one of scale/rotation pair or precomputed 3D covariance", who); return TGS_EINVAL; }
    }
    return 0;
}

}  // namespace

extern "C" int tgs_project_gaussians(const TgsSettings* s, const TgsGaussians* g, tgs_alloc_fn alloc, void* user,
                                     int32_t* radii, TgsSaved* saved, void* stream) {
    int rc = check_geometry(s, g, "tgs_project_gaussians");
    if (rc) return rc;
    if (!alloc || !saved || (g->N > 0 && !radii)) { tgs_set_error("tgs_project_gaussians: NULL output / allocator"); return TGS_EINVAL; }
    const TgsCam cam = tgs_make_cam(s);
    TgsGeomLayout gl; tgs_geom_layout(g->N, &gl);
    void* geom = alloc(user, TGS_BUF_GEOM, gl.total);
    if (!geom) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    TgsGaussians gg = *g;
    gg.shs = nullptr; gg.colors_precomp = nullptr; gg.opacities = nullptr;     // geometry only
    if (g->N > 0) { rc = tgs_launch_preprocess(cam, s, &gg, tgs_geom_view(geom, g->N), radii, (cudaStream_t)stream); if (rc) return rc; }
    saved->geom = geom; saved->binning = nullptr; saved->image = nullptr; saved->num_rendered = 0; saved->capacity = 0;
    return 0;
}

extern "C" int tgs_project_gaussians_backward(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                              const int32_t* radii, const float* screen_grads, const TgsGrads* grads,
                                              void* stream) {
    int rc = check_geometry(s, g, "tgs_project_gaussians_backward");
    if (rc) return rc;
    if (g->N == 0) return 0;
    if (!saved || !saved->geom) { tgs_set_error("tgs_project_gaussians_backward: saved geometry missing"); return TGS_ESTATE; }
    if (!radii || !screen_grads || !grads || !grads->dmeans2D || !grads->dmeans3D || !grads->dopacity) {
        tgs_set_error("tgs_project_gaussians_backward: NULL gradient buffers"); return TGS_EINVAL; }
    if (g->scales && (!grads->dscales || !grads->drotations)) { tgs_set_error("dscales/drotations required"); return TGS_EINVAL; }
    if (g->cov3D_precomp && !grads->dcov3D) { tgs_set_error("dcov3D required when cov3D_precomp given"); return TGS_EINVAL; }
    const TgsCam cam = tgs_make_cam(s);
    TgsGaussians gg = *g;
    gg.shs = nullptr; gg.colors_precomp = nullptr;
    TgsGrads gr = *grads;
    gr.dshs = nullptr; gr.dcolors = nullptr;
    return tgs_launch_preprocess_bwd(cam, s, &gg, tgs_geom_view(saved->geom, g->N), radii, screen_grads, nullptr, nullptr, 0,
                                     false, &gr, (cudaStream_t)stream);
}

extern "C" int tgs_rasterize_screen_forward(const TgsSettings* s, int32_t N, const float* xys, const float* depths,
                                            const int32_t* radii, const float* conics, const float* colors,
                                            const float* opacities, float pixel_offset, tgs_alloc_fn alloc, void* user,
                                            float* out_color, float* out_depth, float* out_alpha, TgsSaved* saved,
                                            void* stream) {
    if (!s || !alloc || !saved || !out_color || !out_depth || !out_alpha || !s->bg || s->image_width <= 0 || s->image_height <= 0 || N < 0) {
        tgs_set_error("tgs_rasterize_screen_forward: bad arguments"); return TGS_EINVAL; }
    if (N > 0 && (!xys || !depths || !radii || !conics || !colors || !opacities)) {
        tgs_set_error("tgs_rasterize_screen_forward: NULL per-Gaussian tensor"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    TgsGeomLayout gl; tgs_geom_layout(N, &gl);
    TgsImageLayout il; tgs_image_layout(cam.W, cam.H, &il);
    void* geom = alloc(user, TGS_BUF_GEOM, gl.total);
    void* image = alloc(user, TGS_BUF_IMAGE, il.total);
    if (!geom || !image) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    GeomView gv = tgs_geom_view(geom, N);
    ImageView iv = tgs_image_view(image, cam.W, cam.H);
    int64_t I = 0;
    void* temp = nullptr;
    if (N > 0) {
        k_screen_records<<<(N + 255) / 256, 256, 0, st>>>(N, xys, depths, radii, conics, colors, opacities, pixel_offset, cam,
                                                          gv.records, gv.tiles_touched, gv.rect, gv.depth_keys, gv.ids);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
        int rc = tgs_depth_order(gv, N, st); if (rc) return rc;
        temp = alloc(user, TGS_BUF_TEMP, tgs_bin_temp_bytes(N, cam.Tx, cam.Ty));
        if (!temp) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    }
    int rc = tgs_bin_count(gv, N, cam.Tx, cam.Ty, cam.row0, cam.row1, temp, iv.ranges, iv.count, st); if (rc) return rc;
    if (N > 0) {
        uint32_t h_I[2] = {0, 0};
        TGS_CUDA(cudaMemcpyAsync(h_I, iv.count, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TGS_CUDA(cudaStreamSynchronize(st));
        if (h_I[1]) { tgs_set_error("num_rendered does not fit 32 bits"); return TGS_EINVAL; }
        I = (int64_t)h_I[0];
    }
    TgsBinningLayout bl; tgs_binning_layout(I, &bl);
    void* binning = alloc(user, TGS_BUF_BINNING, bl.total);
    if (!binning) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    BinView bv = tgs_bin_view(binning, I);
    rc = tgs_bin_scatter(gv, bv, N, I, I, false, cam.Tx, cam.Ty, cam.row0, cam.row1, temp, iv.ranges, iv.count, st); if (rc) return rc;
    TgsSettings s2 = *s;
    s2.depth_normalize = 0;
    rc = tgs_launch_render_fwd(cam, &s2, gv.records, bv, iv, I, out_color, out_depth, out_alpha, nullptr, nullptr, st); if (rc) return rc;
    saved->geom = geom; saved->binning = binning; saved->image = image; saved->num_rendered = I; saved->capacity = I;
    return 0;
}

extern "C" int tgs_rasterize_screen_backward(const TgsSettings* s, int32_t N, const TgsSaved* saved, const float* dL_dcolor,
                                             const float* dL_ddepth, const float* dL_dalpha, float* screen_grads,
                                             void* stream) {
    if (!s || !saved || !saved->binning || !saved->image || !dL_dcolor || (N > 0 && !screen_grads) || !s->bg) {
        tgs_set_error("tgs_rasterize_screen_backward: bad arguments"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    BinView bv = tgs_bin_view(saved->binning, saved->capacity > 0 ? saved->capacity : saved->num_rendered);
    ImageView iv = tgs_image_view(saved->image, cam.W, cam.H);
    if (N > 0) TGS_CUDA(cudaMemsetAsync(screen_grads, 0, sizeof(float) * TGS_NGRAD * (size_t)N, st));
    TgsSettings s2 = *s;
    s2.depth_normalize = 0;
    if (!saved->geom) { tgs_set_error("tgs_rasterize_screen_backward: saved geometry missing"); return TGS_ESTATE; }
    GeomView gvb = tgs_geom_view(saved->geom, N);
    return tgs_launch_render_bwd(cam, &s2, gvb.records, bv, iv, saved->num_rendered, dL_dcolor, dL_ddepth, dL_dalpha, nullptr, nullptr,
                                 screen_grads, nullptr, st);
}

extern "C" int tgs_spherical_harmonics(int32_t N, int32_t degree, int32_t K, const float* dirs, const float* coeffs,
                                       float* colors, void* stream) {
    if (N < 0 || degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || (N > 0 && (!dirs || !coeffs || !colors))) {
        tgs_set_error("tgs_spherical_harmonics: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    k_sh_eval<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, (degree + 1) * (degree + 1), K, dirs, coeffs, colors);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_spherical_harmonics_backward(int32_t N, int32_t degree, int32_t K, const float* dirs,
                                                const float* v_colors, float* v_coeffs, void* stream) {
    if (N < 0 || degree < 0 || degree > 3 || K < (degree + 1) * (degree + 1) || K > 16 || (N > 0 && (!dirs || !v_colors || !v_coeffs))) {
        tgs_set_error("tgs_spherical_harmonics_backward: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    k_sh_eval_bwd<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, (degree + 1) * (degree + 1), K, dirs, v_colors, v_coeffs);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
