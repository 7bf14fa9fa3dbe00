// TEST-ONLY: compiles the __host__ __device__ per-Gaussian math of csrc/tgs_math.cuh for the
// CPU so that `pytest -m "not gpu"` can check it against the oracle without a GPU.
// Built by tests/conftest.py with: g++ -O1 -ffp-contract=off -shared -fPIC.
// Never linked into libtgs.so; the product path is CUDA only.
#include "../touch-gs_b200/csrc/tgs_math.cuh"
#include "../touch-gs_b200/csrc/touch_inputs_math.cuh"

extern "C" {

// out_f: [N, 6] = px, py, depth, conA, conB, conC ; out_i: [N, 6] = radius, rminx, rminy, rmaxx, rmaxy, tiles
// out_rgb: [N,3]; out_clamped [N]; cov_out [N,6]
void hm_preprocess(int N, const float* means, const float* scales, const float* rots,
                   const float* shs, const float* cov_pre, const float* vm, const float* pm,
                   const float* campos, const TgsCam* cam,
                   float* out_f, int* out_i, float* out_rgb, unsigned* out_clamped, float* cov_out) {
    for (int i = 0; i < N; ++i) {
        float cov[6];
        if (cov_pre) for (int k = 0; k < 6; ++k) cov[k] = cov_pre[6 * i + k];
        else tgs_cov3d(scales + 3 * i, cam->mod, rots + 4 * i, cov);
        for (int k = 0; k < 6; ++k) cov_out[6 * i + k] = cov[k];
        TgsProj p;
        bool vis = tgs_project(vm, pm, *cam, means[3 * i], means[3 * i + 1], means[3 * i + 2], cov, p);
        out_f[6 * i + 0] = p.px; out_f[6 * i + 1] = p.py; out_f[6 * i + 2] = p.depth;
        out_f[6 * i + 3] = p.conA; out_f[6 * i + 4] = p.conB; out_f[6 * i + 5] = p.conC;
        out_i[6 * i + 0] = p.radius; out_i[6 * i + 1] = p.rminx; out_i[6 * i + 2] = p.rminy;
        out_i[6 * i + 3] = p.rmaxx; out_i[6 * i + 4] = p.rmaxy; out_i[6 * i + 5] = p.tiles;
        out_rgb[3 * i] = out_rgb[3 * i + 1] = out_rgb[3 * i + 2] = 0.0f;
        out_clamped[i] = 0;
        if (vis && shs) {
            float sh48[48];
            int nb = (cam->deg + 1) * (cam->deg + 1);
            for (int k = 0; k < 48; ++k) sh48[k] = (k < 3 * nb) ? shs[3 * cam->K * i + k] : 0.0f;
            tgs_sh_forward(cam->deg, sh48, means[3 * i] - campos[0],
                           means[3 * i + 1] - campos[1], means[3 * i + 2] - campos[2],
                           out_rgb + 3 * i, out_clamped[i]);
        }
    }
}

// screen grads sg [N,10] -> dmeans [N,3], dscales [N,3], drots [N,4], dsh [N,K,3], dcov [N,6]
void hm_backward(int N, const float* means, const float* scales, const float* rots,
                 const float* shs, const float* cov_pre, const float* vm, const float* pm,
                 const float* campos, const TgsCam* cam, const int* radii,
                 const unsigned* clamped, const float* sg,
                 float* dmeans, float* dscales, float* drots, float* dsh, float* dcov_out) {
    for (int i = 0; i < N; ++i) {
        float dm[3] = {0, 0, 0}, dc[6] = {0, 0, 0, 0, 0, 0}, ds[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0};
        if (radii[i] > 0) {
            float cov[6];
            if (cov_pre) for (int k = 0; k < 6; ++k) cov[k] = cov_pre[6 * i + k];
            else tgs_cov3d(scales + 3 * i, cam->mod, rots + 4 * i, cov);
            tgs_project_backward(vm, pm, *cam, means[3 * i], means[3 * i + 1], means[3 * i + 2], cov,
                                 sg + 10 * i, dm, dc);
            if (shs)
                tgs_sh_backward(cam->deg, cam->K, shs + 3 * cam->K * i, means[3 * i] - campos[0],
                                means[3 * i + 1] - campos[1], means[3 * i + 2] - campos[2],
                                sg + 10 * i + 6, clamped[i], dsh + 3 * cam->K * i, dm);
            if (!cov_pre) tgs_cov3d_backward(scales + 3 * i, cam->mod, rots + 4 * i, dc, ds, dq);
        } else if (shs) {
            for (int k = 0; k < 3 * cam->K; ++k) dsh[3 * cam->K * i + k] = 0.0f;
        }
        for (int k = 0; k < 3; ++k) { dmeans[3 * i + k] = dm[k]; dscales[3 * i + k] = ds[k]; }
        for (int k = 0; k < 4; ++k) drots[4 * i + k] = dq[k];
        for (int k = 0; k < 6; ++k) dcov_out[6 * i + k] = dc[k];
    }
}

// per-pixel touch / vision fusion (fp64) on the host: out6 = va, ds, fu, fs (uint16) ; target, weight (float)
void hm_fuse(long n, const unsigned short* touch, const unsigned short* vision, const unsigned short* tsig,
             double scale, double offset, double offset2, int real_world, double scene_scale,
             unsigned short* va, unsigned short* ds, unsigned short* fu, unsigned short* fs, float* target, float* weight) {
    FuseParams p; p.scale = scale; p.offset = offset; p.offset2 = offset2; p.unit = 1e-3 * scene_scale; p.real_world = real_world;
    for (long i = 0; i < n; ++i) {
        PixelOut o = fuse_pixel(touch[i], vision[i], tsig[i], p);
        va[i] = o.va; ds[i] = o.ds; fu[i] = o.fu; fs[i] = o.fs; target[i] = o.target; weight[i] = o.weight;
    }
}

}  // extern "C"
