// host_step.cu -- tgs_train_step_host: the whole forward + backward of the operator driven from
// HOST buffers (upload -> forward -> photometric L1 + fused touch loss -> backward -> download).
// This is the C-ABI call a non-PyTorch trainer would make and the path the end-to-end bench times
// with every host<->device copy inside the timed region.
#include "tgs_common.cuh"
#include <vector>

namespace {

constexpr unsigned kFull = 0xffffffffu;

// dL/dcolor = sign(color - gt) / (3HW);  loss += sum |color - gt| / (3HW)
__global__ void k_l1_photometric(const float* __restrict__ color, const float* __restrict__ gt, int64_t n,
                                 float inv_n, float* __restrict__ dcolor, float* __restrict__ loss) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.0f;
    for (; i < n; i += stride) {
        float d = color[i] - gt[i];
        acc += fabsf(d);
        dcolor[i] = (float)((d > 0.0f) - (d < 0.0f)) * inv_n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc * inv_n);
}

struct Pool {
    cudaStream_t st;
    std::vector<void*> ptrs;
    void* get(size_t bytes) {
        void* p = nullptr;
        if (cudaMallocAsync(&p, bytes ? bytes : 256, st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return p;
    }
    void release() { for (void* p : ptrs) cudaFreeAsync(p, st); ptrs.clear(); }
};

void* pool_alloc(void* user, int /*which*/, size_t bytes) { return static_cast<Pool*>(user)->get(bytes); }

template <typename T>
int upload(Pool& pool, const T* host, size_t count, const T** dev) {
    *dev = nullptr;
    if (!host || count == 0) return 0;
    T* d = static_cast<T*>(pool.get(count * sizeof(T)));
    if (!d) { tgs_set_error("cudaMallocAsync failed"); return TGS_ENOMEM; }
    TGS_CUDA(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, pool.st));
    *dev = d;
    return 0;
}
int download(cudaStream_t st, float* host, const float* dev, size_t count) {
    if (!host || !dev || count == 0) return 0;
    TGS_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(float), cudaMemcpyDeviceToHost, st));
    return 0;
}

int run_step(Pool& pool, const TgsSettings* sh, const TgsGaussians* gh, const float* gt_rgb_host,
             const float* target_host, const float* weight_host, int32_t loss_mode, float mult,
             const TgsGrads* grh, float* out_color_host, float* out_depth_host, int32_t* radii_host,
             float* loss_host, int64_t* num_rendered_host) {
    cudaStream_t st = pool.st;
    const int N = gh->N;
    const size_t n = (size_t)N, P = (size_t)sh->image_width * sh->image_height, K = (size_t)sh->sh_coeffs;
    int rc;
    TgsSettings s = *sh;
    TgsGaussians g = *gh;
#define UP(field, cnt) if ((rc = upload(pool, gh->field, (cnt), &g.field))) return rc
    UP(means3D, 3 * n); UP(opacities, n); UP(shs, 3 * K * n); UP(colors_precomp, 3 * n);
    UP(scales, 3 * n); UP(rotations, 4 * n); UP(cov3D_precomp, 6 * n);
#undef UP
    if ((rc = upload(pool, sh->viewmatrix, 16, &s.viewmatrix))) return rc;
    if ((rc = upload(pool, sh->projmatrix, 16, &s.projmatrix))) return rc;
    if ((rc = upload(pool, sh->campos, 3, &s.campos))) return rc;
    if ((rc = upload(pool, sh->bg, 3, &s.bg))) return rc;
    const float *gt = nullptr, *target = nullptr, *weight = nullptr;
    if ((rc = upload(pool, gt_rgb_host, 3 * P, &gt))) return rc;
    if ((rc = upload(pool, target_host, P, &target))) return rc;
    if ((rc = upload(pool, weight_host, P, &weight))) return rc;

    float* color = (float*)pool.get(3 * P * 4);
    float* depth = (float*)pool.get(P * 4);
    float* alpha = (float*)pool.get(P * 4);
    float* dcolor = (float*)pool.get(3 * P * 4);
    int32_t* radii = (int32_t*)pool.get((n ? n : 1) * 4);
    s.contrib_flags = 0;                  // one GPU: plain rows (the chain rule skips the all-zero ones)
    float* sgrad = (float*)pool.get(n ? tgs_screen_grad_bytes(N, 0) : 64);
    float* scal = (float*)pool.get(16);   // [0..1] touch scale + counter, [2] photometric loss
    if (!color || !depth || !alpha || !dcolor || !radii || !sgrad || !scal) { tgs_set_error("cudaMallocAsync failed"); return TGS_ENOMEM; }
    TGS_CUDA(cudaMemsetAsync(scal, 0, 16, st));

    TgsSaved saved{};
    rc = tgs_forward(&s, &g, pool_alloc, &pool, color, depth, alpha, radii, nullptr, nullptr, &saved, st);
    if (rc) return rc;
    if (num_rendered_host) *num_rendered_host = saved.num_rendered;

    if (gt) {
        int64_t cnt = (int64_t)(3 * P);
        int blocks = (int)((cnt + 256 * 8 - 1) / (256 * 8));
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_l1_photometric<<<blocks, 256, 0, st>>>(color, gt, cnt, 1.0f / (float)cnt, dcolor, scal + 2);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    } else {
        TGS_CUDA(cudaMemsetAsync(dcolor, 0, 3 * P * 4, st));
    }
    TgsTouch touch{};
    const TgsTouch* tp = nullptr;
    if (target && loss_mode != TGS_LOSS_NONE) {
        rc = tgs_touch_loss_scale(target, (int64_t)P, mult, 0.0f, scal, st);
        if (rc) return rc;
        touch.target = target; touch.weight = weight; touch.scale = scal; touch.mode = loss_mode;
        tp = &touch;
    }
    TgsGrads gd{};
    if (N > 0) {
        gd.dmeans2D = (float*)pool.get(3 * n * 4); gd.dmeans3D = (float*)pool.get(3 * n * 4);
        gd.dopacity = (float*)pool.get(n * 4);
        if (g.shs) gd.dshs = (float*)pool.get(3 * K * n * 4);
        if (g.colors_precomp) gd.dcolors = (float*)pool.get(3 * n * 4);
        if (g.scales) { gd.dscales = (float*)pool.get(3 * n * 4); gd.drotations = (float*)pool.get(4 * n * 4); }
        if (g.cov3D_precomp) gd.dcov3D = (float*)pool.get(6 * n * 4);
    }
    rc = tgs_backward(&s, &g, &saved, radii, dcolor, nullptr, nullptr, tp, nullptr, sgrad, &gd, st);
    if (rc) return rc;

    if (grh) {
        if ((rc = download(st, grh->dmeans2D, gd.dmeans2D, 3 * n))) return rc;
        if ((rc = download(st, grh->dmeans3D, gd.dmeans3D, 3 * n))) return rc;
        if ((rc = download(st, grh->dopacity, gd.dopacity, n))) return rc;
        if ((rc = download(st, grh->dshs, gd.dshs, 3 * K * n))) return rc;
        if ((rc = download(st, grh->dcolors, gd.dcolors, 3 * n))) return rc;
        if ((rc = download(st, grh->dscales, gd.dscales, 3 * n))) return rc;
        if ((rc = download(st, grh->drotations, gd.drotations, 4 * n))) return rc;
        if ((rc = download(st, grh->dcov3D, gd.dcov3D, 6 * n))) return rc;
    }
    if ((rc = download(st, out_color_host, color, 3 * P))) return rc;
    if ((rc = download(st, out_depth_host, depth, P))) return rc;
    if (radii_host && N > 0) TGS_CUDA(cudaMemcpyAsync(radii_host, radii, n * 4, cudaMemcpyDeviceToHost, st));
    if ((rc = download(st, loss_host, scal + 2, 1))) return rc;
    return 0;
}

}  // namespace

extern "C" int tgs_train_step_host(const TgsSettings* s_host, const TgsGaussians* g_host,
                                   const float* gt_rgb_host, const float* touch_target_host,
                                   const float* touch_weight_host, int32_t loss_mode, float depth_loss_mult,
                                   const TgsGrads* grads_host, float* out_color_host, float* out_depth_host,
                                   int32_t* radii_host, float* loss_host, int64_t* num_rendered_host,
                                   void* stream) {
    if (!s_host || !g_host) { tgs_set_error("tgs_train_step_host: NULL settings / gaussians"); return TGS_EINVAL; }
    Pool pool; pool.st = (cudaStream_t)stream;
    int rc = run_step(pool, s_host, g_host, gt_rgb_host, touch_target_host, touch_weight_host, loss_mode,
                      depth_loss_mult, grads_host, out_color_host, out_depth_host, radii_host, loss_host,
                      num_rendered_host);
    cudaError_t e = cudaStreamSynchronize(pool.st);
    pool.release();
    if (rc) return rc;
    return tgs_check_cuda(e, "cudaStreamSynchronize", __FILE__, __LINE__);
}
