// train_ops.cu -- the per-step pieces AROUND the rasterizer in a Touch-GS train step (SURVEY.md §8f row N1,
// BASELINE config c5 "full Touch-GS train step (Adam + densify)"): fused photometric loss (L1 + SSIM) forward and
// gradient, parameter activations and their backward, one-launch multi-group Adam, refine statistics and the
// refine (densify / cull) stream compaction.  The trainer these replace lives in the reference's empty nerfstudio
// submodule (reference .gitmodules:7-9); the knobs that ARE pinned in the tree are cited in include/tgs.h.
//
// Every kernel here is a streaming HBM-bound pass (roofline: HBM): one read of its inputs, one write of its
// outputs, 16-byte accesses where the layout allows.  No tensor cores (no dense contraction anywhere).
#include "tgs_common.cuh"
#include <cub/cub.cuh>
#include <cmath>

namespace {

constexpr unsigned kFull = 0xffffffffu;

// --------------------------------------------------------------------------- photometric loss
constexpr int kST = 16;            // output tile edge
constexpr int kSR = 5;             // window radius (11 taps)
constexpr int kSH = kST + 2 * kSR; // haloed tile edge = 26
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct SsimTaps { float w[2 * kSR + 1]; };

SsimTaps make_taps() {             // same construction as oracle/train_oracle.py::ssim_window (float64, then cast)
    SsimTaps t; double g[11], s = 0.0;
    for (int i = 0; i < 11; ++i) { double x = (double)(i - 5); g[i] = std::exp(-(x * x) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    for (int i = 0; i < 11; ++i) t.w[i] = (float)(g[i] / s);
    return t;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    }
    return t;      // valid in thread 0
}

// Forward: one CTA = one 16x16 tile of ONE channel.  Loads the 26x26 haloed tiles of the rendered image and the
// ground truth (zero outside the IMAGE), separable 11-tap blur of (x, y, x^2, y^2, xy), then per pixel the SSIM
// value and its three partial derivatives (w.r.t. blurred x, blurred x^2, blurred xy), which the backward kernel
// blurs again.  Only rows [y_begin, y_end) are loss pixels (a rank's band); sums go to two double accumulators.
__global__ void __launch_bounds__(256)
k_ssim_fwd(const float* __restrict__ color, const float* __restrict__ gt, int W, int H, int y_begin, int y_end,
           SsimTaps taps, float* __restrict__ dmaps, double* __restrict__ sums) {
    __shared__ float sx[kSH][kSH + 1], sy[kSH][kSH + 1];
    __shared__ float hz[5][kSH][kST];
    __shared__ float red[8];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * kST, y0 = y_begin + blockIdx.y * kST;
    const size_t HW = (size_t)W * H;
    const float* cx = color + ch * HW;
    const float* cy = gt + ch * HW;
    for (int i = tid; i < kSH * kSH; i += 256) {
        const int r = i / kSH, c = i - r * kSH;
        const int gy = y0 - kSR + r, gx = x0 - kSR + c;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        sx[r][c] = in ? cx[(size_t)gy * W + gx] : 0.0f;
        sy[r][c] = in ? cy[(size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kSH * kST; i += 256) {
        const int r = i >> 4, c = i & 15;
        float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * kSR + 1; ++k) {
            const float w = taps.w[k], x = sx[r][c + k], y = sy[r][c + k];
            const float wx = w * x, wy = w * y;
            m1 += wx; m2 += wy; xx = fmaf(wx, x, xx); yy = fmaf(wy, y, yy); xy = fmaf(wx, y, xy);
        }
        hz[0][r][c] = m1; hz[1][r][c] = m2; hz[2][r][c] = xx; hz[3][r][c] = yy; hz[4][r][c] = xy;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kSR + 1; ++k) {
        const float w = taps.w[k];
        mu1 = fmaf(w, hz[0][ty + k][tx], mu1); mu2 = fmaf(w, hz[1][ty + k][tx], mu2);
        exx = fmaf(w, hz[2][ty + k][tx], exx); eyy = fmaf(w, hz[3][ty + k][tx], eyy);
        exy = fmaf(w, hz[4][ty + k][tx], exy);
    }
    const int px = x0 + tx, py = y0 + ty;
    const bool valid = px < W && py < y_end && py < H;
    float l1 = 0.0f, ss = 0.0f;
    if (valid) {
        const float s11 = exx - mu1 * mu1, s22 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
        const float A = 2.0f * mu1 * mu2 + kC1, B = 2.0f * s12 + kC2;
        const float Cc = mu1 * mu1 + mu2 * mu2 + kC1, Dd = s11 + s22 + kC2;
        const float rCD = 1.0f / (Cc * Dd);
        ss = A * B * rCD;
        // d/dmu1 (through A, Cc directly and through s11 = E[xx]-mu1^2, s12 = E[xy]-mu1 mu2)
        const float dmu = (2.0f * mu2 * B - 2.0f * mu2 * A) * rCD - ss * (2.0f * mu1 * Dd - 2.0f * mu1 * Cc) * rCD;
        const float dxx = -ss / Dd;
        const float dxy = 2.0f * A * rCD;
        const size_t o = (size_t)ch * HW + (size_t)py * W + px;
        dmaps[o] = dmu; dmaps[3 * HW + o] = dxx; dmaps[6 * HW + o] = dxy;
        l1 = fabsf(sx[ty + kSR][tx + kSR] - sy[ty + kSR][tx + kSR]);
    }
    const float bl1 = block_sum_256(l1, red);
    const float bss = block_sum_256(ss, red);
    if (tid == 0) { atomicAdd(sums, (double)bl1); atomicAdd(sums + 1, (double)bss); }
}

// loss = (1-l) * sum|x-y| / n + l * (cnt - sum ssim) / n ; n = 3*W*H (full image), cnt = loss pixels of this band
__global__ void k_loss_finish(const double* sums, double cnt, double n, float lambda, float* loss_out) {
    loss_out[0] = (float)((1.0 - (double)lambda) * sums[0] / n + (double)lambda * (cnt - sums[1]) / n);
}

// Backward: dL/dx_q = g * [ -(l/n) * ( blur(dmu)_q + 2 x_q blur(dxx)_q + y_q blur(dxy)_q ) + ((1-l)/n) sign(x_q-y_q) ]
// with the three derivative maps taken as ZERO outside the loss rows / the image.  Writes rows [o_begin, o_end).
__global__ void __launch_bounds__(256)
k_ssim_bwd(const float* __restrict__ color, const float* __restrict__ gt, const float* __restrict__ dmaps, int W, int H,
           int y_begin, int y_end, int o_begin, int o_end, SsimTaps taps, float lambda, float inv_n,
           const float* __restrict__ grad_out, float* __restrict__ dcolor) {
    __shared__ float sm[3][kSH][kSH + 1];
    __shared__ float hz[3][kSH][kST];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * kST, y0 = o_begin + blockIdx.y * kST;
    const size_t HW = (size_t)W * H;
    for (int i = tid; i < kSH * kSH; i += 256) {
        const int r = i / kSH, c = i - r * kSH;
        const int gy = y0 - kSR + r, gx = x0 - kSR + c;
        const bool in = gx >= 0 && gx < W && gy >= y_begin && gy < y_end && gy < H;
        const size_t o = (size_t)ch * HW + (size_t)gy * W + gx;
        sm[0][r][c] = in ? dmaps[o] : 0.0f;
        sm[1][r][c] = in ? dmaps[3 * HW + o] : 0.0f;
        sm[2][r][c] = in ? dmaps[6 * HW + o] : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < kSH * kST; i += 256) {
        const int r = i >> 4, c = i & 15;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * kSR + 1; ++k) {
            const float w = taps.w[k];
            a = fmaf(w, sm[0][r][c + k], a); b = fmaf(w, sm[1][r][c + k], b); d = fmaf(w, sm[2][r][c + k], d);
        }
        hz[0][r][c] = a; hz[1][r][c] = b; hz[2][r][c] = d;
    }
    __syncthreads();
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kSR + 1; ++k) {
        const float w = taps.w[k];
        a = fmaf(w, hz[0][ty + k][tx], a); b = fmaf(w, hz[1][ty + k][tx], b); d = fmaf(w, hz[2][ty + k][tx], d);
    }
    const int px = x0 + tx, py = y0 + ty;
    if (px < W && py < o_end && py < H) {
        const size_t o = (size_t)ch * HW + (size_t)py * W + px;
        const float x = color[o], y = gt[o];
        const float g = grad_out ? grad_out[0] : 1.0f;
        float v = -(lambda * inv_n) * (a + 2.0f * x * b + y * d);
        if (py >= y_begin && py < y_end) v += (1.0f - lambda) * inv_n * (float)((x > y) - (x < y));
        dcolor[o] = g * v;
    }
}

// lambda_dssim == 0: the loss is the plain L1 mean -- two streaming passes (sum, then sign), 16-byte accesses.
// Plain L1 (lambda_dssim = 0): HBM-bound streaming kernels.  blockIdx.y = channel; a channel's loss rows are one
// contiguous run of floats, read with 16-byte loads when the run is 16-byte aligned (scalar otherwise): no per-element
// integer division, 8 floats in flight per thread.
__global__ void __launch_bounds__(256)
k_l1_fwd(const float* __restrict__ color, const float* __restrict__ gt, int W, int H, int y_begin, int y_end,
         double* __restrict__ sums) {
    __shared__ float red[8];
    const size_t base = (size_t)blockIdx.y * W * H + (size_t)y_begin * W;
    const int64_t n = (int64_t)W * (y_end - y_begin);
    const float* c = color + base;
    const float* g = gt + base;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    if ((((uintptr_t)c | (uintptr_t)g) & 15) == 0 && (n & 3) == 0) {
        const float4* c4 = reinterpret_cast<const float4*>(c);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (int64_t i = t0; i < (n >> 2); i += stride) {
            const float4 a = __ldcs(c4 + i), b = __ldcs(g4 + i);
            acc += (fabsf(a.x - b.x) + fabsf(a.y - b.y)) + (fabsf(a.z - b.z) + fabsf(a.w - b.w));
        }
    } else {
        for (int64_t i = t0; i < n; i += stride) acc += fabsf(c[i] - g[i]);
    }
    const float b = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(sums, (double)b);
}

// dcolor rows [o_begin, o_end) <- g * sign(color - gt) on the loss rows [y_begin, y_end), 0 on the other rows written
__global__ void __launch_bounds__(256)
k_l1_bwd(const float* __restrict__ color, const float* __restrict__ gt, int W, int H, int y_begin, int y_end,
         int o_begin, int o_end, float inv_n, const float* __restrict__ grad_out, float* __restrict__ dcolor) {
    const size_t base = (size_t)blockIdx.y * W * H + (size_t)o_begin * W;
    const int64_t n = (int64_t)W * (o_end - o_begin);
    // loss rows as a range of the run's element indices
    const int64_t l0 = (int64_t)W * (max(y_begin, o_begin) - o_begin), l1 = (int64_t)W * (min(y_end, o_end) - o_begin);
    const float g = (grad_out ? grad_out[0] : 1.0f) * inv_n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto sgn = [&](float x, float y, int64_t e) { return (e >= l0 && e < l1) ? g * (float)((x > y) - (x < y)) : 0.0f; };
    if ((((uintptr_t)(color + base) | (uintptr_t)(gt + base) | (uintptr_t)(dcolor + base)) & 15) == 0 && (n & 3) == 0) {
        const float4* c4 = reinterpret_cast<const float4*>(color + base);
        const float4* g4 = reinterpret_cast<const float4*>(gt + base);
        float4* d4 = reinterpret_cast<float4*>(dcolor + base);
        for (int64_t i = t0; i < (n >> 2); i += stride) {
            const float4 a = __ldcs(c4 + i), b = __ldcs(g4 + i);
            const int64_t e = i << 2;
            d4[i] = make_float4(sgn(a.x, b.x, e), sgn(a.y, b.y, e + 1), sgn(a.z, b.z, e + 2), sgn(a.w, b.w, e + 3));
        }
    } else {
        for (int64_t i = t0; i < n; i += stride) dcolor[base + i] = sgn(color[base + i], gt[base + i], i);
    }
}

// -------------------------------------------------------------------------------- activations
__global__ void __launch_bounds__(256)
k_activate_fwd(int N, const float* __restrict__ scales_log, const float* __restrict__ quats,
               const float* __restrict__ opacity_logit, float* __restrict__ scales, float* __restrict__ rot,
               float* __restrict__ opac) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) scales[3 * i + k] = expf(scales_log[3 * i + k]);
    const float4 q = reinterpret_cast<const float4*>(quats)[i];
    const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    reinterpret_cast<float4*>(rot)[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    opac[i] = 1.0f / (1.0f + expf(-opacity_logit[i]));
}

// gradients w.r.t. the activated values -> gradients w.r.t. the raw parameters (outputs may alias the inputs)
__global__ void __launch_bounds__(256)
k_activate_bwd(int N, const float* __restrict__ scales_log, const float* __restrict__ quats,
               const float* __restrict__ opacity_logit, const float* dscales, const float* drot, const float* dopac,
               float* dscales_log, float* dquats, float* dopacity_logit) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) dscales_log[3 * i + k] = dscales[3 * i + k] * expf(scales_log[3 * i + k]);
    const float4 q = reinterpret_cast<const float4*>(quats)[i];
    const float4 g = reinterpret_cast<const float4*>(drot)[i];
    const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    const float rx = q.x * inv, ry = q.y * inv, rz = q.z * inv, rw = q.w * inv;
    const float dot = rx * g.x + ry * g.y + rz * g.z + rw * g.w;
    reinterpret_cast<float4*>(dquats)[i] = make_float4((g.x - rx * dot) * inv, (g.y - ry * dot) * inv,
                                                       (g.z - rz * dot) * inv, (g.w - rw * dot) * inv);
    const float s = 1.0f / (1.0f + expf(-opacity_logit[i]));
    dopacity_logit[i] = dopac[i] * s * (1.0f - s);
}

// -------------------------------------------------------------------------------------- Adam
struct AdamArgs {
    TgsAdamGroup g[TGS_ADAM_MAX_GROUPS];
    float step_size[TGS_ADAM_MAX_GROUPS], step_size_tail[TGS_ADAM_MAX_GROUPS];
    int n;
    float beta1, beta2, eps, inv_bc2_sqrt;
    float omb1, omb2;               // 1-beta1, 1-beta2 formed in DOUBLE on the host (as torch does), then rounded
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step, const AdamArgs& a) {
    const float eps = a.eps, inv_bc2 = a.inv_bc2_sqrt;
    m = m + (g - m) * a.omb1;                     // lerp(m, g, 1-b1)
    v = v * a.beta2 + (a.omb2 * g) * g;           // mul_(b2).addcmul_(g, g, 1-b2)
    const float denom = sqrtf(v) * inv_bc2 + eps;
    p = p - step * (m / denom);                   // addcdiv_(m, denom, -step)
}

// One launch for ALL parameter groups: blockIdx.y = group, grid-stride over 16-byte chunks.
__global__ void __launch_bounds__(256)
k_adam(const __grid_constant__ AdamArgs a) {
    const int gi = blockIdx.y;
    const TgsAdamGroup G = a.g[gi];
    const int64_t n4 = G.numel >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* p4 = reinterpret_cast<float4*>(G.param);
    const float4* g4 = reinterpret_cast<const float4*>(G.grad);
    float4* m4 = reinterpret_cast<float4*>(G.exp_avg);
    float4* v4 = reinterpret_cast<float4*>(G.exp_avg_sq);
    const float s0 = a.step_size[gi], s1 = a.step_size_tail[gi];
    const bool two = G.period > 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
        float st[4] = {s0, s0, s0, s0};
        if (two) {
            const int r = (int)((i * 4) % G.period);
#pragma unroll
            for (int k = 0; k < 4; ++k) { int rr = r + k; if (rr >= G.period) rr -= G.period; st[k] = rr < G.head ? s0 : s1; }
        }
        adam_one(p.x, g.x, m.x, v.x, st[0], a);
        adam_one(p.y, g.y, m.y, v.y, st[1], a);
        adam_one(p.z, g.z, m.z, v.z, st[2], a);
        adam_one(p.w, g.w, m.w, v.w, st[3], a);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (G.numel & 3)) {          // scalar tail
        const int64_t i = (n4 << 2) + threadIdx.x;
        float st = s0;
        if (two) st = (int)(i % G.period) < G.head ? s0 : s1;
        float p = G.param[i], m = G.exp_avg[i], v = G.exp_avg_sq[i];
        adam_one(p, G.grad[i], m, v, st, a);
        G.param[i] = p; G.exp_avg[i] = m; G.exp_avg_sq[i] = v;
    }
}

// ------------------------------------------------------------------------------------ refine
__global__ void __launch_bounds__(256)
k_densify_stats(int N, const float* __restrict__ dmeans2D, const int32_t* __restrict__ radii,
                float* __restrict__ grad_accum, int32_t* __restrict__ vis_count, int32_t* __restrict__ max_radii) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = dmeans2D[3 * i], gy = dmeans2D[3 * i + 1];
    grad_accum[i] += sqrtf(gx * gx + gy * gy);
    vis_count[i] += 1;
    max_radii[i] = max(max_radii[i], r);
}

__global__ void __launch_bounds__(256)
k_densify_classify(int N, const float* __restrict__ opacity_logit, const float* __restrict__ scales_log,
                   const float* __restrict__ grad_accum, const int32_t* __restrict__ vis_count,
                   const int32_t* __restrict__ max_radii, TgsDensifyConfig cfg, int allow, uint32_t* __restrict__ counts) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int vc = vis_count[i];
    const float avg = vc > 0 ? grad_accum[i] / (float)vc : 0.0f;
    const float smax = fmaxf(fmaxf(expf(scales_log[3 * i]), expf(scales_log[3 * i + 1])), expf(scales_log[3 * i + 2]));
    const float op = 1.0f / (1.0f + expf(-opacity_logit[i]));
    // screen-size rules (largest screen radius seen since the last refine; a threshold of 0 switches the rule off)
    const float rmax = max_radii ? (float)max_radii[i] : 0.0f;
    const bool big_screen = cfg.split_screen_radius > 0.0f && rmax > cfg.split_screen_radius;
    const bool cull = (op < cfg.cull_alpha_thresh) || (smax > cfg.cull_scale_thresh) ||
                      (cfg.cull_screen_radius > 0.0f && rmax > cfg.cull_screen_radius);
    const bool high = allow && (avg > cfg.grad_thresh);
    const bool split = !cull && allow && ((high && smax > cfg.size_thresh) || big_screen);
    uint32_t c = 1;
    if (cull) c = 0;
    else if (split) c = (uint32_t)cfg.n_split_samples;
    else if (high) c = 2u;
    // bit 31 marks "split" so that apply does not have to recompute the classification
    counts[i] = c | (split ? 0x80000000u : 0u);
}

struct CountOnly {
    __device__ __forceinline__ uint32_t operator()(uint32_t v) const { return v & 0x7FFFFFFFu; }
};

// One WARP per source Gaussian: lanes copy the rows of every tensor (value + both Adam moments) cooperatively.
__global__ void __launch_bounds__(256)
k_densify_apply(int N, int K, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                const float* __restrict__ noise, TgsDensifyConfig cfg, TgsParamSet ip, TgsParamSet im, TgsParamSet iv,
                TgsParamSet op, TgsParamSet om, TgsParamSet ov, int32_t* __restrict__ src_out) {
    const int i = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const uint32_t cw = counts[i];
    const uint32_t cnt = cw & 0x7FFFFFFFu;
    if (cnt == 0) return;
    const bool split = (cw >> 31) != 0;
    const uint32_t o0 = offsets[i];
    const int widths[5] = {3, 3 * K, 1, 3, 4};
    const float* pin[5] = {ip.means, ip.shs, ip.opacity, ip.scales, ip.quats};
    const float* min_[5] = {im.means, im.shs, im.opacity, im.scales, im.quats};
    const float* vin[5] = {iv.means, iv.shs, iv.opacity, iv.scales, iv.quats};
    float* pout[5] = {op.means, op.shs, op.opacity, op.scales, op.quats};
    float* mout[5] = {om.means, om.shs, om.opacity, om.scales, om.quats};
    float* vout[5] = {ov.means, ov.shs, ov.opacity, ov.scales, ov.quats};
    // split samples: mean + R(q/|q|) (exp(s) * noise), scale = log(exp(s) / shrink)
    float R[9], sc[3], mean[3];
    if (split) {
        const float4 q = reinterpret_cast<const float4*>(ip.quats)[i];
        const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float w = q.x * inv, x = q.y * inv, y = q.z * inv, z = q.w * inv;
        R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
        R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
        R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
#pragma unroll
        for (int k = 0; k < 3; ++k) { sc[k] = expf(ip.scales[3 * i + k]); mean[k] = ip.means[3 * i + k]; }
    }
    for (uint32_t j = 0; j < cnt; ++j) {
        const size_t o = (size_t)o0 + j;
        const bool fresh = split || j == 1;               // new entries start with zero Adam moments
        if (lane == 0 && src_out) src_out[o] = fresh ? -(i + 1) : i;
        for (int t = 0; t < 5; ++t) {
            const int wd = widths[t];
            if (pin[t] == nullptr) continue;
            for (int c = lane; c < wd; c += 32) {
                float v = pin[t][(size_t)i * wd + c];
                if (split && t == 0) {
                    const float* nz = noise + ((size_t)i * cfg.n_split_samples + j) * 3;
                    const float a0 = sc[0] * nz[0], a1 = sc[1] * nz[1], a2 = sc[2] * nz[2];
                    v = mean[c] + ((R[3 * c] * a0 + R[3 * c + 1] * a1) + R[3 * c + 2] * a2);
                } else if (split && t == 3) {
                    v = logf(sc[c] / cfg.split_shrink);
                }
                pout[t][o * wd + c] = v;
                if (mout[t]) mout[t][o * wd + c] = fresh ? 0.0f : min_[t][(size_t)i * wd + c];
                if (vout[t]) vout[t][o * wd + c] = fresh ? 0.0f : vin[t][(size_t)i * wd + c];
            }
        }
    }
}

}  // namespace

// =============================================================================== C ABI
extern "C" size_t tgs_photometric_scratch_floats(int32_t W, int32_t H) { return (size_t)9 * W * H; }

extern "C" int tgs_photometric_loss_forward(const float* color, const float* gt, int32_t W, int32_t H,
                                            int32_t row_begin, int32_t row_end, float lambda_dssim, float* dmaps,
                                            double* sums, float* loss_out, void* stream) {
    if (!color || !gt || !dmaps || !sums || !loss_out || W <= 0 || H <= 0) { tgs_set_error("tgs_photometric_loss_forward: bad arguments"); return TGS_EINVAL; }
    if (row_end <= row_begin) { row_begin = 0; row_end = H; }
    if (row_begin < 0 || row_end > H) { tgs_set_error("tgs_photometric_loss_forward: rows outside the image"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    static const SsimTaps taps = make_taps();
    TgsProfScope prof(TGS_STAGE_PHOTO_FWD, st);
    TGS_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    dim3 grid((W + kST - 1) / kST, (row_end - row_begin + kST - 1) / kST, 3);
    if (lambda_dssim == 0.0f) {                     // plain L1: no SSIM pass, no derivative maps
        const int64_t n = 1ll * W * (row_end - row_begin);         // per channel
        int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
        if (blocks > 148 * 4) blocks = 148 * 4;
        if (blocks < 1) blocks = 1;
        k_l1_fwd<<<dim3((unsigned)blocks, 3), 256, 0, st>>>(color, gt, W, H, row_begin, row_end, sums);
    } else
    k_ssim_fwd<<<grid, 256, 0, st>>>(color, gt, W, H, row_begin, row_end, taps, dmaps, sums);
    k_loss_finish<<<1, 1, 0, st>>>(sums, 3.0 * W * (double)(row_end - row_begin), 3.0 * W * (double)H, lambda_dssim, loss_out);
    tgs_count_own(2);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_photometric_loss_backward(const float* color, const float* gt, const float* dmaps, int32_t W, int32_t H,
                                             int32_t row_begin, int32_t row_end, int32_t out_row_begin, int32_t out_row_end,
                                             float lambda_dssim, const float* grad_out, float* dL_dcolor, void* stream) {
    if (!color || !gt || !dmaps || !dL_dcolor || W <= 0 || H <= 0) { tgs_set_error("tgs_photometric_loss_backward: bad arguments"); return TGS_EINVAL; }
    if (row_end <= row_begin) { row_begin = 0; row_end = H; }
    if (out_row_end <= out_row_begin) { out_row_begin = 0; out_row_end = H; }
    if (row_begin < 0 || row_end > H || out_row_begin < 0 || out_row_end > H) { tgs_set_error("tgs_photometric_loss_backward: rows outside the image"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    static const SsimTaps taps = make_taps();
    dim3 grid((W + kST - 1) / kST, (out_row_end - out_row_begin + kST - 1) / kST, 3);
    TgsProfScope prof(TGS_STAGE_PHOTO_BWD, st);
    if (lambda_dssim == 0.0f) {
        const int64_t n = 1ll * W * (out_row_end - out_row_begin);  // per channel
        int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
        if (blocks > 148 * 4) blocks = 148 * 4;
        if (blocks < 1) blocks = 1;
        k_l1_bwd<<<dim3((unsigned)blocks, 3), 256, 0, st>>>(color, gt, W, H, row_begin, row_end, out_row_begin, out_row_end,
                                                   (float)(1.0 / (3.0 * W * (double)H)), grad_out, dL_dcolor);
    } else
    k_ssim_bwd<<<grid, 256, 0, st>>>(color, gt, dmaps, W, H, row_begin, row_end, out_row_begin, out_row_end, taps,
                                     lambda_dssim, (float)(1.0 / (3.0 * W * (double)H)), grad_out, dL_dcolor);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_activate_forward(int32_t N, const float* scales_log, const float* quats, const float* opacity_logit,
                                    float* scales, float* rotations, float* opacities, void* stream) {
    if (N < 0 || (N > 0 && (!scales_log || !quats || !opacity_logit || !scales || !rotations || !opacities))) {
        tgs_set_error("tgs_activate_forward: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    TgsProfScope prof(TGS_STAGE_ACTIVATE, (cudaStream_t)stream);
    k_activate_fwd<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, scales_log, quats, opacity_logit, scales, rotations, opacities);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_activate_backward(int32_t N, const float* scales_log, const float* quats, const float* opacity_logit,
                                     const float* dscales, const float* drotations, const float* dopacities,
                                     float* dscales_log, float* dquats, float* dopacity_logit, void* stream) {
    if (N < 0 || (N > 0 && (!scales_log || !quats || !opacity_logit || !dscales || !drotations || !dopacities ||
                            !dscales_log || !dquats || !dopacity_logit))) {
        tgs_set_error("tgs_activate_backward: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    TgsProfScope prof(TGS_STAGE_ACTIVATE, (cudaStream_t)stream);
    k_activate_bwd<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, scales_log, quats, opacity_logit, dscales, drotations,
                                                                       dopacities, dscales_log, dquats, dopacity_logit);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_adam_step(const TgsAdamGroup* groups, int32_t n_groups, int32_t step, double beta1, double beta2,
                             double eps, void* stream) {
    if (!groups || n_groups <= 0 || n_groups > TGS_ADAM_MAX_GROUPS || step <= 0) { tgs_set_error("tgs_adam_step: bad arguments"); return TGS_EINVAL; }
    AdamArgs a;
    a.n = n_groups; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    const double bc1 = 1.0 - std::pow(beta1, (double)step);
    const double bc2 = 1.0 - std::pow(beta2, (double)step);
    a.inv_bc2_sqrt = (float)(1.0 / std::sqrt(bc2));
    int64_t most = 0;
    for (int i = 0; i < n_groups; ++i) {
        const TgsAdamGroup& g = groups[i];
        if (g.numel < 0 || (g.numel > 0 && (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq))) { tgs_set_error("tgs_adam_step: group %d has NULL tensors", i); return TGS_EINVAL; }
        if (g.period < 0 || g.head < 0 || g.head > g.period) { tgs_set_error("tgs_adam_step: group %d bad period/head", i); return TGS_EINVAL; }
        if ((((uintptr_t)g.param | (uintptr_t)g.grad | (uintptr_t)g.exp_avg | (uintptr_t)g.exp_avg_sq) & 15) != 0) { tgs_set_error("tgs_adam_step: group %d not 16-byte aligned", i); return TGS_EINVAL; }
        a.g[i] = g;
        a.step_size[i] = (float)((double)g.lr / bc1);
        a.step_size_tail[i] = (float)((double)g.lr_tail / bc1);
        if (g.numel > most) most = g.numel;
    }
    if (most == 0) return 0;
    int64_t blocks = ((most >> 2) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;                  // persistent-style grid: 8 CTAs per SM per group
    if (blocks < 1) blocks = 1;
    TgsProfScope prof(TGS_STAGE_ADAM, (cudaStream_t)stream);
    k_adam<<<dim3((unsigned)blocks, n_groups), 256, 0, (cudaStream_t)stream>>>(a);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_densify_stats(int32_t N, const float* dmeans2D, const int32_t* radii, float* grad_accum,
                                 int32_t* vis_count, int32_t* max_radii, void* stream) {
    if (N < 0 || (N > 0 && (!dmeans2D || !radii || !grad_accum || !vis_count || !max_radii))) { tgs_set_error("tgs_densify_stats: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    TgsProfScope prof(TGS_STAGE_REFINE, (cudaStream_t)stream);
    k_densify_stats<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(N, dmeans2D, radii, grad_accum, vis_count, max_radii);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t tgs_densify_temp_bytes(int32_t N) {
    size_t b = 0;
    cub::TransformInputIterator<uint32_t, CountOnly, const uint32_t*> it(nullptr, CountOnly());
    cub::DeviceScan::ExclusiveSum(nullptr, b, it, (uint32_t*)nullptr, N > 0 ? N : 1);
    return b + 256;
}

extern "C" int tgs_densify_plan(int32_t N, const float* opacity_logit, const float* scales_log, const float* grad_accum,
                                const int32_t* vis_count, const int32_t* max_radii, const TgsDensifyConfig* cfg,
                                int32_t allow_split_dup,
                                uint32_t* counts, uint32_t* offsets, void* temp, size_t temp_bytes,
                                int64_t* total_host, void* stream) {
    if (N < 0 || !cfg || !total_host || (N > 0 && (!opacity_logit || !scales_log || !grad_accum || !vis_count || !counts || !offsets || !temp))) {
        tgs_set_error("tgs_densify_plan: bad arguments"); return TGS_EINVAL; }
    if (N == 0) { *total_host = 0; return 0; }          // an empty population stays empty
    if (cfg->n_split_samples < 1 || cfg->n_split_samples > 8) { tgs_set_error("tgs_densify_plan: n_split_samples out of range"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    TgsProfScope prof(TGS_STAGE_REFINE, st);
    k_densify_classify<<<(N + 255) / 256, 256, 0, st>>>(N, opacity_logit, scales_log, grad_accum, vis_count, max_radii,
                                                        *cfg, allow_split_dup, counts);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    cub::TransformInputIterator<uint32_t, CountOnly, const uint32_t*> it(counts, CountOnly());
    size_t tb = temp_bytes;
    TGS_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, it, offsets, N, st));
    tgs_count_cub(1);
    uint32_t last[2] = {0, 0};
    TGS_CUDA(cudaMemcpyAsync(&last[0], offsets + (N - 1), 4, cudaMemcpyDeviceToHost, st));
    TGS_CUDA(cudaMemcpyAsync(&last[1], counts + (N - 1), 4, cudaMemcpyDeviceToHost, st));
    TGS_CUDA(cudaStreamSynchronize(st));          // refine runs every ~100 steps: the host must size the new tensors
    *total_host = (int64_t)last[0] + (int64_t)(last[1] & 0x7FFFFFFFu);
    return 0;
}

extern "C" int tgs_densify_apply(int32_t N, int32_t K, const uint32_t* counts, const uint32_t* offsets,
                                 const float* noise, const TgsDensifyConfig* cfg, const TgsParamSet* in_pmv,
                                 const TgsParamSet* out_pmv, int32_t* src_out, void* stream) {
    if (N < 0 || K < 0 || !cfg || !in_pmv || !out_pmv || (N > 0 && (!counts || !offsets))) { tgs_set_error("tgs_densify_apply: bad arguments"); return TGS_EINVAL; }
    if (N == 0) return 0;
    const TgsParamSet& p = in_pmv[0]; const TgsParamSet& o = out_pmv[0];
    if (!p.means || !p.scales || !p.quats || !o.means || !o.scales || !o.quats || !noise) { tgs_set_error("tgs_densify_apply: means / scales / quats / noise required"); return TGS_EINVAL; }
    const int64_t threads = (int64_t)N * 32;
    TgsProfScope prof(TGS_STAGE_REFINE, (cudaStream_t)stream);
    k_densify_apply<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        N, K, counts, offsets, noise, *cfg, in_pmv[0], in_pmv[1], in_pmv[2], out_pmv[0], out_pmv[1], out_pmv[2], src_out);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
