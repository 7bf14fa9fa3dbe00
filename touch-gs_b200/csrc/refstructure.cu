// refstructure.cu -- the "reference-structure CUDA" comparison path (SURVEY.md §8d
// "Reference CUDA path beside it").
//
// The reference's own CUDA rasterizer is not in its tree (reference .gitmodules:7-9 -> empty submodule) and
// cannot be installed here, so the column "reference CUDA path" of every table is THIS build: the same
// algorithm laid out the way the public splat rasterizers of that era structure it,
// written from the spec and compiled for sm_100a:
//   * 1 thread per Gaussian preprocess (shared with the product path: preprocess.cu already has that shape);
//   * inclusive scan of tiles_touched in Gaussian-id order, one blocking D2H read of num_rendered;
//   * one thread per Gaussian loops over its tile rectangle writing 64-bit keys (tile << 32 | bits(depth))
//     and 32-bit values;
//   * ONE stable radix sort of the 12-byte pairs over 32 + ceil(log2 T) bits (A3); boundary detection (A4);
//   * 16x16 thread block per tile, one pixel per thread, cooperative GATHER of 256 instances per round from
//     the per-Gaussian arrays into shared memory, no culling, block-wide early exit (A5);
//   * backward: same tiling, back-to-front, TEN global fp32 atomics per (pixel, Gaussian) pair (A6);
//   * depth is an extra composited channel; the touch-depth loss is NOT fused: the caller computes it in
//     PyTorch from the returned raw depth / alpha and passes dL/ddepth, dL/dalpha images back in.
// It is a measurement and cross-checking arm only (bench.py "reference_structure_cuda", tests): the product
// operator never calls it.  The per-pair arithmetic is render_math.cuh, shared with the product kernels, so
// n_contrib must agree BIT-EXACTLY with the culling / multi-pixel product kernels at any size -- the
// full-size parity test that the CPU oracle is too slow for.
#include "tgs_common.cuh"
#include "render_math.cuh"
#include <cub/cub.cuh>

namespace {

constexpr int kBlk = 256;

struct RefBinView {
    uint32_t* offsets;                       // [N] inclusive scan of tiles_touched, id order
    uint64_t* keys_unsorted; uint64_t* keys_sorted;
    uint32_t* vals_unsorted; uint32_t* vals_sorted;
    uint2* ranges;                           // [T]
    void* temp; size_t temp_bytes;
};

size_t ref_temp_bytes(int N, int64_t I) {
    size_t a = 0, b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, N > 0 ? N : 1);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, I > 0 ? I : 1, 0, 64);
    return a > b ? a : b;
}

RefBinView ref_bin_view(void* base, int N, int64_t I, int T) {
    TgsRefBinningLayout l; tgs_refstructure_binning_layout(N, I, T, &l);
    char* b = (char*)base; RefBinView v;
    v.offsets = (uint32_t*)(b + l.offsets);
    v.keys_unsorted = (uint64_t*)(b + l.keys_unsorted); v.keys_sorted = (uint64_t*)(b + l.keys_sorted);
    v.vals_unsorted = (uint32_t*)(b + l.vals_unsorted); v.vals_sorted = (uint32_t*)(b + l.vals_sorted);
    v.ranges = (uint2*)(b + l.ranges);
    v.temp = b + l.temp; v.temp_bytes = l.temp_bytes;
    return v;
}

// per Gaussian, row-major walk of its tile rectangle.
__global__ void __launch_bounds__(kBlk)
k_ref_duplicate(int N, const TgsRecord* __restrict__ rec, const uint32_t* __restrict__ tiles,
                const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect, int Tx,
                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * kBlk + threadIdx.x;
    if (i >= N || tiles[i] == 0) return;
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    const uint2 rc = rect[i];
    const uint32_t x0 = rc.x & 0xFFFF, x1 = rc.x >> 16, y0 = rc.y & 0xFFFF, y1 = rc.y >> 16;
    const uint64_t dbits = (uint64_t)__float_as_uint(rec[i].a.z);
    for (uint32_t y = y0; y < y1; ++y)
        for (uint32_t x = x0; x < x1; ++x) {
            keys[off] = ((uint64_t)(y * Tx + x) << 32) | dbits;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

// SURVEY A4
__global__ void __launch_bounds__(kBlk)
k_ref_ranges(int64_t I, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int64_t j = (int64_t)blockIdx.x * kBlk + threadIdx.x;
    if (j >= I) return;
    const uint32_t tile = (uint32_t)(keys[j] >> 32);
    if (j == 0) ranges[tile].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[j - 1] >> 32);
        if (prev != tile) { ranges[prev].y = (uint32_t)j; ranges[tile].x = (uint32_t)j; }
    }
    if (j == I - 1) ranges[tile].y = (uint32_t)I;
}

// SURVEY A5, upstream structure: block = tile, thread = pixel, rounds of 256 gathered instances.
__global__ void __launch_bounds__(kBlk)
k_ref_render_fwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ list,
                 const TgsRecord* __restrict__ rec, int W, int H, int Tx, const float* __restrict__ bg,
                 float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                 float* __restrict__ final_T, uint32_t* __restrict__ n_contrib) {
    __shared__ uint32_t s_id[kBlk];
    __shared__ float4 s_a[kBlk];      // x, y, depth, id bits
    __shared__ float4 s_q[kBlk];      // conic A, B, C, opacity
    const int tile = blockIdx.y * Tx + blockIdx.x;
    const int px = blockIdx.x * TGS_TILE + threadIdx.x, py = blockIdx.y * TGS_TILE + threadIdx.y;
    const int rank = threadIdx.y * TGS_TILE + threadIdx.x;
    const bool inside = px < W && py < H;
    const int pix = py * W + px;
    const float fx = (float)px, fy = (float)py;
    const uint2 rng = ranges[tile];
    int todo = (int)(rng.y - rng.x);
    const int rounds = (todo + kBlk - 1) / kBlk;
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t contributor = 0, last = 0;
    for (int r = 0; r < rounds; ++r, todo -= kBlk) {
        if (__syncthreads_count(done) == kBlk) break;
        const int progress = r * kBlk + rank;
        if (rng.x + progress < rng.y) {
            const uint32_t id = list[rng.x + progress];
            s_id[rank] = id; s_a[rank] = rec[id].a; s_q[rank] = rec[id].b;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(kBlk, todo); ++j) {
            ++contributor;
            const float4 a = s_a[j], q = s_q[j];
            const float power = splat_power(q, a.x - fx, a.y - fy);
            if (power > 0.0f) continue;
            const float alpha = splat_alpha(q.w, splat_exp(power));
            if (alpha < TGS_ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < TGS_T_EPS) { done = true; continue; }
            const float4 c = rec[s_id[j]].c;           // colour fetched from global per contributor
            const float w = alpha * T;
            C0 += c.x * w; C1 += c.y * w; C2 += c.z * w; D += a.z * w;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t HW = (size_t)W * H;
        out_color[pix] = C0 + T * bg[0];
        out_color[HW + pix] = C1 + T * bg[1];
        out_color[2 * HW + pix] = C2 + T * bg[2];
        out_depth[pix] = D;
        out_alpha[pix] = 1.0f - T;
        final_T[pix] = T;
        n_contrib[pix] = last;
    }
}

// SURVEY A6 without the fusion, upstream structure: per-thread global atomics.
__global__ void __launch_bounds__(kBlk)
k_ref_render_bwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ list,
                 const TgsRecord* __restrict__ rec, int W, int H, int Tx, const float* __restrict__ bg,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
                 const float* __restrict__ dL_dalpha, float* __restrict__ sgrad) {
    __shared__ uint32_t s_id[kBlk];
    __shared__ float4 s_a[kBlk], s_q[kBlk], s_c[kBlk];
    const int tile = blockIdx.y * Tx + blockIdx.x;
    const int px = blockIdx.x * TGS_TILE + threadIdx.x, py = blockIdx.y * TGS_TILE + threadIdx.y;
    const int rank = threadIdx.y * TGS_TILE + threadIdx.x;
    const bool inside = px < W && py < H;
    const int pix = py * W + px;
    const float fx = (float)px, fy = (float)py;
    const uint2 rng = ranges[tile];
    int todo = (int)(rng.y - rng.x);
    const int rounds = (todo + kBlk - 1) / kBlk;
    const bool done = !inside;
    const float T_final = inside ? final_T[pix] : 0.0f;
    float T = T_final;
    uint32_t contributor = (uint32_t)todo;
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        const size_t HW = (size_t)W * H;
        g0 = dL_dcolor[pix]; g1 = dL_dcolor[HW + pix]; g2 = dL_dcolor[2 * HW + pix];
        if (dL_ddepth) gD = dL_ddepth[pix];
        if (dL_dalpha) gA = dL_dalpha[pix];
    }
    const float bg_dot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;
    for (int r = 0; r < rounds; ++r, todo -= kBlk) {
        __syncthreads();
        const int progress = r * kBlk + rank;
        if (rng.x + progress < rng.y) {
            const uint32_t id = list[rng.y - progress - 1];
            s_id[rank] = id; s_a[rank] = rec[id].a; s_q[rank] = rec[id].b; s_c[rank] = rec[id].c;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(kBlk, todo); ++j) {
            --contributor;
            if (contributor >= last_contributor) continue;
            const float4 a = s_a[j], q = s_q[j];
            const float dx = a.x - fx, dy = a.y - fy;
            const float power = splat_power(q, dx, dy);
            if (power > 0.0f) continue;
            const float G = splat_exp(power);
            const float alpha = splat_alpha(q.w, G);
            if (alpha < TGS_ALPHA_MIN) continue;
            T = T / (1.0f - alpha);
            const float w = alpha * T;
            const float4 c = s_c[j];
            float* dst = sgrad + (size_t)s_id[j] * TGS_NGRAD;
            float dLda = 0.0f;
            acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0; lc0 = c.x; dLda += (c.x - acc0) * g0;
            acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1; lc1 = c.y; dLda += (c.y - acc1) * g1;
            acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2; lc2 = c.z; dLda += (c.z - acc2) * g2;
            accD = last_alpha * lD + (1.0f - last_alpha) * accD; lD = a.z; dLda += (a.z - accD) * gD;
            atomicAdd(dst + 6, w * g0); atomicAdd(dst + 7, w * g1); atomicAdd(dst + 8, w * g2);
            atomicAdd(dst + 9, w * gD);
            dLda *= T;
            last_alpha = alpha;
            dLda += (T_final / (1.0f - alpha)) * (gA - bg_dot);     // background (rgb) and alpha channel
            const float dLdG = q.w * dLda;                          // straight-through alpha clamp
            const float gdx = G * dx, gdy = G * dy;
            atomicAdd(dst + 0, dLdG * (-gdx * q.x - gdy * q.y));
            atomicAdd(dst + 1, dLdG * (-gdy * q.z - gdx * q.y));
            atomicAdd(dst + 2, -0.5f * gdx * dx * dLdG);
            atomicAdd(dst + 3, -gdx * dy * dLdG);
            atomicAdd(dst + 4, -0.5f * gdy * dy * dLdG);
            atomicAdd(dst + 5, G * dLda);
        }
    }
}

int ceil_log2_u(uint32_t v) { int b = 0; while ((1u << b) < v) ++b; return b; }

}  // namespace

extern "C" int tgs_refstructure_binning_layout(int32_t N, int64_t I, int32_t T, TgsRefBinningLayout* o) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off = tgs_align_up(off + bytes); return r; };
    const size_t n = (size_t)(N > 0 ? N : 0), m = (size_t)(I > 0 ? I : 0);
    o->offsets = take(n * 4);
    o->keys_unsorted = take(m * 8); o->keys_sorted = take(m * 8);
    o->vals_unsorted = take(m * 4); o->vals_sorted = take(m * 4);
    o->ranges = take((size_t)(T > 0 ? T : 1) * sizeof(uint2));
    o->temp_bytes = ref_temp_bytes(N, I);
    o->temp = take(o->temp_bytes);
    o->total = off;
    return 0;
}

extern "C" int tgs_refstructure_forward(const TgsSettings* s, const TgsGaussians* g, tgs_alloc_fn alloc, void* user,
                                        float* out_color, float* out_depth_raw, float* out_alpha, int32_t* radii,
                                        TgsSaved* saved, void* stream) {
    if (!s || !g || !alloc || !saved || !out_color || !out_depth_raw || !out_alpha || (g->N > 0 && !radii)) {
        tgs_set_error("tgs_refstructure_forward: NULL argument"); return TGS_EINVAL; }
    if (g->N <= 0) { tgs_set_error("tgs_refstructure_forward: N must be > 0"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    if (cam.row0 != 0 || cam.row1 != cam.Ty) { tgs_set_error("tgs_refstructure_forward: whole image only"); return TGS_EINVAL; }
    const int N = g->N, T = cam.Tx * cam.Ty;
    TgsGeomLayout gl; tgs_geom_layout(N, &gl);
    TgsImageLayout il; tgs_image_layout(cam.W, cam.H, &il);
    void* geom = alloc(user, TGS_BUF_GEOM, gl.total);
    void* image = alloc(user, TGS_BUF_IMAGE, il.total);
    if (!geom || !image) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    GeomView gv = tgs_geom_view(geom, N);
    ImageView iv = tgs_image_view(image, cam.W, cam.H);
    int rc = tgs_launch_preprocess(cam, s, g, gv, radii, st); if (rc) return rc;
    // scan in id order needs a place before the binning buffer exists: the geometry buffer's `order` array (this arm
    // does not depth-sort the Gaussians, so it is free)
    size_t tb = gv.temp_bytes;
    TGS_CUDA(cub::DeviceScan::InclusiveSum(gv.temp, tb, gv.tiles_touched, gv.order, N, st));
    tgs_count_cub(1);
    uint32_t h_I = 0;
    TGS_CUDA(cudaMemcpyAsync(&h_I, gv.order + (N - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TGS_CUDA(cudaStreamSynchronize(st));               // the blocking read of num_rendered
    const int64_t I = (int64_t)h_I;
    TgsRefBinningLayout bl; tgs_refstructure_binning_layout(N, I, T, &bl);
    void* binning = alloc(user, TGS_BUF_BINNING, bl.total);
    if (!binning) { tgs_set_error("allocator returned NULL"); return TGS_ENOMEM; }
    RefBinView bv = ref_bin_view(binning, N, I, T);
    TGS_CUDA(cudaMemsetAsync(bv.ranges, 0, sizeof(uint2) * (size_t)T, st));
    if (I > 0) {
        k_ref_duplicate<<<(N + kBlk - 1) / kBlk, kBlk, 0, st>>>(N, gv.records, gv.tiles_touched, gv.order, gv.rect,
                                                                cam.Tx, bv.keys_unsorted, bv.vals_unsorted);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
        size_t bytes = bv.temp_bytes;
        TGS_CUDA(cub::DeviceRadixSort::SortPairs(bv.temp, bytes, bv.keys_unsorted, bv.keys_sorted, bv.vals_unsorted,
                                                 bv.vals_sorted, I, 0, 32 + ceil_log2_u((uint32_t)T), st));
        tgs_count_cub(1);
        k_ref_ranges<<<(unsigned)((I + kBlk - 1) / kBlk), kBlk, 0, st>>>(I, bv.keys_sorted, bv.ranges);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    }
    k_ref_render_fwd<<<dim3(cam.Tx, cam.Ty), dim3(TGS_TILE, TGS_TILE), 0, st>>>(
        bv.ranges, bv.vals_sorted, gv.records, cam.W, cam.H, cam.Tx, s->bg, out_color, out_depth_raw, out_alpha,
        iv.final_T, iv.n_contrib);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    saved->geom = geom; saved->binning = binning; saved->image = image; saved->num_rendered = I; saved->capacity = I;
    return 0;
}

extern "C" int tgs_refstructure_backward_render(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                                const float* dL_dcolor, const float* dL_ddepth_raw,
                                                const float* dL_dalpha, float* screen_grads, void* stream) {
    if (!s || !g || !saved || !saved->geom || !saved->binning || !saved->image || !dL_dcolor || !screen_grads) {
        tgs_set_error("tgs_refstructure_backward_render: NULL argument"); return TGS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const TgsCam cam = tgs_make_cam(s);
    const int N = g->N, T = cam.Tx * cam.Ty;
    GeomView gv = tgs_geom_view(saved->geom, N);
    ImageView iv = tgs_image_view(saved->image, cam.W, cam.H);
    RefBinView bv = ref_bin_view(saved->binning, N, saved->num_rendered, T);
    TGS_CUDA(cudaMemsetAsync(screen_grads, 0, sizeof(float) * TGS_NGRAD * (size_t)N, st));
    k_ref_render_bwd<<<dim3(cam.Tx, cam.Ty), dim3(TGS_TILE, TGS_TILE), 0, st>>>(
        bv.ranges, bv.vals_sorted, gv.records, cam.W, cam.H, cam.Tx, s->bg, iv.final_T, iv.n_contrib, dL_dcolor,
        dL_ddepth_raw, dL_dalpha, screen_grads);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    return 0;
}
