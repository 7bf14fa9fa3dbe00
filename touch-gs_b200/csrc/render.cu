// render.cu -- per-tile alpha compositing and its backward with the touch-depth
// gradient fused into the same traversal (A6), plus the tiny loss-scale reduction.
//
// B200 design
//   * one CTA per 16x16 tile, one thread per pixel, a warp covers an 8x4 pixel patch;
//   * the tile's depth-sorted list is a CONTIGUOUS run of 48-byte records (binning.cu packs it), so
//     each batch of 256 records (12 KB) is brought into shared memory by ONE TMA bulk copy
//     (cp.async.bulk ... mbarrier::complete_tx::bytes), double-buffered: the copy of batch b+2 is
//     issued as soon as batch b has been consumed, so HBM/L2 latency is hidden behind the blend;
//   * EXACT warp-level culling: each lane tests one splat's alpha-threshold ellipse against the warp's
//     8x4 pixel patch (closed-form minimum of the quadratic over the rectangle), a ballot gives the
//     survivors, and only those are evaluated per pixel -- most (pixel, splat) pairs of the 3-sigma
//     square bounding boxes never reach alpha >= 1/255, so this removes the bulk of the arithmetic
//     while leaving every result bit-identical;
//   * forward: RGB, expected depth and alpha are composited in ONE traversal;
//   * backward: back-to-front replay, one warp per 16x8 half tile with four pixels per thread; the
//     per-pixel depth / alpha gradients are computed IN-KERNEL from the touch target, its weight and
//     the loss scale (no autograd round trip through HBM); per-thread moment accumulation, then the
//     10 per-Gaussian gradient values are reduced across the warp with a 12-shuffle reduce-scatter
//     butterfly (instead of 10 x 5 shuffles) and land in HBM as 10 REDs per (half tile, Gaussian)
//     instead of 1280 per-thread atomics.
//   No tensor cores: the path is gather/blend, not a dense contraction.
//
// Roofline: HBM.  Algorithmic bytes: forward 48 B/instance + 24 B/pixel written;
// backward 48 B/instance read + 40 B/instance accumulated + 32 B/pixel.
#include "tgs_common.cuh"
#include <cstdlib>
#include "render_math.cuh"

namespace {

constexpr int kBatch = 256;                 // records per TMA stage
constexpr uint32_t kRecBytes = 48;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16-byte asynchronous copy global -> shared (SASS: LDGSTS), grouped per thread
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct PixelMap {
    int px, py, pix; bool inside; float fx, fy;
    float x0, x1, y0, y1;           // pixel-centre rectangle of this warp's 8x4 patch
};
__device__ __forceinline__ PixelMap map_pixel(int tile, int Tx, int W, int H) {
    PixelMap m;
    int tx = tile % Tx, ty = tile / Tx;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int wx = tx * TGS_TILE + (warp & 1) * 8, wy = ty * TGS_TILE + (warp >> 1) * 4;
    m.px = wx + (lane & 7);
    m.py = wy + (lane >> 3);
    m.inside = (m.px < W) && (m.py < H);
    m.pix = m.py * W + m.px;
    m.fx = (float)m.px; m.fy = (float)m.py;
    m.x0 = (float)wx; m.x1 = (float)(wx + 7); m.y0 = (float)wy; m.y1 = (float)(wy + 3);
    return m;
}

// EXACT warp-level cull.  A splat can only contribute to a pixel if alpha = o*exp(power) >= 1/255,
// i.e. power >= thr := -ln(255 o)  (thr is precomputed per splat in record.c.w).  power = -q/2 with
// q(p) = (p-mu)^T Q (p-mu) convex, so over the warp's pixel rectangle R the maximum power is -q_min/2
// where q_min is 0 if mu lies in R and otherwise the minimum over the four edges (each a clamped 1-D
// quadratic).  If even that maximum is below thr (minus a safety margin that dominates the fp32
// rounding of this bound and of the per-pixel evaluation), NO pixel of the patch would pass the
// alpha test, so skipping the splat for the whole warp is bit-identical to evaluating it.
__device__ __forceinline__ bool patch_may_touch(const float4 a, const float4 q, float thr, const PixelMap& pm) {
    const float ex0 = pm.x0 - a.x, ex1 = pm.x1 - a.x, ey0 = pm.y0 - a.y, ey1 = pm.y1 - a.y;
    if (ex0 <= 0.0f && ex1 >= 0.0f && ey0 <= 0.0f && ey1 >= 0.0f) return thr <= 0.05f;
    const float A = q.x, B = q.y, Cc = q.z;
    // the closed-form bound below assumes a convex quadratic; fp32 cancellation in det = a*c - b*b of an extremely
    // anisotropic splat can leave an indefinite conic, for which nothing may be culled (forward and backward use
    // different patch sizes and must still agree on every blended pair)
    if (!(A > 0.0f) || !(A * Cc > B * B)) return true;
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

// ------------------------------------------------------------------------------- forward
// Staging of a tile's depth-sorted list.  The list is the contiguous run ids[rng.x .. rng.y) of Gaussian ids (binning.cu);
// the 48-byte records live once per GAUSSIAN in the geometry buffer (48 MB at 1M splats: L2 resident).  Per batch of
// 256 list entries:
//   * one TMA bulk copy (cp.async.bulk + mbarrier complete_tx) brings the batch's 1 KB of ids into shared memory
//     (3-deep ring, issued two batches ahead; the window is widened to 16-byte alignment);
//   * every thread gathers ONE record by id with three 16-byte cp.async copies (LDGSTS) into the 2-stage record buffer,
//     one batch ahead of the blend.
// Only the batches a tile actually composites before all its pixels saturate are ever gathered (22 % of the instances at
// c3), and no per-instance copy of the records is written to HBM at all (the packed-record array of the previous
// design cost 96 B/instance of traffic and a kernel of its own: 0.175 ms at c3, 2.5 ms at c5).
constexpr int kIdRing = 3;
constexpr int kIdSlots = kBatch + 4;             // a 16-byte aligned window around 256 ids

// kMbar = true (product): the arrival of a batch's records is tracked by an mbarrier -- every thread posts
// cp.async.mbarrier.arrive.noinc behind its three copies, the consumers wait on the barrier's phase (PTX: when the wait
// returns, all cp.async operations the participating threads requested before their cp.async.mbarrier.arrive are performed
// and visible).  kMbar = false (TGS_FWD_RECORD_SYNC=waitgroup, diagnostics): the classic cp.async.wait_group + block
// barrier instead, one more __syncthreads per batch.  compute-sanitizer's racecheck does not model the mbarrier form and
// reports the record stage as racing; with the wait_group form it is clean (profiles/r04i_sanitizer_racecheck*.log).
template <bool kMbar>
__global__ void __launch_bounds__(256)
k_render_fwd(const uint2* __restrict__ ranges, uint32_t cap, const uint32_t* __restrict__ ids,
             const float4* __restrict__ table, int W, int H, int Tx,
             int row0, const float* __restrict__ bg, int normalize, float alpha_max, float* __restrict__ out_color,
             float* __restrict__ out_depth, float* __restrict__ out_alpha, float* __restrict__ final_T,
             uint32_t* __restrict__ n_contrib, float* __restrict__ depth_raw, float* __restrict__ color_acc,
             float* __restrict__ ckpt, uint32_t* __restrict__ slot_tile, uint32_t* __restrict__ ckpt_list,
             uint32_t* __restrict__ ckpt_count,
             const float* __restrict__ t_target, float* __restrict__ residual) {
    __shared__ __align__(128) float4 sbuf[2][kBatch * 3];
    __shared__ __align__(16) uint32_t sid[kIdRing][kIdSlots];
    __shared__ __align__(8) uint64_t idbar[kIdRing];
    __shared__ __align__(8) uint64_t full[2];        // record stages: 256 arrivals, one per thread when its copies land
    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x + row0 * Tx;
    const PixelMap pm = map_pixel(tile, Tx, W, H);
    uint2 rng = ranges[tile];
    // speculative sizing: the buffers hold `cap` instances; if the real count overflows them this launch's result is
    // discarded (the host re-runs with the exact size), but it must stay inside the buffers
    rng.x = min(rng.x, cap); rng.y = min(rng.y, cap);
    const int len = (int)(rng.y - rng.x);
    const int nb = (len + kBatch - 1) / kBatch;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kIdRing; ++k) mbar_init(&idbar[k], 1);
        mbar_init(&full[0], 256); mbar_init(&full[1], 256);
        mbar_fence_init();
    }
    __syncthreads();
    // ids of batch b: TMA bulk copy of the 16-byte aligned window that contains list positions [p, p + cnt)
    auto issue_ids = [&](int b) {
        const uint32_t p = rng.x + (uint32_t)b * kBatch, pa = p & ~3u;
        const uint32_t cnt = (uint32_t)min(kBatch, len - b * kBatch);
        const uint32_t bytes = (((p - pa) + cnt) * 4u + 15u) & ~15u;
        mbar_expect_tx(&idbar[b % kIdRing], bytes);
        tma_bulk_g2s(sid[b % kIdRing], ids + pa, bytes, &idbar[b % kIdRing]);
    };
    // records of batch b: every thread gathers the record of ITS list entry
    auto gather = [&](int b) {
        mbar_wait(&idbar[b % kIdRing], (uint32_t)((b / kIdRing) & 1));
        const uint32_t p = rng.x + (uint32_t)b * kBatch;
        if (tid < min(kBatch, len - b * kBatch)) {
            const uint32_t id = sid[b % kIdRing][(p & 3u) + tid];
            const float4* src = table + (size_t)3 * id;
            float4* dst = sbuf[b & 1] + 3 * tid;
            cp_async16(dst, src); cp_async16(dst + 1, src + 1); cp_async16(dst + 2, src + 2);
        }
        // arrive on the stage's mbarrier when THIS thread's copies have landed (immediately if it issued none)
        if (kMbar) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[b & 1])) : "memory");
        else cp_async_commit();
    };
    if (tid == 0) { for (int b = 0; b < kIdRing && b < nb; ++b) issue_ids(b); }
    if (nb > 0) gather(0);
    if (nb > 1) gather(1);

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
    uint32_t last = 0;
    bool done = !pm.inside;
    int b = 0;
    for (; b < nb; ++b) {
        if (kMbar) {
            mbar_wait(&full[b & 1], (uint32_t)((b >> 1) & 1));           // all 256 threads' copies of batch b have landed
        } else {
            if (b + 1 < nb) cp_async_wait<1>(); else cp_async_wait<0>(); // this thread's copies of batch b ...
            __syncthreads();                                             // ... and everybody else's
        }
        const float4* s = sbuf[b & 1];
        const int cnt = min(kBatch, len - b * kBatch);
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            if (__all_sync(kFull, done)) break;                 // whole warp saturated
            const int jl = c0 + lane;
            bool pass = false;
            if (jl < cnt) pass = patch_may_touch(s[3 * jl], s[3 * jl + 1], s[3 * jl + 2].w, pm);
            unsigned mask = __ballot_sync(kFull, pass);
            while (mask) {
                const int j = c0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 a = s[3 * j], q = s[3 * j + 1];
                const float dx = a.x - pm.fx, dy = a.y - pm.fy;
                const float power = splat_power(q, dx, dy);
                const float alpha = splat_alpha(q.w, splat_exp(power), alpha_max);
                bool valid = !done && (power <= 0.0f) && (alpha >= TGS_ALPHA_MIN);
                if (!__any_sync(kFull, valid)) continue;
                const float test_T = T * (1.0f - alpha);
                if (valid && test_T < TGS_T_EPS) { done = true; valid = false; }
                if (valid) {
                    const float4 c = s[3 * j + 2];
                    const float w = alpha * T;
                    C0 += c.x * w; C1 += c.y * w; C2 += c.z * w; D += a.z * w;
                    T = test_T;
                    last = (uint32_t)(b * kBatch + j + 1);
                }
            }
        }
        const int nd = __syncthreads_count(done);
        if (nd == 256) break;
        if (tid == 0 && b + kIdRing < nb) issue_ids(b + kIdRing);        // its ring slot (batch b's ids) is free now
        if (b + 2 < nb) gather(b + 2);                                    // into the record stage batch b just released
        if (b + 1 < nb) {
            // the list continues: CHECKPOINT the per-pixel state in front of list position (b+1)*256, so that the
            // backward can replay the tile's list in independent 256-record segments (one warp per segment)
            const uint32_t slot = (rng.x + (uint32_t)(b + 1) * kBatch) >> 8;   // unique per (tile, boundary)
            float* ck = ckpt + (size_t)slot * TGS_CKPT_FLOATS + ((pm.py & 15) * 16 + (pm.px & 15));
            ck[0] = T; ck[256] = C0; ck[512] = C1; ck[768] = C2; ck[1024] = D;
            if (tid == 0) { slot_tile[slot] = (uint32_t)tile; ckpt_list[atomicAdd(ckpt_count, 1u)] = slot; }
        }
    }
    // copies may still be in flight if we broke out early: the CTA must not retire (and free its shared memory)
    // before they land.  Records: this thread's cp.async copies; ids: batches issued but not yet waited for by gather().
    asm volatile("cp.async.wait_all;" ::: "memory");
    // (broke out of iteration b: ids were issued through batch b + 2 and gather() waited through b + 1)
    if (tid == 0 && b < nb && b + 2 < nb) mbar_wait(&idbar[(b + 2) % kIdRing], (uint32_t)(((b + 2) / kIdRing) & 1));

    if (pm.inside) {
        const size_t HW = (size_t)W * H;
        out_color[pm.pix] = C0 + T * bg[0];
        out_color[HW + pm.pix] = C1 + T * bg[1];
        out_color[2 * HW + pm.pix] = C2 + T * bg[2];
        if (last > (uint32_t)kBatch) {      // only pixels that blend beyond the first segment are re-based in backward
            color_acc[pm.pix] = C0; color_acc[HW + pm.pix] = C1; color_acc[2 * HW + pm.pix] = C2;
        }
        const float A = 1.0f - T;
        out_alpha[pm.pix] = A;
        const float dhat = normalize ? (A > 0.0f ? D / A : 0.0f) : D;
        out_depth[pm.pix] = dhat;
        if (residual) {
            const float tgt = t_target[pm.pix];
            residual[pm.pix] = (tgt > 0.0f && A > 0.0f) ? dhat - tgt : 0.0f;
        }
        final_T[pm.pix] = T;
        n_contrib[pm.pix] = last;
        depth_raw[pm.pix] = D;
    }
}

// ------------------------------------------------------------------------------ backward
// Reduce-scatter of 10 per-lane values across the warp in 12 shuffles.  On return lane L holds in
// `out` the warp-wide sum of value `slot` if `valid`.
__device__ __forceinline__ void warp_reduce_scatter10(const float (&v)[TGS_NGRAD], int lane, float& out,
                                                      int& slot, bool& valid) {
    bool h = lane & 16;
    float r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float send = h ? v[k] : v[k + 5];
        float keep = h ? v[k + 5] : v[k];
        r[k] = keep + __shfl_xor_sync(kFull, send, 16);
    }
    h = lane & 8;
    float s0 = (h ? r[3] : r[0]) + __shfl_xor_sync(kFull, h ? r[0] : r[3], 8);
    float s1 = (h ? r[4] : r[1]) + __shfl_xor_sync(kFull, h ? r[1] : r[4], 8);
    float s2 = (h ? 0.0f : r[2]) + __shfl_xor_sync(kFull, h ? r[2] : 0.0f, 8);
    h = lane & 4;
    float t0 = (h ? s2 : s0) + __shfl_xor_sync(kFull, h ? s0 : s2, 4);
    float t1 = (h ? 0.0f : s1) + __shfl_xor_sync(kFull, h ? s1 : 0.0f, 4);
    h = lane & 2;
    float u = (h ? t1 : t0) + __shfl_xor_sync(kFull, h ? t0 : t1, 2);
    u += __shfl_xor_sync(kFull, u, 1);
    out = u;
    slot = ((lane & 16) ? 5 : 0) + ((lane & 8) ? 3 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
    valid = !(lane & 1) && ((lane & 8) ? !(lane & 4) : !((lane & 4) && (lane & 2)));
}

// Backward: ONE WARP PER (16x8 HALF TILE, 256-RECORD SEGMENT of the tile's list), FOUR PIXELS PER THREAD (a vertical
// strip x, y0..y0+3).  Measured at c3 (1M splats, 1080p): 5.16 M contributing (8x4 patch, splat) pairs but only 1.84 M
// contributing (16x8 patch, splat) pairs, so the warp reduction + REDs -- a third of the work of a one-pixel-per-thread
// kernel -- are paid 2.8x less often.  dx is shared by a thread's four pixels, so the five geometric gradients and
// dL/dopacity collapse into three per-thread moments
//   U0 = sum u,  U1 = sum u*dy,  U2 = sum u*dy^2     with u = o*G*dL/dalpha
// from which  d/dx = -(A dx U0 + B U1), d/dy = -(C U1 + B dx U0), dA = -dx^2 U0/2, dB = -dx U1,
// dC = -U2/2, do = U0/o.   The colour / depth composited BEHIND a splat enters dL/dalpha only through its dot product
// with the pixel's gradient vector g = (g_r, g_g, g_b, g_D), so it is tracked as ONE scalar per pixel:
//   h_i = c_i . g,   dL/dalpha_i = h_i T_i + Q_i / (1 - alpha_i),   Q_i = tail - sum_{j>i} h_j alpha_j T_j
// (tail = T_final (g_A - bg . g): the background and alpha-channel terms), updated as Q <- Q - h_i w_i.
//
// SEGMENTS.  A long list no longer serialises on one warp: the forward checkpoints (T, C, D) per pixel at every
// 256-record boundary it crosses (binning buffer `ckpt`, slot = list position >> 8), so the replay of segment
// [s0, s1) can start from   T = T_ckpt(s1),  Q = tail - g . (C_final - C_ckpt(s1))   for every pixel whose last
// contributor lies beyond s1 (and from T_final, Q = tail for the pixels that end inside it) -- exactly the state the
// sequential back-to-front walk has when it reaches s1.  Work units = 2 per tile (first segment of every tile) + 2 per
// checkpoint slot; PERSISTENT warps (4 independent warps per CTA, own shared-memory stages and mbarriers, no block
// barrier) fetch units from a global counter, so long and short lists balance across the 148 SMs.
constexpr int kBwdWarps = 4;
constexpr int kBwdThreads = 32 * kBwdWarps;
constexpr int kBwdBatch = 64;
constexpr int kPix = 4;
constexpr int kSeg = 256;                       // == kBatch: the forward checkpoints once per staged batch

// FLAGS: also write the contributor bytes (TgsSettings.contrib_flags).  A template parameter, not a run-time test: the
// kernel is bound by instruction issue, and even a never-taken branch on a NULL pointer cost 2.5 % (0.554 -> 0.568 ms).
template <bool FLAGS>
__global__ void __launch_bounds__(kBwdThreads, 5)
k_render_bwd(const uint2* __restrict__ ranges, const uint32_t* __restrict__ ids, const float4* __restrict__ table,
             int W, int H, int Tx,
             int row0, const float* __restrict__ bg, int normalize, float alpha_max, const float* __restrict__ final_T,
             const uint32_t* __restrict__ n_contrib, const float* __restrict__ depth_raw,
             const float* __restrict__ color_acc, const float* __restrict__ ckpt,
             const uint32_t* __restrict__ slot_tile, const uint32_t* __restrict__ ckpt_list,
             uint32_t* __restrict__ work_counter, int n_first,
             const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
             const float* __restrict__ dL_dalpha, const float* __restrict__ t_target,
             const float* __restrict__ t_weight, const float* __restrict__ t_scale,
             const float* __restrict__ t_gscale, int t_mode,
             int t_row0, int t_row1, float* __restrict__ residual, float* __restrict__ sgrad,
             uint8_t* __restrict__ contrib) {
    __shared__ __align__(128) float4 sbuf_all[kBwdWarps][2][kBwdBatch * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 (*sbuf)[kBwdBatch * 3] = sbuf_all[warp];
    const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
    // loss scale (mult / Z) times the upstream gradient of the touch-loss scalar (NULL = 1: the loss enters the
    // caller's objective with unit weight)
    const float tscale = (t_target != nullptr && t_mode != TGS_LOSS_NONE)
                             ? t_scale[0] * (t_gscale ? t_gscale[0] : 1.0f) : 0.0f;
    const size_t HW = (size_t)W * H;
    // work units: the first segment of both half tiles of every tile, then both halves of every checkpointed slot
    const uint32_t n_units = (uint32_t)n_first + 2u * work_counter[1];

    for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(work_counter, 1u);
        u = __shfl_sync(kFull, u, 0);
        if (u >= n_units) break;
        // ---- decode the work unit: (tile, half, segment)
        int tile, half, seg;
        uint2 rng;
        if (u < (uint32_t)n_first) {
            tile = (int)(u >> 1) + row0 * Tx; half = (int)(u & 1u); seg = 0;
            rng = ranges[tile];
        } else {
            const uint32_t slot = ckpt_list[(u - (uint32_t)n_first) >> 1];
            half = (int)((u - (uint32_t)n_first) & 1u);
            tile = (int)slot_tile[slot];
            rng = ranges[tile];
            seg = (int)(slot - (rng.x >> 8));
        }
        const int len = (int)(rng.y - rng.x);
        const int s0 = seg * kSeg;                     // this unit replays list positions [s0, s1)
        const int s1 = min(len, s0 + kSeg);
        const int tx = tile % Tx, ty = tile / Tx;
        const int px = tx * TGS_TILE + (lane & 15);
        const int py0 = ty * TGS_TILE + half * 8 + (lane >> 4) * kPix;
        PixelMap pm;                                   // only the cull rectangle of this warp is used
        pm.x0 = (float)(tx * TGS_TILE); pm.x1 = pm.x0 + 15.0f;
        pm.y0 = (float)(ty * TGS_TILE + half * 8); pm.y1 = pm.y0 + 7.0f;
        const float fx = (float)px, fy0 = (float)py0;

        // ---- per-pixel state and the FUSED touch-depth gradient (SURVEY A6 "Fusion")
        float T[kPix], g0[kPix], g1[kPix], g2[kPix], gD[kPix], Q[kPix];
        uint32_t nc[kPix];
        uint32_t wmax = 0;
        const float* ck = ckpt + (size_t)((rng.x + (uint32_t)s1) >> 8) * TGS_CKPT_FLOATS;
#pragma unroll
        for (int r = 0; r < kPix; ++r) {
            const int py = py0 + r;
            T[r] = 1.0f; g0[r] = g1[r] = g2[r] = gD[r] = Q[r] = 0.0f; nc[r] = 0;
            if (px < W && py < H) {
                const int pix = py * W + px;
                const uint32_t ncp = n_contrib[pix];
                const float Tf = final_T[pix];
                const float A = 1.0f - Tf;
                const float D = depth_raw[pix];
                float res = 0.0f;
                float gDhat = dL_ddepth ? dL_ddepth[pix] : 0.0f;
                const bool touch_px = t_target != nullptr && A > 0.0f && py >= t_row0 && py < t_row1;
                if (ncp > (uint32_t)s0 || (seg == 0 && residual != nullptr)) {
                    if (touch_px) {
                        const float tgt = t_target[pix];
                        if (tgt > 0.0f) {
                            const float dhat = normalize ? D / A : D;
                            res = dhat - tgt;
                            if (t_mode != TGS_LOSS_NONE) {
                                const float wgt = (t_weight ? t_weight[pix] : 1.0f) * tscale;
                                gDhat += (t_mode == TGS_LOSS_L1) ? wgt * (float)((res > 0.0f) - (res < 0.0f))
                                                                 : 2.0f * wgt * res;
                            }
                        }
                    }
                    if (seg == 0 && residual) residual[pix] = res;
                }
                if (ncp > (uint32_t)s0) {              // this pixel blends something inside [s0, s1)
                    nc[r] = ncp;
                    g0[r] = dL_dcolor[pix]; g1[r] = dL_dcolor[HW + pix]; g2[r] = dL_dcolor[2 * HW + pix];
                    float gA = dL_dalpha ? dL_dalpha[pix] : 0.0f;
                    if (normalize) {
                        if (A > 0.0f) { gD[r] = gDhat / A; gA -= gDhat * D / (A * A); }
                    } else {
                        gD[r] = gDhat;
                    }
                    // colour: d(T_final*bg)/dalpha_i = -T_final/(1-alpha_i)*bg ; alpha: dA/dalpha_i = +T_final/(1-alpha_i)
                    Q[r] = Tf * (gA - (bg0 * g0[r] + bg1 * g1[r] + bg2 * g2[r]));
                    if (ncp <= (uint32_t)s1) {
                        T[r] = Tf;                     // the pixel's last contributor lies in this segment
                    } else {
                        // state of the sequential walk when it arrives at s1, from the forward's checkpoint
                        const int pidx = ((py & 15) << 4) + (px & 15);
                        T[r] = ck[pidx];
                        float behind = g0[r] * (color_acc[pix] - ck[256 + pidx]);
                        behind = fmaf(g1[r], color_acc[HW + pix] - ck[512 + pidx], behind);
                        behind = fmaf(g2[r], color_acc[2 * HW + pix] - ck[768 + pidx], behind);
                        behind = fmaf(gD[r], D - ck[1024 + pidx], behind);
                        Q[r] -= behind;
                    }
                    wmax = max(wmax, min(ncp, (uint32_t)s1));
                }
            }
        }
        // ---- nothing beyond the deepest contributor of any pixel of this half tile needs replaying
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(kFull, wmax, o));
        const int leff = (int)wmax - s0;               // records of the segment to replay (<= 0: none)
        const int nb = (leff + kBwdBatch - 1) / kBwdBatch;
        if (nb <= 0) continue;

        // staging: the segment's records are gathered by Gaussian id from the per-Gaussian table (L2 resident), two per
        // lane and batch, with 16-byte cp.async copies two batches ahead; the ids of the batch after that are prefetched
        // into registers so that no gather waits for its ids
        const uint32_t* seg_ids = ids + rng.x + s0;
        auto load_ids = [&](int q, uint32_t& i0, uint32_t& i1) {      // sequence step q stages batch nb-1-q (back to front)
            const int b = nb - 1 - q;
            const int cnt = min(kBwdBatch, leff - b * kBwdBatch);
            const uint32_t* p = seg_ids + b * kBwdBatch;
            i0 = lane < cnt ? __ldg(p + lane) : 0u;
            i1 = lane + 32 < cnt ? __ldg(p + lane + 32) : 0u;
        };
        auto gather = [&](int q, uint32_t i0, uint32_t i1) {
            const int b = nb - 1 - q;
            const int cnt = min(kBwdBatch, leff - b * kBwdBatch);
            float4* dst = sbuf[q & 1];
            if (lane < cnt) {
                const float4* src = table + (size_t)3 * i0;
                cp_async16(dst + 3 * lane, src); cp_async16(dst + 3 * lane + 1, src + 1); cp_async16(dst + 3 * lane + 2, src + 2);
            }
            if (lane + 32 < cnt) {
                const float4* src = table + (size_t)3 * i1;
                float4* d = dst + 3 * (lane + 32);
                cp_async16(d, src); cp_async16(d + 1, src + 1); cp_async16(d + 2, src + 2);
            }
            cp_async_commit();
        };
        uint32_t n0 = 0, n1 = 0;
        {
            uint32_t a0, a1;
            load_ids(0, a0, a1);
            if (nb > 1) load_ids(1, n0, n1);
            gather(0, a0, a1);
            if (nb > 1) gather(1, n0, n1);
            if (nb > 2) load_ids(2, n0, n1);
        }

        for (int q = 0; q < nb; ++q) {
            if (q + 1 < nb) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            const int b = nb - 1 - q;
            const int cnt = min(kBwdBatch, leff - b * kBwdBatch);
            const float4* s = sbuf[q & 1];
            const int base = s0 + b * kBwdBatch;
            for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                const int jl = c0 + lane;
                bool pass = false;
                if (jl < cnt) pass = patch_may_touch(s[3 * jl], s[3 * jl + 1], s[3 * jl + 2].w, pm);
                unsigned mask = __ballot_sync(kFull, pass);
                while (mask) {
                    const int hb = 31 - __clz(mask);
                    mask &= ~(1u << hb);
                    const int j = c0 + hb;
                    const uint32_t idx = (uint32_t)(base + j);
                    const float4 a = s[3 * j], cq = s[3 * j + 1];
                    // same pinned rounding sequence as splat_power(): ax, ax*dx and B*dx are shared by the 4 pixels
                    const float dx = a.x - fx;
                    const float t1 = __fmul_rn(__fmul_rn(cq.x, dx), dx);
                    const float bxd = __fmul_rn(cq.y, dx);
                    float og[kPix], dy[kPix];             // og = o*G where the pixel blends this splat, else 0
                    bool anyv = false;
#pragma unroll
                    for (int r = 0; r < kPix; ++r) {
                        dy[r] = a.y - (fy0 + (float)r);
                        const float sq = __fmaf_rn(__fmul_rn(cq.z, dy[r]), dy[r], t1);
                        const float power = __fmaf_rn(-0.5f, sq, -__fmul_rn(bxd, dy[r]));
                        const float oG = __fmul_rn(cq.w, splat_exp(power));
                        const bool valid = (idx < nc[r]) && (power <= 0.0f) && (oG >= TGS_ALPHA_MIN);   // min(0.99,oG) >= 1/255 <=> oG >= 1/255
                        og[r] = valid ? oG : 0.0f;
                        anyv |= valid;
                    }
                    if (!__any_sync(kFull, anyv)) continue;
                    const float4 c = s[3 * j + 2];
                    float U0 = 0.f, U1 = 0.f, U2 = 0.f, V0 = 0.f, V1 = 0.f, V2 = 0.f, VD = 0.f;
#pragma unroll
                    for (int r = 0; r < kPix; ++r) {
                        // pixels that do not blend this splat run with og == 0: alpha == 0, inv == 1, T and Q untouched,
                        // u == 0 and w == 0, so every gradient term is exactly 0 without any branch
                        const float h = fmaf(a.z, gD[r], fmaf(c.z, g2[r], fmaf(c.y, g1[r], c.x * g0[r])));
                        const float am = fminf(alpha_max, og[r]);
                        const float inv = fast_rcp(1.0f - am);
                        T[r] *= inv;                               // transmittance in front of this splat
                        const float w = am * T[r];
                        const float dLda = fmaf(h, T[r], Q[r] * inv);
                        Q[r] = fmaf(-h, w, Q[r]);
                        const float uu = og[r] * dLda;             // straight-through alpha clamp: o*G, not alpha
                        U0 += uu;
                        const float udy = uu * dy[r];
                        U1 += udy;
                        U2 = fmaf(udy, dy[r], U2);
                        V0 = fmaf(w, g0[r], V0); V1 = fmaf(w, g1[r], V1); V2 = fmaf(w, g2[r], V2);
                        VD = fmaf(w, gD[r], VD);
                    }
                    float v[TGS_NGRAD];
                    const float dxU0 = dx * U0;
                    v[0] = -(cq.x * dxU0 + cq.y * U1);
                    v[1] = -(cq.z * U1 + cq.y * dxU0);
                    v[2] = -0.5f * dx * dxU0;
                    v[3] = -dx * U1;
                    v[4] = -0.5f * U2;
                    v[5] = U0 * fast_rcp(cq.w);
                    v[6] = V0; v[7] = V1; v[8] = V2; v[9] = VD;
                    float sum; int slot; bool ok;
                    warp_reduce_scatter10(v, lane, sum, slot, ok);
                    if (ok) atomicAdd(sgrad + (size_t)__float_as_int(a.w) * TGS_NGRAD + slot, sum);
                    // contributor byte (TgsSettings.contrib_flags): some pixel blended this Gaussian, its row is live.
                    // One idle lane of the reduction stores it; every writer stores the same value.
                    if (FLAGS && lane == 31) contrib[__float_as_int(a.w)] = 1;
                }
            }
            __syncwarp();
            if (q + 2 < nb) {
                gather(q + 2, n0, n1);
                if (q + 3 < nb) load_ids(q + 3, n0, n1);
            }
        }
    }
}

// -------------------------------------------------------------------- touch loss scale
__global__ void k_count_valid(const float* __restrict__ target, int64_t P, unsigned int* __restrict__ counter) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int c = 0;
    for (; i < P; i += stride) c += target[i] > 0.0f ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, c);
}
__global__ void k_finish_scale(float mult, float norm, float* out) {
    unsigned int cnt = reinterpret_cast<unsigned int*>(out)[1];
    float Z = norm > 0.0f ? norm : fmaxf(1.0f, (float)cnt);
    out[0] = mult / Z;
}

// value of the fused touch loss from the residual image (0 where invalid): scale * sum w*|r| (l1) or w*r^2 (l2)
__global__ void k_touch_loss_value(const float* __restrict__ residual, const float* __restrict__ weight, int64_t i0,
                                   int64_t i1, int mode, double* __restrict__ acc) {
    int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float s = 0.0f;
    for (; i < i1; i += stride) {
        const float r = residual[i], w = weight ? weight[i] : 1.0f;
        s += (mode == TGS_LOSS_L1) ? w * fabsf(r) : w * r * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0f) atomicAdd(acc, (double)s);
}
__global__ void k_touch_loss_finish(const double* acc, const float* scale, float* out) { out[0] = (float)(acc[0] * (double)scale[0]); }

}  // namespace

int tgs_launch_touch_loss_value(const float* residual, const float* weight, int64_t i0, int64_t i1, int mode,
                                const float* scale, double* acc, float* out, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_LOSS_SCALE, st);
    TGS_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
    if (i1 > i0 && mode != TGS_LOSS_NONE) {
        int blocks = (int)((i1 - i0 + 256 * 8 - 1) / (256 * 8));
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_touch_loss_value<<<blocks, 256, 0, st>>>(residual, weight, i0, i1, mode, acc);
        tgs_count_own(1);
    }
    k_touch_loss_finish<<<1, 1, 0, st>>>(acc, scale, out);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

int tgs_launch_render_fwd(const TgsCam& cam, const TgsSettings* s, const TgsRecord* gv_records, BinView bv, ImageView iv,
                          int64_t capacity, float* out_color, float* out_depth, float* out_alpha,
                          const float* touch_target, float* residual_out, cudaStream_t st) {
    int nt = cam.Tx * (cam.row1 - cam.row0);
    if (nt <= 0) return 0;
    TgsProfScope prof(TGS_STAGE_RENDER_FWD, st);
    TGS_CUDA(cudaMemsetAsync(bv.work_counter, 0, 2 * sizeof(uint32_t), st));
    static int waitgroup = -1;
    if (waitgroup < 0) {
        const char* e = getenv("TGS_FWD_RECORD_SYNC");
        waitgroup = (e && e[0] == 'w') ? 1 : 0;
    }
    auto kern = waitgroup ? k_render_fwd<false> : k_render_fwd<true>;
    kern<<<nt, 256, 0, st>>>(iv.ranges, (uint32_t)(capacity > 0xFFFFFFFFll ? 0xFFFFFFFFll : capacity), bv.vals_sorted,
                                     reinterpret_cast<const float4*>(gv_records), cam.W, cam.H, cam.Tx, cam.row0, s->bg,
                                     s->depth_normalize, cam.alpha_max, out_color, out_depth, out_alpha, iv.final_T,
                                     iv.n_contrib, iv.depth_raw, iv.color_acc, bv.ckpt, bv.slot_tile, bv.ckpt_list,
                                     bv.work_counter + 1, touch_target, residual_out);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    return 0;
}

int tgs_launch_render_bwd(const TgsCam& cam, const TgsSettings* s, const TgsRecord* gv_records, BinView bv, ImageView iv,
                          int64_t num_rendered,
                          const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                          const TgsTouch* touch, float* residual, float* screen_grads, uint8_t* contrib_flags,
                          cudaStream_t st) {
    int nt = cam.Tx * (cam.row1 - cam.row0);
    if (nt <= 0) return 0;
    const float* tt = nullptr; const float* tw = nullptr; const float* ts = nullptr; const float* tg = nullptr;
    int mode = TGS_LOSS_NONE;
    int tr0 = 0, tr1 = cam.H;
    if (touch && touch->target) {
        tt = touch->target; tw = touch->weight; ts = touch->scale; tg = touch->grad_scale; mode = touch->mode;
        if (touch->row_end > touch->row_begin) { tr0 = touch->row_begin; tr1 = touch->row_end; }
        if (mode != TGS_LOSS_NONE && ts == nullptr) { tgs_set_error("touch loss enabled but scale pointer is NULL"); return TGS_EINVAL; }
    }
    TgsProfScope prof(TGS_STAGE_RENDER_BWD, st);
    // work units: two half tiles per tile (first segment) + two per checkpointed 256-record boundary (counted on the
    // device by the forward); the grid is sized for the upper bound
    const int64_t slots = (num_rendered + 255) >> 8;
    const int64_t units = 2 * (int64_t)nt + 2 * slots;
    if (units > 0x7FFFFFFFll) { tgs_set_error("render backward: too many work units"); return TGS_EINVAL; }
    static int ctas_per_sm[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    TGS_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && ctas_per_sm[dev] == 0) {
        TGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[dev], k_render_bwd<false>, kBwdThreads, 0));
        TGS_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
    }
    const int cps = (dev >= 0 && dev < 64 && ctas_per_sm[dev] > 0) ? ctas_per_sm[dev] : 4;
    const int sms = (dev >= 0 && dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
    int64_t grid = (units + kBwdWarps - 1) / kBwdWarps;
    if (grid > (int64_t)cps * sms) grid = (int64_t)cps * sms;        // persistent: one resident wave
    TGS_CUDA(cudaMemsetAsync(bv.work_counter, 0, sizeof(uint32_t), st));
    auto kern = contrib_flags ? k_render_bwd<true> : k_render_bwd<false>;
    kern<<<(unsigned)grid, kBwdThreads, 0, st>>>(iv.ranges, bv.vals_sorted, reinterpret_cast<const float4*>(gv_records), cam.W, cam.H, cam.Tx, cam.row0, s->bg,
                                     s->depth_normalize, cam.alpha_max, iv.final_T, iv.n_contrib, iv.depth_raw,
                                     iv.color_acc, bv.ckpt, bv.slot_tile, bv.ckpt_list, bv.work_counter, 2 * nt, dL_dcolor,
                                     dL_ddepth, dL_dalpha, tt, tw, ts, tg, mode, tr0, tr1, residual, screen_grads,
                                     contrib_flags);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    return 0;
}

int tgs_launch_loss_scale(const float* target, int64_t P, float mult, float norm, float* out, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_LOSS_SCALE, st);
    TGS_CUDA(cudaMemsetAsync(out, 0, 8, st));
    if (norm <= 0.0f && P > 0) {
        int blocks = (int)((P + 256 * 8 - 1) / (256 * 8));
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_count_valid<<<blocks, 256, 0, st>>>(target, P, reinterpret_cast<unsigned int*>(out) + 1);
        tgs_count_own(1);
    }
    k_finish_scale<<<1, 1, 0, st>>>(mult, norm, out);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
