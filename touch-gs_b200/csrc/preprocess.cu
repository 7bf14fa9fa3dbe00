// preprocess.cu -- per-Gaussian kernels: forward projection, its backward (A9)
// and markVisible.  THIS TRANSLATION UNIT IS COMPILED WITH --fmad=false so that the integer
// results derived from float math (radius, tile rectangle, depth sort key) are bit-identical to
// the oracle (oracle/gs_oracle.py::preprocess), which rounds every operation separately.
//
// Roofline: HBM.  Algorithmic bytes per Gaussian: forward reads 4*(11+3K) and writes
// a 48 B record + 24 B cov3D + 4 B tiles + 4 B radii + 8 B rect + 1 B clamp mask; backward reads the
// parameters again plus 40 B screen gradients and writes 4*(11+3K)+12 B of gradients.
#include "tgs_common.cuh"
#include <cstdlib>

namespace {

constexpr int kBlock = 256;

struct CamMats { float vm[16]; float pm[16]; float cp[3]; };

__device__ __forceinline__ void load_cam(const float* vm, const float* pm, const float* cp, CamMats* sm) {
    int t = threadIdx.x;
    if (t < 16) sm->vm[t] = vm[t];
    else if (t < 32) sm->pm[t - 16] = pm[t - 16];
    else if (t < 35) sm->cp[t - 32] = cp ? cp[t - 32] : 0.0f;
    __syncthreads();
}

// ---- coalesced staging of the SH block (K = 16: 192 B per Gaussian, the bulk of this path's HBM traffic)
// One thread per Gaussian reading its own 192 bytes makes every warp-wide 16-byte load touch 32 different lines.
// Instead each WARP moves the 6 KB block of its 32 consecutive Gaussians between HBM and shared memory with fully
// coalesced 16-byte accesses (12 independent loads in flight per lane), and the threads work on their own row in
// shared memory.  Rows are padded from 12 to 13 float4 so that the per-thread float4 accesses of a quarter warp hit
// 32 distinct banks.
constexpr int kShRowF4 = 12;                 // float4 per Gaussian at K = 16
constexpr int kShRowPad = kShRowF4 + 1;
constexpr size_t kShStageBytes = (size_t)kBlock * kShRowPad * sizeof(float4);   // 53,248 B per CTA

__device__ __forceinline__ void warp_stage_in(const float* __restrict__ g_base, int first, int N, float4* s_rows, int lane) {
    const float4* src = reinterpret_cast<const float4*>(g_base) + (size_t)first * kShRowF4;
    const int total = min(32, N - first) * kShRowF4;
#pragma unroll
    for (int it = 0; it < kShRowF4; ++it) {
        const int e = it * 32 + lane;
        if (e < total) s_rows[(e / kShRowF4) * kShRowPad + (e % kShRowF4)] = __ldg(src + e);
    }
    __syncwarp();
}
__device__ __forceinline__ void warp_stage_out(float* __restrict__ g_base, int first, int N, const float4* s_rows, int lane) {
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(g_base) + (size_t)first * kShRowF4;
    const int total = min(32, N - first) * kShRowF4;
#pragma unroll
    for (int it = 0; it < kShRowF4; ++it) {
        const int e = it * 32 + lane;
        if (e < total) dst[e] = s_rows[(e / kShRowF4) * kShRowPad + (e % kShRowF4)];
    }
}

// Row-wise staging of a FEW Gaussians of the warp (bit k of `mask` = lane k's Gaussian): 12 lanes move one 192-byte row.
// Used when most of the warp's Gaussians do not need their colour (culled, or outside this rank's tile-row band).
__device__ __forceinline__ void warp_stage_rows(const float* __restrict__ g_base, int first, float4* s_rows, unsigned mask, int lane) {
    const float4* src = reinterpret_cast<const float4*>(g_base) + (size_t)first * kShRowF4;
    while (mask) {
        const int k0 = __ffs(mask) - 1;
        mask &= mask - 1;
        int k1 = -1;
        if (mask) { k1 = __ffs(mask) - 1; mask &= mask - 1; }
        // lanes 0..11 take row k0, lanes 16..27 row k1
        const int k = lane < 16 ? k0 : k1, e = lane & 15;
        if (k >= 0 && e < kShRowF4) s_rows[k * kShRowPad + e] = __ldg(src + k * kShRowF4 + e);
    }
    __syncwarp();
}

// colour sums of one staged row: four coefficients' worth of floats at a time, each channel accumulated in ascending-k
// order exactly like tgs_sh_forward (bit-identical), without a 48-register copy of the row.  Shared by the forward and
// by the backward's recomputation of the clamp mask, so both see the same bits.
__device__ __forceinline__ void staged_row_color(const float4* s_row, const float* b, int deg, float* acc) {
    acc[0] = acc[1] = acc[2] = 0.0f;
    const int nf = 3 * (deg + 1) * (deg + 1);
#pragma unroll
    for (int k4 = 0; k4 < kShRowF4; ++k4) {
        if (4 * k4 < nf) {
            const float4 v = s_row[k4];
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 4 * k4 + j;
                if (e < nf) acc[e % 3] += b[e / 3] * vv[j];
            }
        }
    }
}
#define TGS_CLAMP_UNKNOWN 0x80u   /* clamp-mask byte of a visible Gaussian whose colour this rank did not evaluate */

template <bool STAGE>
__global__ void __launch_bounds__(kBlock)
k_preprocess(int N, const float* __restrict__ means, const float* __restrict__ scales,
             const float* __restrict__ rots, const float* __restrict__ opac,
             const float* __restrict__ shs, const float* __restrict__ colors,
             const float* __restrict__ covpre, const float* __restrict__ vm,
             const float* __restrict__ pm, const float* __restrict__ campos, TgsCam cam,
             TgsRecord* __restrict__ rec, float* __restrict__ cov3D,
             uint32_t* __restrict__ tiles, uint8_t* __restrict__ clamped,
             uint2* __restrict__ rect, uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ ids,
             int32_t* __restrict__ radii) {
    __shared__ CamMats cm;
    extern __shared__ __align__(16) float4 s_stage[];
    load_cam(vm, pm, campos, &cm);
    int i = blockIdx.x * kBlock + threadIdx.x;
    const bool inN = i < N;
    float x = 0.f, y = 0.f, z = 0.f;
    float cov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    TgsProj p;
    bool vis = false;
    if (inN) {
        x = means[3 * i]; y = means[3 * i + 1]; z = means[3 * i + 2];
        if (covpre) {
#pragma unroll
            for (int k = 0; k < 6; ++k) cov[k] = covpre[6 * i + k];
        } else {
            float4 q = reinterpret_cast<const float4*>(rots)[i];
            float sc[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
            float qq[4] = {q.x, q.y, q.z, q.w};
            tgs_cov3d(sc, cam.mod, qq, cov);
        }
        vis = tgs_project(cm.vm, cm.pm, cam, x, y, z, cov, p);
    }
    float rgb[3] = {0.f, 0.f, 0.f};
    unsigned cl = 0;
    if (STAGE) {
        // The colour is consumed only by the instances a Gaussian emits ON THIS RANK (p.tiles > 0: inside the tile-row
        // band), so only those read their 192-byte SH row (K = 16) -- culled Gaussians and, on a tile-row shard, the
        // (g-1)/g of the scene outside the band skip it.  The warp stages its whole 6 KB block with coalesced loads when
        // most rows are needed, else row by row.  A visible Gaussian without colour gets the clamp mask "unknown": the
        // backward (which runs for every visible Gaussian on every rank) recomputes it from the row it stages anyway.
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float4* s_rows = s_stage + (size_t)warp * 32 * kShRowPad;
        const int first = blockIdx.x * kBlock + warp * 32;
        const bool need = vis && p.tiles > 0;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (__popc(m) > 12) warp_stage_in(shs, first, N, s_rows, lane);
        else if (m) warp_stage_rows(shs, first, s_rows, m, lane);
        if (need) {
            float b[16], acc[3];
            tgs_sh_bases(cam.deg, x - cm.cp[0], y - cm.cp[1], z - cm.cp[2], b);
            staged_row_color(s_rows + lane * kShRowPad, b, cam.deg, acc);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float a = acc[c] + 0.5f;
                if (a < 0.0f) { cl |= (1u << c); a = 0.0f; }
                rgb[c] = a;
            }
        } else if (vis) {
            cl = TGS_CLAMP_UNKNOWN;
        }
    }
    if (!inN) return;
    if (STAGE) {
    } else if (vis) {
        if (shs && !STAGE) {
            // K*3 contiguous floats per Gaussian; 16-byte loads (K*12 B is a multiple of 16 for K in {1,4,9,16}
            // only when K*3 % 4 == 0, so fall back to scalar loads otherwise)
            float sh[48];
#pragma unroll
            for (int k = 0; k < 48; ++k) sh[k] = 0.0f;
            const float* src = shs + (size_t)3 * cam.K * i;
            int nb = (cam.deg + 1) * (cam.deg + 1);
            if (((3 * cam.K) & 3) == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(src);
                int n4 = (3 * nb + 3) >> 2;
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    if (k < n4) {
                        float4 v = __ldg(s4 + k);
                        // entries above 3*nb may belong to inactive bands: mask them to zero
                        sh[4 * k] = (4 * k < 3 * nb) ? v.x : 0.0f;
                        sh[4 * k + 1] = (4 * k + 1 < 3 * nb) ? v.y : 0.0f;
                        sh[4 * k + 2] = (4 * k + 2 < 3 * nb) ? v.z : 0.0f;
                        sh[4 * k + 3] = (4 * k + 3 < 3 * nb) ? v.w : 0.0f;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 48; ++k)
                    if (k < 3 * nb) sh[k] = __ldg(src + k);
            }
            tgs_sh_forward(cam.deg, sh, x - cm.cp[0], y - cm.cp[1], z - cm.cp[2], rgb, cl);
        } else if (colors) {                  // (neither: geometry-only projection, tgs_project_gaussians)
            rgb[0] = colors[3 * i]; rgb[1] = colors[3 * i + 1]; rgb[2] = colors[3 * i + 2];
        }
    }
    const float o = opac ? opac[i] : 1.0f;
    TgsRecord r;
    r.a = make_float4(p.px, p.py, p.depth, __int_as_float(i));
    r.b = make_float4(p.conA, p.conB, p.conC, vis ? o : 0.0f);
    // c.w = power threshold of the alpha >= 1/255 test: o*exp(power) >= 1/255  <=>  power >= -ln(255 o)
    r.c = make_float4(rgb[0], rgb[1], rgb[2], vis ? -logf(255.0f * o) : 3.0e38f);
    rec[i] = r;
#pragma unroll
    for (int k = 0; k < 6; ++k) cov3D[6 * i + k] = cov[k];
    tiles[i] = (uint32_t)p.tiles;
    clamped[i] = (uint8_t)cl;
    rect[i] = make_uint2((uint32_t)p.rminx | ((uint32_t)p.rmaxx << 16),
                         (uint32_t)p.rminy | ((uint32_t)p.rmaxy << 16));
    radii[i] = p.radius;
    depth_keys[i] = p.tiles > 0 ? __float_as_uint(p.depth) : 0xFFFFFFFFu;   // depth > 0.2: bit order == float order
    ids[i] = (uint32_t)i;
}

// Multi-GPU exchange FUSED into the consumer: instead of an all-reduce of the [N,10] screen-space
// gradient buffers between BACKWARD::render and this kernel, every rank reads the partial sums straight out of its
// peers' buffers over NVLink (P2P loads through NVSwitch) while it does the chain rule.  Every rank preprocesses all N
// Gaussians, so each can tell from a Gaussian's (unclipped) tile-row span which ranks' bands it touches: only those
// ranks' rows are read -- about one 40-byte row per Gaussian instead of the 2 x 40 B x (g-1)/g an all-reduce moves --
// and they are summed in ascending rank order on every rank, so the replicas stay bit-identical.
struct PeerGather {
    int world;                                   // 0: single buffer (`sgrad`), no exchange
    const float* ptr[TGS_MAX_PEERS];             // peer r's [N,10] partial screen-gradient buffer
    int row0[TGS_MAX_PEERS], row1[TGS_MAX_PEERS];  // tile-row band rendered by rank r
    size_t flag_off;                             // SKIP == 2: byte offset of the contributor bytes in every buffer
};
// 32 bytes that are 0 or 1 -> one bit each (byte k -> bit k)
__device__ __forceinline__ unsigned flag_bits(uint4 lo, uint4 hi) {
    const unsigned w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) m |= (((w[k] & 0x01010101u) * 0x10204080u) >> 28) << (4 * k);
    return m;
}

// GATHER (multi-GPU): the remote rows are REQUESTED before anything else is read, so the NVLink round trip (a few
// microseconds through the switch) runs under the HBM round trips of the camera, the staging and the parameter loads.
// SKIP: a Gaussian whose screen-gradient row is exactly zero -- culled, off screen, or (the common case in a dense
// scene: 86 % at c3) hidden behind the depth at which its tiles saturate, so that no pixel ever blended it -- has zero
// parameter gradients: its 236 bytes of parameters and SH coefficients are not read and no chain rule is evaluated;
// zeros are written.  What the warp stages (the whole 6 KB SH block when most lanes need it, else row by row, exactly
// like the forward does for the Gaussians outside a rank's band) is decided
//   SKIP = 1: from the row itself, which therefore comes first;
//   SKIP = 2: from the CONTRIBUTOR BYTES BACKWARD::render wrote behind the rows (TgsSettings.contrib_flags): one
//             coalesced byte per Gaussian instead of a 40-byte row -- and in the multi-GPU gather one 32-byte sector per
//             (warp, peer) tells which of the peer's rows are worth a round trip at all;
//   SKIP = 0: nothing is skipped.
// (Exact zeros in give exact zeros out except for non-finite parameters, where 0 x inf would have produced NaN.)
template <bool STAGE, bool GATHER, int SKIP>
__global__ void __launch_bounds__(kBlock, 3)
k_preprocess_bwd(int N, PeerGather pg, const TgsRecord* __restrict__ rec, const float* __restrict__ means, const float* __restrict__ scales,
                 const float* __restrict__ rots, const float* __restrict__ shs,
                 const float* __restrict__ covpre, const float* __restrict__ vm,
                 const float* __restrict__ pm, const float* __restrict__ campos, TgsCam cam,
                 const float* __restrict__ cov3D, const uint8_t* __restrict__ clamped,
                 const int32_t* __restrict__ radii, const float* __restrict__ sgrad,
                 float* __restrict__ dmeans2D, float* __restrict__ dmeans3D,
                 float* __restrict__ dopacity, float* __restrict__ dshs, float* __restrict__ dcolors,
                 float* __restrict__ dscales, float* __restrict__ drots, float* __restrict__ dcov3D) {
    __shared__ CamMats cm;
    extern __shared__ __align__(16) float4 s_stage[];
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const bool inN = i < N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = blockIdx.x * kBlock + warp * 32;
    int rad = 0;
    float py = 0.0f;                                      // rec[i].a.y (0 when culled)
    unsigned live = 1;                                    // SKIP == 2, single buffer: this Gaussian's contributor byte
    uint4 fl_lo = make_uint4(0u, 0u, 0u, 0u), fl_hi = fl_lo;   // SKIP == 2, gather: lane r holds peer r's 32 bytes
    float2 pre[TGS_NGRAD / 2];                            // the (first) screen-gradient row, in flight
#pragma unroll
    for (int k = 0; k < TGS_NGRAD / 2; ++k) pre[k] = make_float2(0.f, 0.f);
    if (GATHER && SKIP == 2 && first < N && lane < pg.world) {
        // the contributor bytes of this warp's 32 Gaussians on peer `lane`: one 32-byte sector over NVLink
        const uint4* f = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(pg.ptr[lane]) + pg.flag_off + first);
        fl_lo = f[0]; fl_hi = f[1];
    }
    if (inN) {
        rad = __ldg(radii + i);
        if (GATHER) {
            // the second value that decides which peers are asked.  (The branch further down tests BOTH values with a
            // bitwise &: otherwise ptxas sinks this load into the `rad > 0` branch and serialises the two round trips.)
            py = __ldg(reinterpret_cast<const float*>(rec + i) + 1);
        } else if (SKIP == 2) {
            live = reinterpret_cast<const uint8_t*>(sgrad)[pg.flag_off + i];
        } else {
            // rows of culled Gaussians are zero (the buffer is zeroed before BACKWARD::render): read without waiting for `rad`
            const float2* s2 = reinterpret_cast<const float2*>(sgrad + (size_t)TGS_NGRAD * i);
#pragma unroll
            for (int k = 0; k < TGS_NGRAD / 2; ++k) pre[k] = s2[k];
        }
    }
    load_cam(vm, pm, campos, &cm);
    // STAGE (K = 16): the warp's SH block comes in through shared memory, each thread turns its row into dL/dSH in
    // place, and the block goes back out to `dshs` with coalesced stores
    float4* s_rows = STAGE ? s_stage + (size_t)warp * 32 * kShRowPad : nullptr;
    unsigned rmask = 0;                                   // GATHER: the ranks that hold a partial row of this Gaussian
    if (GATHER) {
        if (inN && ((rad > 0) & (py > -3.0e38f))) {
            int y0, y1;                                   // the Gaussian's tile rows in the FULL image (getRect, unclipped)
            tgs_rect1(py, (float)rad, cam.Ty, y0, y1);
            for (int r = 0; r < pg.world; ++r)
                if (y0 < pg.row1[r] && y1 > pg.row0[r]) rmask |= 1u << r;
        }
        if (SKIP == 2 && first < N) {
            // ... of which only those that blended it have anything but zeros to give
            const unsigned mine = flag_bits(fl_lo, fl_hi);
            for (int r = 0; r < pg.world; ++r) {
                const unsigned m = __shfl_sync(0xffffffffu, mine, r);
                if (!((m >> lane) & 1u)) rmask &= ~(1u << r);
            }
        }
        if (rmask) {
            const float2* s2 = reinterpret_cast<const float2*>(pg.ptr[__ffs(rmask) - 1] + (size_t)TGS_NGRAD * i);
#pragma unroll
            for (int k = 0; k < TGS_NGRAD / 2; ++k) pre[k] = s2[k];
        }
    }
    const bool vis = inN && rad > 0;
    bool need = vis;
    if (SKIP == 2) {
        need = vis && (GATHER ? rmask != 0 : live != 0);
        if (!GATHER && need) {
            const float2* s2 = reinterpret_cast<const float2*>(sgrad + (size_t)TGS_NGRAD * i);
#pragma unroll
            for (int k = 0; k < TGS_NGRAD / 2; ++k) pre[k] = s2[k];
        }
    }
    if (STAGE && first < N) {
        if (SKIP == 0) {
            warp_stage_in(shs, first, N, s_rows, lane);
        } else if (SKIP == 2) {                           // known before the rows arrive: the staging overlaps them
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (__popc(m) > 12) warp_stage_in(shs, first, N, s_rows, lane);
            else if (m) warp_stage_rows(shs, first, s_rows, m, lane);
        }
    }
    float sg[TGS_NGRAD];
#pragma unroll
    for (int k = 0; k < TGS_NGRAD; ++k) sg[k] = 0.0f;
    if (need) {
        if (GATHER) {
            // ascending rank order, starting from 0.0f: bit-identical on every rank
            if (rmask) {
#pragma unroll
                for (int k = 0; k < TGS_NGRAD / 2; ++k) { sg[2 * k] += pre[k].x; sg[2 * k + 1] += pre[k].y; }
                rmask &= rmask - 1;
            }
            while (rmask) {
                const float2* s2 = reinterpret_cast<const float2*>(pg.ptr[__ffs(rmask) - 1] + (size_t)TGS_NGRAD * i);
                rmask &= rmask - 1;
#pragma unroll
                for (int k = 0; k < TGS_NGRAD / 2; ++k) { float2 v = s2[k]; sg[2 * k] += v.x; sg[2 * k + 1] += v.y; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < TGS_NGRAD / 2; ++k) { sg[2 * k] = pre[k].x; sg[2 * k + 1] = pre[k].y; }
        }
    }
    if (SKIP == 1) {
        bool nz = false;
#pragma unroll
        for (int k = 0; k < TGS_NGRAD; ++k) nz |= sg[k] != 0.0f;          // NaN != 0: a poisoned row is not skipped
        need = vis && nz;
        if (STAGE && first < N) {
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (__popc(m) > 12) warp_stage_in(shs, first, N, s_rows, lane);
            else if (m) warp_stage_rows(shs, first, s_rows, m, lane);
        }
    }
    if (inN) {
    float dm[3] = {0.f, 0.f, 0.f}, dc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    if (need) {
        float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
        float cov[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) cov[k] = cov3D[6 * i + k];
        tgs_project_backward(cm.vm, cm.pm, cam, x, y, z, cov, sg, dm, dc);
        if (shs) {
            if (STAGE) {
                float* row = reinterpret_cast<float*>(s_rows + lane * kShRowPad);
                unsigned cl = clamped[i];
                if (cl & TGS_CLAMP_UNKNOWN) {
                    // this rank's forward did not evaluate the colour (the Gaussian lies outside its tile-row band):
                    // recompute the clamp mask from the staged row with the forward's own arithmetic
                    float b[16], acc[3];
                    tgs_sh_bases(cam.deg, x - cm.cp[0], y - cm.cp[1], z - cm.cp[2], b);
                    staged_row_color(s_rows + lane * kShRowPad, b, cam.deg, acc);
                    cl = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (acc[c] + 0.5f < 0.0f) cl |= (1u << c);
                }
                tgs_sh_backward(cam.deg, cam.K, row, x - cm.cp[0], y - cm.cp[1], z - cm.cp[2], sg + 6, cl, row, dm,
                                false);
            } else {
                // streams the K coefficients in groups of 4 straight from / to global memory
                tgs_sh_backward(cam.deg, cam.K, shs + (size_t)3 * cam.K * i, x - cm.cp[0], y - cm.cp[1], z - cm.cp[2],
                                sg + 6, clamped[i], dshs + (size_t)3 * cam.K * i, dm);
            }
        }
        if (!covpre) {
            float4 q = reinterpret_cast<const float4*>(rots)[i];
            float sc[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
            float qq[4] = {q.x, q.y, q.z, q.w};
            tgs_cov3d_backward(sc, cam.mod, qq, dc, ds, dq);
        }
    } else if (dshs) {
        if (STAGE) {
            float4* row = s_rows + lane * kShRowPad;
#pragma unroll
            for (int k = 0; k < kShRowF4; ++k) row[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float* dst = dshs + (size_t)3 * cam.K * i;
            for (int k = 0; k < 3 * cam.K; ++k) dst[k] = 0.0f;
        }
    }
    dmeans2D[3 * i] = sg[0] * 0.5f * (float)cam.W;
    dmeans2D[3 * i + 1] = sg[1] * 0.5f * (float)cam.H;
    dmeans2D[3 * i + 2] = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) dmeans3D[3 * i + k] = dm[k];
    dopacity[i] = sg[5];
    if (dcolors) { dcolors[3 * i] = sg[6]; dcolors[3 * i + 1] = sg[7]; dcolors[3 * i + 2] = sg[8]; }
    if (dscales) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dscales[3 * i + k] = ds[k];
    }
    if (drots) reinterpret_cast<float4*>(drots)[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
    if (dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) dcov3D[6 * i + k] = dc[k];
    }
    }   // inN
    if (STAGE && first < N) warp_stage_out(dshs, first, N, s_rows, lane);
}

__global__ void k_mark_visible(int N, const float* __restrict__ means, const float* __restrict__ vm,
                               uint8_t* __restrict__ present) {   // markVisible is an Inria-convention entry point: 0.2
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float tz = tgs_xform(vm, means[3 * i], means[3 * i + 1], means[3 * i + 2], 2);
    present[i] = tz > TGS_NEAR_Z ? 1 : 0;
}

}  // namespace

// opt in to > 48 KB of dynamic shared memory for the staged kernels, once per device
static cudaError_t enable_stage_smem() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(k_preprocess<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_preprocess_bwd<true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kShStageBytes);
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

int tgs_launch_preprocess(const TgsCam& cam, const TgsSettings* s, const TgsGaussians* g,
                          GeomView gv, int32_t* radii, cudaStream_t st) {
    int N = g->N;
    if (N == 0) return 0;
    TgsProfScope prof(TGS_STAGE_PREPROCESS, st);
    const bool stage = g->shs != nullptr && cam.K == 16;
    TGS_CUDA(enable_stage_smem());
    auto kern = stage ? k_preprocess<true> : k_preprocess<false>;
    kern<<<(N + kBlock - 1) / kBlock, kBlock, stage ? kShStageBytes : 0, st>>>(
        N, g->means3D, g->scales, g->rotations, g->opacities, g->shs, g->colors_precomp,
        g->cov3D_precomp, s->viewmatrix, s->projmatrix, s->campos, cam, gv.records, gv.cov3D,
        gv.tiles_touched, gv.clamped, gv.rect, gv.depth_keys, gv.ids, radii);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    return 0;
}

int tgs_launch_preprocess_bwd(const TgsCam& cam, const TgsSettings* s, const TgsGaussians* g,
                              GeomView gv, const int32_t* radii, const float* screen_grads,
                              const float* const* peer_grads, const int32_t* peer_rows, int world, bool with_flags,
                              const TgsGrads* gr, cudaStream_t st) {
    int N = g->N;
    if (N == 0) return 0;
    PeerGather pg;
    pg.world = 0;
    pg.flag_off = TGS_SCREEN_GRAD_FLAG_OFFSET(N);
    for (int r = 0; r < TGS_MAX_PEERS; ++r) { pg.ptr[r] = nullptr; pg.row0[r] = pg.row1[r] = 0; }
    if (world > 0) {
        pg.world = world;
        for (int r = 0; r < world; ++r) { pg.ptr[r] = peer_grads[r]; pg.row0[r] = peer_rows[2 * r]; pg.row1[r] = peer_rows[2 * r + 1]; }
    }
    TgsProfScope prof(TGS_STAGE_PREPROCESS_BWD, st);
    const bool stage = g->shs != nullptr && cam.K == 16;
    TGS_CUDA(enable_stage_smem());
    // Which shortcut (see the kernel): contributor bytes when the buffers carry them; else the zero-row test for a single
    // buffer, and nothing for a gather over plain rows (waiting for the remote rows before the SH block is staged costs more
    // than it saves: 0.150 vs 0.138 ms on 2 GPUs).  TGS_ZERO_SKIP=0|1 overrides (experiments; 2 needs the bytes).
    static int skip_env = -2;
    if (skip_env == -2) {
        const char* e = getenv("TGS_ZERO_SKIP");
        skip_env = e ? atoi(e) : -1;
        if (skip_env < -1 || skip_env > 1) skip_env = -1;
    }
    const bool gather = world > 0;
    const int skip = skip_env >= 0 ? skip_env : (with_flags ? 2 : (gather ? 0 : 1));
#define TGS_PICK(ST, GA) (skip == 2 ? k_preprocess_bwd<ST, GA, 2> : skip == 1 ? k_preprocess_bwd<ST, GA, 1> : k_preprocess_bwd<ST, GA, 0>)
    auto kern = stage ? (gather ? TGS_PICK(true, true) : TGS_PICK(true, false))
                      : (gather ? TGS_PICK(false, true) : TGS_PICK(false, false));
#undef TGS_PICK
    kern<<<(N + kBlock - 1) / kBlock, kBlock, stage ? kShStageBytes : 0, st>>>(
        N, pg, gv.records, g->means3D, g->scales, g->rotations, g->shs, g->cov3D_precomp, s->viewmatrix,
        s->projmatrix, s->campos, cam, gv.cov3D, gv.clamped, radii, screen_grads, gr->dmeans2D,
        gr->dmeans3D, gr->dopacity, g->shs ? gr->dshs : nullptr, gr->dcolors, gr->dscales,
        gr->drotations, gr->dcov3D);
    tgs_count_own(1);
    TGS_KERNEL_CHECK(st, s->debug);
    return 0;
}

int tgs_launch_mark_visible(int N, const float* means, const float* vm, uint8_t* present, cudaStream_t st) {
    if (N == 0) return 0;
    k_mark_visible<<<(N + 255) / 256, 256, 0, st>>>(N, means, vm, present);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
