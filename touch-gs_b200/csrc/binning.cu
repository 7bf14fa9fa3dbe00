// binning.cu -- tile binning of the visible Gaussians, hand-written end to end except the
// depth sort of the N Gaussians.
//
// The spec'd RESULT is the list of (tile, Gaussian) instances sorted by the 64-bit key (tile << 32 | bits(depth)),
// ties in ascending Gaussian id (stable sort of a Gaussian-major emission), plus one [start,end) range per tile.
// We produce exactly that list (bit-exact against oracle/gs_oracle.py::bin_and_sort) WITHOUT ever materialising or
// sorting per-instance keys.  Every instance of a Gaussian comes from its tile RECTANGLE, so the position of
// instance (g, t) in the final list is
//       ranges[t].x + #{ g' before g in (depth, id) order : t in rect(g') },
// a counting problem over rectangles:
//   0. radix-sort the N Gaussians once by depth bits (CUB, 32-bit keys, N items; stable => equal depths keep
//      ascending id; Gaussians that emit nothing get key 0xFFFFFFFF) -> `order`;
//   1. k_bin_count: the depth order is cut into chunks of C Gaussians; one CTA per chunk counts how many of the
//      chunk's rectangles cover each tile (4 corner updates per rectangle into a shared-memory difference array +
//      a 2-D prefix sum) and writes its row of the [chunks x T] count matrix;
//   2. k_bin_prefix: per tile column, exclusive prefix over the chunks (in place) + the column total;
//      k_bin_ranges: exclusive scan of the totals -> ranges[t] and num_rendered (64-bit, overflow-checked);
//   3. k_bin_scatter: one WARP per chunk loads its row (+ the tile starts) into shared memory as per-tile cursors and
//      walks its C Gaussians IN ORDER; the lanes take the tiles of the current Gaussian's rectangle, bump the cursors
//      and write the Gaussian id to its final position.  Order inside a tile = order of the walk = depth order.
// Every tile's list is then one contiguous run of ids; the compositing kernels (render.cu) TMA-copy the ids and gather
// the 48-byte per-Gaussian records by id, so no per-instance copy of the records is ever written.
// Gaussians that emit nothing sort to the end of the depth order (key 0xFFFFFFFF), so only the leading chunks hold
// emitters: k_bin_count finds the number of LIVE chunks on the device (no host round trip), and the prefix and the
// scatter visit only those -- on a rank of the tile-row shard the whole chain costs what its band's emitters cost,
// not what the N replicated Gaussians cost.
// Tiles are processed in bands of tile rows (count: <= 8192 tiles = 32 KB of counters per CTA; scatter: ~1024 tiles
// per single-warp unit) so any image size fits.
//
// Replaces: emission of (tile id, Gaussian id) pairs + two CUB onesweep passes over I pairs + boundary detection
// (0.048 + 0.190 ms at c3, and 12 B/instance of key/value traffic per pass) by three small kernels over N Gaussians.
// Roofline: HBM.  Algorithmic bytes per instance: key/val write 12, sort 24, range detect 8; what
// actually moves: 4 B (the id) written once per instance, plus 8 B x live chunks x T of count-matrix traffic (0.06 GB
// at c3) -- the compositing kernels gather the 48-byte records by id, no per-instance copy of them exists.
#include "tgs_common.cuh"
#include <cub/cub.cuh>
#include <cstdlib>

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kCountCells = 16384;      // 64 KB difference array per count CTA (c3: 69 x 121 cells = 33 KB, one band)

struct Bands { int rows; int n; int tiles; };   // tile rows per band, number of bands (over the RENDERED rows), rows * Tx
struct Plan {
    int chunk;           // Gaussians per chunk (depth ranks)
    int nchunks;
    Bands count;         // k_bin_count: wide bands (<= TGS_BIN_BAND_TILES tiles of shared-memory counters per CTA)
    Bands scatter;       // k_bin_scatter: narrow bands -- the walk of a (chunk, band) unit is sequential, so many small
                         // units (a few tile rows, ~4 KB of cursors, 30+ resident warps per SM) beat few large ones
};
Bands make_bands(int max_tiles, int Tx, int Ty) {
    Bands b;
    b.rows = max_tiles / Tx;
    if (b.rows < 1) b.rows = 1;
    if (b.rows > Ty) b.rows = Ty;
    b.n = (Ty + b.rows - 1) / b.rows;
    b.tiles = b.rows * Tx;
    return b;
}
// Ty = number of tile rows this rank renders (its band of the tile-row shard, or the whole image)
Plan make_plan(int N, int Tx, int Ty, bool whole_image = true) {
    Plan p;
    p.chunk = 1024;
    while (p.chunk < 8192 && (N + p.chunk - 1) / p.chunk > 1536) p.chunk *= 2;
    p.nchunks = (N + p.chunk - 1) / p.chunk;
    if (p.nchunks < 1) p.nchunks = 1;
    // count kernel: a (rows + 1) x (Tx + 1) difference array of int32 in shared memory, <= kCountCells cells
    p.count.rows = kCountCells / (Tx + 1) - 1;
    if (p.count.rows < 1) p.count.rows = 1;
    if (p.count.rows > Ty) p.count.rows = Ty;
    p.count.n = (Ty + p.count.rows - 1) / p.count.rows;
    p.count.tiles = (p.count.rows + 1) * (Tx + 1);       // cells, not tiles
    // scatter bands: every (chunk, band) unit scans its whole chunk for hits, so the number of bands is kept around
    // 32 whatever the image size (256 tiles at c3, 1024 at 4K); TGS_SCATTER_TILES=<tiles per band> overrides (experiments)
    static int env_tiles = -1;
    if (env_tiles < 0) {
        const char* e = getenv("TGS_SCATTER_TILES");
        env_tiles = e ? atoi(e) : 0;
        if (env_tiles < 32 || env_tiles > TGS_BIN_SCATTER_MAX_TX) env_tiles = 0;
    }
    // measured at c3: whole image (one warp per unit) best with ~32 bands; a band of a tile-row shard (8 warps per
    // unit) best with ~4 bands (half image: 1024 tiles 0.09 ms vs 256 tiles 0.136; an eighth: 256 tiles 0.035 ms)
    int scatter_tiles = (Tx * Ty) / (whole_image ? 32 : 4);
    if (scatter_tiles < TGS_BIN_SCATTER_TILES) scatter_tiles = TGS_BIN_SCATTER_TILES;
    if (scatter_tiles > (whole_image ? 4096 : 2048)) scatter_tiles = whole_image ? 4096 : 2048;
    if (env_tiles) scatter_tiles = env_tiles;
    p.scatter = make_bands(scatter_tiles, Tx, Ty);
    return p;
}

// the part of a Gaussian's tile rectangle inside tile rows [r0, r1)
struct RectClip { uint32_t x0, w, y0, h; };
__device__ __forceinline__ RectClip clip_rect(uint2 rc, int r0, int r1) {
    RectClip c;
    c.x0 = rc.x & 0xFFFF; c.w = (rc.x >> 16) - c.x0;
    int y0 = (int)(rc.y & 0xFFFF), y1 = (int)(rc.y >> 16);
    y0 = max(y0, r0); y1 = min(y1, r1);
    c.y0 = (uint32_t)y0; c.h = y1 > y0 ? (uint32_t)(y1 - y0) : 0u;
    return c;
}
// l / w for l < 2^32 with a precomputed m = floor(2^32 / w): the estimate is at most 1 too small
__device__ __forceinline__ uint32_t div_by(uint32_t l, uint32_t w, uint32_t m, uint32_t& rem) {
    uint32_t q = w == 1 ? l : __umulhi(l, m);
    rem = l - q * w;
    if (rem >= w) { ++q; rem -= w; }
    return q;
}
__device__ __forceinline__ uint32_t magic_of(uint32_t w) { return w > 1 ? (uint32_t)(0x100000000ull / w) : 0u; }

// ---- 1. per-chunk tile counts.  grid (nchunks, nbands), 256 threads.
// A rectangle adds 1 to every tile it covers; instead of one shared-memory atomic per covered tile (I of them) each
// rectangle posts FOUR signed corner updates into a (rows+1) x (Tx+1) difference array, and a 2-D prefix sum (along
// x by warp scans, along y by one thread per column, fused with the write of the chunk's row of the count matrix)
// turns them into coverage counts: 4 atomics per Gaussian instead of ~11 (c3) / ~33 (c5).
__global__ void __launch_bounds__(256)
k_bin_count(int N, int chunk, const uint32_t* __restrict__ order, const uint32_t* __restrict__ keys_sorted,
            const uint32_t* __restrict__ tiles, const uint2* __restrict__ rect, int Tx, int row_begin, int row_end,
            int rows_per_band, int T, uint32_t* __restrict__ cnt, uint2* __restrict__ span_sorted,
            uint32_t* __restrict__ live_chunks) {
    extern __shared__ int diff[];                      // [(rows+1)][Tx+1]
    const int band = blockIdx.y;
    {   // the keys are sorted: a chunk whose first key is "emits nothing" is dead, and so is every chunk behind it
        const int head = blockIdx.x * chunk;
        const bool live = keys_sorted[head] != kFull;
        if (band == 0 && threadIdx.x == 0) {
            // exactly one chunk is the last live one (or chunk 0 is dead: nothing is rendered at all)
            if (live && (head + chunk >= N || keys_sorted[head + chunk] == kFull)) *live_chunks = blockIdx.x + 1;
            if (!live && blockIdx.x == 0) *live_chunks = 0;
        }
        if (!live) return;                             // its row of the count matrix is never read
    }
    const int r0 = row_begin + band * rows_per_band, r1 = min(row_end, r0 + rows_per_band);
    const int rows = r1 - r0, S = Tx + 1;
    for (int t = threadIdx.x; t < (rows + 1) * S; t += 256) diff[t] = 0;
    __syncthreads();
    const int first = blockIdx.x * chunk, last = min(N, first + chunk);
    for (int r = first + threadIdx.x; r < last; r += 256) {
        const uint32_t id = order[r];
        uint2 rc = make_uint2(0u, 0u);
        if (tiles[id]) {
            rc = rect[id];
            const RectClip c = clip_rect(rc, r0, r1);
            if (c.w && c.h) {
                const int y0 = (int)c.y0 - r0, x0 = (int)c.x0;
                atomicAdd(&diff[y0 * S + x0], 1);
                atomicAdd(&diff[y0 * S + x0 + (int)c.w], -1);
                atomicAdd(&diff[(y0 + (int)c.h) * S + x0], -1);
                atomicAdd(&diff[(y0 + (int)c.h) * S + x0 + (int)c.w], 1);
            }
        }
        // the rectangles in DEPTH ORDER (empty = (0,0)): the scatter walks them with coalesced loads
        if (band == 0) span_sorted[r] = rc;
    }
    __syncthreads();
    {   // prefix along x: one warp per row
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int y = warp; y < rows; y += 8) {
            int carry = 0;
            for (int c0 = 0; c0 < Tx; c0 += 32) {
                const int x = c0 + lane;
                int v = x < Tx ? diff[y * S + x] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(kFull, v, o);
                    if (lane >= o) v += n;
                }
                v += carry;
                if (x < Tx) diff[y * S + x] = v;
                carry = __shfl_sync(kFull, v, 31);
            }
        }
    }
    __syncthreads();
    // prefix along y, one thread per column, written straight to the chunk's row of the count matrix
    uint32_t* row = cnt + (size_t)blockIdx.x * T + (size_t)r0 * Tx;
    for (int x = threadIdx.x; x < Tx; x += 256) {
        int run = 0;
        for (int y = 0; y < rows; ++y) {
            run += diff[y * S + x];
            row[(size_t)y * Tx + x] = (uint32_t)run;
        }
    }
}

// ---- 2a. per tile column: exclusive prefix over the chunks (in place) and the column total.
// CTA = 32 tile columns x 8 chunk slices (a warp reads one 128-byte row segment per chunk).
__global__ void __launch_bounds__(256)
k_bin_prefix(const uint32_t* __restrict__ live_chunks, int T, int t_begin, int t_end, uint32_t* __restrict__ cnt,
             uint32_t* __restrict__ totals) {
    __shared__ uint32_t part[8][32];
    const int nchunks = (int)*live_chunks;             // written by k_bin_count
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int t = t_begin + blockIdx.x * 32 + lane;
    const int per = (nchunks + 7) / 8;
    const int c0 = slice * per, c1 = min(nchunks, c0 + per);
    uint32_t s = 0;
    if (t < t_end) {
        const uint32_t* p = cnt + (size_t)c0 * T + t;
#pragma unroll 8
        for (int c = c0; c < c1; ++c, p += T) s += *p;
    }
    part[slice][lane] = s;
    __syncthreads();
    uint32_t run = 0;
    for (int k = 0; k < slice; ++k) run += part[k][lane];
    if (t < t_end) {
        if (slice == 7) totals[t] = run + s;
        uint32_t* p = cnt + (size_t)c0 * T + t;
#pragma unroll 4
        for (int c = c0; c < c1; ++c, p += T) { const uint32_t v = *p; *p = run; run += v; }
    }
}

// ---- 2b. tile starts: exclusive scan of the column totals -> ranges, and the instance count.  One CTA.
// count_out[0] = num_rendered (low 32 bits), count_out[1] = 1 if it does not fit 32 bits.
__global__ void __launch_bounds__(1024)
k_bin_ranges(int T, int t_begin, int t_end, const uint32_t* __restrict__ totals, uint2* __restrict__ ranges,
             uint32_t* __restrict__ count_out) {
    __shared__ unsigned long long warp_excl[32];
    __shared__ unsigned long long block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;                      // running total of the tiles in front (same in every thread)
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        const unsigned long long v = (t >= t_begin && t < t_end) ? totals[t] : 0u;   // nothing outside the rendered rows
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long n = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_excl[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = warp_excl[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long n = __shfl_up_sync(kFull, wi, o);
                if (lane >= o) wi += n;
            }
            warp_excl[lane] = wi - w;
            if (lane == 31) block_total = wi;
        }
        __syncthreads();
        const unsigned long long start = carry + warp_excl[warp] + (inc - v);
        if (t < T) {
            const unsigned long long end = start + v;
            ranges[t] = v == 0 ? make_uint2(0u, 0u)          // untouched tiles = (0,0)
                               : make_uint2((uint32_t)(start > 0xFFFFFFFFull ? 0xFFFFFFFFull : start),
                                            (uint32_t)(end > 0xFFFFFFFFull ? 0xFFFFFFFFull : end));
        }
        carry += block_total;
        __syncthreads();                               // warp_excl / block_total are rewritten by the next round
    }
    if (threadIdx.x == 0) {
        count_out[0] = (uint32_t)(carry > 0xFFFFFFFFull ? 0xFFFFFFFFull : carry);
        count_out[1] = carry > 0xFFFFFFFFull ? 1u : 0u;
    }
}

// ---- 3. ordered scatter.  One CTA of kSub warps per (chunk, band); dynamic shared memory = kSub cursor arrays of the
// band's tiles + the chunk's hit lists.
// The walk that assigns positions is sequential by construction (order inside a tile = order of the walk), and
// depth-consecutive Gaussians cluster on screen, so a few (chunk, band) units hold several times the average number
// of hits: a single walker per unit leaves the kernel waiting for its longest chain.  The chunk is therefore cut into
// kSub sub-chunks of consecutive ranks, one per warp, with the SAME counting idea one level down:
//   phase 1 (throughput): each warp finds the Gaussians of its sub-chunk that touch the band (coalesced loads of the
//            depth-ordered rectangles, ordered compaction by ballot) and COUNTS their coverage per tile into its own
//            array (order-free: one lane per hit, shared-memory atomics);
//   prefix:  per tile, exclusive prefix over the warps + the tile's start + the chunks in front (count matrix row)
//            turns the count arrays into the warps' private cursors;
//   phase 2 (latency): each warp walks ITS hits in order, 32 staged at a time (one broadcast LDS.128 per Gaussian);
//            the tiles of one rectangle are distinct, so the lanes bump the cursors with plain LDS / STS; hits
//            (2j, 2j+1) with <= 16 tiles each and disjoint rectangles are bumped together, one per half warp.
struct __align__(16) Staged { uint32_t base, w, magic, area; };   // base = (y0 - r0) * Tx + x0
// kSub = warps (sub-chunks) per (chunk, band) unit.  8 when a rank renders a band of the image (tile-row shard: few,
// skewed units -> the longest chain decides), 1 for a whole image (many units: throughput decides, and a lone warp
// needs no counting pass: its cursors start at the tile starts + the chunks in front).  Measured at c3:
// whole image 0.155 ms (kSub 1) vs 0.19-0.34 (kSub 8); an eighth of the image 0.035 ms (kSub 8) vs 0.09 (kSub 1).
template <int kSub>
__global__ void __launch_bounds__(32 * kSub)
k_bin_scatter(int N, int chunk, int nbands, const uint32_t* __restrict__ order, const uint2* __restrict__ span_sorted,
              int Tx, int row_begin, int row_end, int rows_per_band, int T, const uint32_t* __restrict__ cnt,
              const uint2* __restrict__ ranges, const uint32_t* __restrict__ live_chunks, uint32_t cap,
              uint32_t* __restrict__ vals) {
    extern __shared__ __align__(16) uint32_t smem_u32[];
    __shared__ Staged stage_all[kSub][32];
    __shared__ uint32_t stage_id_all[kSub][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk_id = blockIdx.x / nbands, band = blockIdx.x - chunk_id * nbands;
    if (chunk_id >= (int)*live_chunks) return;          // nothing behind the last emitter of the depth order
    Staged* stage = stage_all[warp];
    uint32_t* stage_id = stage_id_all[warp];
    const int band_tiles = rows_per_band * Tx;
    const int sub = chunk / kSub;                       // ranks per warp (chunk is a multiple of 32 * kSub)
    uint32_t* cursor = smem_u32 + (size_t)warp * band_tiles;                                        // [kSub][band tiles]
    uint16_t* hits = reinterpret_cast<uint16_t*>(smem_u32 + (size_t)kSub * band_tiles) + (size_t)warp * sub;   // [kSub][sub]
    const int r0 = row_begin + band * rows_per_band, r1 = min(row_end, r0 + rows_per_band);
    const int nt = (r1 - r0) * Tx;
    if (kSub > 1) {
        for (int t = lane; t < nt; t += 32) cursor[t] = 0;
        __syncwarp();
    }
    const int first = chunk_id * chunk + warp * sub, last = min(N, first + sub);
    // ---- phase 1: hits of this warp's sub-chunk + their coverage counts
    int nh = 0;
    for (int base = first; base < last; base += 128) {
        uint2 sp[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                  // four independent coalesced loads in flight
            const int r = base + 32 * k + lane;
            sp[k] = r < last ? span_sorted[r] : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const RectClip c = clip_rect(sp[k], r0, r1);
            const bool hit = c.w != 0 && c.h != 0;
            const unsigned m = __ballot_sync(kFull, hit);
            const uint32_t area = c.w * c.h;
            const uint32_t t0 = (c.y0 - (uint32_t)r0) * (uint32_t)Tx + c.x0;
            if (hit) hits[nh + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(base - first + 32 * k + lane);
            nh += __popc(m);
            if (kSub == 1) continue;                       // a lone walker needs no counts
            if (hit) {
                if (area <= 64) {                          // small rectangles: the owning lane counts its own tiles
                    uint32_t t = t0;
                    for (uint32_t y = 0; y < c.h; ++y, t += Tx)
                        for (uint32_t x = 0; x < c.w; ++x) atomicAdd(&cursor[t + x], 1u);
                }
            }
            unsigned big = __ballot_sync(kFull, hit && area > 64);
            while (big) {                                  // large rectangles: the whole warp counts them together
                const int j = __ffs(big) - 1;
                big &= big - 1;
                const uint32_t bt = __shfl_sync(kFull, t0, j), bw = __shfl_sync(kFull, c.w, j), ba = __shfl_sync(kFull, area, j);
                const uint32_t bm = magic_of(bw);
                for (uint32_t l = lane; l < ba; l += 32) {
                    uint32_t xx; const uint32_t yy = div_by(l, bw, bm, xx);
                    atomicAdd(&cursor[bt + yy * (uint32_t)Tx + xx], 1u);
                }
            }
        }
    }
    if (kSub > 1) __syncthreads(); else __syncwarp();
    {   // ---- prefix over the warps: cursor_w[t] = start of tile t + chunks in front + sub-chunks in front
        const uint32_t* row = cnt + (size_t)chunk_id * T + (size_t)r0 * Tx;
        const uint2* rg = ranges + (size_t)r0 * Tx;
#pragma unroll 4
        for (int t = threadIdx.x; t < nt; t += 32 * kSub) {
            uint32_t run = rg[t].x + row[t];
#pragma unroll
            for (int w = 0; w < kSub; ++w) {
                uint32_t* p = smem_u32 + (size_t)w * band_tiles + t;
                const uint32_t v = kSub > 1 ? *p : 0u;
                *p = run;
                run += v;
            }
        }
    }
    if (kSub > 1) __syncthreads(); else __syncwarp();
    // ---- phase 2: walk this warp's hits IN ORDER
    for (int hb = 0; hb < nh; hb += 32) {
        const int e = hb + lane;
        const int n = min(32, nh - hb);
        RectClip c; c.w = c.h = c.x0 = c.y0 = 0;
        uint32_t id = 0;
        if (e < nh) {
            const int r = first + (int)hits[e];
            c = clip_rect(span_sorted[r], r0, r1);
            id = order[r];
        }
        const uint32_t area = c.w * c.h;
        {   // can this hit share a step with its pair partner (lane ^ 1)?
            const uint32_t px0 = __shfl_xor_sync(kFull, c.x0, 1), pw = __shfl_xor_sync(kFull, c.w, 1);
            const uint32_t py0 = __shfl_xor_sync(kFull, c.y0, 1), ph = __shfl_xor_sync(kFull, c.h, 1);
            const bool disjoint = (c.x0 + c.w <= px0) || (px0 + pw <= c.x0) || (c.y0 + c.h <= py0) || (py0 + ph <= c.y0);
            const bool fuse = disjoint && area <= 16 && pw * ph <= 16 && area != 0 && pw * ph != 0;
            Staged sg;
            sg.base = (c.y0 - (uint32_t)r0) * (uint32_t)Tx + c.x0; sg.w = c.w; sg.magic = magic_of(c.w);
            sg.area = area | (fuse ? 0x80000000u : 0u);
            stage[lane] = sg;
            stage_id[lane] = id;
        }
        __syncwarp();
        for (int k = 0; k < n; k += 2) {
            const uint32_t fused = stage[k].area >> 31;               // uniform: both partners carry the same flag
            if (fused) {
                const int half = lane >> 4;
                const Staged q = stage[k + half];
                const uint32_t gid = stage_id[k + half];
                const uint32_t l = (uint32_t)(lane & 15);
                if (l < (q.area & 0x7FFFFFFFu)) {
                    uint32_t xx; const uint32_t yy = div_by(l, q.w, q.magic, xx);
                    const uint32_t t = q.base + yy * (uint32_t)Tx + xx;
                    const uint32_t pos = cursor[t];
                    cursor[t] = pos + 1;
                    if (pos < cap) vals[pos] = gid;                  // speculative mode: never write past the hint
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (k + h < n) {
                        const Staged q = stage[k + h];
                        const uint32_t gid = stage_id[k + h];
                        const uint32_t a = q.area & 0x7FFFFFFFu;
                        for (uint32_t l = lane; l < a; l += 32) {    // one round for rectangles of up to 32 tiles
                            uint32_t xx; const uint32_t yy = div_by(l, q.w, q.magic, xx);
                            const uint32_t t = q.base + yy * (uint32_t)Tx + xx;
                            const uint32_t pos = cursor[t];
                            cursor[t] = pos + 1;
                            if (pos < cap) vals[pos] = gid;
                        }
                        __syncwarp();                                // the next Gaussian's bumps come after this one's
                    }
                }
            }
        }
    }
}

}  // namespace

size_t tgs_depth_sort_temp_bytes(int N) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, N, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, N);   // refstructure arm
    return a > b ? a : b;
}

size_t tgs_bin_temp_bytes(int N, int Tx, int Ty) {
    const Plan p = make_plan(N, Tx, Ty);               // nchunks does not depend on the rendered rows
    const size_t T = (size_t)Tx * Ty;
    return tgs_align_up((size_t)p.nchunks * T * sizeof(uint32_t)) + tgs_align_up(T * sizeof(uint32_t)) +
           tgs_align_up(sizeof(uint32_t));                // count matrix, column totals, number of live chunks
}
static uint32_t* live_chunks_of(const void* temp, int nchunks, int T) {
    return (uint32_t*)((char*)temp + tgs_align_up((size_t)nchunks * T * sizeof(uint32_t)) + tgs_align_up((size_t)T * sizeof(uint32_t)));
}

// Phase 0: depth order of the Gaussians.
int tgs_depth_order(GeomView gv, int N, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_SCAN, st);
    size_t bytes = gv.temp_bytes;
    TGS_CUDA(cub::DeviceRadixSort::SortPairs(gv.temp, bytes, gv.depth_keys, gv.depth_keys_sorted, gv.ids, gv.order, N,
                                             0, 32, st));
    tgs_count_cub(1);
    return 0;
}

// Phases 1-2: count matrix, per-tile prefixes, tile ranges and the instance count (on the device: count_out[0..1]).
int tgs_bin_count(GeomView gv, int N, int Tx, int Ty, int row0, int row1, void* temp, uint2* ranges, uint32_t* count_out,
                  cudaStream_t st) {
    const int T = Tx * Ty;
    if (N == 0 || row1 <= row0) {
        TGS_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)T, st));
        TGS_CUDA(cudaMemsetAsync(count_out, 0, 2 * sizeof(uint32_t), st));
        return 0;
    }
    const Plan p = make_plan(N, Tx, row1 - row0, row0 == 0 && row1 == Ty);      // bands over the rows this rank renders
    uint32_t* cnt = (uint32_t*)temp;
    uint32_t* totals = (uint32_t*)((char*)temp + tgs_align_up((size_t)p.nchunks * T * sizeof(uint32_t)));
    uint32_t* live = live_chunks_of(temp, p.nchunks, T);
    const int t_begin = row0 * Tx, t_end = row1 * Tx;
    static bool attr_done[64] = {};
    int dev = 0;
    TGS_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        TGS_CUDA(cudaFuncSetAttribute(k_bin_count, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (TGS_BIN_BAND_TILES + 1) * 4));
        // one tile row of the widest image per warp + the hit lists
        TGS_CUDA(cudaFuncSetAttribute(k_bin_scatter<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TGS_BIN_SCATTER_MAX_TX * 4 + 8192 * 2));
        TGS_CUDA(cudaFuncSetAttribute(k_bin_scatter<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (TGS_BIN_SCATTER_MAX_TX * 4) + 8192 * 2));
        attr_done[dev] = true;
    }
    {
        TgsProfScope prof(TGS_STAGE_DUPLICATE, st);
        k_bin_count<<<dim3(p.nchunks, p.count.n), 256, (size_t)p.count.tiles * 4, st>>>(
            N, p.chunk, gv.order, gv.depth_keys_sorted, gv.tiles_touched, gv.rect, Tx, row0, row1, p.count.rows, T, cnt,
            gv.span_sorted, live);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    }
    {
        TgsProfScope prof(TGS_STAGE_SORT, st);
        k_bin_prefix<<<(t_end - t_begin + 31) / 32, 256, 0, st>>>(live, T, t_begin, t_end, cnt, totals);
        k_bin_ranges<<<1, 1024, 0, st>>>(T, t_begin, t_end, totals, ranges, count_out);
        tgs_count_own(2);
        TGS_CUDA(cudaGetLastError());
    }
    return 0;
}

// Phase 3.  `cap` = instances the binning buffer holds (speculative mode: positions beyond it are not written).
int tgs_bin_scatter(GeomView gv, BinView bv, int N, int64_t count, int64_t cap, bool speculative, int Tx, int Ty,
                         int row0, int row1, const void* temp, const uint2* ranges, const uint32_t* count_dev, cudaStream_t st) {
    if (count == 0 || N == 0 || row1 <= row0) return 0;
    const int T = Tx * Ty;
    const Plan p = make_plan(N, Tx, row1 - row0, row0 == 0 && row1 == Ty);
    const uint32_t* cnt = (const uint32_t*)temp;
    {
        TgsProfScope prof(TGS_STAGE_BIN_SCATTER, st);
        const bool whole = (row0 == 0 && row1 == Ty);
        const int ksub = whole ? 1 : 8;
        const size_t smem = ((size_t)ksub * p.scatter.tiles + (size_t)(p.chunk + 1) / 2) * 4;
        auto kern = whole ? k_bin_scatter<1> : k_bin_scatter<8>;
        kern<<<p.nchunks * p.scatter.n, 32 * ksub, smem, st>>>(
            N, p.chunk, p.scatter.n, gv.order, gv.span_sorted, Tx, row0, row1, p.scatter.rows, T, cnt, ranges,
            live_chunks_of(temp, p.nchunks, T), (uint32_t)(cap > 0xFFFFFFFFll ? 0xFFFFFFFFll : cap), bv.vals_sorted);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    }
    return 0;
}
