// binning.cu -- tile binning of the visible Gaussians.
//
// The spec'd RESULT is the list of (tile, Gaussian) instances sorted by the 64-bit key
// (tile << 32 | bits(depth)), ties in ascending Gaussian id (stable sort of a Gaussian-major emission),
// plus one [start,end) range per tile.  We produce exactly that list (bit-exact against
// oracle/gs_oracle.py::bin_and_sort) with a cheaper TWO-PHASE sort:
//   1. radix-sort the N Gaussians once by depth bits (32-bit keys, N items; stable => equal depths keep
//      ascending id; invisible Gaussians get key 0xFFFFFFFF and emit nothing);
//   2. scan tiles_touched in that depth order, emit the instances depth-major (warp-cooperative,
//      coalesced) with the TILE ID as the only key (16 bits when T <= 65536);
//   3. stable radix sort of I (tile, id) pairs over ceil(log2 T) bits: 2 onesweep passes on 6-byte pairs
//      instead of 6 passes on 12-byte pairs.  Stability carries the depth order into every tile.
//   4. fused tile-range detection + packing of the sorted 48-byte records (gather -> shared memory ->
//      one TMA bulk store per 256 records, so the packed array is written with full-line stores).
// The scans / sorts are CUB device primitives (library code, like cuBLAS for a plain GEMM).
// Roofline: HBM.  Algorithmic bytes per instance: key/val write 12, sort 24, range detect 8,
// record gather 48 + write 48 (we move 6-byte pairs, so real traffic is below the algorithmic figure).
#include "tgs_common.cuh"
#include <cub/cub.cuh>

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct TilesInOrder {
    const uint32_t* tiles; const uint32_t* order;
    __device__ __forceinline__ uint32_t operator()(uint32_t r) const { return tiles[order[r]]; }
};
using TilesIt = cub::TransformInputIterator<uint32_t, TilesInOrder, cub::CountingInputIterator<uint32_t>>;

// Emission in depth order.  One warp per 32 consecutive depth ranks.  The scan makes the instances of those 32
// Gaussians ONE contiguous output range [B, E), so the warp walks that range 32 slots at a time -- every store is a
// full coalesced line -- and each lane finds the Gaussian that owns its slot by a 5-step binary search over the 32
// inclusive offsets held one per lane (shuffles), instead of the warp serialising over its Gaussians with a third of
// the lanes busy.
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_emit(int N, const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles,
       const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect, int Tx, uint32_t cap,
       KeyT* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;          // depth rank
    const int r_first = r - lane;
    if (r_first >= N) return;
    const int r_last = min(r_first + 31, N - 1);
    // inclusive offset of this lane's rank (lanes past N repeat the last one: they own nothing)
    const uint32_t end = offsets[min(r, r_last)];
    uint32_t id = 0, cnt = 0;
    uint2 rc = make_uint2(0, 0);
    if (r < N) {
        id = order[r];
        cnt = tiles[id];
        if (cnt) rc = rect[id];
    }
    const uint32_t base = end - cnt;                               // exclusive offset of this lane's Gaussian
    const uint32_t B = __shfl_sync(kFull, base, 0);
    const uint32_t E = __shfl_sync(kFull, end, 31);
    for (uint32_t p0 = B; p0 < E; p0 += 32) {
        const uint32_t p = p0 + lane;
        // owner = number of lanes whose inclusive offset is <= p (offsets are monotone)
        int own = 0;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const uint32_t e = __shfl_sync(kFull, end, own + s - 1);
            if (e <= p) own += s;
        }
        own = min(own, 31);
        const uint32_t gbase = __shfl_sync(kFull, base, own);
        const uint32_t gid = __shfl_sync(kFull, id, own);
        const uint32_t rx = __shfl_sync(kFull, rc.x, own), ry = __shfl_sync(kFull, rc.y, own);
        if (p < E && p < cap) {                                    // speculative mode: never write past the hint
            const uint32_t x0 = rx & 0xFFFF, w = (rx >> 16) - x0, y0 = ry & 0xFFFF;
            const uint32_t t = p - gbase;
            const uint32_t yy = t / w, xx = t - yy * w;            // row-major: y outer, x inner
            keys[p] = (KeyT)((y0 + yy) * Tx + x0 + xx);
            vals[p] = gid;
        }
    }
}

// Fused A4 + record packing.  256 instances per CTA: each thread gathers its instance's 48-byte record
// (3 x LDG.128, mostly L2 hits: the per-Gaussian record array is 48 MB at 1M splats) into shared memory,
// performs the tile-boundary test of identifyTileRanges, and one thread writes the 12 KB block back with
// a single TMA bulk store.
template <typename KeyT>
__global__ void __launch_bounds__(256)
k_pack_ranges(int64_t I_host, const uint32_t* __restrict__ I_dev, int64_t cap, const KeyT* __restrict__ keys,
              const uint32_t* __restrict__ vals, const float4* __restrict__ rec_in, float4* __restrict__ rec_out,
              uint2* __restrict__ ranges) {
    __shared__ __align__(128) float4 sm[256 * 3];
    // exact mode: I_host; speculative mode: the real count lives on the device (last scan element), clamped
    // to the buffer capacity (an overflowing speculation is discarded and re-run by the host)
    int64_t I = I_host;
    if (I_dev) { I = (int64_t)*I_dev; if (I > cap) I = cap; }
    const int64_t j0 = (int64_t)blockIdx.x * 256;
    if (j0 >= I) return;
    const int64_t j = j0 + threadIdx.x;
    if (j < I) {
        const uint32_t id = vals[j];
        const float4* src = rec_in + (size_t)3 * id;
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        sm[3 * threadIdx.x] = a; sm[3 * threadIdx.x + 1] = b; sm[3 * threadIdx.x + 2] = c;
        const uint32_t tile = (uint32_t)keys[j];
        if (j == 0) ranges[tile].x = 0;
        else {
            const uint32_t prev = (uint32_t)keys[j - 1];
            if (prev != tile) { ranges[prev].y = (uint32_t)j; ranges[tile].x = (uint32_t)j; }
        }
        if (j == I - 1) ranges[tile].y = (uint32_t)I;
    }
    // make the generic-proxy shared-memory writes visible to the async proxy, then bulk-store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const int64_t n = (I - j0) < 256 ? (I - j0) : 256;
        const uint32_t bytes = (uint32_t)n * 48u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rec_out + 3 * j0),
                     "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released after the read
    }
}

template <typename KeyT>
int emit_sort_pack(GeomView gv, BinView bv, int N, int64_t I, int64_t cap, bool spec, int T, int Tx, int tile_bits,
                   cudaStream_t st) {
    KeyT* ku = reinterpret_cast<KeyT*>(bv.tile_unsorted);
    KeyT* ks = reinterpret_cast<KeyT*>(bv.tile_sorted);
    {
        TgsProfScope prof(TGS_STAGE_DUPLICATE, st);
        // speculative mode sorts `cap` slots: the unused tail must sort behind every real tile id
        if (spec) TGS_CUDA(cudaMemsetAsync(ku, 0xFF, sizeof(KeyT) * (size_t)cap, st));
        k_emit<KeyT><<<(N + 255) / 256, 256, 0, st>>>(N, gv.order, gv.tiles_touched, gv.offsets, gv.rect, Tx,
                                                      (uint32_t)(cap > 0xFFFFFFFFll ? 0xFFFFFFFFll : cap), ku,
                                                      bv.vals_unsorted);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    }
    {
        TgsProfScope prof(TGS_STAGE_SORT, st);
        size_t bytes = bv.cub_temp_bytes;
        TGS_CUDA(cub::DeviceRadixSort::SortPairs(bv.cub_temp, bytes, ku, ks, bv.vals_unsorted, bv.vals_sorted, I, 0,
                                                 tile_bits, st));
        tgs_count_cub(1);
    }
    {
        TgsProfScope prof(TGS_STAGE_PACK, st);
        k_pack_ranges<KeyT><<<(unsigned)((I + 255) / 256), 256, 0, st>>>(
            I, spec ? gv.offsets + (N - 1) : nullptr, cap, ks, bv.vals_sorted,
            reinterpret_cast<const float4*>(gv.records), reinterpret_cast<float4*>(bv.records), bv.ranges);
        tgs_count_own(1);
        TGS_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace

size_t tgs_depth_sort_temp_bytes(int N) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, N, 0, 32);
    TilesInOrder f{nullptr, nullptr};
    TilesIt it(cub::CountingInputIterator<uint32_t>(0), f);
    cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, N);
    return a > b ? a : b;
}

size_t tgs_tile_sort_temp_bytes(int64_t I, int T) {
    size_t bytes = 0;
    if (T < 65535)
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, I, 0, 16);
    else
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, I, 0, 32);
    return bytes;
}

// Phase 1 + scan: depth order of the Gaussians and the inclusive scan of tiles_touched in that order.
int tgs_depth_order_and_scan(GeomView gv, int N, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_SCAN, st);
    size_t bytes = gv.temp_bytes;
    TGS_CUDA(cub::DeviceRadixSort::SortPairs(gv.temp, bytes, gv.depth_keys, gv.depth_keys_sorted, gv.ids, gv.order, N,
                                             0, 32, st));
    TilesInOrder f{gv.tiles_touched, gv.order};
    TilesIt it(cub::CountingInputIterator<uint32_t>(0), f);
    bytes = gv.temp_bytes;
    TGS_CUDA(cub::DeviceScan::InclusiveSum(gv.temp, bytes, it, gv.offsets, N, st));
    tgs_count_cub(2);
    return 0;
}

int tgs_emit_sort_pack(GeomView gv, BinView bv, int N, int64_t count, int64_t cap, bool speculative, int T, int Tx,
                       cudaStream_t st) {
    TGS_CUDA(cudaMemsetAsync(bv.ranges, 0, sizeof(uint2) * (size_t)T, st));
    if (count == 0 || N == 0) return 0;
    int bits = 1;
    while ((1 << bits) < T) ++bits;
    if (speculative && (1 << bits) == T) ++bits;     // the all-ones pad key must compare above every tile id
    if (T < 65535) return emit_sort_pack<uint16_t>(gv, bv, N, count, cap, speculative, T, Tx, bits, st);
    return emit_sort_pack<uint32_t>(gv, bv, N, count, cap, speculative, T, Tx, bits > 32 ? 32 : bits, st);
}
