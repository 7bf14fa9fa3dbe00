// binning.cu -- tile binning of the visible Gaussians:
//   inclusive scan of tiles_touched  ->  duplicateWithKeys  ->  stable 64-bit radix sort
//   ->  (fused) tile-range detection + packing of the sorted per-instance records.
//
// Integer work, bit-exact against oracle/gs_oracle.py::bin_and_sort.  The scan and the sort are
// CUB device primitives (library code, like calling cuBLAS for a plain GEMM); the kernels around
// them are ours.  Roofline: HBM.  Algorithmic bytes per instance: key/val write 12,
// sort 24 (one idealised pass), range detect 8, packed-record gather 48 + write 48.
#include "tgs_common.cuh"
#include <cub/cub.cuh>

namespace {

// One warp per Gaussian would waste lanes on the many small splats; one thread per Gaussian
// serialises the few huge ones.  v1: one thread per Gaussian (A2's definition); emission order is
// Gaussian-major, then tile row, then tile column -- that order is part of the sort-stability spec.
__global__ void __launch_bounds__(256)
k_duplicate(int N, const TgsRecord* __restrict__ rec, const uint32_t* __restrict__ tiles,
            const uint32_t* __restrict__ offsets, const uint2* __restrict__ rect, int Tx,
            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (tiles[i] == 0) return;
    uint32_t off = (i == 0) ? 0u : offsets[i - 1];
    uint2 r = rect[i];
    int x0 = r.x & 0xFFFF, x1 = r.x >> 16, y0 = r.y & 0xFFFF, y1 = r.y >> 16;
    uint32_t dbits = __float_as_uint(rec[i].a.z);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            uint64_t key = ((uint64_t)(uint32_t)(y * Tx + x) << 32) | dbits;
            keys[off] = key;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

// Fused A4 + record packing.  Three threads per instance, one 16-byte quarter of the 48-byte
// record each, so the packed writes are perfectly coalesced; the thread holding quarter 0 also
// performs the tile-boundary test of identifyTileRanges.
__global__ void __launch_bounds__(256)
k_pack_ranges(int64_t I, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
              const float4* __restrict__ rec_in, float4* __restrict__ rec_out,
              uint2* __restrict__ ranges) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * I) return;
    int64_t j = t / 3;
    int part = (int)(t - 3 * j);
    uint32_t id = vals[j];
    rec_out[t] = __ldg(rec_in + (size_t)3 * id + part);
    if (part == 0) {
        uint32_t tile = (uint32_t)(keys[j] >> 32);
        if (j == 0) ranges[tile].x = 0;
        else {
            uint32_t prev = (uint32_t)(keys[j - 1] >> 32);
            if (prev != tile) { ranges[prev].y = (uint32_t)j; ranges[tile].x = (uint32_t)j; }
        }
        if (j == I - 1) ranges[tile].y = (uint32_t)I;
    }
}

}  // namespace

size_t tgs_scan_temp_bytes(int N) {
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, N);
    return bytes;
}

size_t tgs_sort_temp_bytes(int64_t I, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, I, 0, end_bit);
    return bytes;
}

int tgs_scan_tiles(GeomView gv, int N, void* temp, size_t temp_bytes, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_SCAN, st);
    TGS_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, gv.tiles_touched, gv.offsets, N, st));
    tgs_count_cub(1);
    return 0;
}

int tgs_launch_duplicate(GeomView gv, int N, int Tx, BinView bv, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_DUPLICATE, st);
    k_duplicate<<<(N + 255) / 256, 256, 0, st>>>(N, gv.records, gv.tiles_touched, gv.offsets, gv.rect, Tx,
                                                 bv.keys_unsorted, bv.vals_unsorted);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

int tgs_sort_instances(BinView bv, int64_t I, int end_bit, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_SORT, st);
    TGS_CUDA(cub::DeviceRadixSort::SortPairs(bv.cub_temp, bv.cub_temp_bytes, bv.keys_unsorted, bv.keys_sorted,
                                             bv.vals_unsorted, bv.vals_sorted, I, 0, end_bit, st));
    tgs_count_cub(1);
    return 0;
}

int tgs_launch_pack_ranges(GeomView gv, BinView bv, int64_t I, int T, cudaStream_t st) {
    TgsProfScope prof(TGS_STAGE_PACK, st);
    TGS_CUDA(cudaMemsetAsync(bv.ranges, 0, sizeof(uint2) * (size_t)T, st));
    if (I == 0) return 0;
    int64_t threads = 3 * I;
    k_pack_ranges<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
        I, bv.keys_sorted, bv.vals_sorted, reinterpret_cast<const float4*>(gv.records),
        reinterpret_cast<float4*>(bv.records), bv.ranges);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
