// touch_inputs.cu -- the per-pixel part of the reference's touch / vision depth fusion on the GPU
// (the data formats feeding the hot path's touch target and weight).
//
// One elementwise kernel replaces, per image, reference utils/fuse_touch_vision.py:270-276 (uint16 mm
// decode), :288-306 (apply the fitted alignment), :310-313 + utils/create_uncertainty_from_depth.py:21
// (vision sigma heuristic), :76-202 (inverse-variance fusion), :360-361 (clips), :373-376 (uint16 encode)
// and the trainer-side decode (reference legacy/dataparser_tactile.py:65-66) into the fp32 target /
// weight tensors the rasterizer consumes.  The two L-BFGS-B fits (:285,:301) stay on the CPU; their
// results enter as scalars.
//
// All arithmetic is FLOAT64 in the reference's operation order and this translation unit is compiled
// with --fmad=false, so the uint16 outputs are BIT-IDENTICAL to the PNGs the reference writes
// (tests/golden/fusion_reference.npz is produced by the reference's own code).
// Roofline: HBM.  Algorithmic bytes per pixel: 3 x 2 read + 4 x 2 + 2 x 4 written = 22.
#include "tgs_common.cuh"
#include "touch_inputs_math.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_fuse_touch_vision(int64_t n, const unsigned short* __restrict__ touch, const unsigned short* __restrict__ vision,
                    const unsigned short* __restrict__ tsig, FuseParams p, unsigned short* __restrict__ out_va,
                    unsigned short* __restrict__ out_ds, unsigned short* __restrict__ out_fu,
                    unsigned short* __restrict__ out_fs, float* __restrict__ target, float* __restrict__ weight) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const ushort4 t4 = reinterpret_cast<const ushort4*>(touch)[i];
        const ushort4 v4 = reinterpret_cast<const ushort4*>(vision)[i];
        const ushort4 s4 = reinterpret_cast<const ushort4*>(tsig)[i];
        const PixelOut a = fuse_pixel(t4.x, v4.x, s4.x, p), b = fuse_pixel(t4.y, v4.y, s4.y, p);
        const PixelOut c = fuse_pixel(t4.z, v4.z, s4.z, p), d = fuse_pixel(t4.w, v4.w, s4.w, p);
        if (out_va) reinterpret_cast<ushort4*>(out_va)[i] = make_ushort4(a.va, b.va, c.va, d.va);
        if (out_ds) reinterpret_cast<ushort4*>(out_ds)[i] = make_ushort4(a.ds, b.ds, c.ds, d.ds);
        if (out_fu) reinterpret_cast<ushort4*>(out_fu)[i] = make_ushort4(a.fu, b.fu, c.fu, d.fu);
        if (out_fs) reinterpret_cast<ushort4*>(out_fs)[i] = make_ushort4(a.fs, b.fs, c.fs, d.fs);
        if (target) reinterpret_cast<float4*>(target)[i] = make_float4(a.target, b.target, c.target, d.target);
        if (weight) reinterpret_cast<float4*>(weight)[i] = make_float4(a.weight, b.weight, c.weight, d.weight);
    }
    // tail (n not a multiple of 4)
    const int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const PixelOut a = fuse_pixel(touch[i], vision[i], tsig[i], p);
        if (out_va) out_va[i] = a.va;
        if (out_ds) out_ds[i] = a.ds;
        if (out_fu) out_fu[i] = a.fu;
        if (out_fs) out_fs[i] = a.fs;
        if (target) target[i] = a.target;
        if (weight) weight[i] = a.weight;
    }
}

// Trainer-side decode of the on-disk maps (reference legacy/dataparser_tactile.py:65-66: depth_unit_scale_factor 1e-3;
// :229-235: the pose scale factor also scales the depths): target = mm * unit, weight = f(sigma) with sigma = mm * 1e-3
// (the uncertainty PNG is sigma x 1000, reference utils/fuse_touch_vision.py:376,387; sigma is NOT scaled with the scene).
//   weight_mode 0: 1            (SIMPLE_LOSS)
//   weight_mode 1: 1 / (uw * sigma)       weight_mode 2: 1 / (uw * sigma)^2       (0 where sigma == 0)
__global__ void __launch_bounds__(256)
k_decode_touch_maps(int64_t n, const unsigned short* __restrict__ depth_mm, const unsigned short* __restrict__ sigma_mm,
                    float unit, float uw, int mode, float* __restrict__ target, float* __restrict__ weight) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (target) target[i] = (float)depth_mm[i] * unit;
        if (weight) {
            float w = 1.0f;
            if (mode != 0 && sigma_mm) {
                const float sg = (float)sigma_mm[i] * 1e-3f * uw;
                w = sg > 0.0f ? (mode == 1 ? 1.0f / sg : 1.0f / (sg * sg)) : 0.0f;
            }
            weight[i] = w;
        }
    }
}

}  // namespace

extern "C" int tgs_decode_touch_maps(const uint16_t* depth_mm, const uint16_t* sigma_mm, int64_t num_pixels,
                                     float depth_unit, float uncertainty_weight, int32_t weight_mode, float* target,
                                     float* weight, void* stream) {
    if (num_pixels < 0 || weight_mode < 0 || weight_mode > 2 || (num_pixels > 0 && !depth_mm && target) ||
        (num_pixels > 0 && weight && weight_mode != 0 && !sigma_mm) || !(uncertainty_weight > 0.0f)) {
        tgs_set_error("tgs_decode_touch_maps: bad arguments"); return TGS_EINVAL; }
    if (num_pixels == 0) return 0;
    int64_t blocks = (num_pixels + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_decode_touch_maps<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(num_pixels, depth_mm, sigma_mm, depth_unit,
                                                                             uncertainty_weight, weight_mode, target, weight);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tgs_fuse_touch_vision(const uint16_t* touch_mm, const uint16_t* vision_mm, const uint16_t* touch_sigma_mm,
                                     int64_t num_pixels, double scale, double offset, double offset2,
                                     int32_t is_real_world, double scene_scale, uint16_t* vision_aligned_mm,
                                     uint16_t* ds_gs_mm, uint16_t* fused_mm, uint16_t* fused_sigma_mm, float* target,
                                     float* weight, void* stream) {
    if (num_pixels < 0 || (num_pixels > 0 && (!touch_mm || !vision_mm || !touch_sigma_mm))) {
        tgs_set_error("tgs_fuse_touch_vision: bad arguments"); return TGS_EINVAL; }
    if (num_pixels == 0) return 0;
    const uintptr_t al = (uintptr_t)touch_mm | (uintptr_t)vision_mm | (uintptr_t)touch_sigma_mm |
                         (uintptr_t)vision_aligned_mm | (uintptr_t)ds_gs_mm | (uintptr_t)fused_mm |
                         (uintptr_t)fused_sigma_mm;
    if ((al & 7) || ((uintptr_t)target & 15) || ((uintptr_t)weight & 15)) {
        tgs_set_error("tgs_fuse_touch_vision: buffers must be 8-byte (uint16) / 16-byte (float) aligned"); return TGS_EINVAL; }
    FuseParams p; p.scale = scale; p.offset = offset; p.offset2 = offset2; p.unit = 1e-3 * scene_scale;
    p.real_world = is_real_world;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t work = (num_pixels >> 2) > 0 ? (num_pixels >> 2) : 1;
    int64_t blocks = (work + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;          // grid-stride: 16 resident CTAs x 148 SMs
    if (blocks < 1) blocks = 1;
    k_fuse_touch_vision<<<(unsigned)blocks, 256, 0, st>>>(num_pixels, touch_mm, vision_mm, touch_sigma_mm, p,
                                                        vision_aligned_mm, ds_gs_mm, fused_mm, fused_sigma_mm, target,
                                                        weight);
    tgs_count_own(1);
    TGS_CUDA(cudaGetLastError());
    return 0;
}
