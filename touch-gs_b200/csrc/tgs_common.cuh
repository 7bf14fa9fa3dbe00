// tgs_common.cuh -- shared declarations of libtgs.so (layouts of the saved buffers, the 48-byte
// per-Gaussian record, error plumbing, kernel-launcher prototypes between translation units).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/tgs.h"
#include "tgs_math.cuh"

// ---------------------------------------------------------------------------------- records
// One 48-byte record per Gaussian (geometry buffer).  The compositing kernels stage a tile's depth-sorted list by
// TMA-copying its contiguous run of Gaussian ids and gathering these records by id (render.cu).
//   a = (x_pix, y_pix, depth, bits(gaussian id))
//   b = (conic A, conic B, conic C, opacity)
//   c = (r, g, b, thr)   thr = -ln(255*opacity): power threshold of the alpha >= 1/255 test
struct __align__(16) TgsRecord { float4 a, b, c; };
static_assert(sizeof(TgsRecord) == 48, "record must be 48 bytes");

#define TGS_ALIGN 256
static inline size_t tgs_align_up(size_t v) { return (v + TGS_ALIGN - 1) / TGS_ALIGN * TGS_ALIGN; }

// ------------------------------------------------------------------------------------ errors
void tgs_set_error(const char* fmt, ...);
int tgs_check_cuda(cudaError_t e, const char* what, const char* file, int line);
#define TGS_CUDA(expr)                                                           \
    do {                                                                         \
        int _rc = tgs_check_cuda((expr), #expr, __FILE__, __LINE__);             \
        if (_rc) return _rc;                                                     \
    } while (0)
// after a kernel launch: always peek the launch error; in debug mode also synchronise
#define TGS_KERNEL_CHECK(stream, debug)                                          \
    do {                                                                         \
        TGS_CUDA(cudaGetLastError());                                            \
        if (debug) TGS_CUDA(cudaStreamSynchronize(stream));                      \
    } while (0)

void tgs_count_own(int n);
// stage timers (api.cu): no-ops unless tgs_profile_enable(1)
void tgs_prof_begin(int stage, cudaStream_t st);
void tgs_prof_end(int stage, cudaStream_t st);
struct TgsProfScope {
    int stage; cudaStream_t st;
    TgsProfScope(int s, cudaStream_t t) : stage(s), st(t) { tgs_prof_begin(stage, st); }
    ~TgsProfScope() { tgs_prof_end(stage, st); }
};
void tgs_count_cub(int n);

// ---------------------------------------------------------------------- typed buffer views
struct GeomView {
    TgsRecord* records;      // [N]
    float* cov3D;            // [N,6]
    uint32_t* tiles_touched; // [N]
    uint8_t* clamped;        // [N] bit c = colour channel c clamped at 0
    uint2* rect;             // [N] (rminx | rmaxx<<16, rminy | rmaxy<<16)
    uint32_t* depth_keys;    // [N] bits(depth) or 0xFFFFFFFF
    uint32_t* ids;           // [N] iota
    uint32_t* depth_keys_sorted;
    uint32_t* order;         // [N] ids in (depth, id) order
    uint2* span_sorted;      // [N] rect of the Gaussians in that order
    void* temp; size_t temp_bytes;
};
struct BinView {
    uint32_t* vals_sorted;   // [I] Gaussian ids in final (tile, depth, id) order: every tile's list is one contiguous run
    float* ckpt;             // [slots][5][256] forward checkpoints at 256-record boundaries of the tile lists
    uint32_t* slot_tile;     // [slots] owning tile of the boundary in a slot (valid for the slots in ckpt_list)
    uint32_t* ckpt_list;     // [slots] slots the forward checkpointed, in completion order
    uint32_t* work_counter;  // [0] dynamic work-unit counter of the backward, [1] length of ckpt_list
};
#define TGS_CKPT_FLOATS (5 * 256)
struct ImageView {
    float* final_T; uint32_t* n_contrib; float* depth_raw;
    float* color_acc;        // [3][H*W] composited colour without the background term
    uint2* ranges;           // [T] per-tile [start, end) into the sorted instance list
    uint32_t* count;         // [2] num_rendered (device copy), overflow flag
};
#define TGS_BIN_BAND_TILES 8192     /* tiles per band of the count kernel: 32 KB of shared-memory counters */
#define TGS_BIN_SCATTER_MAX_TX 6144  /* widest image (in tiles) the scatter's shared memory holds: 8 cursor rows of 24 KB */
#define TGS_BIN_SCATTER_TILES 256   /* tiles per band of the ordered scatter (one warp per (chunk, band)): measured best at c3 */
GeomView tgs_geom_view(void* base, int N);
BinView tgs_bin_view(void* base, int64_t I);
ImageView tgs_image_view(void* base, int W, int H);
size_t tgs_depth_sort_temp_bytes(int N);
size_t tgs_bin_temp_bytes(int N, int Tx, int Ty);

// ------------------------------------------------------------------------ kernel launchers
// preprocess.cu
int tgs_launch_preprocess(const TgsCam& cam, const TgsSettings* s, const TgsGaussians* g,
                          GeomView gv, int32_t* radii, cudaStream_t st);
// world > 0: gather the partial screen gradients from `peer_grads[0..world)` (bands in peer_rows[2r], [2r+1]).
// `with_flags`: the buffer(s) carry contributor bytes at TGS_SCREEN_GRAD_FLAG_OFFSET(N) (TgsSettings.contrib_flags)
int tgs_launch_preprocess_bwd(const TgsCam& cam, const TgsSettings* s, const TgsGaussians* g,
                              GeomView gv, const int32_t* radii, const float* screen_grads,
                              const float* const* peer_grads, const int32_t* peer_rows, int world, bool with_flags,
                              const TgsGrads* grads, cudaStream_t st);
int tgs_launch_mark_visible(int N, const float* means, const float* vm, uint8_t* present, cudaStream_t st);
// binning.cu
int tgs_depth_order(GeomView gv, int N, cudaStream_t st);
// count matrix + per-tile prefixes + ranges + instance count (device: count_out[0] = I, count_out[1] = overflow flag)
int tgs_bin_count(GeomView gv, int N, int Tx, int Ty, int row0, int row1, void* temp, uint2* ranges, uint32_t* count_out,
                  cudaStream_t st);
// `cap` = instances the binning buffer holds; `count` = instances to place (== I in exact mode, == cap in speculative
// mode, where the real count is read on the device from count_dev)
int tgs_bin_scatter(GeomView gv, BinView bv, int N, int64_t count, int64_t cap, bool speculative, int Tx, int Ty,
                         int row0, int row1, const void* temp, const uint2* ranges, const uint32_t* count_dev, cudaStream_t st);
// render.cu
int tgs_launch_render_fwd(const TgsCam& cam, const TgsSettings* s, const TgsRecord* gv_records, BinView bv, ImageView iv,
                          int64_t capacity, float* out_color, float* out_depth, float* out_alpha,
                          const float* touch_target, float* residual_out, cudaStream_t st);
int tgs_launch_render_bwd(const TgsCam& cam, const TgsSettings* s, const TgsRecord* gv_records, BinView bv, ImageView iv,
                          int64_t num_rendered,
                          const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                          const TgsTouch* touch, float* residual, float* screen_grads, uint8_t* contrib_flags,
                          cudaStream_t st);
int tgs_launch_loss_scale(const float* target, int64_t P, float mult, float norm, float* out, cudaStream_t st);
int tgs_launch_touch_loss_value(const float* residual, const float* weight, int64_t i0, int64_t i1, int mode,
                                const float* scale, double* acc, float* out, cudaStream_t st);

static inline TgsCam tgs_make_cam(const TgsSettings* s) {
    TgsCam c;
    c.W = s->image_width; c.H = s->image_height;
    c.Tx = (c.W + TGS_TILE - 1) / TGS_TILE; c.Ty = (c.H + TGS_TILE - 1) / TGS_TILE;
    c.fx = (float)c.W / (2.0f * s->tanfovx);
    c.fy = (float)c.H / (2.0f * s->tanfovy);
    c.limx = 1.3f * s->tanfovx; c.limy = 1.3f * s->tanfovy;
    c.mod = s->scale_modifier;
    c.row0 = s->tile_row_begin; c.row1 = s->tile_row_end;
    if (c.row1 <= c.row0) { c.row0 = 0; c.row1 = c.Ty; }
    if (c.row0 < 0) c.row0 = 0;
    if (c.row1 > c.Ty) c.row1 = c.Ty;
    c.deg = s->sh_degree; c.K = s->sh_coeffs;
    c.near_z = s->near_z > 0.0f ? s->near_z : TGS_NEAR_Z;
    c.alpha_max = s->alpha_max > 0.0f ? s->alpha_max : TGS_ALPHA_MAX;
    c.ppx = s->principal_dx; c.ppy = s->principal_dy;
    return c;
}
