/*
 * tgs.h -- torch-free C ABI of the B200-native Touch-GS rasterizer (libtgs.so).
 *
 * This is the drop-in boundary of the hot path.  The reference does not
 * vendor its rasterizer (reference .gitmodules:7-9 -> empty nerfstudio/ submodule), so the
 * interface each entry point replaces is the *operator* surface the reference's trainer calls
 * (`ns-train depth-gaussian-splatting`, reference scripts/train_bunny_real.sh:52,
 * scripts/train_block_data.sh:50), i.e. the public Inria-style C++ entry points
 * RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible named in SURVEY.md
 * §2.2 rows H2-H10 and §8(b), extended with expected depth, alpha and the fused touch-depth
 * loss.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 / int32 unless its name ends in
 *     `_host`;  matrices are 16 floats, passed TRANSPOSED (row-vector convention) exactly as
 *     the operator's GaussianRasterizationSettings holds them;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - every function returns 0 on success, a negative TGS_E* code or a positive cudaError_t
 *     otherwise; tgs_last_error() gives the message (thread-local);
 *   - kernels never own memory: scratch comes from the caller through `tgs_alloc_fn`.
 */
#ifndef TGS_H_
#define TGS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGS_ABI_VERSION 8

#define TGS_EINVAL   (-1)   /* bad argument combination / shape */
#define TGS_ENOMEM   (-2)   /* allocator callback returned NULL */
#define TGS_ESTATE   (-3)   /* saved buffers do not match the call */
#define TGS_EOVERFLOW (-4)  /* a deferred speculative forward rendered more instances than its buffers held */

/* which scratch buffer an allocation is for (the three opaque byte buffers of SURVEY §8b) */
#define TGS_BUF_GEOM    0   /* per-Gaussian state, lives fwd -> bwd */
#define TGS_BUF_BINNING 1   /* per-instance state (the sorted id list, checkpoints of the segmented backward) */
#define TGS_BUF_IMAGE   2   /* per-pixel state (final_T, n_contrib, raw depth) */
#define TGS_BUF_TEMP    3   /* temporaries not needed by backward */

/* Returns a device pointer to >= bytes of memory aligned to >= 256 B, or NULL. */
typedef void* (*tgs_alloc_fn)(void* user, int which, size_t bytes);

/* depth-loss modes of the fused touch gradient */
#define TGS_LOSS_NONE 0
#define TGS_LOSS_L1   1
#define TGS_LOSS_L2   2

/* Camera + raster settings: mirrors GaussianRasterizationSettings. */
typedef struct TgsSettings {
    int32_t image_width;
    int32_t image_height;
    float   tanfovx;
    float   tanfovy;
    float   scale_modifier;
    int32_t sh_degree;        /* active degree 0..3 */
    int32_t sh_coeffs;        /* K = coefficients stored per Gaussian (stride of shs) */
    int32_t prefiltered;
    int32_t debug;            /* 1: synchronise + check after every kernel */
    int32_t tile_row_begin;   /* tile-row band rendered by this rank (multi-GPU shard) */
    int32_t tile_row_end;     /* exclusive; (0, ceil(H/16)) = whole image; (0,0) is also whole image */
    int32_t depth_normalize;  /* 1: returned depth = D/alpha (expected depth), 0: raw sum */
    int32_t defer_count;      /* with rendered_hint > 0: do NOT wait for num_rendered in tgs_forward at all.  saved->num_rendered
                               * is then a negative TICKET; tgs_backward_render (or tgs_forward_resolve) redeems it once the count
                               * has long arrived.  If the count exceeded the hint the forward's outputs were computed from a
                               * truncated list: the redeeming call returns TGS_EOVERFLOW and the caller must redo the step with
                               * a larger hint (a trainer's hint = the view's previous count + a margin never overflows in steady
                               * state).  Keeps the host a full step ahead of the GPU: what a tile-row shard with little work per
                               * rank needs. */
    int32_t contrib_flags;    /* 1: every screen-gradient buffer handed to tgs_backward_render / tgs_backward_preprocess /
                               * tgs_backward_preprocess_gather is tgs_screen_grad_bytes(N, 1) bytes long: the [N,10] rows
                               * followed (at TGS_SCREEN_GRAD_FLAG_OFFSET(N)) by one CONTRIBUTOR byte per Gaussian, set by
                               * BACKWARD::render for every Gaussian some pixel blended.  The chain rule then reads neither
                               * the row nor the parameters of a non-contributor (its gradients are exactly zero), and the
                               * multi-GPU gather asks a peer only for the rows that peer flagged: in a dense scene most
                               * Gaussians lie behind the depth at which their tiles saturate (c3: 86 %).
                               * 0: plain [N,10] buffers (required when the buffers are summed by an all-reduce; also the
                               * better choice on ONE GPU, where the chain rule spots the all-zero rows itself and the
                               * bytes would only cost BACKWARD::render 3 %).
                               * (occupies what was alignment padding: the struct layout is unchanged) */
    int64_t rendered_hint;    /* 0: synchronous sizing (read num_rendered, then bin).  > 0: SPECULATIVE mode: the
                               * binning buffers are sized for this many instances and the binning + render kernels
                               * are enqueued BEFORE the host waits for the real count (the wait is on an event
                               * recorded right after the scan, so the GPU always has work queued behind it).
                               * If the real count exceeds the hint, the tail is re-run with the exact size:
                               * results are identical in both modes. */
    const float* viewmatrix;  /* [16] device */
    const float* projmatrix;  /* [16] device */
    const float* campos;      /* [3]  device */
    const float* bg;          /* [3]  device */
    /* convention switches; 0 = the primary (Inria) convention */
    float   alpha_max;        /* alpha clamp: 0 -> 0.99;  gsplat 0.1.x: 0.999 */
    float   near_z;           /* near-plane cull, view z <= near_z: 0 -> 0.2;  gsplat 0.1.x clip_thresh: 0.01 */
    float   principal_dx;     /* principal point offset from the image centre in pixels (cx - W/2, cy - H/2): added to */
    float   principal_dy;     /* the pixel mean after the NDC -> pixel map (gsplat 0.1.x ndc2pix(x, W, cx)) */
} TgsSettings;

/* Per-Gaussian inputs (replaces the tensor arguments of rasterize_gaussians). */
typedef struct TgsGaussians {
    int32_t N;
    const float* means3D;        /* [N,3] */
    const float* opacities;      /* [N]   */
    const float* shs;            /* [N,K,3] or NULL */
    const float* colors_precomp; /* [N,3]   or NULL  (exactly one of shs / colors_precomp) */
    const float* scales;         /* [N,3]   or NULL */
    const float* rotations;      /* [N,4]   or NULL  (w,x,y,z), used as given */
    const float* cov3D_precomp;  /* [N,6]   or NULL  (exactly one of (scales,rotations) / cov3D_precomp) */
} TgsGaussians;

/* Fused touch-depth supervision. target==NULL or mode==NONE disables it. */
typedef struct TgsTouch {
    const float* target;   /* [H,W] metres (x scene scale); 0 = invalid  (reference utils/fuse_touch_vision.py:372-388) */
    const float* weight;   /* [H,W] per-pixel weight (e.g. 1/sigma) or NULL = 1 */
    const float* scale;    /* device scalar: depth_loss_mult / Z  (see tgs_touch_loss_scale) */
    int32_t mode;          /* TGS_LOSS_* */
    int32_t row_begin;     /* pixel rows [row_begin,row_end) where the loss applies; (0,0) = every row.  A rank of the */
    int32_t row_end;       /* tile-row shard that renders a halo around its band restricts the loss to its own rows. */
    const float* grad_scale; /* device scalar or NULL (= 1): upstream gradient of the touch-loss scalar, multiplied into
                              * the fused gradient on the device (loss scaling, gradient accumulation, GradScaler) */
} TgsTouch;

/* Saved state handed from forward to backward. */
typedef struct TgsSaved {
    void*   geom;          /* TGS_BUF_GEOM    */
    void*   binning;       /* TGS_BUF_BINNING */
    void*   image;         /* TGS_BUF_IMAGE   */
    int64_t num_rendered;  /* I */
    int64_t capacity;      /* instances the binning buffer was laid out for (>= num_rendered) */
} TgsSaved;

/* Gradient outputs of backward (all device, fully written by the kernels: no memset needed). */
typedef struct TgsGrads {
    float* dmeans2D;   /* [N,3] NDC-scaled mean gradient (x*0.5W, y*0.5H, 0): densifier statistic */
    float* dmeans3D;   /* [N,3] */
    float* dopacity;   /* [N]   */
    float* dshs;       /* [N,K,3] or NULL */
    float* dcolors;    /* [N,3]   or NULL */
    float* dscales;    /* [N,3]   or NULL */
    float* drotations; /* [N,4]   or NULL */
    float* dcov3D;     /* [N,6]   or NULL */
} TgsGrads;

int         tgs_abi_version(void);
const char* tgs_last_error(void);
/* number of kernels this library has launched in this process (own kernels, CUB kernels) */
void        tgs_launch_counts(uint64_t* own, uint64_t* cub);

/*
 * Stage timers: when enabled, every kernel stage is bracketed by cudaEventRecord on the launching
 * stream (what bench.py's roofline uses: per-launch durations measured live inside the timed region).
 * tgs_profile_read synchronises nothing: call it after the stream is idle.  It returns, per stage,
 * the accumulated milliseconds and launch count since the last read, then resets.
 */
#define TGS_STAGE_PREPROCESS      0
#define TGS_STAGE_SCAN            1   /* depth sort of the N Gaussians (CUB) */
#define TGS_STAGE_DUPLICATE       2   /* per-chunk tile counts (k_bin_count) */
#define TGS_STAGE_SORT            3   /* per-tile prefixes over the chunks + tile ranges (k_bin_prefix, k_bin_ranges) */
#define TGS_STAGE_PACK            4
#define TGS_STAGE_RENDER_FWD      5
#define TGS_STAGE_LOSS_SCALE      6
#define TGS_STAGE_RENDER_BWD      7
#define TGS_STAGE_PREPROCESS_BWD  8
#define TGS_STAGE_PHOTO_FWD       9   /* train step: fused L1+SSIM loss forward */
#define TGS_STAGE_PHOTO_BWD       10
#define TGS_STAGE_ACTIVATE        11  /* activations forward + backward */
#define TGS_STAGE_ADAM            12
#define TGS_STAGE_REFINE          13  /* refine statistics + densify plan / apply */
#define TGS_STAGE_BIN_SCATTER     14  /* ordered scatter of the Gaussian ids into the per-tile lists */
#define TGS_NUM_STAGES            15
int tgs_profile_enable(int32_t on);
int tgs_profile_read(float* ms_per_stage, int32_t* launches_per_stage);

/* Redeem the ticket a deferred forward left in saved->num_rendered (see TgsSettings.defer_count): waits for the count's
 * event (normally long complete), returns the instance count.  Idempotent until the ticket slot is reused (8 forwards later). */
int tgs_forward_resolve(int64_t ticket, int64_t capacity, int64_t* num_rendered_out);

/* replaces markVisible: present[i] = view-space z > 0.2 */
int tgs_mark_visible(int32_t N, const float* means3D, const float* viewmatrix,
                     uint8_t* present, void* stream);

/*
 * replaces RasterizeGaussiansCUDA.
 * Outputs: out_color [3,H,W], out_depth [H,W], out_alpha [H,W], radii [N] (int32),
 * saved (buffers come from `alloc`).  One stream-synchronising D2H read of num_rendered.
 * touch_target / residual_out [H,W] (both or neither NULL): residual = depth - target where
 * target > 0 and alpha > 0, else 0 (logging value of the touch-depth loss).
 * Pixels outside this rank's tile-row band are left untouched.
 */
int tgs_forward(const TgsSettings* s, const TgsGaussians* g,
                tgs_alloc_fn alloc, void* alloc_user,
                float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
                const float* touch_target, float* residual_out,
                TgsSaved* saved, void* stream);

/*
 * replaces BACKWARD::render with the touch-depth gradient fused in.
 * dL_dcolor [3,H,W]; dL_ddepth / dL_dalpha [H,W] or NULL (external autograd grads on the
 * returned depth / alpha);  residual_out [H,W] or NULL.
 * screen_grads [N,10] = (dx,dy (pixel units), dA,dB,dC, dopacity, dr,dg,db, ddepth) is
 * ZEROED then accumulated: this is the buffer the multi-GPU path exchanges.
 * With s->contrib_flags the buffer is tgs_screen_grad_bytes(N, 1) bytes: rows, then the contributor bytes.
 */
#define TGS_SCREEN_GRAD_FLAG_OFFSET(N) ((((size_t)(N) * 40u) + 127u) / 128u * 128u)   /* bytes; flags are 128-byte aligned */
size_t tgs_screen_grad_bytes(int32_t N, int32_t contrib_flags);
int tgs_backward_render(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                        const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                        const TgsTouch* touch, float* residual_out,
                        float* screen_grads, void* stream);

/* replaces BACKWARD::preprocess: screen_grads -> parameter gradients. */
int tgs_backward_preprocess(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                            const int32_t* radii, const float* screen_grads,
                            const TgsGrads* grads, void* stream);

/*
 * Multi-GPU form of tgs_backward_preprocess with the exchange FUSED into the kernel: rank r rendered
 * the tile rows [peer_tile_rows_host[2r], peer_tile_rows_host[2r+1]) into ITS screen-gradient buffer
 * peer_screen_grads_host[r] (device pointers valid on THIS device: peer-mapped / symmetric memory over NVLink,
 * the local rank's own buffer included).  For every Gaussian the kernel reads the rows of exactly those ranks whose
 * band its tile-row span touches and sums them in ascending rank order -- no all-reduce, no second kernel.
 * With s->contrib_flags (every peer buffer carries its contributor bytes) a warp first fetches the 32 flag bytes of its
 * Gaussians from each peer (one sector per peer) and then asks only for the rows that peer actually wrote.
 * The caller must have synchronised the ranks (all tgs_backward_render calls finished) before this runs.
 */
#define TGS_MAX_PEERS 8
int tgs_backward_preprocess_gather(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                   const int32_t* radii, const float* const* peer_screen_grads_host,
                                   const int32_t* peer_tile_rows_host, int32_t world,
                                   const TgsGrads* grads, void* stream);

/* replaces RasterizeGaussiansBackwardCUDA: tgs_backward_render + tgs_backward_preprocess. */
int tgs_backward(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                 const int32_t* radii,
                 const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                 const TgsTouch* touch, float* residual_out,
                 float* screen_grads, const TgsGrads* grads, void* stream);

/*
 * scale_out[0] = mult / max(1, #(target > 0))   (or mult / norm when norm > 0);
 * scale_out[1] is an integer workspace.  scale_out: 2 x 4 bytes, device.
 * (loss multiplier semantics: reference legacy/model_tactile.py:162; validity = depth > 0:
 *  reference utils/fuse_touch_vision.py:51,109)
 */
int tgs_touch_loss_scale(const float* target, int64_t num_pixels, float mult, float norm,
                         float* scale_out, void* stream);

/*
 * Value of the fused touch loss, for logging and for callers that want it as a term of their autograd graph:
 * loss_out[0] = scale[0] * sum over pixel rows [row_begin,row_end) ((0,0) = all) of weight*|residual| (L1) or
 * weight*residual^2 (L2); residual as written by tgs_forward (0 where the pixel is invalid), weight NULL = 1.
 * acc: 8 bytes of device workspace (a double).
 */
int tgs_touch_loss_value(const float* residual, const float* weight, int32_t W, int32_t H, int32_t row_begin,
                         int32_t row_end, int32_t mode, const float* scale, double* acc, float* loss_out, void* stream);

/*
 * SURVEY §8(f) N2 -- the per-pixel part of the reference's touch / vision depth fusion, fp64, bit-identical
 * to the uint16 PNGs the reference writes.  Replaces (per image) reference utils/fuse_touch_vision.py
 * :270-276 (uint16-mm decode), :288-306 (apply the fitted alignment: scale/offset from the first fit,
 * offset2 from the second), :310-313 (vision sigma = clip(0.05*depth,0,10)+5), :76-202 (inverse-variance
 * fusion), :360-361 (clips), :373-376 (uint16 encode) and the trainer-side decode (reference
 * legacy/dataparser_tactile.py:65-66): target = mm*1e-3*scene_scale, weight = 1/sigma (0 where sigma = 0).
 * Inputs / uint16 outputs: device, 8-byte aligned; target / weight: device fp32, 16-byte aligned.
 * Any output pointer may be NULL.
 */
int tgs_fuse_touch_vision(const uint16_t* touch_mm, const uint16_t* vision_mm, const uint16_t* touch_sigma_mm,
                          int64_t num_pixels, double scale, double offset, double offset2,
                          int32_t is_real_world, double scene_scale,
                          uint16_t* vision_aligned_mm, uint16_t* ds_gs_mm, uint16_t* fused_mm,
                          uint16_t* fused_sigma_mm, float* target, float* weight, void* stream);

/*
 * SURVEY §8(f) N2 -- trainer-side decode of the on-disk touch maps: uint16 millimetre depth PNG -> fp32 target
 * (mm * depth_unit; depth_unit = 1e-3 * pose scale: reference legacy/dataparser_tactile.py:65-66,229-235), uint16
 * uncertainty PNG (sigma x 1000, reference utils/fuse_touch_vision.py:376,387) -> per-pixel weight
 * (weight_mode 0: 1 [SIMPLE_LOSS]; 1: 1/(uw*sigma); 2: 1/(uw*sigma)^2; 0 where sigma == 0), uw = uncertainty_weight
 * (reference scripts/train_bunny_real.sh:52).  Either output may be NULL.
 */
int tgs_decode_touch_maps(const uint16_t* depth_mm, const uint16_t* sigma_mm, int64_t num_pixels, float depth_unit,
                          float uncertainty_weight, int32_t weight_mode, float* target, float* weight, void* stream);

/*
 * Host-buffer entry point (the call a non-PyTorch trainer makes; timed by the end-to-end bench with
 * every host<->device copy inside the timed region).  EVERY pointer reachable from s_host / g_host /
 * grads_host and every *_host argument is a HOST pointer (NULL = absent / not wanted).  The call
 * uploads, runs forward, forms dL/dcolor = sign(color - gt_rgb)/(3HW) (L1 photometric; zero when
 * gt_rgb_host is NULL) plus the fused touch loss (mode/mult as in TgsTouch, Z = #(target>0)), runs
 * backward and downloads what was asked for.  Scratch comes from cudaMallocAsync on `stream`.
 * *loss_host receives the photometric loss.  Synchronises the stream before returning.
 */
int tgs_train_step_host(const TgsSettings* s_host, const TgsGaussians* g_host,
                        const float* gt_rgb_host, const float* touch_target_host,
                        const float* touch_weight_host, int32_t loss_mode, float depth_loss_mult,
                        const TgsGrads* grads_host, float* out_color_host, float* out_depth_host,
                        int32_t* radii_host, float* loss_host, int64_t* num_rendered_host,
                        void* stream);

/* introspection used by tests: layout of the saved buffers (byte offsets from the base) */
typedef struct TgsGeomLayout {
    size_t records;           /* TgsRecord[N]: 3 x float4 = (x,y,depth,id) (A,B,C,opacity) (r,g,b,thr) */
    size_t cov3D;             /* float[N,6] */
    size_t tiles_touched;     /* uint32[N] */
    size_t clamped;           /* uint8[N] bit c = colour channel c clamped */
    size_t rect;              /* uint32[N,2]: (rminx | rmaxx<<16, rminy | rmaxy<<16) */
    size_t depth_keys;        /* uint32[N] bits(depth), 0xFFFFFFFF when nothing is emitted */
    size_t ids;               /* uint32[N] iota */
    size_t depth_keys_sorted; /* uint32[N] */
    size_t order;             /* uint32[N] Gaussian ids in ascending (depth, id) order */
    size_t span_sorted;       /* uint32[N,2]: `rect` of the Gaussians in that order ((0,0) = emits nothing) */
    size_t temp;              /* CUB temp for the depth sort */
    size_t temp_bytes;
    size_t total;
} TgsGeomLayout;
typedef struct TgsBinningLayout {
    size_t vals_sorted;       /* uint32[I] Gaussian ids, final (tile, depth, id) order: each tile's list is one contiguous run
                               * (the compositing kernels TMA-copy the ids and gather the per-Gaussian records by id) */
    size_t ckpt;              /* float[slots][5][256]: per-pixel (T, r, g, b, D) composited BEFORE list position
                               * ranges[tile].x + 256k, written by the forward for every 256-record boundary it crosses;
                               * slot = that position >> 8 (unique per boundary).  Lets the backward replay a tile's list
                               * in independent 256-record segments. */
    size_t slot_tile;         /* uint32[slots]: tile that owns the boundary in this slot (valid for listed slots) */
    size_t ckpt_list;         /* uint32[slots]: the slots the forward wrote, in completion order */
    size_t work_counter;      /* uint32[2]: [0] dynamic work-unit counter of the backward, [1] length of ckpt_list */
    size_t slots;             /* (I >> 8) + 2 */
    size_t total;
} TgsBinningLayout;
typedef struct TgsImageLayout {
    size_t final_T;        /* float[H,W] */
    size_t n_contrib;      /* uint32[H,W] */
    size_t depth_raw;      /* float[H,W] un-normalised sum depth*alpha*T */
    size_t color_acc;      /* float[3,H,W] composited colour WITHOUT the background term */
    size_t ranges;         /* uint32[T,2]: per-tile [start, end) into the sorted instance list */
    size_t count;          /* uint32[2]: num_rendered (device copy), 32-bit overflow flag */
    size_t total;
} TgsImageLayout;
int tgs_geom_layout(int32_t N, TgsGeomLayout* out);
int tgs_binning_layout(int64_t num_rendered, TgsBinningLayout* out);
int tgs_image_layout(int32_t W, int32_t H, TgsImageLayout* out);

/*
 * "Reference-structure CUDA" comparison arm (BASELINE.md §3 column 2; SURVEY.md §8d "Reference CUDA path beside
 * it").  The reference's own rasterizer is not in its tree (reference .gitmodules:7-9), so the comparison is the
 * same algorithm in the upstream kernels' structure, written from the spec: id-order scan + blocking count read,
 * per-Gaussian 64-bit key emission, ONE 12-byte-pair radix sort, 16x16 one-pixel-per-thread compositing with
 * gathered per-Gaussian data and no culling, ten per-thread global atomics per (pixel, Gaussian) pair in backward,
 * touch-depth loss NOT fused (the caller forms dL/ddepth_raw and dL/dalpha).  Used by bench.py and the full-size
 * cross-check tests only; the operator never calls it.  out_depth_raw = sum depth*alpha*T (un-normalised).
 * saved->geom / image have the layouts of tgs_geom_layout / tgs_image_layout (tgs_backward_preprocess accepts
 * them); saved->binning has the layout below.  Whole image only (no tile-row band).
 */
typedef struct TgsRefBinningLayout {
    size_t offsets;        /* uint32[N] inclusive scan of tiles_touched in Gaussian-id order */
    size_t keys_unsorted;  /* uint64[I] (tile << 32) | bits(depth), emission order */
    size_t keys_sorted;    /* uint64[I] */
    size_t vals_unsorted;  /* uint32[I] */
    size_t vals_sorted;    /* uint32[I] */
    size_t ranges;         /* uint32[T,2] */
    size_t temp;
    size_t temp_bytes;
    size_t total;
} TgsRefBinningLayout;
int tgs_refstructure_binning_layout(int32_t N, int64_t num_rendered, int32_t num_tiles, TgsRefBinningLayout* out);
int tgs_refstructure_forward(const TgsSettings* s, const TgsGaussians* g, tgs_alloc_fn alloc, void* alloc_user,
                             float* out_color, float* out_depth_raw, float* out_alpha, int32_t* radii,
                             TgsSaved* saved, void* stream);
int tgs_refstructure_backward_render(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                     const float* dL_dcolor, const float* dL_ddepth_raw, const float* dL_dalpha,
                                     float* screen_grads, void* stream);

/* ====================================================================================================
 * SURVEY.md §8(f) row N1 -- the rest of a Touch-GS TRAIN STEP around the rasterizer (BASELINE config c5:
 * "full Touch-GS train step (Adam + densify)").  The trainer is `ns-train depth-gaussian-splatting`
 * (reference scripts/train_bunny_real.sh:52) from the reference's EMPTY nerfstudio submodule (reference
 * .gitmodules:7-9), so these entry points replace the per-step torch ops of that model's loss / optimizer /
 * refine code as publicly documented for the splat trainers of that era; what the tree
 * pins: AdamOptimizerConfig(lr=..., eps=1e-15) (reference legacy/config_tactile.py:43-50), 30000 iterations
 * (reference legacy/config_tactile.py:28), the depth-loss knobs (reference scripts/train_block_data.sh:50).
 * All pointers are DEVICE pointers unless named *_host.
 * ==================================================================================================== */

/* floats of scratch tgs_photometric_loss_forward needs (three derivative maps per channel) */
size_t tgs_photometric_scratch_floats(int32_t W, int32_t H);
/*
 * loss = (1-l) * mean|C - C*| + l * (1 - mean SSIM(C, C*)); 11x11 Gaussian window (sigma 1.5), zero padding,
 * C1 = 0.01^2, C2 = 0.03^2; means over all 3*H*W elements.  color / gt: [3,H,W].  Only rows
 * [row_begin,row_end) are loss pixels ((0,0) = all): the partial losses of the disjoint bands of a tile-row
 * shard add up to the full-image loss.  sums: 2 doubles (workspace), loss_out: 1 float.
 */
int tgs_photometric_loss_forward(const float* color, const float* gt, int32_t W, int32_t H,
                                 int32_t row_begin, int32_t row_end, float lambda_dssim, float* dmaps,
                                 double* sums, float* loss_out, void* stream);
/* dL_dcolor rows [out_row_begin,out_row_end) <- grad_out[0] (device scalar, NULL = 1) * dloss/dcolor;
 * pixels within 5 rows of the loss band receive SSIM gradient (the band's halo). */
int tgs_photometric_loss_backward(const float* color, const float* gt, const float* dmaps, int32_t W, int32_t H,
                                  int32_t row_begin, int32_t row_end, int32_t out_row_begin, int32_t out_row_end,
                                  float lambda_dssim, const float* grad_out, float* dL_dcolor, void* stream);

/* raw parameters -> rasterizer inputs: scales = exp(scales_log), rotations = quats/|quats|, opacities = sigmoid */
int tgs_activate_forward(int32_t N, const float* scales_log, const float* quats, const float* opacity_logit,
                         float* scales, float* rotations, float* opacities, void* stream);
/* chain rule of the above; every output may alias the matching d* input */
int tgs_activate_backward(int32_t N, const float* scales_log, const float* quats, const float* opacity_logit,
                          const float* dscales, const float* drotations, const float* dopacities,
                          float* dscales_log, float* dquats, float* dopacity_logit, void* stream);

/* One-launch Adam over up to TGS_ADAM_MAX_GROUPS parameter groups (torch.optim.Adam arithmetic, no weight decay,
 * no amsgrad).  period > 0: element i uses `lr` when (i % period) < head, else `lr_tail` (one [N,K,3] SH tensor
 * with the DC / higher-band learning rates of the features_dc / features_rest groups, no torch.cat per step). */
#define TGS_ADAM_MAX_GROUPS 8
typedef struct TgsAdamGroup {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;   /* 16-byte aligned */
    int64_t numel;
    float lr, lr_tail;
    int32_t period, head;
} TgsAdamGroup;
int tgs_adam_step(const TgsAdamGroup* groups_host, int32_t n_groups, int32_t step, double beta1, double beta2,
                  double eps, void* stream);   /* doubles: 1-beta and the bias corrections are formed in double, as torch does */

/* refine statistics of one step: visible (radii > 0) Gaussians accumulate |dL/dmean2D| (NDC-scaled, as returned in
 * TgsGrads.dmeans2D), a visit count and their maximum screen radius */
int tgs_densify_stats(int32_t N, const float* dmeans2D, const int32_t* radii, float* grad_accum,
                      int32_t* vis_count, int32_t* max_radii, void* stream);

typedef struct TgsDensifyConfig {
    float grad_thresh;        /* mean |dL/dmean2D| above which a Gaussian is duplicated / split (0.0002) */
    float size_thresh;        /* max world scale above which it is split instead of duplicated (0.01) */
    float cull_alpha_thresh;  /* cull when sigmoid(opacity) < this (0.1) */
    float cull_scale_thresh;  /* cull when max world scale > this (0.5) */
    float split_shrink;       /* split samples get scale / this (1.6) */
    int32_t n_split_samples;  /* 2 */
    float split_screen_radius; /* > 0: split when the largest screen radius (pixels) since the last refine exceeds this */
    float cull_screen_radius;  /* > 0: cull when it exceeds this */
} TgsDensifyConfig;
typedef struct TgsParamSet { float* means; float* shs; float* opacity; float* scales; float* quats; } TgsParamSet;
size_t tgs_densify_temp_bytes(int32_t N);
/* counts[i] (bit 31 = split) / offsets[i] (exclusive scan) of the outputs of Gaussian i: culled 0, kept 1,
 * duplicated 2 (itself, copy), split n_split_samples (the original is dropped).  Synchronises to return the total.
 * N == 0 is legal (total 0). */
int tgs_densify_plan(int32_t N, const float* opacity_logit, const float* scales_log, const float* grad_accum,
                     const int32_t* vis_count, const int32_t* max_radii /* NULL: no screen-size rules */,
                     const TgsDensifyConfig* cfg, int32_t allow_split_dup,
                     uint32_t* counts, uint32_t* offsets, void* temp, size_t temp_bytes,
                     int64_t* total_host, void* stream);
/* in_pmv / out_pmv: 3 TgsParamSet each = (values, exp_avg, exp_avg_sq); raw parameters (opacity logit [N],
 * scales log [N,3], quats [N,4], shs [N,K,3], means [N,3]).  noise: [N, n_split_samples, 3] ~ N(0,1).
 * src_out[o] = source id for carried-over entries, -(source id + 1) for new ones (zero Adam moments). */
int tgs_densify_apply(int32_t N, int32_t K, const uint32_t* counts, const uint32_t* offsets, const float* noise,
                      const TgsDensifyConfig* cfg, const TgsParamSet* in_pmv, const TgsParamSet* out_pmv,
                      int32_t* src_out, void* stream);

/* ====================================================================================================
 * SURVEY.md §8(f) row N3 -- the hot path split at the screen-space boundary, as the gsplat-0.1-style three-call API
 * of the nerfstudio splat model the Touch-GS fork builds on expects it (reference .gitmodules:7-9 -> empty submodule;
 * trainer entry reference scripts/train_bunny_real.sh:52): project_gaussians -> caller's colours -> rasterize_gaussians.
 * Same kernels as the fused path; convention differences are the TgsSettings switches alpha_max / near_z /
 * principal_dx,dy and the pixel_offset argument.
 * ==================================================================================================== */
/* A1 alone (geometry only: shs / colors_precomp / opacities of `g` are ignored).  Results live in saved->geom
 * (layout: tgs_geom_layout): records (x, y, depth | conic A,B,C), cov3D, tiles_touched; radii [N] is written. */
int tgs_project_gaussians(const TgsSettings* s, const TgsGaussians* g, tgs_alloc_fn alloc, void* alloc_user,
                          int32_t* radii, TgsSaved* saved, void* stream);
/* chain rule of the projection: screen_grads [N,10] (slots 0,1 = d/dxy in pixels, 2..4 = d/dconic, 9 = d/ddepth; the
 * others are ignored) -> grads->dmeans3D, dscales, drotations (or dcov3D), dmeans2D. */
int tgs_project_gaussians_backward(const TgsSettings* s, const TgsGaussians* g, const TgsSaved* saved,
                                   const int32_t* radii, const float* screen_grads, const TgsGrads* grads, void* stream);
/* A2-A5 from caller-provided screen-space tensors: xys [N,2] pixels, depths [N], radii [N] (<= 0: skipped),
 * conics [N,3], colors [N,3], opacities [N].  Sample point of pixel (x,y) = (x + pixel_offset, y + pixel_offset).
 * out_color [3,H,W] (incl. background), out_depth [H,W] = sum depth*alpha*T, out_alpha [H,W].
 * Only image_width / image_height / bg / alpha_max / tile rows of `s` are used. */
int tgs_rasterize_screen_forward(const TgsSettings* s, int32_t N, const float* xys, const float* depths,
                                 const int32_t* radii, const float* conics, const float* colors,
                                 const float* opacities, float pixel_offset, tgs_alloc_fn alloc, void* alloc_user,
                                 float* out_color, float* out_depth, float* out_alpha, TgsSaved* saved, void* stream);
/* A6 without the touch fusion: screen_grads [N,10] = (dxy pixels, dconic, dopacity, dcolor, ddepth) */
int tgs_rasterize_screen_backward(const TgsSettings* s, int32_t N, const TgsSaved* saved, const float* dL_dcolor,
                                  const float* dL_ddepth, const float* dL_dalpha, float* screen_grads, void* stream);
/* colours [N,3] = sum_k Y_k(dir/|dir|) coeffs[N,K,3] over the (degree+1)^2 active bases, RAW (no +0.5, no clamp) */
int tgs_spherical_harmonics(int32_t N, int32_t degree, int32_t K, const float* dirs, const float* coeffs,
                            float* colors, void* stream);
int tgs_spherical_harmonics_backward(int32_t N, int32_t degree, int32_t K, const float* dirs, const float* v_colors,
                                     float* v_coeffs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TGS_H_ */
