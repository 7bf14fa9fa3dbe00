// torch_ext.cpp -- the PyTorch C++ extension module `_C` of the operator (SURVEY.md §8(b) "C++/C-ABI surface").
//
// Exports, over the torch-free C ABI of include/tgs.h (libtgs.so, hand-written sm_100a kernels):
//   rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
//                       projmatrix, tanfovx, tanfovy, H, W, sh, degree, campos, prefiltered, debug, ...extension)
//       -> (num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer, residual, capacity)
//   rasterize_gaussians_backward(...saved..., dL_dcolor, dL_ddepth, dL_dalpha, touch_depth, touch_weight, loss_mode,
//                                loss_scale, ...) -> (dmeans2D, dcolors, dopacity, dmeans3D, dcov3D, dsh, dscales, drot)
//   mark_visible(means3D, viewmatrix, projmatrix) -> bool[N]
// i.e. the entry points of the reference-era binding (RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA /
// markVisible; the rasterizer itself is not in the reference tree: reference .gitmodules:7-9), extended
// with expected depth, alpha, the tile-row band and the fused touch-depth loss.  Plus the two halves of the backward
// (backward_render / backward_preprocess[_gather]) that the tile-row shard puts its exchange between.
//
// Conventions: shape / dtype / device mismatches are TORCH_CHECK errors; outputs and the three saved
// byte buffers are torch-allocated (at::empty on the inputs' device, handed to the kernels through the C ABI's
// allocator callback); kernels run on at::cuda::getCurrentCUDAStream(); a CUDAGuard pins the inputs' device (backward
// may run on autograd's thread).  No compute happens here and there is no CPU path.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <exception>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/tgs.h"

namespace {

using at::Tensor;
using OptT = c10::optional<Tensor>;

struct Alloc {
    Tensor bufs[4];
    c10::Device dev{c10::kCUDA, 0};
    std::exception_ptr err;
};
void* alloc_cb(void* user, int which, size_t bytes) {
    auto* a = static_cast<Alloc*>(user);
    try {
        if (which < 0 || which > 3) return nullptr;
        a->bufs[which] = at::empty({(int64_t)(bytes < 256 ? 256 : bytes)}, at::TensorOptions().dtype(at::kByte).device(a->dev));
        return a->bufs[which].data_ptr();
    } catch (...) {               // must not propagate through the C frame
        a->err = std::current_exception();
        return nullptr;
    }
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, " failed (code ", rc, "): ", tgs_last_error());
}

// contiguous fp32 CUDA tensor on `dev` with the given shape (-1 = any extent); returns the tensor to keep alive
Tensor f32(const Tensor& t, const char* name, std::vector<int64_t> shape, const c10::Device& dev) {
    TORCH_CHECK(t.defined(), name, " must be a tensor");
    TORCH_CHECK(t.device() == dev, name, " must be on ", dev, ", got ", t.device());
    TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32, got ", t.scalar_type());
    TORCH_CHECK((size_t)t.dim() == shape.size(), name, " must have ", shape.size(), " dimensions, got ", t.dim());
    for (size_t i = 0; i < shape.size(); ++i)
        TORCH_CHECK(shape[i] < 0 || t.size(i) == shape[i], name, " must have shape ", at::IntArrayRef(shape), ", got ", t.sizes());
    return t.contiguous();
}
// an absent optional tensor is passed as undefined or as the 1-D zero-element placeholder (the reference-era module's
// torch.Tensor([])); a [0,K,3] tensor of an EMPTY scene is present
bool present(const Tensor& t) { return t.defined() && !(t.dim() <= 1 && t.numel() == 0); }
const float* fp(const Tensor& t) { return (t.defined() && t.numel() > 0) ? t.data_ptr<float>() : nullptr; }

struct Inputs {           // validated, contiguous views + the C structs that point into them
    Tensor bg, means3D, colors, opacity, scales, rotations, cov3D, view, proj, sh, campos;
    TgsSettings s{};
    TgsGaussians g{};
};

Inputs make_inputs(const Tensor& bg, const Tensor& means3D, const Tensor& colors, const Tensor& opacity, const Tensor& scales,
                   const Tensor& rotations, double scale_modifier, const Tensor& cov3D_precomp, const Tensor& viewmatrix,
                   const Tensor& projmatrix, double tanfovx, double tanfovy, int64_t H, int64_t W, const Tensor& sh,
                   int64_t degree, const Tensor& campos, bool prefiltered, bool debug, int64_t row0, int64_t row1,
                   bool depth_normalize, int64_t rendered_hint, bool defer_count = false) {
    TORCH_CHECK(means3D.defined() && means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "touchgs_b200 rasterizer is CUDA-only (no CPU fallback); means3D is on ", means3D.device());
    const c10::Device dev = means3D.device();
    const int64_t N = means3D.size(0);
    Inputs in;
    in.means3D = f32(means3D, "means3D", {N, 3}, dev);
    in.bg = f32(bg.reshape({-1}), "bg", {3}, dev);
    in.view = f32(viewmatrix, "viewmatrix", {4, 4}, dev);
    in.proj = f32(projmatrix, "projmatrix", {4, 4}, dev);
    in.opacity = f32(opacity.reshape({-1}), "opacities", {N}, dev);
    TORCH_CHECK(present(sh) != present(colors), "Please provide exactly one of either SHs or precomputed colors!");
    const bool sr = present(scales) && present(rotations);
    TORCH_CHECK((present(scales) == present(rotations)) && (sr != present(cov3D_precomp)),
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    int64_t K = 0;
    if (present(sh)) {
        TORCH_CHECK(sh.dim() == 3 && sh.size(0) == N && sh.size(2) == 3, "shs must have shape [N,K,3], got ", sh.sizes());
        K = sh.size(1);
        in.sh = f32(sh, "shs", {N, K, 3}, dev);
        in.campos = f32(campos.reshape({-1}), "campos", {3}, dev);
    } else {
        in.colors = f32(colors, "colors_precomp", {N, 3}, dev);
        if (present(campos)) in.campos = f32(campos.reshape({-1}), "campos", {3}, dev);
    }
    if (sr) {
        in.scales = f32(scales, "scales", {N, 3}, dev);
        in.rotations = f32(rotations, "rotations", {N, 4}, dev);
    } else {
        in.cov3D = f32(cov3D_precomp, "cov3D_precomp", {N, 6}, dev);
    }
    TgsSettings& s = in.s;
    s.image_width = (int32_t)W; s.image_height = (int32_t)H;
    s.tanfovx = (float)tanfovx; s.tanfovy = (float)tanfovy; s.scale_modifier = (float)scale_modifier;
    s.sh_degree = (int32_t)degree; s.sh_coeffs = (int32_t)K;
    s.prefiltered = prefiltered; s.debug = debug;
    s.tile_row_begin = (int32_t)row0; s.tile_row_end = (int32_t)row1;
    s.depth_normalize = depth_normalize; s.rendered_hint = rendered_hint > 0 ? rendered_hint : 0;
    s.defer_count = (defer_count && rendered_hint > 0) ? 1 : 0;
    s.contrib_flags = 0;
    s.viewmatrix = fp(in.view); s.projmatrix = fp(in.proj); s.campos = fp(in.campos); s.bg = fp(in.bg);
    TgsGaussians& g = in.g;
    g.N = (int32_t)N;
    g.means3D = fp(in.means3D); g.opacities = fp(in.opacity); g.shs = fp(in.sh); g.colors_precomp = fp(in.colors);
    g.scales = fp(in.scales); g.rotations = fp(in.rotations); g.cov3D_precomp = fp(in.cov3D);
    return in;
}

void* cur_stream(const c10::Device& dev) { return (void*)at::cuda::getCurrentCUDAStream(dev.index()).stream(); }

// ------------------------------------------------------------------------------------------------ forward
std::tuple<int64_t, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, int64_t>
rasterize_gaussians(const Tensor& bg, const Tensor& means3D, const Tensor& colors, const Tensor& opacity, const Tensor& scales,
                    const Tensor& rotations, double scale_modifier, const Tensor& cov3D_precomp, const Tensor& viewmatrix,
                    const Tensor& projmatrix, double tanfovx, double tanfovy, int64_t H, int64_t W, const Tensor& sh,
                    int64_t degree, const Tensor& campos, bool prefiltered, bool debug,
                    int64_t tile_row_begin, int64_t tile_row_end, bool depth_normalize, int64_t rendered_hint,
                    const OptT& touch_depth, bool defer_count) {
    Inputs in = make_inputs(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                            projmatrix, tanfovx, tanfovy, H, W, sh, degree, campos, prefiltered, debug, tile_row_begin,
                            tile_row_end, depth_normalize, rendered_hint, defer_count);
    const c10::Device dev = in.means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t N = in.means3D.size(0), Ty = (H + 15) / 16;
    TORCH_CHECK(0 <= tile_row_begin && tile_row_begin <= tile_row_end && tile_row_end <= Ty, "tile rows (", tile_row_begin,
                ", ", tile_row_end, ") outside [0, ", Ty, "]");
    const bool full = (tile_row_begin == 0 && (tile_row_end == Ty || tile_row_end == 0));
    auto fo = at::TensorOptions().dtype(at::kFloat).device(dev);
    Tensor color = full ? at::empty({3, H, W}, fo) : at::zeros({3, H, W}, fo);
    Tensor depth = full ? at::empty({1, H, W}, fo) : at::zeros({1, H, W}, fo);
    Tensor alpha = full ? at::empty({1, H, W}, fo) : at::zeros({1, H, W}, fo);
    Tensor radii = at::zeros({N}, fo.dtype(at::kInt));
    Tensor target, resid;
    if (touch_depth.has_value() && touch_depth->defined()) {
        TORCH_CHECK(touch_depth->numel() == H * W, "touch_depth must have H*W = ", H * W, " elements, got ", touch_depth->sizes());
        target = f32(touch_depth->reshape({H, W}), "touch_depth", {H, W}, dev);
        resid = full ? at::empty({1, H, W}, fo) : at::zeros({1, H, W}, fo);
    } else {
        resid = at::zeros({1, H, W}, fo);
    }
    Alloc al;
    al.dev = dev;
    TgsSaved saved{};
    const int rc = tgs_forward(&in.s, &in.g, alloc_cb, &al, color.data_ptr<float>(), depth.data_ptr<float>(),
                               alpha.data_ptr<float>(), N > 0 ? radii.data_ptr<int32_t>() : nullptr, fp(target),
                               target.defined() ? resid.data_ptr<float>() : nullptr, &saved, cur_stream(dev));
    if (al.err) std::rethrow_exception(al.err);
    check_rc(rc, "tgs_forward");
    return std::make_tuple((int64_t)saved.num_rendered, color, depth, alpha, radii, al.bufs[TGS_BUF_GEOM],
                           al.bufs[TGS_BUF_BINNING], al.bufs[TGS_BUF_IMAGE], resid, (int64_t)saved.capacity);
}

// ----------------------------------------------------------------------------------------------- backward
struct Touch {
    Tensor target, weight, scale, gscale;
    TgsTouch t{};
    bool on = false;
};
Touch make_touch(const OptT& touch_depth, const OptT& touch_weight, int64_t loss_mode, const OptT& loss_scale,
                 const OptT& grad_scale, int64_t row0, int64_t row1, int64_t H, int64_t W, const c10::Device& dev) {
    Touch r;
    if (!(touch_depth.has_value() && touch_depth->defined()) || loss_mode == TGS_LOSS_NONE) return r;
    TORCH_CHECK(loss_mode == TGS_LOSS_L1 || loss_mode == TGS_LOSS_L2, "loss_mode must be 0 (none), 1 (l1) or 2 (l2)");
    r.target = f32(touch_depth->reshape({H, W}), "touch_depth", {H, W}, dev);
    if (touch_weight.has_value() && touch_weight->defined()) r.weight = f32(touch_weight->reshape({H, W}), "touch_weight", {H, W}, dev);
    TORCH_CHECK(loss_scale.has_value() && loss_scale->defined() && loss_scale->numel() >= 1, "loss_scale (device scalar) is required with a touch loss");
    r.scale = f32(loss_scale->reshape({-1}), "loss_scale", {-1}, dev);
    if (grad_scale.has_value() && grad_scale->defined()) r.gscale = f32(grad_scale->reshape({-1}), "grad_scale", {-1}, dev);
    r.t.target = fp(r.target); r.t.weight = fp(r.weight); r.t.scale = fp(r.scale); r.t.grad_scale = fp(r.gscale);
    r.t.mode = (int32_t)loss_mode; r.t.row_begin = (int32_t)row0; r.t.row_end = (int32_t)row1;
    r.on = true;
    return r;
}

TgsSaved make_saved(const Tensor& geom, const Tensor& binning, const Tensor& img, int64_t num_rendered, int64_t capacity) {
    TORCH_CHECK(geom.defined() && img.defined() && geom.is_cuda(), "saved buffers missing (forward not run?)");
    TgsSaved s{};
    s.geom = geom.data_ptr();
    s.binning = binning.defined() ? binning.data_ptr() : nullptr;
    s.image = img.data_ptr();
    s.num_rendered = num_rendered; s.capacity = capacity;
    return s;
}

struct GradOut {
    Tensor dmeans2D, dcolors, dopacity, dmeans3D, dcov3D, dsh, dscales, drot;
    TgsGrads g{};
};
GradOut make_grads(const Inputs& in) {
    const int64_t N = in.means3D.size(0);
    auto fo = at::TensorOptions().dtype(at::kFloat).device(in.means3D.device());
    GradOut o;
    o.dmeans2D = at::empty({N, 3}, fo); o.dmeans3D = at::empty({N, 3}, fo); o.dopacity = at::empty({N}, fo);
    if (in.sh.defined()) o.dsh = at::empty({N, in.s.sh_coeffs, 3}, fo);
    if (in.colors.defined()) o.dcolors = at::empty({N, 3}, fo);
    if (in.scales.defined()) { o.dscales = at::empty({N, 3}, fo); o.drot = at::empty({N, 4}, fo); }
    if (in.cov3D.defined()) o.dcov3D = at::empty({N, 6}, fo);
    o.g.dmeans2D = (float*)fp(o.dmeans2D); o.g.dmeans3D = (float*)fp(o.dmeans3D); o.g.dopacity = (float*)fp(o.dopacity);
    o.g.dshs = (float*)fp(o.dsh); o.g.dcolors = (float*)fp(o.dcolors); o.g.dscales = (float*)fp(o.dscales);
    o.g.drotations = (float*)fp(o.drot); o.g.dcov3D = (float*)fp(o.dcov3D);
    return o;
}
using GradTuple = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>;
GradTuple grad_tuple(GradOut& o) {
    return std::make_tuple(o.dmeans2D, o.dcolors, o.dopacity, o.dmeans3D, o.dcov3D, o.dsh, o.dscales, o.drot);
}

// floats a screen-gradient buffer must hold: the [N,10] rows (+ the contributor bytes, TgsSettings.contrib_flags)
int64_t screen_grad_floats(int64_t N, bool contrib_flags) {
    return (int64_t)((tgs_screen_grad_bytes((int32_t)N, contrib_flags ? 1 : 0) + 3) / 4);
}

// BACKWARD::render half: zeroes and fills `screen_grads` [N,10] (caller-owned: it may be a peer-mapped buffer)
void backward_render(const Tensor& bg, const Tensor& means3D, const Tensor& colors, const Tensor& opacity, const Tensor& scales,
                     const Tensor& rotations, double scale_modifier, const Tensor& cov3D_precomp, const Tensor& viewmatrix,
                     const Tensor& projmatrix, double tanfovx, double tanfovy, int64_t H, int64_t W, const Tensor& sh,
                     int64_t degree, const Tensor& campos, bool debug, int64_t tile_row_begin, int64_t tile_row_end,
                     bool depth_normalize, const Tensor& geom, const Tensor& binning, const Tensor& img, int64_t num_rendered,
                     int64_t capacity, const Tensor& dL_dcolor, const OptT& dL_ddepth, const OptT& dL_dalpha,
                     const OptT& touch_depth, const OptT& touch_weight, int64_t loss_mode, const OptT& loss_scale,
                     const OptT& grad_scale, int64_t touch_row_begin, int64_t touch_row_end, Tensor screen_grads,
                     bool contrib_flags) {
    Inputs in = make_inputs(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                            projmatrix, tanfovx, tanfovy, H, W, sh, degree, campos, false, debug, tile_row_begin,
                            tile_row_end, depth_normalize, 0);
    const c10::Device dev = in.means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t N = in.means3D.size(0);
    Tensor gc = f32(dL_dcolor, "grad_color", {3, H, W}, dev);
    Tensor gd, ga;
    if (dL_ddepth.has_value() && dL_ddepth->defined()) gd = f32(dL_ddepth->reshape({H, W}), "grad_depth", {H, W}, dev);
    if (dL_dalpha.has_value() && dL_dalpha->defined()) ga = f32(dL_dalpha->reshape({H, W}), "grad_alpha", {H, W}, dev);
    Touch th = make_touch(touch_depth, touch_weight, loss_mode, loss_scale, grad_scale, touch_row_begin, touch_row_end, H, W, dev);
    in.s.contrib_flags = contrib_flags ? 1 : 0;
    TORCH_CHECK(screen_grads.defined() && screen_grads.is_contiguous() && screen_grads.scalar_type() == at::kFloat &&
                screen_grads.device() == dev && screen_grads.numel() >= screen_grad_floats(N, contrib_flags),
                "screen_grads must be a contiguous float32 buffer of screen_grad_floats(N, contrib_flags) elements on ", dev);
    TgsSaved saved = make_saved(geom, binning, img, num_rendered, capacity);
    check_rc(tgs_backward_render(&in.s, &in.g, &saved, gc.data_ptr<float>(), fp(gd), fp(ga), th.on ? &th.t : nullptr, nullptr,
                                 N > 0 ? screen_grads.data_ptr<float>() : nullptr, cur_stream(dev)), "tgs_backward_render");
}

// BACKWARD::preprocess half.  peer_ptrs empty: `screen_grads` holds the (summed) gradients; otherwise the fused gather
// over the peers' buffers (device addresses valid on this device) with the ranks' tile-row bands.
GradTuple backward_preprocess(const Tensor& means3D, const Tensor& radii, const Tensor& colors, const Tensor& opacity,
                              const Tensor& scales, const Tensor& rotations, double scale_modifier, const Tensor& cov3D_precomp,
                              const Tensor& viewmatrix, const Tensor& projmatrix, double tanfovx, double tanfovy, int64_t H,
                              int64_t W, const Tensor& sh, int64_t degree, const Tensor& campos, bool debug, const Tensor& geom,
                              const OptT& screen_grads, const std::vector<int64_t>& peer_ptrs,
                              const std::vector<int64_t>& peer_rows, bool contrib_flags) {
    Tensor bg = at::zeros({3}, means3D.options());
    Inputs in = make_inputs(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                            tanfovx, tanfovy, H, W, sh, degree, campos, false, debug, 0, 0, true, 0);
    const c10::Device dev = in.means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t N = in.means3D.size(0);
    TORCH_CHECK(radii.defined() && radii.scalar_type() == at::kInt && radii.device() == dev && radii.numel() == N, "radii must be int32 [N] on ", dev);
    Tensor rd = radii.contiguous();
    GradOut go = make_grads(in);
    TgsSaved saved{};
    TORCH_CHECK(geom.defined() && geom.is_cuda(), "saved geometry buffer missing");
    saved.geom = geom.data_ptr();
    in.s.contrib_flags = contrib_flags ? 1 : 0;
    if (peer_ptrs.empty()) {
        TORCH_CHECK(screen_grads.has_value() && screen_grads->defined(), "screen_grads required");
        Tensor sg = f32(screen_grads->reshape({-1}), "screen_grads", {-1}, dev);
        TORCH_CHECK(sg.numel() >= screen_grad_floats(N, contrib_flags), "screen_grads must hold screen_grad_floats(N, contrib_flags) floats");
        check_rc(tgs_backward_preprocess(&in.s, &in.g, &saved, rd.data_ptr<int32_t>(), fp(sg), &go.g, cur_stream(dev)),
                 "tgs_backward_preprocess");
    } else {
        const int world = (int)peer_ptrs.size();
        TORCH_CHECK(world <= TGS_MAX_PEERS && (int)peer_rows.size() == 2 * world, "peer_ptrs / peer_rows: up to ", TGS_MAX_PEERS,
                    " ranks with one (row_begin, row_end) pair each");
        const float* ptrs[TGS_MAX_PEERS];
        int32_t rows[2 * TGS_MAX_PEERS];
        for (int r = 0; r < world; ++r) {
            ptrs[r] = reinterpret_cast<const float*>((uintptr_t)peer_ptrs[r]);
            rows[2 * r] = (int32_t)peer_rows[2 * r]; rows[2 * r + 1] = (int32_t)peer_rows[2 * r + 1];
        }
        check_rc(tgs_backward_preprocess_gather(&in.s, &in.g, &saved, rd.data_ptr<int32_t>(), ptrs, rows, world, &go.g,
                                                cur_stream(dev)), "tgs_backward_preprocess_gather");
    }
    return grad_tuple(go);
}

// RasterizeGaussiansBackwardCUDA: both halves on one GPU
GradTuple rasterize_gaussians_backward(const Tensor& bg, const Tensor& means3D, const Tensor& radii, const Tensor& colors,
                                       const Tensor& opacity, const Tensor& scales, const Tensor& rotations, double scale_modifier,
                                       const Tensor& cov3D_precomp, const Tensor& viewmatrix, const Tensor& projmatrix,
                                       double tanfovx, double tanfovy, const Tensor& dL_dcolor, const OptT& dL_ddepth,
                                       const OptT& dL_dalpha, const Tensor& sh, int64_t degree, const Tensor& campos,
                                       const Tensor& geom, int64_t num_rendered, const Tensor& binning, const Tensor& img,
                                       int64_t capacity, bool debug, int64_t tile_row_begin, int64_t tile_row_end,
                                       bool depth_normalize, const OptT& touch_depth, const OptT& touch_weight,
                                       int64_t loss_mode, const OptT& loss_scale, const OptT& grad_scale,
                                       int64_t touch_row_begin, int64_t touch_row_end) {
    TORCH_CHECK(dL_dcolor.defined() && dL_dcolor.dim() == 3, "dL_dcolor must be [3,H,W]");
    const int64_t H = dL_dcolor.size(1), W = dL_dcolor.size(2);
    Inputs in = make_inputs(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                            projmatrix, tanfovx, tanfovy, H, W, sh, degree, campos, false, debug, tile_row_begin,
                            tile_row_end, depth_normalize, 0);
    const c10::Device dev = in.means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t N = in.means3D.size(0);
    TORCH_CHECK(radii.defined() && radii.scalar_type() == at::kInt && radii.device() == dev && radii.numel() == N, "radii must be int32 [N] on ", dev);
    Tensor rd = radii.contiguous();
    Tensor gc = f32(dL_dcolor, "grad_color", {3, H, W}, dev);
    Tensor gd, ga;
    if (dL_ddepth.has_value() && dL_ddepth->defined()) gd = f32(dL_ddepth->reshape({H, W}), "grad_depth", {H, W}, dev);
    if (dL_dalpha.has_value() && dL_dalpha->defined()) ga = f32(dL_dalpha->reshape({H, W}), "grad_alpha", {H, W}, dev);
    Touch th = make_touch(touch_depth, touch_weight, loss_mode, loss_scale, grad_scale, touch_row_begin, touch_row_end, H, W, dev);
    TgsSaved saved = make_saved(geom, binning, img, num_rendered, capacity);
    GradOut go = make_grads(in);
    // plain rows: on one GPU the chain rule finds the Gaussians no pixel blended from the zero rows themselves, which is
    // cheaper than having BACKWARD::render write contributor bytes (they pay off in the multi-GPU gather)
    Tensor sgrad = at::empty({N > 0 ? N : 1, 10}, at::TensorOptions().dtype(at::kFloat).device(dev));
    check_rc(tgs_backward(&in.s, &in.g, &saved, rd.data_ptr<int32_t>(), gc.data_ptr<float>(), fp(gd), fp(ga),
                          th.on ? &th.t : nullptr, nullptr, sgrad.data_ptr<float>(), &go.g, cur_stream(dev)), "tgs_backward");
    return grad_tuple(go);
}

Tensor mark_visible(const Tensor& means3D, const Tensor& viewmatrix, const Tensor& projmatrix) {
    TORCH_CHECK(means3D.defined() && means3D.dim() == 2 && means3D.size(1) == 3, "means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D.is_cuda(), "touchgs_b200 rasterizer is CUDA-only (no CPU fallback)");
    const c10::Device dev = means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t N = means3D.size(0);
    Tensor m = f32(means3D, "means3D", {N, 3}, dev);
    Tensor v = f32(viewmatrix, "viewmatrix", {4, 4}, dev);
    (void)projmatrix;                                   // kept for signature parity with the reference-era binding
    Tensor present_u8 = at::zeros({N}, at::TensorOptions().dtype(at::kByte).device(dev));
    check_rc(tgs_mark_visible((int32_t)N, fp(m), fp(v), N > 0 ? present_u8.data_ptr<uint8_t>() : nullptr, cur_stream(dev)),
             "tgs_mark_visible");
    return present_u8.to(at::kBool);
}

// scale[0] = mult / Z on the device (Z = #(target > 0) or `norm`); scale is a 2-float workspace tensor
Tensor touch_loss_scale(const Tensor& touch_depth, double mult, double norm) {
    TORCH_CHECK(touch_depth.defined() && touch_depth.is_cuda(), "touch_depth must be a CUDA tensor");
    const c10::Device dev = touch_depth.device();
    c10::cuda::CUDAGuard guard(dev);
    Tensor t = f32(touch_depth.reshape({-1}), "touch_depth", {-1}, dev);
    Tensor scale = at::empty({2}, t.options());
    check_rc(tgs_touch_loss_scale(fp(t), t.numel(), (float)mult, (float)norm, scale.data_ptr<float>(), cur_stream(dev)),
             "tgs_touch_loss_scale");
    return scale;
}

Tensor touch_loss_value(const Tensor& residual, const OptT& weight, int64_t H, int64_t W, int64_t row0, int64_t row1,
                        int64_t mode, const Tensor& scale) {
    const c10::Device dev = residual.device();
    c10::cuda::CUDAGuard guard(dev);
    Tensor r = f32(residual.reshape({H, W}), "residual", {H, W}, dev), w;
    if (weight.has_value() && weight->defined()) w = f32(weight->reshape({H, W}), "touch_weight", {H, W}, dev);
    Tensor sc = f32(scale.reshape({-1}), "loss_scale", {-1}, dev);
    Tensor acc = at::empty({1}, r.options().dtype(at::kDouble));
    Tensor out = at::empty({}, r.options());
    check_rc(tgs_touch_loss_value(fp(r), fp(w), (int32_t)W, (int32_t)H, (int32_t)row0, (int32_t)row1, (int32_t)mode, fp(sc),
                                  acc.data_ptr<double>(), out.data_ptr<float>(), cur_stream(dev)), "tgs_touch_loss_value");
    return out;
}

// (1-l) * mean|C - C*| + l * (1 - mean SSIM) on the loss rows; returns (loss scalar, derivative maps for the backward)
std::tuple<Tensor, Tensor> photometric_loss_forward(const Tensor& color, const Tensor& gt, int64_t row_begin, int64_t row_end,
                                                    double lambda_dssim) {
    TORCH_CHECK(color.defined() && color.is_cuda(), "color: touchgs_b200 train ops are CUDA-only (no CPU fallback)");
    const c10::Device dev = color.device();
    c10::cuda::CUDAGuard guard(dev);
    TORCH_CHECK(color.dim() == 3 && color.size(0) == 3 && gt.sizes() == color.sizes(), "color / gt must both be [3,H,W], got ",
                color.sizes(), " / ", gt.sizes());
    const int64_t H = color.size(1), W = color.size(2);
    Tensor c = f32(color, "color", {3, H, W}, dev), g = f32(gt, "gt", {3, H, W}, dev);
    auto fo = at::TensorOptions().dtype(at::kFloat).device(dev);
    Tensor dmaps = at::empty({(int64_t)tgs_photometric_scratch_floats((int32_t)W, (int32_t)H)}, fo);
    Tensor sums = at::empty({2}, fo.dtype(at::kDouble));
    Tensor loss = at::empty({}, fo);
    check_rc(tgs_photometric_loss_forward(fp(c), fp(g), (int32_t)W, (int32_t)H, (int32_t)row_begin, (int32_t)row_end,
                                          (float)lambda_dssim, dmaps.data_ptr<float>(), sums.data_ptr<double>(),
                                          loss.data_ptr<float>(), cur_stream(dev)), "tgs_photometric_loss_forward");
    return std::make_tuple(loss, dmaps);
}

Tensor photometric_loss_backward(const Tensor& color, const Tensor& gt, const Tensor& dmaps, int64_t row_begin, int64_t row_end,
                                 int64_t out_row_begin, int64_t out_row_end, double lambda_dssim, const Tensor& grad_out) {
    const c10::Device dev = color.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t H = color.size(1), W = color.size(2);
    Tensor c = f32(color, "color", {3, H, W}, dev), g = f32(gt, "gt", {3, H, W}, dev);
    Tensor go = f32(grad_out.reshape({-1}), "grad_out", {-1}, dev);
    const bool full = (out_row_begin == 0 && out_row_end == H);
    Tensor dcolor = full ? at::empty_like(c) : at::zeros_like(c);
    check_rc(tgs_photometric_loss_backward(fp(c), fp(g), fp(dmaps), (int32_t)W, (int32_t)H, (int32_t)row_begin, (int32_t)row_end,
                                           (int32_t)out_row_begin, (int32_t)out_row_end, (float)lambda_dssim, fp(go),
                                           dcolor.data_ptr<float>(), cur_stream(dev)), "tgs_photometric_loss_backward");
    return dcolor;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "touchgs_b200._C: PyTorch binding of libtgs.so (B200-native Touch-GS rasterizer)";
    m.def("rasterize_gaussians", &rasterize_gaussians);
    m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward);
    m.def("backward_render", &backward_render);
    m.def("backward_preprocess", &backward_preprocess);
    m.def("screen_grad_floats", &screen_grad_floats);
    m.def("mark_visible", &mark_visible);
    m.def("resolve_count", [](int64_t ticket, int64_t capacity) {
        int64_t n = 0;
        check_rc(tgs_forward_resolve(ticket, capacity, &n), "tgs_forward_resolve");
        return n;
    });
    m.def("touch_loss_scale", &touch_loss_scale);
    m.def("touch_loss_value", &touch_loss_value);
    m.def("photometric_loss_forward", &photometric_loss_forward);
    m.def("photometric_loss_backward", &photometric_loss_backward);
    m.def("abi_version", []() { return tgs_abi_version(); });
}
