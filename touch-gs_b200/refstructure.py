"""The "reference-structure CUDA" comparison arm as an autograd operator (BASELINE.md §3 column 2).

The reference's own CUDA rasterizer is not in its tree (reference ``.gitmodules:7-9`` -> empty submodule;
SURVEY.md §0), so the comparison column of every table is the same algorithm in the UPSTREAM kernels' structure
(``csrc/refstructure.cu``): id-order scan + blocking count read, per-Gaussian 64-bit key emission, one 12-byte-pair
radix sort, one-pixel-per-thread compositing without culling, per-thread atomics in backward, and the touch-depth
loss computed OUTSIDE the kernels in PyTorch (``touch_depth_loss_unfused``), the way the fork's model does
(``depth-loss-mult`` knob: reference ``scripts/train_block_data.sh:50``).

Never used by ``GaussianRasterizer``: a measurement and full-size cross-checking arm only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .rasterizer import (GaussianRasterizationSettings, TouchOptions, _Scratch, _chk, _make_gaussians,
                         _make_settings, _ptr, _stream_ptr)


class _RefStructureRasterize(torch.autograd.Function):
    """(means3D, opacities, shs, scales, rotations, settings) -> (color [3,H,W], radii, depth_raw [1,H,W],
    alpha [1,H,W]); depth_raw = sum depth*alpha*T (upstream depth forks return the raw sum)."""

    @staticmethod
    def forward(ctx, means3D, opacities, shs, scales, rotations, rs: GaussianRasterizationSettings):
        lib = L.load()
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("refstructure arm is CUDA-only")
        N, K = int(means3D.shape[0]), int(shs.shape[1])
        H, W = int(rs.image_height), int(rs.image_width)
        means3D = _chk(means3D, "means3D", (N, 3), dev)
        opacity_shape = tuple(opacities.shape)
        opacities = _chk(opacities.reshape(-1), "opacities", (N,), dev)
        shs = _chk(shs, "shs", (N, K, 3), dev)
        scales = _chk(scales, "scales", (N, 3), dev)
        rotations = _chk(rotations, "rotations", (N, 4), dev)
        keep = []
        with torch.cuda.device(dev):
            s, _ = _make_settings(rs, TouchOptions(), K, keep)
            g = _make_gaussians(means3D, opacities, shs, None, scales, rotations, None)
            color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            alpha = torch.empty((1, H, W), dtype=torch.float32, device=dev)
            radii = torch.zeros((N,), dtype=torch.int32, device=dev)
            scratch = _Scratch(dev)
            saved = L.TgsSaved()
            rc = lib.tgs_refstructure_forward(C.byref(s), C.byref(g), scratch.cb, None, _ptr(color), _ptr(depth),
                                              _ptr(alpha), _ptr(radii), C.byref(saved), _stream_ptr(dev))
            scratch.disarm()
            if scratch.error is not None:
                raise scratch.error
            L.check(rc, "tgs_refstructure_forward")
        ctx.rs, ctx.K, ctx.opacity_shape = rs, K, opacity_shape
        ctx.num_rendered = int(saved.num_rendered)
        ctx.save_for_backward(means3D, opacities, shs, scales, rotations, radii, scratch.bufs[L.BUF_GEOM],
                              scratch.bufs[L.BUF_BINNING], scratch.bufs[L.BUF_IMAGE])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth, g_alpha):
        lib = L.load()
        means3D, opacities, shs, scales, rots, radii, geom, binning, image = ctx.saved_tensors
        rs, K = ctx.rs, ctx.K
        dev = means3D.device
        N = int(means3D.shape[0])
        H, W = int(rs.image_height), int(rs.image_width)
        keep = []
        with torch.cuda.device(dev):
            s, _ = _make_settings(rs, TouchOptions(), K, keep)
            g = _make_gaussians(means3D, opacities, shs, None, scales, rots, None)
            saved = L.TgsSaved(geom=geom.data_ptr(), binning=binning.data_ptr(), image=image.data_ptr(),
                               num_rendered=ctx.num_rendered, capacity=ctx.num_rendered)
            g_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev) if g_color is None \
                else _chk(g_color, "grad_color", (3, H, W), dev)
            g_depth = None if g_depth is None else _chk(g_depth.reshape(H, W), "grad_depth", (H, W), dev)
            g_alpha = None if g_alpha is None else _chk(g_alpha.reshape(H, W), "grad_alpha", (H, W), dev)
            sgrad = torch.empty((N, L.NGRAD), dtype=torch.float32, device=dev)
            L.check(lib.tgs_refstructure_backward_render(C.byref(s), C.byref(g), C.byref(saved), _ptr(g_color),
                                                         _ptr(g_depth), _ptr(g_alpha), _ptr(sgrad), _stream_ptr(dev)),
                    "tgs_refstructure_backward_render")
            dmeans2D = torch.empty((N, 3), dtype=torch.float32, device=dev)
            dmeans3D = torch.empty((N, 3), dtype=torch.float32, device=dev)
            dopac = torch.empty((N,), dtype=torch.float32, device=dev)
            dsh = torch.empty((N, K, 3), dtype=torch.float32, device=dev)
            dsc = torch.empty((N, 3), dtype=torch.float32, device=dev)
            drot = torch.empty((N, 4), dtype=torch.float32, device=dev)
            gr = L.TgsGrads(dmeans2D=dmeans2D.data_ptr(), dmeans3D=dmeans3D.data_ptr(), dopacity=dopac.data_ptr(),
                            dshs=dsh.data_ptr(), dcolors=None, dscales=dsc.data_ptr(), drotations=drot.data_ptr(),
                            dcov3D=None)
            L.check(lib.tgs_backward_preprocess(C.byref(s), C.byref(g), C.byref(saved), _ptr(radii), _ptr(sgrad),
                                                C.byref(gr), _stream_ptr(dev)), "tgs_backward_preprocess")
        ctx.sgrad = sgrad
        return dmeans3D, dopac.reshape(ctx.opacity_shape), dsh, dsc, drot, None


def rasterize_refstructure(means3D, opacities, shs, scales, rotations, raster_settings):
    return _RefStructureRasterize.apply(means3D, opacities, shs, scales, rotations, raster_settings)


def touch_depth_loss_unfused(depth_raw, alpha, touch_depth, touch_weight, mult: float, mode: str = "l1"):
    """The touch-depth loss as PLAIN PyTorch ops on the rendered images -- what the fused backward replaces.
    Same definition as DESIGN.md §2: valid = (target > 0) & (alpha > 0); expected depth = D / alpha;
    loss = mult / #(target > 0) * sum valid * w * |r|  (l1)  or  r^2  (l2)."""
    d, a = depth_raw.reshape(-1), alpha.reshape(-1)
    t = touch_depth.reshape(-1)
    w = torch.ones_like(t) if touch_weight is None else touch_weight.reshape(-1)
    valid = (t > 0) & (a > 0)
    dhat = d / torch.where(valid, a, torch.ones_like(a))
    r = torch.where(valid, dhat - t, torch.zeros_like(t))
    z = (t > 0).sum().clamp_min(1).to(torch.float32)
    per = r.abs() if mode == "l1" else r * r
    return (mult / z) * (w * per).sum()


def forward_state(means3D, opacities, shs, scales, rotations, rs: GaussianRasterizationSettings):
    """Forward of the reference-structure arm (no grad) + its decoded internal state, for the cross-checks:
    64-bit sorted keys, sorted Gaussian ids, tile ranges, final_T, n_contrib."""
    lib = L.load()

    class Ctx:
        def save_for_backward(self, *t): self.saved = t
        def set_materialize_grads(self, v): pass
        def mark_non_differentiable(self, *t): pass

    ctx = Ctx()
    with torch.no_grad():
        color, radii, depth, alpha = _RefStructureRasterize.forward(ctx, means3D, opacities, shs, scales, rotations, rs)
    binning, image = ctx.saved[-2], ctx.saved[-1]
    N, I = int(means3D.shape[0]), ctx.num_rendered
    H, W = int(rs.image_height), int(rs.image_width)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    bl, il = L.TgsRefBinningLayout(), L.TgsImageLayout()
    lib.tgs_refstructure_binning_layout(N, I, T, C.byref(bl))
    lib.tgs_image_layout(W, H, C.byref(il))

    def view(buf, off, count, dtype):
        return buf[off:off + count * torch.empty((), dtype=dtype).element_size()].view(dtype)

    return dict(color=color, radii=radii, depth_raw=depth, alpha=alpha, num_rendered=I,
                keys=view(binning, bl.keys_sorted, I, torch.int64), vals=view(binning, bl.vals_sorted, I, torch.int32),
                keys_emitted=view(binning, bl.keys_unsorted, I, torch.int64),
                vals_emitted=view(binning, bl.vals_unsorted, I, torch.int32),
                ranges=view(binning, bl.ranges, 2 * T, torch.int32).view(T, 2),
                final_T=view(image, il.final_T, H * W, torch.float32).view(H, W),
                n_contrib=view(image, il.n_contrib, H * W, torch.int32).view(H, W))
