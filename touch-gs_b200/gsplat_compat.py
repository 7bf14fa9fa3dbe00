"""gsplat-0.1-style three-call surface on the B200 kernels.

The nerfstudio splat model of early 2024 -- the code base the Touch-GS fork's ``depth-gaussian-splatting`` method
extends (reference ``.gitmodules:7-9`` -> empty submodule; trainer entry reference ``scripts/train_bunny_real.sh:52``)
-- does not call one fused operator: it calls ``project_gaussians`` -> its own colour code (``spherical_harmonics``,
+0.5, clamp) -> ``rasterize_gaussians`` (once for RGB, once more with depth as the colour), each an autograd op.  This
module serves those three names, with that argument order, from ``libtgs.so`` so that such a model drops in unchanged.

Conventions (all [NIT] -- the fork's pinned gsplat version is not in the reference tree, so
they are module-level switches, defaults = later 0.1.x):

* ``viewmat`` / ``projmat`` are world->camera / full projection in COLUMN-vector convention (not transposed);
* near plane = ``clip_thresh`` (0.01); quaternions are normalised inside; ``glob_scale`` multiplies the scales;
* pixel mean = 0.5*W*ndc.x + cx - 0.5 (principal point ``cx, cy``); a pixel is sampled at its centre (+0.5);
* alpha clamp 0.999; images are HWC; SH returns the raw sum (the caller adds 0.5 and clamps);
* gradients w.r.t. ``xys`` are in pixels.
Tile rectangles (``num_tiles_hit``) follow OUR rule (SURVEY A1 ``getRect`` on the un-offset pixel mean), which may
differ from gsplat's by a border tile; images and gradients do not depend on it beyond the 3-sigma cut-off itself.

All compute is CUDA (no CPU fallback); torch is memory / streams / autograd plumbing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .rasterizer import _Scratch, _ptr, _stream_ptr

PIXEL_CENTER_OFFSET = 0.5      # 0.0 for the early 0.1.x releases that sampled at integer (x, y)
ALPHA_MAX = 0.999


def _f32(t, name, shape=None):
    if t.device.type != "cuda":
        raise RuntimeError(f"{name}: touchgs_b200 is CUDA-only (no CPU fallback); tensor is on {t.device}")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {list(shape)}, got {list(t.shape)}")
    return t.contiguous()


def _mat4(m, name, dev):
    m = m.to(device=dev, dtype=torch.float32)
    if m.shape == (3, 4):
        m = torch.cat([m, torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=dev)], 0)
    if m.shape != (4, 4):
        raise ValueError(f"{name} must be [4,4] (or [3,4]), got {list(m.shape)}")
    return m.t().contiguous()          # the C ABI reads matrices TRANSPOSED (row-vector convention)


def _settings(H, W, keep, *, dev, fx=None, fy=None, cx=None, cy=None, view=None, proj=None, glob_scale=1.0,
              clip_thresh=0.0, bg=None):
    zero3 = torch.zeros(3, dtype=torch.float32, device=dev)
    bg = zero3 if bg is None else bg
    keep.extend([t for t in (view, proj, bg, zero3) if t is not None])
    return L.TgsSettings(
        image_width=int(W), image_height=int(H),
        tanfovx=1.0 if fx is None else float(W) / (2.0 * float(fx)),
        tanfovy=1.0 if fy is None else float(H) / (2.0 * float(fy)),
        scale_modifier=float(glob_scale), sh_degree=0, sh_coeffs=0, prefiltered=0, debug=0,
        tile_row_begin=0, tile_row_end=0, depth_normalize=0, defer_count=0, rendered_hint=0,
        viewmatrix=None if view is None else view.data_ptr(), projmatrix=None if proj is None else proj.data_ptr(),
        campos=zero3.data_ptr(), bg=bg.data_ptr(), alpha_max=ALPHA_MAX, near_z=float(clip_thresh),
        principal_dx=0.0 if cx is None else float(cx) - 0.5 * float(W),
        principal_dy=0.0 if cy is None else float(cy) - 0.5 * float(H))


class _ProjectGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3d, scales, rot, glob_scale, viewmat, projmat, fx, fy, cx, cy, H, W, clip_thresh):
        lib = L.load()
        N = int(means3d.shape[0])
        dev = means3d.device
        means3d, scales, rot = _f32(means3d, "means3d", (N, 3)), _f32(scales, "scales", (N, 3)), _f32(rot, "quats", (N, 4))
        keep = []
        with torch.cuda.device(dev):
            s = _settings(H, W, keep, dev=dev, fx=fx, fy=fy, cx=cx, cy=cy, view=_mat4(viewmat, "viewmat", dev),
                          proj=_mat4(projmat, "projmat", dev), glob_scale=glob_scale, clip_thresh=clip_thresh)
            g = L.TgsGaussians(N=N, means3D=means3d.data_ptr(), opacities=None, shs=None, colors_precomp=None,
                               scales=scales.data_ptr(), rotations=rot.data_ptr(), cov3D_precomp=None)
            radii = torch.zeros(N, dtype=torch.int32, device=dev)
            scratch = _Scratch(dev)
            saved = L.TgsSaved()
            rc = lib.tgs_project_gaussians(C.byref(s), C.byref(g), scratch.cb, None, _ptr(radii), C.byref(saved), _stream_ptr(dev))
            scratch.disarm()
            if scratch.error is not None:
                raise scratch.error
            L.check(rc, "tgs_project_gaussians")
            geom = scratch.bufs[L.BUF_GEOM]
            gl = L.TgsGeomLayout()
            lib.tgs_geom_layout(N, C.byref(gl))
            rec = geom[gl.records:gl.records + N * 48].view(torch.float32).view(N, 12)
            xys, depths, conics = rec[:, 0:2].clone(), rec[:, 2].clone(), rec[:, 4:7].clone()
            cov3d = geom[gl.cov3D:gl.cov3D + N * 24].view(torch.float32).view(N, 6).clone()
            tiles = geom[gl.tiles_touched:gl.tiles_touched + N * 4].view(torch.int32).clone()
        ctx.settings, ctx.keep = s, keep
        ctx.save_for_backward(means3d, scales, rot, radii, geom)
        ctx.mark_non_differentiable(radii, tiles)
        ctx.set_materialize_grads(False)
        return xys, depths, radii, conics, tiles, cov3d

    @staticmethod
    def backward(ctx, v_xys, v_depths, _vr, v_conics, _vt, _vc):
        lib = L.load()
        means3d, scales, rot, radii, geom = ctx.saved_tensors
        N, dev = int(means3d.shape[0]), means3d.device
        with torch.cuda.device(dev):
            sg = torch.zeros((max(N, 1), L.NGRAD), dtype=torch.float32, device=dev)
            if v_xys is not None:
                sg[:N, 0:2] = v_xys
            if v_conics is not None:
                sg[:N, 2:5] = v_conics
            if v_depths is not None:
                sg[:N, 9] = v_depths
            g = L.TgsGaussians(N=N, means3D=means3d.data_ptr(), opacities=None, shs=None, colors_precomp=None,
                               scales=scales.data_ptr(), rotations=rot.data_ptr(), cov3D_precomp=None)
            saved = L.TgsSaved(geom=geom.data_ptr(), binning=None, image=None, num_rendered=0, capacity=0)
            d2, dm = torch.empty((N, 3), device=dev), torch.empty((N, 3), device=dev)
            do, ds, dr = torch.empty(N, device=dev), torch.empty((N, 3), device=dev), torch.empty((N, 4), device=dev)
            gr = L.TgsGrads(dmeans2D=d2.data_ptr(), dmeans3D=dm.data_ptr(), dopacity=do.data_ptr(), dshs=None, dcolors=None,
                            dscales=ds.data_ptr(), drotations=dr.data_ptr(), dcov3D=None)
            L.check(lib.tgs_project_gaussians_backward(C.byref(ctx.settings), C.byref(g), C.byref(saved), _ptr(radii), _ptr(sg),
                                                       C.byref(gr), _stream_ptr(dev)), "tgs_project_gaussians_backward")
        return (dm, ds, dr) + (None,) * 10


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                      tile_bounds=None, clip_thresh: float = 0.01):
    """-> (xys [N,2], depths [N], radii [N] int32, conics [N,3], num_tiles_hit [N] int32, cov3d [N,6]).
    ``tile_bounds`` is accepted for signature compatibility and ignored (derived from the image size)."""
    rot = quats / quats.norm(dim=-1, keepdim=True)          # normalised inside, differentiable
    return _ProjectGaussians.apply(means3d, scales, rot, float(glob_scale), viewmat, projmat, fx, fy, cx, cy,
                                   int(img_height), int(img_width), float(clip_thresh))


class _RasterizeScreen(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xys, depths, radii, conics, colors3, opacity, H, W, bg3):
        lib = L.load()
        N, dev = int(xys.shape[0]), xys.device
        xys, depths, conics = _f32(xys, "xys", (N, 2)), _f32(depths, "depths", (N,)), _f32(conics, "conics", (N, 3))
        colors3, opacity = _f32(colors3, "colors", (N, 3)), _f32(opacity.reshape(-1), "opacity", (N,))
        radii = radii.to(torch.int32).contiguous()
        keep = []
        with torch.cuda.device(dev):
            s = _settings(H, W, keep, dev=dev, bg=_f32(bg3, "background", (3,)))
            color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
            depth = torch.empty((H, W), dtype=torch.float32, device=dev)
            alpha = torch.empty((H, W), dtype=torch.float32, device=dev)
            scratch = _Scratch(dev)
            saved = L.TgsSaved()
            rc = lib.tgs_rasterize_screen_forward(C.byref(s), N, _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics),
                                                  _ptr(colors3), _ptr(opacity), float(PIXEL_CENTER_OFFSET), scratch.cb, None,
                                                  _ptr(color), _ptr(depth), _ptr(alpha), C.byref(saved), _stream_ptr(dev))
            scratch.disarm()
            if scratch.error is not None:
                raise scratch.error
            L.check(rc, "tgs_rasterize_screen_forward")
        ctx.settings, ctx.keep, ctx.N, ctx.hw = s, keep, N, (H, W)
        ctx.num_rendered = int(saved.num_rendered)
        ctx.save_for_backward(scratch.bufs[L.BUF_GEOM], scratch.bufs[L.BUF_BINNING], scratch.bufs[L.BUF_IMAGE])
        ctx.set_materialize_grads(False)
        return color, alpha

    @staticmethod
    def backward(ctx, g_color, g_alpha):
        lib = L.load()
        geom, binning, image = ctx.saved_tensors
        N, (H, W) = ctx.N, ctx.hw
        dev = geom.device
        with torch.cuda.device(dev):
            g_color = torch.zeros((3, H, W), device=dev) if g_color is None else g_color.contiguous()
            g_alpha = None if g_alpha is None else g_alpha.contiguous()
            sg = torch.empty((max(N, 1), L.NGRAD), dtype=torch.float32, device=dev)
            saved = L.TgsSaved(geom=geom.data_ptr(), binning=binning.data_ptr(), image=image.data_ptr(),
                               num_rendered=ctx.num_rendered, capacity=ctx.num_rendered)
            L.check(lib.tgs_rasterize_screen_backward(C.byref(ctx.settings), N, C.byref(saved), _ptr(g_color), None,
                                                      _ptr(g_alpha), _ptr(sg), _stream_ptr(dev)), "tgs_rasterize_screen_backward")
            sg = sg[:N]
        return sg[:, 0:2], None, None, sg[:, 2:5], sg[:, 6:9], sg[:, 5], None, None, None


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                        background: Optional[torch.Tensor] = None, return_alpha: bool = False):
    """-> out_img [H,W,C] (and out_alpha [H,W] when ``return_alpha``).  ``colors`` [N,C] with C <= 3."""
    if colors.dim() != 2 or not 1 <= colors.shape[1] <= 3:
        raise ValueError(f"colors must be [N,C] with 1 <= C <= 3, got {list(colors.shape)}")
    Cn = int(colors.shape[1])
    dev = colors.device
    if background is None:
        background = torch.ones(Cn, dtype=torch.float32, device=dev)
    if background.numel() != Cn:
        raise ValueError(f"background must have {Cn} elements")
    pad = 3 - Cn
    c3 = colors if pad == 0 else torch.cat([colors, colors.new_zeros(colors.shape[0], pad)], 1)
    bg3 = background.to(dev).float() if pad == 0 else torch.cat([background.to(dev).float(), torch.zeros(pad, device=dev)])
    opacity_flat = opacity.reshape(-1)
    color, alpha = _RasterizeScreen.apply(xys, depths, radii, conics, c3, opacity_flat, int(img_height), int(img_width), bg3)
    img = color[:Cn].permute(1, 2, 0)
    return (img, alpha) if return_alpha else img


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree, viewdirs, coeffs):
        lib = L.load()
        N, K = int(coeffs.shape[0]), int(coeffs.shape[1])
        dev = coeffs.device
        viewdirs, coeffs = _f32(viewdirs, "viewdirs", (N, 3)), _f32(coeffs, "coeffs", (N, K, 3))
        with torch.cuda.device(dev):
            out = torch.empty((N, 3), dtype=torch.float32, device=dev)
            L.check(lib.tgs_spherical_harmonics(N, int(degree), K, _ptr(viewdirs), _ptr(coeffs), _ptr(out), _stream_ptr(dev)),
                    "tgs_spherical_harmonics")
        ctx.save_for_backward(viewdirs)
        ctx.args = (int(degree), K)
        return out

    @staticmethod
    def backward(ctx, v):
        lib = L.load()
        (viewdirs,) = ctx.saved_tensors
        degree, K = ctx.args
        N, dev = int(viewdirs.shape[0]), viewdirs.device
        with torch.cuda.device(dev):
            vc = torch.empty((N, K, 3), dtype=torch.float32, device=dev)
            L.check(lib.tgs_spherical_harmonics_backward(N, degree, K, _ptr(viewdirs), _ptr(v.contiguous()), _ptr(vc),
                                                         _stream_ptr(dev)), "tgs_spherical_harmonics_backward")
        return None, None, vc


def spherical_harmonics(degrees_to_use: int, viewdirs, coeffs):
    """Raw SH colour sum over the (degrees_to_use+1)^2 active bases; gradient w.r.t. ``coeffs`` only (as the
    0.1.x op).  The caller adds 0.5 and clamps."""
    return _SphericalHarmonics.apply(int(degrees_to_use), viewdirs, coeffs)
