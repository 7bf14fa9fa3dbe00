"""The operator surface: ``GaussianRasterizationSettings`` / ``GaussianRasterizer`` /
``rasterize_gaussians`` -> ``_RasterizeGaussians`` (a ``torch.autograd.Function``).

Mirrors the Inria-style operator the Touch-GS trainer calls once per step
(``ns-train depth-gaussian-splatting``, reference ``scripts/train_bunny_real.sh:52``; the
rasterizer itself is not vendored in the reference -- SURVEY.md §0, §8(b)), extended with

* rendered expected depth and alpha (one traversal with RGB), and
* the touch-depth loss whose gradient is FUSED into the backward kernel: pass
  ``touch_depth`` (metres, 0 = invalid: reference ``utils/fuse_touch_vision.py:372-388``),
  ``touch_weight`` (e.g. 1/sigma from the uncertainty PNG, reference
  ``utils/fuse_touch_vision.py:376,387``), ``depth_loss in {'none','l1','l2'}`` and
  ``depth_loss_mult`` (reference ``scripts/train_block_data.sh:50``: ``--pipeline.model.depth-loss-mult``).
  The caller must NOT add that loss term to its autograd graph again; ``depth_residual`` is returned
  (non-differentiable) so the loss value can still be logged.

All compute happens in ``libtgs.so`` (hand-written sm_100a CUDA, C ABI in ``include/tgs.h``).
PyTorch is used for device memory, streams and ``torch.distributed`` only.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import torch

from . import _lib as L

TILE = 16


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor      # [4,4] transposed world->view
    projmatrix: torch.Tensor      # [4,4] transposed full projection
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False     # accepted for signature parity: the reference-era kernels only use it to ASSERT that a
                                  # caller who pre-culled with markVisible passes no point behind the near plane; results
                                  # never depend on it, and here a culled point is simply skipped
    debug: bool = False


class TouchOptions(NamedTuple):
    """Keyword-only extension of the operator (default = off: the base call is unchanged)."""
    touch_depth: Optional[torch.Tensor] = None     # [H,W]
    touch_weight: Optional[torch.Tensor] = None    # [H,W]
    depth_loss: str = "none"                       # 'none' | 'l1' | 'l2'
    depth_loss_mult: float = 1.0
    depth_normalize: bool = True                   # returned depth = D/alpha
    depth_loss_norm: Optional[float] = None        # Z; None -> #(touch_depth > 0)
    tile_rows: Optional[Tuple[int, int]] = None    # tile-row band of this rank
    process_group: object = None                   # all-reduce group for the screen-space gradients
    rendered_hint: int = 0                         # > 0: speculative sizing (hides the forward's host sync)
    info: Optional[dict] = None                    # filled with num_rendered / capacity of the call
    touch_rows: Optional[Tuple[int, int]] = None   # pixel rows where the touch loss applies (None = all rows)
    peer_exchange: object = None                   # sharding.PeerScreenGrads: fused P2P gather instead of all-reduce
    return_touch_loss: bool = False                # True: the operator also returns the touch loss as a DIFFERENTIABLE
                                                   # scalar; the fused gradient is then scaled by that scalar's upstream
                                                   # gradient (0 if the caller leaves it out of the objective)
    defer_count: bool = False                      # with rendered_hint > 0: the forward never waits for num_rendered (the host
                                                   # stays a full step ahead of the GPU); the count is checked when the backward
                                                   # runs and an overflowed hint raises there (redo the step with a larger hint)
    loss_grad_scale: object = None                 # injected mode (return_touch_loss=False): explicit upstream scale of
                                                   # the objective (float or 0-dim tensor), e.g. a GradScaler's scale or
                                                   # 1/accumulation_steps; None = 1


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: Optional[torch.Tensor], name: str, shape, device, dtype=torch.float32):
    if t is None:
        return None
    if t.device != device:
        raise ValueError(f"{name} must be on {device}, got {t.device}")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None:
        if t.dim() != len(shape) or any(s is not None and int(t.shape[i]) != s for i, s in enumerate(shape)):
            raise ValueError(f"{name} must have shape {list(shape)}, got {list(t.shape)}")
    return t.contiguous()


import threading

_tls = threading.local()


def _thread_alloc():
    """ONE ctypes callback per thread (building a CFUNCTYPE object per call costs ~10 us).  The callback writes into
    the thread-local `state` dict that `_Scratch` re-arms for every call; it never references a `_Scratch` instance,
    so no reference cycle can keep ~1 GB of scratch per call alive until the cyclic GC runs."""
    cb = getattr(_tls, "cb", None)
    if cb is None:
        state = {"bufs": None, "err": None, "device": None}

        def _alloc(_user, which, nbytes, _state=state):
            try:
                t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=_state["device"])
                _state["bufs"][int(which)] = t
                return t.data_ptr()
            except Exception as e:  # noqa: BLE001 - must not propagate through the C frame
                _state["err"] = e
                return None

        _tls.state, _tls.cb = state, L.ALLOC_FN(_alloc)
        cb = _tls.cb
    return cb, _tls.state


class _Scratch:
    """Allocator handed to the C ABI: the three saved byte buffers are torch tensors, so they come from torch's
    caching allocator on the caller's device and are kept alive by ``ctx``."""

    def __init__(self, device):
        self.cb, self._state = _thread_alloc()
        self.bufs = {}
        self._state["bufs"], self._state["err"], self._state["device"] = self.bufs, None, device

    @property
    def error(self):
        return self._state["err"]

    def disarm(self):
        """Drop the callback's reference to this call's buffers (they now belong to ``self.bufs`` / ``ctx`` only)."""
        if self._state["bufs"] is self.bufs:
            self._state["bufs"] = None

    def __del__(self):
        self.disarm()


def _stream_ptr(device):
    """Raw cudaStream_t of torch's current stream on `device` (the C accessor: several times cheaper than building a
    torch.cuda.Stream object; this is called for every C-ABI entry point of every step)."""
    idx = device.index
    if idx is None:
        idx = torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


def _make_settings(rs: GaussianRasterizationSettings, opt: TouchOptions, K: int, keep):
    dev = rs.viewmatrix.device
    vm = _chk(rs.viewmatrix, "viewmatrix", (4, 4), dev)
    pmx = _chk(rs.projmatrix, "projmatrix", (4, 4), dev)
    cam = _chk(rs.campos.reshape(-1), "campos", (3,), dev)
    bg = _chk(rs.bg.reshape(-1), "bg", (3,), dev)
    keep.extend([vm, pmx, cam, bg])
    Ty = (rs.image_height + TILE - 1) // TILE
    r0, r1 = (0, Ty) if opt.tile_rows is None else (int(opt.tile_rows[0]), int(opt.tile_rows[1]))
    if not (0 <= r0 <= r1 <= Ty):
        raise ValueError(f"tile_rows {opt.tile_rows} outside [0, {Ty}]")
    s = L.TgsSettings(
        image_width=int(rs.image_width), image_height=int(rs.image_height),
        tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), scale_modifier=float(rs.scale_modifier),
        sh_degree=int(rs.sh_degree), sh_coeffs=int(K), prefiltered=int(bool(rs.prefiltered)),
        debug=int(bool(rs.debug)), tile_row_begin=r0, tile_row_end=r1,
        depth_normalize=int(bool(opt.depth_normalize)),
        defer_count=int(bool(opt.defer_count) and int(opt.rendered_hint or 0) > 0),
        rendered_hint=max(0, int(opt.rendered_hint or 0)),
        viewmatrix=vm.data_ptr(), projmatrix=pmx.data_ptr(), campos=cam.data_ptr(), bg=bg.data_ptr())
    return s, (r0, r1, Ty)


def _make_gaussians(means3D, opacities, sh, colors, scales, rots, cov3D):
    return L.TgsGaussians(
        N=int(means3D.shape[0]), means3D=means3D.data_ptr(), opacities=opacities.data_ptr(),
        shs=None if sh is None else sh.data_ptr(),
        colors_precomp=None if colors is None else colors.data_ptr(),
        scales=None if scales is None else scales.data_ptr(),
        rotations=None if rots is None else rots.data_ptr(),
        cov3D_precomp=None if cov3D is None else cov3D.data_ptr())


_EMPTY = {}


def _empty_on(device):
    """Cached zero-element placeholder for absent optional tensors (the reference-era module builds a CPU tensor and
    copies it to the device on every call)."""
    t = _EMPTY.get(device)
    if t is None:
        t = _EMPTY[device] = torch.empty(0, device=device)
    return t


_ZERO = {}


def _zero_scalar(device):
    t = _ZERO.get(device)
    if t is None:
        t = _ZERO[device] = torch.zeros((), dtype=torch.float32, device=device)
    return t


def _empty_if(t):
    """None for an absent optional tensor: None itself or the 1-D zero-element placeholder of the reference-era module
    (``torch.Tensor([])``).  The [0,K,3] tensors of an EMPTY scene are present."""
    return None if (t is None or (t.numel() == 0 and t.dim() <= 1)) else t


# ---------------------------------------------------------------------------------------------------------------------
# Default binding: the PyTorch C++ extension `_C` (csrc/torch_ext.cpp): one call per direction, argument checking
# (TORCH_CHECK), output / scratch allocation (at::empty), stream and device guard all happen in C++.
def _forward_ext(ext, ctx, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, opt):
    dev = means3D.device
    e = _empty_on(dev)
    sh, colors_precomp = _empty_if(sh), _empty_if(colors_precomp)
    scales, rotations, cov3Ds_precomp = _empty_if(scales), _empty_if(rotations), _empty_if(cov3Ds_precomp)
    H, W = int(rs.image_height), int(rs.image_width)
    Ty = (H + TILE - 1) // TILE
    r0, r1 = (0, Ty) if opt.tile_rows is None else (int(opt.tile_rows[0]), int(opt.tile_rows[1]))
    if not (0 <= r0 <= r1 <= Ty):
        raise ValueError(f"tile_rows {opt.tile_rows} outside [0, {Ty}]")
    touch_depth, touch_weight = opt.touch_depth, opt.touch_weight
    if opt.depth_loss != "none" and touch_depth is None:
        raise ValueError("depth_loss != 'none' requires touch_depth")
    if touch_weight is not None and touch_weight.numel() != H * W:
        raise ValueError(f"touch_weight must have H*W = {H * W} elements, got {list(touch_weight.shape)}")
    sh_ = e if sh is None else sh
    col_ = e if colors_precomp is None else colors_precomp
    sc_ = e if scales is None else scales
    rot_ = e if rotations is None else rotations
    cov_ = e if cov3Ds_precomp is None else cov3Ds_precomp
    args = (rs.bg, means3D, col_, opacities, sc_, rot_, float(rs.scale_modifier), cov_, rs.viewmatrix, rs.projmatrix,
            float(rs.tanfovx), float(rs.tanfovy), H, W, sh_, int(rs.sh_degree), rs.campos, bool(rs.prefiltered), bool(rs.debug),
            r0, r1, bool(opt.depth_normalize), max(0, int(opt.rendered_hint or 0)), touch_depth, bool(opt.defer_count))
    try:
        num_rendered, color, depth, alpha, radii, geom, binning, image, resid, capacity = ext.rasterize_gaussians(*args)
    except Exception:
        if rs.debug:
            # reference-era operator habit (SURVEY §8b "Errors"): with debug=True a failing forward leaves its arguments
            # behind for post-mortem inspection
            try:
                torch.save(dict(means3D=means3D, opacities=opacities, shs=sh, colors_precomp=colors_precomp, scales=scales,
                                rotations=rotations, cov3D_precomp=cov3Ds_precomp, settings=tuple(rs)), "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
            except Exception:  # noqa: BLE001 - the dump is best effort; the real error is raised below
                pass
        raise
    tscale, tloss = None, None
    if touch_depth is not None and opt.depth_loss != "none":
        norm = float(opt.depth_loss_norm) if opt.depth_loss_norm is not None else 0.0
        tscale = ext.touch_loss_scale(touch_depth, float(opt.depth_loss_mult), norm)
        if opt.return_touch_loss:
            full = (r0 == 0 and r1 == Ty)
            tr = (0, 0) if opt.touch_rows is None else (int(opt.touch_rows[0]), int(opt.touch_rows[1]))
            if opt.touch_rows is None and not full:
                tr = (min(r0 * TILE, H), min(r1 * TILE, H))
            tloss = ext.touch_loss_value(resid, touch_weight, H, W, tr[0], tr[1], L.LOSS_MODES[opt.depth_loss], tscale)
    ctx.ext = ext
    ctx.rs, ctx.opt, ctx.rows = rs, opt, (r0, r1)
    ctx.tscale = tscale
    ctx.num_rendered, ctx.capacity = int(num_rendered), int(capacity)
    ctx.opacity_shape = tuple(opacities.shape)
    if opt.info is not None:
        opt.info["num_rendered"] = ctx.num_rendered
        opt.info["capacity"] = ctx.capacity
    ctx.has = (sh is not None, colors_precomp is not None, scales is not None, cov3Ds_precomp is not None)
    ctx.touch = (touch_depth, touch_weight)
    ctx.save_for_backward(means3D, opacities, sh_, col_, sc_, rot_, cov_, radii, geom, binning, image)
    ctx.set_materialize_grads(False)
    if tloss is None:
        tloss = _zero_scalar(dev)
        ctx.mark_non_differentiable(radii, resid, tloss)
    else:
        ctx.mark_non_differentiable(radii, resid)
    return color, radii, depth, alpha, resid, tloss


def _backward_ext(ctx, g_color, g_depth, g_alpha, g_tloss):
    ext = ctx.ext
    (means3D, opacities, sh, colors, scales, rots, cov3D, radii, geom, binning, image) = ctx.saved_tensors
    rs, opt = ctx.rs, ctx.opt
    dev = means3D.device
    N = int(means3D.shape[0])
    H, W = int(rs.image_height), int(rs.image_width)
    r0, r1 = ctx.rows
    if g_color is None:
        g_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    touch_depth, touch_weight = ctx.touch
    mode, gs = L.LOSS_NONE, None
    if touch_depth is not None and opt.depth_loss != "none":
        # upstream gradient of the touch-loss scalar, applied on the device (see the ctypes path for the semantics)
        mode = L.LOSS_MODES[opt.depth_loss]
        if opt.return_touch_loss:
            if g_tloss is None:
                mode = L.LOSS_NONE
            else:
                gs = g_tloss.detach().reshape(1).to(device=dev, dtype=torch.float32)
        elif opt.loss_grad_scale is not None:
            gs = torch.as_tensor(opt.loss_grad_scale, dtype=torch.float32, device=dev).detach().reshape(1)
    tr = (0, 0) if opt.touch_rows is None else (int(opt.touch_rows[0]), int(opt.touch_rows[1]))
    cam = (float(rs.scale_modifier), cov3D, rs.viewmatrix, rs.projmatrix, float(rs.tanfovx), float(rs.tanfovy))
    peer = opt.peer_exchange if N > 0 else None
    if peer is not None and N > peer.capacity:
        if opt.process_group is None:
            raise ValueError(f"PeerScreenGrads capacity {peer.capacity} < {N} Gaussians and no process_group to fall back to")
        peer = None
    if peer is None and opt.process_group is None:
        # single GPU: RasterizeGaussiansBackwardCUDA in one call
        grads = ext.rasterize_gaussians_backward(
            rs.bg, means3D, radii, colors, opacities, scales, rots, *cam, g_color, g_depth, g_alpha, sh, int(rs.sh_degree),
            rs.campos, geom, ctx.num_rendered, binning, image, ctx.capacity, bool(rs.debug), r0, r1, bool(opt.depth_normalize),
            touch_depth if mode != L.LOSS_NONE else None, touch_weight, mode, ctx.tscale, gs, tr[0], tr[1])
    else:
        if peer is not None:
            sgrad, peer_ptrs, peer_handle = peer.acquire(N)         # this rank's peer-mapped buffer of the step
        else:
            sgrad = torch.empty((max(N, 1), L.NGRAD), dtype=torch.float32, device=dev)
        # peer buffers carry contributor bytes behind the rows (the gather asks a peer only for rows it wrote); a buffer
        # that goes through an all-reduce must be plain [N,10]
        flags = peer is not None
        ext.backward_render(rs.bg, means3D, colors, opacities, scales, rots, *cam, H, W, sh, int(rs.sh_degree), rs.campos,
                            bool(rs.debug), r0, r1, bool(opt.depth_normalize), geom, binning, image, ctx.num_rendered,
                            ctx.capacity, g_color, g_depth, g_alpha, touch_depth if mode != L.LOSS_NONE else None, touch_weight,
                            mode, ctx.tscale, gs, tr[0], tr[1], sgrad, flags)
        pre = (means3D, radii, colors, opacities, scales, rots, *cam, H, W, sh, int(rs.sh_degree), rs.campos, bool(rs.debug), geom)
        if peer is not None:
            # FUSED exchange: wait until every rank has finished BACKWARD::render, then the chain-rule
            # kernel gathers each Gaussian's partial sums straight from the owning peers' buffers over NVLink
            peer_handle.barrier(channel=0)
            grads = ext.backward_preprocess(*pre, None, [int(p or 0) for p in peer_ptrs], [int(v) for b in peer.bands for v in b],
                                            True)
        else:
            # the ONE exchange step of the multi-GPU path: sum the compact [N,10] screen-space gradients of all
            # tile-row bands, NCCL over NVLink
            import torch.distributed as dist
            dist.all_reduce(sgrad, op=dist.ReduceOp.SUM, group=opt.process_group)
            grads = ext.backward_preprocess(*pre, sgrad, [], [], False)
    dmeans2D, dcol, dopac, dmeans3D, dcov, dsh, dsc, drot = grads
    has_sh, has_col, has_sr, has_cov = ctx.has
    out = (dmeans3D, dmeans2D, dsh if has_sh else None, dcol if has_col else None, dopac.reshape(ctx.opacity_shape),
           dsc if has_sr else None, drot if has_sr else None, dcov if has_cov else None)
    need = ctx.needs_input_grad
    return tuple(g if need[i] else None for i, g in enumerate(out)) + (None, None)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, opt):
        lib = L.load()
        rs: GaussianRasterizationSettings = raster_settings
        opt = opt if opt is not None else TouchOptions()
        if opt.depth_loss not in L.LOSS_MODES:
            raise ValueError(f"depth_loss must be one of {list(L.LOSS_MODES)}, got {opt.depth_loss!r}")
        dev = means3D.device
        if dev.type != "cuda":
            raise RuntimeError("touchgs_b200 rasterizer is CUDA-only (no CPU fallback); tensors are on " + str(dev))
        ext = L.load_ext()
        if ext is not None:
            return _forward_ext(ext, ctx, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, opt)
        sh, colors_precomp = _empty_if(sh), _empty_if(colors_precomp)
        scales, rotations, cov3Ds_precomp = _empty_if(scales), _empty_if(rotations), _empty_if(cov3Ds_precomp)
        N = int(means3D.shape[0])
        H, W = int(rs.image_height), int(rs.image_width)
        means3D = _chk(means3D, "means3D", (N, 3), dev)
        opacity_shape = tuple(opacities.shape)
        opacities = _chk(opacities.reshape(-1), "opacities", (N,), dev)
        K = 0
        if sh is not None:
            if sh.dim() != 3 or sh.shape[0] != N or sh.shape[2] != 3:
                raise ValueError(f"shs must have shape [N,K,3], got {list(sh.shape)}")
            K = int(sh.shape[1])
            sh = _chk(sh, "shs", (N, K, 3), dev)
        colors_precomp = _chk(colors_precomp, "colors_precomp", (N, 3), dev)
        scales = _chk(scales, "scales", (N, 3), dev)
        rotations = _chk(rotations, "rotations", (N, 4), dev)
        cov3Ds_precomp = _chk(cov3Ds_precomp, "cov3D_precomp", (N, 6), dev)
        touch_depth = _chk(opt.touch_depth, "touch_depth", None, dev)
        touch_weight = _chk(opt.touch_weight, "touch_weight", None, dev)
        for nm, t in (("touch_depth", touch_depth), ("touch_weight", touch_weight)):
            if t is not None and t.numel() != H * W:
                raise ValueError(f"{nm} must have H*W = {H * W} elements, got {list(t.shape)}")
        if opt.depth_loss != "none" and touch_depth is None:
            raise ValueError("depth_loss != 'none' requires touch_depth")

        keep = []
        with torch.cuda.device(dev):
            s, (r0, r1, Ty) = _make_settings(rs, opt, K, keep)
            g = _make_gaussians(means3D, opacities, sh, colors_precomp, scales, rotations, cov3Ds_precomp)
            full = (r0 == 0 and r1 == Ty)
            mk = torch.empty if full else torch.zeros
            color = mk((3, H, W), dtype=torch.float32, device=dev)
            depth = mk((1, H, W), dtype=torch.float32, device=dev)
            alpha = mk((1, H, W), dtype=torch.float32, device=dev)
            resid = torch.zeros((1, H, W), dtype=torch.float32, device=dev) if (touch_depth is None or not full) \
                else torch.empty((1, H, W), dtype=torch.float32, device=dev)
            radii = torch.zeros((N,), dtype=torch.int32, device=dev)
            scratch = _Scratch(dev)
            saved = L.TgsSaved()
            rc = lib.tgs_forward(C.byref(s), C.byref(g), scratch.cb, None, _ptr(color), _ptr(depth), _ptr(alpha),
                                 _ptr(radii), _ptr(touch_depth), _ptr(resid) if touch_depth is not None else None,
                                 C.byref(saved), _stream_ptr(dev))
            scratch.disarm()
            if scratch.error is not None:
                raise scratch.error
            if rc != 0 and rs.debug:
                # reference-era operator habit (SURVEY §8b "Errors"): with debug=True a failing forward leaves its
                # arguments behind for post-mortem inspection
                try:
                    torch.save(dict(means3D=means3D, opacities=opacities, shs=sh, colors_precomp=colors_precomp, scales=scales,
                                    rotations=rotations, cov3D_precomp=cov3Ds_precomp, settings=tuple(rs)), "snapshot_fw.dump")
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                except Exception:  # noqa: BLE001 - the dump is best effort; the real error is raised below
                    pass
            L.check(rc, "tgs_forward")

            # scale = depth_loss_mult / Z and the VALUE of the fused touch loss (logging; differentiable on request)
            tscale = None
            tloss = None
            if touch_depth is not None and opt.depth_loss != "none":
                tscale = torch.empty(2, dtype=torch.float32, device=dev)
                norm = float(opt.depth_loss_norm) if opt.depth_loss_norm is not None else 0.0
                L.check(lib.tgs_touch_loss_scale(_ptr(touch_depth), H * W, float(opt.depth_loss_mult), norm,
                                                 _ptr(tscale), _stream_ptr(dev)), "tgs_touch_loss_scale")
                if opt.return_touch_loss:
                    acc = torch.empty(1, dtype=torch.float64, device=dev)
                    tloss = torch.empty((), dtype=torch.float32, device=dev)
                    tr = (0, 0) if opt.touch_rows is None else (int(opt.touch_rows[0]), int(opt.touch_rows[1]))
                    if opt.touch_rows is None and not full:      # a band without explicit loss rows: its own pixel rows
                        tr = (min(r0 * TILE, H), min(r1 * TILE, H))
                    L.check(lib.tgs_touch_loss_value(_ptr(resid), _ptr(touch_weight), W, H, tr[0], tr[1],
                                                     L.LOSS_MODES[opt.depth_loss], _ptr(tscale), _ptr(acc), _ptr(tloss),
                                                     _stream_ptr(dev)), "tgs_touch_loss_value")
        ctx.tscale = tscale

        ctx.rs, ctx.opt, ctx.K = rs, opt, K
        ctx.settings, ctx.keep = s, keep           # the validated C structs are reused by backward
        ctx.opacity_shape = opacity_shape
        ctx.num_rendered = int(saved.num_rendered)
        ctx.capacity = int(saved.capacity)
        if opt.info is not None:
            opt.info["num_rendered"] = ctx.num_rendered
            opt.info["capacity"] = ctx.capacity
        ctx.has = (sh is not None, colors_precomp is not None, scales is not None, cov3Ds_precomp is not None)
        ctx.touch = (touch_depth, touch_weight)
        none = _empty_on(dev)
        ctx.save_for_backward(means3D, opacities, sh if sh is not None else none,
                              colors_precomp if colors_precomp is not None else none,
                              scales if scales is not None else none,
                              rotations if rotations is not None else none,
                              cov3Ds_precomp if cov3Ds_precomp is not None else none,
                              radii, scratch.bufs[L.BUF_GEOM], scratch.bufs[L.BUF_BINNING], scratch.bufs[L.BUF_IMAGE])
        ctx.set_materialize_grads(False)
        if tloss is None:
            tloss = _zero_scalar(dev)                  # placeholder (cached: no fill kernel per call)
            ctx.mark_non_differentiable(radii, resid, tloss)
        else:
            ctx.mark_non_differentiable(radii, resid)
        return color, radii, depth, alpha, resid, tloss

    @staticmethod
    def backward(ctx, g_color, _g_radii, g_depth, g_alpha, _g_resid, g_tloss=None):
        if getattr(ctx, "ext", None) is not None:
            return _backward_ext(ctx, g_color, g_depth, g_alpha, g_tloss)
        lib = L.load()
        (means3D, opacities, sh, colors, scales, rots, cov3D, radii, geom, binning, image) = ctx.saved_tensors
        has_sh, has_col, has_sr, has_cov = ctx.has
        sh = sh if has_sh else None
        colors = colors if has_col else None
        scales, rots = (scales, rots) if has_sr else (None, None)
        cov3D = cov3D if has_cov else None
        rs, opt, K = ctx.rs, ctx.opt, ctx.K
        dev = means3D.device
        N = int(means3D.shape[0])
        H, W = int(rs.image_height), int(rs.image_width)
        keep = ctx.keep
        with torch.cuda.device(dev):
            s = ctx.settings
            g = _make_gaussians(means3D, opacities, sh, colors, scales, rots, cov3D)
            saved = L.TgsSaved(geom=geom.data_ptr(), binning=binning.data_ptr(), image=image.data_ptr(),
                               num_rendered=ctx.num_rendered, capacity=ctx.capacity)
            g_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev) if g_color is None \
                else _chk(g_color, "grad_color", (3, H, W), dev)
            g_depth = None if g_depth is None else _chk(g_depth.reshape(H, W), "grad_depth", (H, W), dev)
            g_alpha = None if g_alpha is None else _chk(g_alpha.reshape(H, W), "grad_alpha", (H, W), dev)
            touch_depth, touch_weight = ctx.touch
            touch = None
            if touch_depth is not None and opt.depth_loss != "none":
                scale = ctx.tscale
                # upstream gradient of the touch-loss scalar, applied on the device (no host sync):
                #   return_touch_loss=True : whatever autograd carries for that output (None = the caller left the loss
                #                            out of its objective -> the fused gradient is off for this backward);
                #   injected mode          : 1, or the caller's explicit `loss_grad_scale`.
                gs = None
                mode = L.LOSS_MODES[opt.depth_loss]
                if opt.return_touch_loss:
                    if g_tloss is None:
                        mode = L.LOSS_NONE
                    else:
                        gs = g_tloss.detach().reshape(1).to(device=dev, dtype=torch.float32).contiguous()
                elif opt.loss_grad_scale is not None:
                    gs = torch.as_tensor(opt.loss_grad_scale, dtype=torch.float32, device=dev).detach().reshape(1).contiguous()
                if gs is not None:
                    keep.append(gs)
                touch = L.TgsTouch(target=touch_depth.data_ptr(),
                                   weight=None if touch_weight is None else touch_weight.data_ptr(),
                                   scale=scale.data_ptr(), mode=mode,
                                   row_begin=0 if opt.touch_rows is None else int(opt.touch_rows[0]),
                                   row_end=0 if opt.touch_rows is None else int(opt.touch_rows[1]),
                                   grad_scale=None if gs is None else gs.data_ptr())
            peer = opt.peer_exchange if N > 0 else None
            if peer is not None and N > peer.capacity:
                # more Gaussians than the peer-mapped buffers were sized for (the population grew at a refine step):
                # every rank sees the same N, so all of them take the all-reduce path for this call
                if opt.process_group is None:
                    raise ValueError(f"PeerScreenGrads capacity {peer.capacity} < {N} Gaussians and no process_group to fall back to")
                peer = None
            if peer is not None:
                sgrad, peer_ptrs, peer_handle = peer.acquire(N)     # this rank's peer-mapped buffer of the step
                s.contrib_flags = 1                                 # ... which carries contributor bytes behind the rows
            else:
                # plain rows: required when the buffer is all-reduced, and on one GPU the chain rule finds the
                # non-contributors from the zero rows themselves (cheaper than writing the bytes in BACKWARD::render)
                sgrad = torch.empty((max(N, 1), L.NGRAD), dtype=torch.float32, device=dev)
                s.contrib_flags = 0
            L.check(lib.tgs_backward_render(C.byref(s), C.byref(g), C.byref(saved), _ptr(g_color), _ptr(g_depth),
                                            _ptr(g_alpha), None if touch is None else C.byref(touch), None,
                                            _ptr(sgrad), _stream_ptr(dev)), "tgs_backward_render")
            if peer is None and opt.process_group is not None:
                # the ONE exchange step of the multi-GPU path: sum the compact [N,10] screen-space
                # gradients of all tile-row bands, NCCL over NVLink
                import torch.distributed as dist
                dist.all_reduce(sgrad, op=dist.ReduceOp.SUM, group=opt.process_group)
            dmeans2D = torch.empty((N, 3), dtype=torch.float32, device=dev)
            dmeans3D = torch.empty((N, 3), dtype=torch.float32, device=dev)
            dopac = torch.empty((N,), dtype=torch.float32, device=dev)
            dsh = torch.empty((N, K, 3), dtype=torch.float32, device=dev) if has_sh else None
            dcol = torch.empty((N, 3), dtype=torch.float32, device=dev) if has_col else None
            dsc = torch.empty((N, 3), dtype=torch.float32, device=dev) if has_sr else None
            drot = torch.empty((N, 4), dtype=torch.float32, device=dev) if has_sr else None
            dcov = torch.empty((N, 6), dtype=torch.float32, device=dev) if has_cov else None
            gr = L.TgsGrads(dmeans2D=dmeans2D.data_ptr(), dmeans3D=dmeans3D.data_ptr(), dopacity=dopac.data_ptr(),
                            dshs=None if dsh is None else dsh.data_ptr(),
                            dcolors=None if dcol is None else dcol.data_ptr(),
                            dscales=None if dsc is None else dsc.data_ptr(),
                            drotations=None if drot is None else drot.data_ptr(),
                            dcov3D=None if dcov is None else dcov.data_ptr())
            if peer is not None:
                # FUSED exchange: wait until every rank has finished BACKWARD::render, then the chain-rule
                # kernel gathers each Gaussian's partial sums straight from the owning peers' buffers over NVLink
                peer_handle.barrier(channel=0)
                L.check(lib.tgs_backward_preprocess_gather(C.byref(s), C.byref(g), C.byref(saved), _ptr(radii), peer_ptrs,
                                                           peer.band_array(), peer.world, C.byref(gr), _stream_ptr(dev)),
                        "tgs_backward_preprocess_gather")
            else:
                L.check(lib.tgs_backward_preprocess(C.byref(s), C.byref(g), C.byref(saved), _ptr(radii), _ptr(sgrad),
                                                    C.byref(gr), _stream_ptr(dev)), "tgs_backward_preprocess")
        # autograd rejects a gradient for an input that was not a Variable (e.g. means2D=None)
        grads = (dmeans3D, dmeans2D, dsh, dcol, dopac.reshape(ctx.opacity_shape), dsc, drot, dcov)
        need = ctx.needs_input_grad
        return tuple(g if need[i] else None for i, g in enumerate(grads)) + (None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, touch: Optional[TouchOptions] = None):
    """Same positional signature as the reference-era ``rasterize_gaussians`` plus the
    optional ``touch`` extension.  Returns (color [3,H,W], radii [N] int32, depth [1,H,W],
    alpha [1,H,W], depth_residual [1,H,W]) -- plus the differentiable touch-loss scalar as a sixth element when
    ``touch.return_touch_loss`` is set."""
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, raster_settings, touch)
    return out if (touch is not None and touch.return_touch_loss) else out[:5]


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self._last = (0, 0)

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """bool[N]: view-space z > 0.2 for the module's camera."""
        lib = L.load()
        rs = self.raster_settings
        dev = positions.device
        if dev.type != "cuda":
            raise RuntimeError("touchgs_b200 rasterizer is CUDA-only (no CPU fallback)")
        ext = L.load_ext()
        if ext is not None:
            with torch.no_grad():
                return ext.mark_visible(positions.detach(), rs.viewmatrix, rs.projmatrix)
        with torch.no_grad(), torch.cuda.device(dev):
            N = int(positions.shape[0])
            pos = _chk(positions.detach(), "positions", (N, 3), dev)
            vm = _chk(rs.viewmatrix, "viewmatrix", (4, 4), dev)
            out = torch.zeros((N,), dtype=torch.uint8, device=dev)
            L.check(lib.tgs_mark_visible(N, _ptr(pos), _ptr(vm), _ptr(out), _stream_ptr(dev)), "tgs_mark_visible")
        return out.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, *, touch_depth=None, touch_weight=None,
                depth_loss: str = "none", depth_loss_mult: float = 1.0, depth_normalize: bool = True,
                depth_loss_norm: Optional[float] = None, tile_rows=None, process_group=None,
                rendered_hint: int = 0, touch_rows=None, peer_exchange=None, return_touch_loss: bool = False,
                loss_grad_scale=None, defer_count: bool = False):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        e = _empty_on(means3D.device)
        info = {}
        opt = TouchOptions(touch_depth, touch_weight, depth_loss, depth_loss_mult, depth_normalize,
                           depth_loss_norm, tile_rows, process_group, rendered_hint, info, touch_rows, peer_exchange,
                           bool(return_touch_loss), bool(defer_count), loss_grad_scale)
        out = rasterize_gaussians(means3D, means2D,
                                  e if shs is None else shs, e if colors_precomp is None else colors_precomp,
                                  opacities, e if scales is None else scales, e if rotations is None else rotations,
                                  e if cov3D_precomp is None else cov3D_precomp, self.raster_settings, opt)
        # instance count of this call: feed it back as `rendered_hint` the next time this view is rendered.  After a deferred
        # forward it is a (negative) ticket that `last_num_rendered` redeems on first access -- read it AFTER backward
        self._last = (info.get("num_rendered", 0), info.get("capacity", 0))
        return out

    @property
    def last_num_rendered(self) -> int:
        n, cap = self._last
        if n < 0:
            import ctypes as _C
            v = _C.c_int64(0)
            L.check(L.load().tgs_forward_resolve(int(n), int(cap), _C.byref(v)), "tgs_forward_resolve")
            n = int(v.value)
            self._last = (n, cap)
        return n

    @last_num_rendered.setter
    def last_num_rendered(self, v):
        self._last = (int(v), 0)
