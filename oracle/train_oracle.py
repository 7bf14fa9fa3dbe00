"""CPU oracle of the TRAIN-STEP pieces around the rasterizer (BASELINE config c5).

STATUS: TEST INFRASTRUCTURE ONLY -- **PARITY UNPINNED**.  The trainer lives in the reference's empty
``nerfstudio/`` submodule (reference ``.gitmodules:7-9``); nothing of it is in the tree.  What the tree pins and
this file follows: the loss knobs of the CLI (``--pipeline.model.depth-loss-mult``, ``depth-loss-type
{SIMPLE_LOSS, DEPTH_UNCERTAINTY_WEIGHTED_LOSS}``, ``uncertainty_weight``: reference
``scripts/train_block_data.sh:50``, ``scripts/train_bunny_blender.sh:50``, ``scripts/train_bunny_real.sh:52``),
``AdamOptimizerConfig(lr=..., eps=1e-15)`` (reference ``legacy/config_tactile.py:43-50``) and the iteration
count (reference ``legacy/config_tactile.py:28``).  Everything else restates the PUBLISHED algorithm of the public
splat trainers of that era: loss = (1-l)*L1 + l*(1-SSIM) with an 11x11 sigma-1.5 Gaussian
window and zero padding, Adam (betas 0.9/0.999) per parameter group, and the refine step (duplicate small /
split large Gaussians whose mean screen-space gradient norm exceeds a threshold, cull transparent / huge ones,
periodic opacity reset).  Adam is ``torch.optim.Adam`` itself.

Only ``tests/`` may import this module.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch
import torch.nn.functional as F

__all__ = ["ssim_window", "photometric_loss", "activate", "adam_reference", "DensifyConfig", "densify_reference",
           "densify_stats_reference", "quat_to_rotmat", "reset_opacity_reference"]

SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2


def ssim_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    """1-D normalised Gaussian taps exp(-(i - size//2)^2 / (2 sigma^2)) (computed in float64, then cast)."""
    x = torch.arange(size, dtype=torch.float64) - size // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
    return (g / g.sum()).to(dtype)


def _blur(img, w1d):
    """Zero-padded separable 11x11 blur, channel-wise.  img [C,H,W]."""
    C = img.shape[0]
    k = (w1d[:, None] * w1d[None, :]).to(img.dtype)
    k = k[None, None].expand(C, 1, -1, -1).contiguous()
    return F.conv2d(img[None], k, padding=w1d.numel() // 2, groups=C)[0]


def photometric_loss(color, gt, lambda_dssim: float = 0.2, rows=None):
    """(1-l) * mean|C - C*| + l * (1 - mean SSIM(C, C*)), means over ALL 3*H*W elements of the image.
    ``rows=(y0,y1)``: only pixels of those rows enter the sums (a rank's band of the tile-row shard); the partial
    losses of disjoint bands add up to the full-image loss.  Differentiable w.r.t. ``color``."""
    C, H, W = color.shape
    w = ssim_window(dtype=color.dtype)
    mu1, mu2 = _blur(color, w), _blur(gt, w)
    s11 = _blur(color * color, w) - mu1 * mu1
    s22 = _blur(gt * gt, w) - mu2 * mu2
    s12 = _blur(color * gt, w) - mu1 * mu2
    smap = ((2 * mu1 * mu2 + SSIM_C1) * (2 * s12 + SSIM_C2)) / ((mu1 * mu1 + mu2 * mu2 + SSIM_C1) * (s11 + s22 + SSIM_C2))
    l1 = (color - gt).abs()
    if rows is not None:
        y0, y1 = rows
        smap, l1 = smap[:, y0:y1], l1[:, y0:y1]
    n = float(C * H * W)
    return (1.0 - lambda_dssim) * l1.sum() / n + lambda_dssim * (smap.numel() - smap.sum()) / n


def activate(scales_log, quats, opacity_logit):
    """Raw parameters -> what the rasterizer consumes: exp, normalise, sigmoid."""
    return torch.exp(scales_log), quats / quats.norm(dim=-1, keepdim=True), torch.sigmoid(opacity_logit)


def adam_reference(params, grads, lrs, steps: int, betas=(0.9, 0.999), eps: float = 1e-15, state=None):
    """``steps`` updates of torch.optim.Adam with the SAME gradient each step; returns (params, exp_avg, exp_avg_sq)."""
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], betas=betas, eps=eps)
    for _ in range(steps):
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        opt.step()
    return ([p.detach() for p in ps], [opt.state[p]["exp_avg"] for p in ps], [opt.state[p]["exp_avg_sq"] for p in ps])


def quat_to_rotmat(q):
    """(w,x,y,z), normalised inside (the refine step samples in the Gaussian's frame)."""
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


class DensifyConfig(NamedTuple):
    grad_thresh: float = 0.0002        # on the mean NDC-space mean2D gradient norm
    size_thresh: float = 0.01          # world-space scale above which a Gaussian is SPLIT instead of duplicated
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    n_split_samples: int = 2
    split_shrink: float = 1.6
    split_screen_radius: float = 0.0   # pixels; > 0: split when the largest screen radius since the last refine exceeds it
    cull_screen_radius: float = 0.0    # pixels; > 0: cull when it exceeds it


def densify_stats_reference(dmeans2D, radii, grad_accum, vis_count, max_radii):
    """Per-step refine statistics: visible (radii > 0) Gaussians accumulate ||dL/dmean2D|| and a visit count."""
    vis = radii > 0
    g = dmeans2D[:, :2].norm(dim=-1)
    return (grad_accum + torch.where(vis, g, torch.zeros_like(g)), vis_count + vis.to(vis_count.dtype),
            torch.maximum(max_radii, torch.where(vis, radii, torch.zeros_like(radii))))


def densify_reference(means, shs, opacity_logit, scales_log, quats, grad_accum, vis_count, noise, cfg: DensifyConfig,
                      allow_split_dup: bool = True, max_radii=None):
    """One refine step.  Output ORDER (ours; the trainer's is not in the tree): Gaussians in id order, each
    replaced by its outputs -- culled: nothing; kept: itself; duplicated: itself then its copy; split: its
    ``n_split_samples`` samples (the original is dropped).  Returns the new parameter tensors plus, per output,
    the source id and a flag "new" (Adam state of new entries is zero, kept entries carry theirs over).
    ``noise`` [N, n_split_samples, 3] ~ N(0,1) supplied by the caller."""
    N = means.shape[0]
    avg = grad_accum / vis_count.clamp_min(1).to(grad_accum.dtype)
    avg = torch.where(vis_count > 0, avg, torch.zeros_like(avg))
    smax = torch.exp(scales_log).max(dim=-1).values
    high = (avg > cfg.grad_thresh) & allow_split_dup
    rmax = max_radii.to(avg.dtype) if max_radii is not None else torch.zeros_like(avg)
    big = (rmax > cfg.split_screen_radius) & (cfg.split_screen_radius > 0) & allow_split_dup
    split = (high & (smax > cfg.size_thresh)) | big
    dup = high & ~split
    cull = (torch.sigmoid(opacity_logit.reshape(-1)) < cfg.cull_alpha_thresh) | (smax > cfg.cull_scale_thresh)
    if cfg.cull_screen_radius > 0:
        cull = cull | (rmax > cfg.cull_screen_radius)
    split, dup = split & ~cull, dup & ~cull
    ns = cfg.n_split_samples
    count = torch.where(cull, 0, torch.where(split, ns, torch.where(dup, 2, 1)))
    offs = torch.cumsum(count, 0) - count
    M = int(count.sum())
    src = torch.repeat_interleave(torch.arange(N), count)
    local = torch.arange(M) - offs[src]
    is_split = split[src]
    is_new = is_split | (dup[src] & (local == 1))
    o_means, o_shs = means[src].clone(), shs[src].clone()
    o_op, o_sc, o_q = opacity_logit.reshape(N, -1)[src].clone(), scales_log[src].clone(), quats[src].clone()
    if bool(is_split.any()):
        R = quat_to_rotmat(quats[src[is_split]])
        nz = noise[src[is_split], local[is_split]]
        sc = torch.exp(scales_log[src[is_split]])
        o_means[is_split] = means[src[is_split]] + torch.einsum("nij,nj->ni", R, sc * nz)
        o_sc[is_split] = torch.log(sc / cfg.split_shrink)
    return dict(means=o_means, shs=o_shs, opacity_logit=o_op.reshape(M, *opacity_logit.shape[1:]), scales_log=o_sc,
                quats=o_q, src=src, is_new=is_new, count=count)


def reset_opacity_reference(opacity_logit, cull_alpha_thresh: float):
    """Periodic opacity reset: clamp the logit to logit(2 * cull_alpha_thresh) from above."""
    v = 2.0 * cull_alpha_thresh
    return torch.clamp(opacity_logit, max=math.log(v / (1.0 - v)))
