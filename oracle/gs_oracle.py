"""CPU oracle: a pure-PyTorch restatement of the Touch-GS rasterizer hot path.

STATUS: TEST INFRASTRUCTURE ONLY -- **PARITY UNPINNED**.

The reference tree (``/root/reference``) does not vendor its rasterizer: training is
``ns-train depth-gaussian-splatting`` (reference ``scripts/train_bunny_real.sh:52``,
``scripts/train_block_data.sh:50``) from an *empty* git submodule (reference
``.gitmodules:7-9``), and there is not one test, golden vector or fixture for this
path anywhere in the tree.  The algorithm therefore follows the
written specification in SURVEY.md §8(a) rows A1-A6 (the public Inria
``diff-gaussian-rasterization`` conventions, extended with the expected-depth /
alpha channels and the fused touch-depth loss).  What *is* pinned in the reference
and followed here: the touch-depth target / uncertainty encodings (uint16 mm PNG,
0 = invalid: reference ``utils/fuse_touch_vision.py:372-388``,
``utils/read_touch_depths.py:48-56``), validity mask ``depth > 0`` (reference
``utils/fuse_touch_vision.py:51,109``), vision sigma heuristic (reference
``utils/fuse_touch_vision.py:310-313``) and the loss multiplier semantics (reference
``legacy/model_tactile.py:162``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.

Design rules of this file
-------------------------
* Every arithmetic step of the *integer-feeding* part of preprocess (projection,
  covariance, radius, tile rectangle, depth key) is written as explicit
  element-wise torch ops in a fixed order.  In float32 each op is IEEE-rounded with
  no fused multiply-add, which is exactly what the CUDA kernel does (its translation
  unit is compiled with ``--fmad=false``), so radii / rects / tiles_touched / sort
  keys are compared BIT-EXACT.
* Gradients come from autograd; nothing here transcribes a backward formula.
* ``dtype=torch.float64`` gives the high-precision variant used for gradcheck and
  for fp32-vs-fp64 self-consistency tests.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import numpy as np
import torch

__all__ = [
    "OracleSettings", "TILE", "NEAR_Z", "ALPHA_MAX", "ALPHA_MIN", "T_EPS",
    "preprocess", "bin_and_sort", "render_tiles", "render_naive", "rasterize",
    "touch_loss", "mark_visible", "sh_to_rgb", "loss_scale_from_target",
    "camera_scalars",
]

TILE = 16                # SURVEY §8 notation: 16x16 tiles
NEAR_Z = 0.2             # A1: cull if view-space z <= 0.2
ALPHA_MAX = 0.99         # A5
ALPHA_MIN = 1.0 / 255.0  # A5
T_EPS = 1e-4             # A5: stop when T*(1-alpha) < 1e-4
COV_BLUR = 0.3           # A1: +0.3 px^2 on the 2D covariance diagonal
FOV_CLAMP = 1.3          # A1: clamp t.x/t.z to +-1.3 tanfov

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435)


class OracleSettings(NamedTuple):
    """Same fields as the operator's ``GaussianRasterizationSettings``."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor            # [3]
    scale_modifier: float
    viewmatrix: torch.Tensor    # [4,4] TRANSPOSED world->view (row-vector convention)
    projmatrix: torch.Tensor    # [4,4] TRANSPOSED full projection
    sh_degree: int
    campos: torch.Tensor        # [3]
    prefiltered: bool = False
    debug: bool = False
    # convention switches; the defaults are the primary (Inria) convention
    alpha_max: float = ALPHA_MAX          # gsplat 0.1.x: 0.999
    near_z: float = NEAR_Z                # gsplat 0.1.x clip_thresh: 0.01
    principal: tuple = (0.0, 0.0)         # (cx - W/2, cy - H/2) in pixels
    pixel_offset: float = 0.0             # sample point (x + offset, y + offset); later gsplat 0.1.x: 0.5


def camera_scalars(S, dtype=torch.float32):
    """focal_x, focal_y, limx, limy computed the way the C-ABI host code does (fp32)."""
    if dtype == torch.float32:
        f = np.float32
        fx = f(S.image_width) / (f(2.0) * f(S.tanfovx))
        fy = f(S.image_height) / (f(2.0) * f(S.tanfovy))
        lx = f(FOV_CLAMP) * f(S.tanfovx)
        ly = f(FOV_CLAMP) * f(S.tanfovy)
        return float(fx), float(fy), float(lx), float(ly)
    fx = S.image_width / (2.0 * S.tanfovx)
    fy = S.image_height / (2.0 * S.tanfovy)
    return fx, fy, FOV_CLAMP * S.tanfovx, FOV_CLAMP * S.tanfovy


# --------------------------------------------------------------------------- SH
def sh_to_rgb(deg: int, dirs: torch.Tensor, sh: torch.Tensor):
    """A1 colour: real SH evaluation (+0.5, clamp >= 0).  dirs [N,3] unit, sh [N,K,3].

    Returns (rgb [N,3], clamped [N,3] bool)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5]
                   + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                   + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9]
                       + SH_C3[1] * xy * z * sh[:, 10]
                       + SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11]
                       + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
                       + SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13]
                       + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                       + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    res = res + 0.5
    clamped = res < 0
    return torch.clamp_min(res, 0.0), clamped


# ------------------------------------------------------------------- preprocess
class Pre(NamedTuple):
    xy: torch.Tensor          # [N,2] pixel-space mean
    depth: torch.Tensor       # [N]   view-space z
    cov3D: torch.Tensor       # [N,6]
    conic: torch.Tensor       # [N,3] (A,B,C)
    opacity: torch.Tensor     # [N]
    rgb: torch.Tensor         # [N,3]
    radii: torch.Tensor       # [N] int32 (0 = invisible in the FULL image)
    rect_min: torch.Tensor    # [N,2] int32 (x,y) tile rect clipped to the band
    rect_max: torch.Tensor    # [N,2] int32
    tiles_touched: torch.Tensor  # [N] int32 (in the band)
    clamped: torch.Tensor     # [N,3] bool
    cov2D: torch.Tensor       # [N,3] (a,b,c) incl. blur, for diagnostics


def _xform(m, x, y, z, col):
    """((m[0][col]*x + m[1][col]*y) + m[2][col]*z) + m[3][col] -- fixed order, no FMA."""
    return ((m[0, col] * x + m[1, col] * y) + m[2, col] * z) + m[3, col]


def mark_visible(means3D: torch.Tensor, viewmatrix: torch.Tensor) -> torch.Tensor:
    """``markVisible``: view-space z > 0.2."""
    x, y, z = means3D.unbind(-1)
    return _xform(viewmatrix.to(means3D.dtype), x, y, z, 2) > NEAR_Z


def preprocess(means3D, scales, rotations, opacities, shs, colors_precomp,
               cov3D_precomp, S: OracleSettings, band: Optional[tuple] = None) -> Pre:
    """SURVEY §8(a) row A1.  ``band=(row_begin,row_end)`` clips tile rects to a tile-row band
    (multi-GPU shard); radii always describe full-image visibility."""
    dt = means3D.dtype
    N = means3D.shape[0]
    W, H = S.image_width, S.image_height
    Tx, Ty = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    vm = S.viewmatrix.to(dt)
    pm = S.projmatrix.to(dt)
    fx, fy, limx, limy = camera_scalars(S, dt)

    x, y, z = means3D.unbind(-1)
    tx, ty, tz = (_xform(vm, x, y, z, c) for c in range(3))
    hx, hy, hw = _xform(pm, x, y, z, 0), _xform(pm, x, y, z, 1), _xform(pm, x, y, z, 3)
    vis = tz > S.near_z
    pw = 1.0 / (hw + 1e-7)
    ndcx, ndcy = hx * pw, hy * pw
    px = ((ndcx + 1.0) * W - 1.0) * 0.5 + float(np.float32(S.principal[0]))
    py = ((ndcy + 1.0) * H - 1.0) * 0.5 + float(np.float32(S.principal[1]))

    # --- 3D covariance
    if cov3D_precomp is not None:
        c3 = cov3D_precomp
    else:
        mod = S.scale_modifier
        sx, sy, sz = (mod * scales[:, i] for i in range(3))
        qr, qx, qy, qz = rotations.unbind(-1)
        R = [[1.0 - 2.0 * (qy * qy + qz * qz), 2.0 * (qx * qy - qr * qz), 2.0 * (qx * qz + qr * qy)],
             [2.0 * (qx * qy + qr * qz), 1.0 - 2.0 * (qx * qx + qz * qz), 2.0 * (qy * qz - qr * qx)],
             [2.0 * (qx * qz - qr * qy), 2.0 * (qy * qz + qr * qx), 1.0 - 2.0 * (qx * qx + qy * qy)]]
        s = (sx, sy, sz)
        M = [[R[i][j] * s[j] for j in range(3)] for i in range(3)]

        def dot3(i, k):
            return (M[i][0] * M[k][0] + M[i][1] * M[k][1]) + M[i][2] * M[k][2]
        c3 = torch.stack([dot3(0, 0), dot3(0, 1), dot3(0, 2), dot3(1, 1), dot3(1, 2), dot3(2, 2)], -1)
    S00, S01, S02, S11, S12, S22 = c3.unbind(-1)
    Sg = [[S00, S01, S02], [S01, S11, S12], [S02, S12, S22]]

    # --- EWA 2D covariance
    safe_tz = torch.where(vis, tz, torch.ones_like(tz))   # culled lanes never divide by <=0.2
    cx = _fov_clamp(tx, safe_tz, limx)
    cy = _fov_clamp(ty, safe_tz, limy)
    itz = 1.0 / safe_tz          # NOTE: torch evaluates scalar/tensor as reciprocal*scalar; be explicit
    J00 = fx * itz
    J11 = fy * itz
    tz2 = safe_tz * safe_tz
    J02 = -(fx * cx) / tz2
    J12 = -(fy * cy) / tz2
    # W[r][k] = V[r][k] = vm[k][r]
    m0 = [J00 * vm[k, 0] + J02 * vm[k, 2] for k in range(3)]
    m1 = [J11 * vm[k, 1] + J12 * vm[k, 2] for k in range(3)]
    u = [(Sg[k][0] * m0[0] + Sg[k][1] * m0[1]) + Sg[k][2] * m0[2] for k in range(3)]
    v = [(Sg[k][0] * m1[0] + Sg[k][1] * m1[1]) + Sg[k][2] * m1[2] for k in range(3)]
    a = ((m0[0] * u[0] + m0[1] * u[1]) + m0[2] * u[2]) + COV_BLUR
    b = (m0[0] * v[0] + m0[1] * v[1]) + m0[2] * v[2]
    c = ((m1[0] * v[0] + m1[1] * v[1]) + m1[2] * v[2]) + COV_BLUR
    det = a * c - b * b
    vis = vis & (det != 0)
    safe_det = torch.where(det != 0, det, torch.ones_like(det))
    det_inv = 1.0 / safe_det
    conic = torch.stack([c * det_inv, -b * det_inv, a * det_inv], -1)
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam = torch.maximum(mid + disc, mid - disc)
    rad_f = torch.ceil(3.0 * torch.sqrt(torch.clamp_min(lam, 0.0)))

    # --- tile rectangle (float clamp, then truncate: identical to trunc-then-clamp for finite input)
    def rect(p, r, n):
        lo = torch.clamp((p - r) / TILE, 0.0, float(n))
        hi = torch.clamp(((p + r) + (TILE - 1.0)) / TILE, 0.0, float(n))
        lo = torch.nan_to_num(lo.detach(), nan=0.0)
        hi = torch.nan_to_num(hi.detach(), nan=0.0)
        return lo.to(torch.int32), hi.to(torch.int32)
    rminx, rmaxx = rect(px, rad_f, Tx)
    rminy, rmaxy = rect(py, rad_f, Ty)
    full_tiles = (rmaxx - rminx) * (rmaxy - rminy)
    vis = vis & (full_tiles > 0)
    if band is not None:
        rminy = torch.clamp(rminy, band[0], band[1])
        rmaxy = torch.clamp(rmaxy, band[0], band[1])
    tiles = torch.where(vis, (rmaxx - rminx) * (rmaxy - rminy), torch.zeros_like(full_tiles))
    radii = torch.where(vis, rad_f.detach().to(torch.int32), torch.zeros_like(full_tiles))

    # --- colour
    if colors_precomp is not None:
        rgb = colors_precomp
        clamped = torch.zeros(N, 3, dtype=torch.bool)
    else:
        cam = S.campos.to(dt)
        dx, dy, dz = x - cam[0], y - cam[1], z - cam[2]
        ln = torch.sqrt((dx * dx + dy * dy) + dz * dz)
        dirs = torch.stack([dx / ln, dy / ln, dz / ln], -1)
        rgb, clamped = sh_to_rgb(S.sh_degree, dirs, shs)

    return Pre(xy=torch.stack([px, py], -1), depth=tz, cov3D=c3, conic=conic,
               opacity=opacities.reshape(-1), rgb=rgb, radii=radii,
               rect_min=torch.stack([rminx, rminy], -1), rect_max=torch.stack([rmaxx, rmaxy], -1),
               tiles_touched=tiles, clamped=clamped, cov2D=torch.stack([a, b, c], -1))


# ---------------------------------------------------------------------- binning
class Bins(NamedTuple):
    offsets: torch.Tensor      # [N] int64 inclusive scan of tiles_touched
    keys_unsorted: torch.Tensor  # [I] int64
    vals_unsorted: torch.Tensor  # [I] int32
    keys: torch.Tensor         # [I] int64 sorted
    vals: torch.Tensor         # [I] int32 sorted
    ranges: torch.Tensor       # [T,2] int32


def bin_and_sort(pre: Pre, S: OracleSettings) -> Bins:
    """inclusive scan, duplicateWithKeys, stable sort, tile ranges."""
    W, H = S.image_width, S.image_height
    Tx, Ty = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    N = pre.radii.shape[0]
    cnt = pre.tiles_touched.to(torch.int64)
    offsets = torch.cumsum(cnt, 0)
    I = int(offsets[-1]) if N > 0 else 0
    g = torch.repeat_interleave(torch.arange(N), cnt)                # Gaussian-major emission
    local = torch.arange(I) - (offsets - cnt)[g]
    rw = (pre.rect_max[:, 0] - pre.rect_min[:, 0]).to(torch.int64)[g]
    ty = pre.rect_min[g, 1].to(torch.int64) + local // torch.clamp_min(rw, 1)   # row-major: y outer
    tx = pre.rect_min[g, 0].to(torch.int64) + local % torch.clamp_min(rw, 1)
    tile = ty * Tx + tx
    dbits = pre.depth.detach().to(torch.float32).contiguous().view(torch.int32).to(torch.int64)[g] & 0xFFFFFFFF
    keys = (tile << 32) | dbits
    vals = g.to(torch.int32)
    order = torch.sort(keys, stable=True).indices
    skeys, svals = keys[order], vals[order]
    stile = skeys >> 32
    T = Tx * Ty
    ranges = torch.zeros(T, 2, dtype=torch.int32)
    if I > 0:
        tiles_ar = torch.arange(T)
        lo = torch.searchsorted(stile, tiles_ar, right=False)
        hi = torch.searchsorted(stile, tiles_ar, right=True)
        has = hi > lo
        ranges[:, 0] = torch.where(has, lo, torch.zeros_like(lo)).to(torch.int32)
        ranges[:, 1] = torch.where(has, hi, torch.zeros_like(hi)).to(torch.int32)
    return Bins(offsets, keys, vals, skeys, svals, ranges)


# ----------------------------------------------------------------------- render
class Img(NamedTuple):
    color: torch.Tensor      # [3,H,W] incl. background
    depth: torch.Tensor      # [H,W] raw  sum depth*alpha*T
    alpha: torch.Tensor      # [H,W] 1 - final_T
    final_T: torch.Tensor    # [H,W]
    n_contrib: torch.Tensor  # [H,W] int32


def _st_clamp_max(x, hi):
    """min(x, hi) in value, identity in gradient: the public rasterizers' backward uses
    dalpha/d(o*G) = 1 even where alpha was clamped to 0.99 (SURVEY A6 writes
    dL/dG = o*dL/dalpha unconditionally); the oracle keeps that behaviour."""
    return x + (torch.clamp_max(x, hi) - x).detach()


def _fov_clamp(t, tz, lim):
    """value = clamp(t/tz, -lim, lim) * tz (A1).  Gradient: 1 w.r.t. t where unclamped, and the
    clamped value is treated as a CONSTANT otherwise (the public backward zeroes dL/dt.x there
    and keeps the clamped t.x as a constant inside dL/dt.z)."""
    q = t / tz
    val = torch.clamp(q, -lim, lim) * tz
    inside = (q >= -lim) & (q <= lim)
    # t + (val - t) == val exactly in floating point when val is within 1 ulp of t (Sterbenz)
    return torch.where(inside, t + (val - t).detach(), val.detach())


def _composite_tile(xy, conic, opac, rgb, depth, pix, alpha_max=ALPHA_MAX):
    """Appendix A.1: vectorised front-to-back compositing reproducing sequential early exit.

    xy [L,2], conic [L,3], opac [L], rgb [L,3], depth [L]; pix [P,2] float pixel coords."""
    dx = xy[None, :, 0] - pix[:, None, 0]
    dy = xy[None, :, 1] - pix[:, None, 1]
    A, B, C = conic[None, :, 0], conic[None, :, 1], conic[None, :, 2]
    power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
    alpha = _st_clamp_max(opac[None, :] * torch.exp(power), alpha_max)
    skip = (power > 0) | (alpha < ALPHA_MIN)
    a_eff = torch.where(skip, torch.zeros_like(alpha), alpha)
    one_m = 1.0 - a_eff
    T_incl = torch.cumprod(one_m, dim=1)
    T_before = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
    term = (~skip) & (T_incl.detach() < T_EPS)
    dead = torch.cummax(term.to(torch.int8), dim=1).values.bool()
    live = ~(dead | skip)
    w = torch.where(live, a_eff * T_before, torch.zeros_like(a_eff))
    col = w @ rgb                       # [P,3]
    dep = w @ depth                     # [P]
    # final_T = product of live (1-alpha): sequential product identical to the running T
    fT = torch.cumprod(torch.where(live, one_m, torch.ones_like(one_m)), dim=1)[:, -1]
    L = xy.shape[0]
    idx = torch.arange(1, L + 1)[None, :].expand_as(w)
    ncon = torch.where(live, idx, torch.zeros_like(idx)).max(dim=1).values
    return col, dep, fT, ncon.to(torch.int32)


def render_tiles(pre: Pre, bins: Bins, S: OracleSettings, tiles=None) -> Img:
    """SURVEY §8(a) row A5 (RGB + expected depth + alpha in ONE traversal).  Differentiable.

    ``tiles``: optional iterable of tile ids to render (others left at background) --
    used for bounded-sample CPU timing."""
    dt = pre.xy.dtype
    W, H = S.image_width, S.image_height
    Tx, Ty = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    bg = S.bg.to(dt)
    color = bg[:, None, None].expand(3, H, W).clone()
    depth = torch.zeros(H, W, dtype=dt)
    fT = torch.ones(H, W, dtype=dt)
    ncon = torch.zeros(H, W, dtype=torch.int32)
    vals = bins.vals.to(torch.int64)
    tile_iter = range(Tx * Ty) if tiles is None else tiles
    for t in tile_iter:
        lo, hi = int(bins.ranges[t, 0]), int(bins.ranges[t, 1])
        if hi <= lo:
            continue
        ty_, tx_ = divmod(t, Tx)
        x0, y0 = tx_ * TILE, ty_ * TILE
        x1, y1 = min(x0 + TILE, W), min(y0 + TILE, H)
        ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).to(dt) + S.pixel_offset
        g = vals[lo:hi]
        col, dep, ft, nc = _composite_tile(pre.xy[g], pre.conic[g], pre.opacity[g],
                                           pre.rgb[g], pre.depth[g], pix, S.alpha_max)
        hh, ww = y1 - y0, x1 - x0
        color[:, y0:y1, x0:x1] = (col + ft[:, None] * bg[None, :]).t().reshape(3, hh, ww)
        depth[y0:y1, x0:x1] = dep.reshape(hh, ww)
        fT[y0:y1, x0:x1] = ft.reshape(hh, ww)
        ncon[y0:y1, x0:x1] = nc.reshape(hh, ww)
    return Img(color, depth, 1.0 - fT, fT, ncon)


def render_naive(pre: Pre, bins: Bins, S: OracleSettings, rows=None) -> Img:
    """The intentionally naive per-pixel / per-Gaussian Python loop (the "pure-PyTorch
    rasterize_gaussians" CPU path of BASELINE config #1).  Differentiable; tiny sizes only.
    ``rows``: optional set of pixel rows to composite (the others stay at the background): bounded-sample timing."""
    dt = pre.xy.dtype
    W, H = S.image_width, S.image_height
    Tx = (W + TILE - 1) // TILE
    bg = S.bg.to(dt)
    rows = None if rows is None else set(int(r) for r in rows)
    rows_c, rows_d, rows_t, rows_n = [], [], [], []
    for yy in range(H):
        rc, rd, rt, rn = [], [], [], []
        for xx in range(W):
            if rows is not None and yy not in rows:
                rc.append(bg.clone()); rd.append(torch.zeros((), dtype=dt)); rt.append(torch.ones((), dtype=dt)); rn.append(0)
                continue
            t = (yy // TILE) * Tx + xx // TILE
            lo, hi = int(bins.ranges[t, 0]), int(bins.ranges[t, 1])
            T = torch.ones((), dtype=dt)
            C = torch.zeros(3, dtype=dt)
            D = torch.zeros((), dtype=dt)
            last = 0
            for k, j in enumerate(range(lo, hi)):
                g = int(bins.vals[j])
                dx = pre.xy[g, 0] - float(xx)
                dy = pre.xy[g, 1] - float(yy)
                con = pre.conic[g]
                power = -0.5 * (con[0] * dx * dx + con[2] * dy * dy) - con[1] * dx * dy
                if power > 0:
                    continue
                alpha = _st_clamp_max(pre.opacity[g] * torch.exp(power), ALPHA_MAX)
                if alpha < ALPHA_MIN:
                    continue
                test_T = T * (1.0 - alpha)
                if test_T < T_EPS:
                    break
                C = C + pre.rgb[g] * (alpha * T)
                D = D + pre.depth[g] * (alpha * T)
                T = test_T
                last = k + 1
            rc.append(C + T * bg)
            rd.append(D)
            rt.append(T)
            rn.append(last)
        rows_c.append(torch.stack(rc, 0))
        rows_d.append(torch.stack(rd, 0))
        rows_t.append(torch.stack(rt, 0))
        rows_n.append(rn)
    color = torch.stack(rows_c, 0).permute(2, 0, 1).contiguous()
    depth = torch.stack(rows_d, 0)
    fT = torch.stack(rows_t, 0)
    return Img(color, depth, 1.0 - fT, fT, torch.tensor(rows_n, dtype=torch.int32))


# ------------------------------------------------------------------- touch loss
def loss_scale_from_target(target: torch.Tensor, mult: float, norm: Optional[float] = None) -> float:
    """scale = depth_loss_mult / Z with Z = #(target > 0) (min 1) unless given.
    Mirrors ``torch.mean(loss[mask])`` times ``depth_loss_mult`` (reference
    ``legacy/model_tactile.py:162`` for the multiplier; validity = depth > 0, reference
    ``utils/fuse_touch_vision.py:51``)."""
    Z = float(norm) if norm is not None else max(1.0, float((target > 0).sum()))
    return float(mult) / Z


def touch_loss(depth_raw, alpha, target, weight, mode: str, scale: float, normalize: bool):
    """SURVEY §8(a) A6 "Fusion": returns (loss scalar, residual [H,W], expected depth [H,W]).

    valid = (target > 0) & (alpha > 0); Dhat = D/A if normalize else D; r = Dhat - target;
    loss = scale * sum(valid*weight*|r|)  (l1)   or   scale * sum(valid*weight*r^2)  (l2)."""
    has = alpha > 0
    safe_a = torch.where(has, alpha, torch.ones_like(alpha))
    dhat = torch.where(has, depth_raw / safe_a, torch.zeros_like(depth_raw)) if normalize else depth_raw
    valid = (target > 0) & has
    r = torch.where(valid, dhat - target.to(dhat.dtype), torch.zeros_like(dhat))
    m = weight.to(dhat.dtype) if weight is not None else torch.ones_like(dhat)
    if mode == "l1":
        loss = scale * (m * r.abs()).sum()
    elif mode == "l2":
        loss = scale * (m * r * r).sum()
    elif mode == "none":
        loss = dhat.sum() * 0.0
    else:
        raise ValueError(mode)
    return loss, r.detach(), dhat


# --------------------------------------------------------------------- top level
class OracleOut(NamedTuple):
    color: torch.Tensor        # [3,H,W]
    radii: torch.Tensor        # [N] int32
    depth: torch.Tensor        # [1,H,W] expected depth (normalised if depth_normalize)
    alpha: torch.Tensor        # [1,H,W]
    residual: torch.Tensor     # [1,H,W]
    touch_loss: torch.Tensor   # scalar: the depth loss whose gradient the CUDA backward fuses
    pre: Pre
    bins: Bins
    img: Img


def rasterize(means3D, opacities, S: OracleSettings, shs=None, colors_precomp=None,
              scales=None, rotations=None, cov3D_precomp=None,
              touch_depth=None, touch_weight=None, depth_loss="none", depth_loss_mult=1.0,
              depth_normalize=True, depth_loss_norm=None, band=None, naive=False, naive_rows=None) -> OracleOut:
    """Full forward of the operator on CPU.  Differentiable w.r.t. all float inputs.

    To compare with the CUDA operator's fused backward, backpropagate
    ``(g_rgb * out.color).sum() + out.touch_loss`` (+ any external depth/alpha terms)."""
    pre = preprocess(means3D, scales, rotations, opacities, shs, colors_precomp, cov3D_precomp, S, band)
    bins = bin_and_sort(pre, S)
    img = render_naive(pre, bins, S, naive_rows) if naive else render_tiles(pre, bins, S)
    H, W = S.image_height, S.image_width
    if touch_depth is not None and depth_loss != "none":
        scale = loss_scale_from_target(touch_depth, depth_loss_mult, depth_loss_norm)
        tl, resid, dhat = touch_loss(img.depth, img.alpha, touch_depth, touch_weight,
                                     depth_loss, scale, depth_normalize)
    else:
        tgt = touch_depth if touch_depth is not None else torch.zeros(H, W, dtype=img.depth.dtype)
        tl, resid, dhat = touch_loss(img.depth, img.alpha, tgt, touch_weight, "none", 0.0, depth_normalize)
    return OracleOut(img.color, pre.radii, dhat[None], img.alpha[None], resid[None], tl, pre, bins, img)
