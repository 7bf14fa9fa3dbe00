"""Reader of a Touch-GS scene directory: what the reference's pre-processing scripts leave
on disk and its trainer's dataparser picks up, turned into the device tensors the rasterizer / trainer consume.

On-disk contract (all pinned in the reference tree):

* ``transforms.json`` -- nerfstudio format: ``fl_x, fl_y, cx, cy, w, h`` and ``frames[i] = {file_path, transform_matrix
  [, depth_file_path, uncertainty_file_path]}``; the last two keys are added by reference
  ``utils/add_depth_file_path_to_transforms.py:36-50`` as ``<template>/<image file name>``; the depth key is read by
  reference ``legacy/dataparser_tactile.py:159-162``.
* depth / uncertainty images: 16-bit grayscale PNG, depth in MILLIMETRES, uncertainty sigma x 1000, 0 = invalid
  (reference ``utils/fuse_touch_vision.py:372-388`` ``save()``; ``utils/read_touch_depths.py:48-56``); decoded with
  ``depth_unit_scale_factor = 1e-3`` (reference ``legacy/dataparser_tactile.py:65-66``) times the pose scale factor
  (reference ``legacy/dataparser_tactile.py:229-235``: ``1 / max|t|`` when ``auto_scale_poses``, times ``scale_factor``).
* train / eval split: ``i_train = linspace(0, n-1, ceil(n * fraction))`` (reference
  ``legacy/dataparser_tactile.py:199-214``).
* ``points_touch.npy`` [M,3] world-space points back-projected from the touch depths and ``points_colors.npy`` [M,3]
  colours x 255 (reference ``utils/create_point_cloud_from_touches.py:171,243-244``): the seed cloud of the Gaussians.
* camera convention: ``transform_matrix`` is camera-to-world, OpenGL axes; the flip to the OpenCV axes the rasterizer
  uses is ``diag(1,-1,-1)`` (reference ``utils/create_point_cloud_from_touches.py:64``).

The PNG decoder is self-contained (zlib + numpy; 8/16-bit grayscale and 8-bit RGB[A], non-interlaced: what ``cv2.imwrite``
produces for these arrays), so the product does not depend on OpenCV.  The uint16 -> fp32 decode and the sigma -> weight
map run on the GPU (``tgs_decode_touch_maps``); there is no CPU fallback for them.
"""
from __future__ import annotations

import json
import math
import os
import struct
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .synth import Camera, SH_C0

DEPTH_UNIT_SCALE_FACTOR = 1e-3        # reference legacy/dataparser_tactile.py:65-66
WEIGHT_MODES = {"SIMPLE_LOSS": 0, "DEPTH_UNCERTAINTY_WEIGHTED_LOSS": 1, "inverse_variance": 2}


# --------------------------------------------------------------------------------------------- PNG
def read_png(path: str) -> np.ndarray:
    """Decode a non-interlaced PNG: grayscale 8/16 bit -> [H,W] uint8/uint16; RGB / RGBA 8 bit -> [H,W,3|4] uint8."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(data):
        (n,), typ = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
        pos += 12 + n
    if hdr is None:
        raise ValueError(f"{path}: no IHDR chunk")
    W, H, depth, ctype, _, _, interlace = hdr
    channels = {0: 1, 2: 3, 6: 4, 4: 2}.get(ctype)
    if channels is None or interlace != 0 or depth not in (8, 16) or (depth == 16 and ctype != 0):
        raise ValueError(f"{path}: unsupported PNG (colour type {ctype}, bit depth {depth}, interlace {interlace})")
    bpp = channels * depth // 8                        # bytes per pixel = filter distance
    stride = W * bpp
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8)
    if raw.size != H * (stride + 1):
        raise ValueError(f"{path}: corrupt image data")
    rows = raw.reshape(H, stride + 1)
    ftype = rows[:, 0]
    cur = rows[:, 1:].astype(np.uint8).copy()
    prev = np.zeros(stride, dtype=np.uint8)
    for y in range(H):
        ft, line = int(ftype[y]), cur[y]
        if ft == 0:
            pass
        elif ft == 2:                                  # Up
            line += prev
        elif ft == 1:                                  # Sub: running sum per byte lane
            for c in range(bpp):
                np.cumsum(line[c::bpp], out=line[c::bpp], dtype=np.uint8)
        elif ft in (3, 4):                             # Average / Paeth: inherently sequential along the row
            l16 = line.astype(np.int32)
            p16 = prev.astype(np.int32)
            for x in range(stride):
                a = l16[x - bpp] if x >= bpp else 0
                b = p16[x]
                if ft == 3:
                    pred = (a + b) >> 1
                else:
                    c = p16[x - bpp] if x >= bpp else 0
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                l16[x] = (l16[x] + pred) & 0xFF
            line[:] = l16.astype(np.uint8)
        else:
            raise ValueError(f"{path}: bad filter type {ft}")
        prev = line
    if depth == 16:                                    # big-endian samples
        be = cur.reshape(H, W, 2).astype(np.uint16)
        return (be[..., 0] << 8) | be[..., 1]
    return cur.reshape(H, W) if channels == 1 else cur.reshape(H, W, channels)


def write_png_u16(path: str, img: np.ndarray) -> None:
    """16-bit grayscale PNG (filter 0), the format of the reference's depth / uncertainty maps (test fixtures, exports)."""
    img = np.ascontiguousarray(img, dtype=np.uint16)
    H, W = img.shape
    be = img.astype(">u2").tobytes()
    rows = b"".join(b"\x00" + be[y * 2 * W:(y + 1) * 2 * W] for y in range(H))

    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 16, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(rows, 6)) + chunk(b"IEND", b""))


# ------------------------------------------------------------------------------------- transforms
@dataclass
class Frame:
    file_path: str
    transform_matrix: np.ndarray                 # [4,4] camera-to-world, OpenGL axes
    depth_file_path: Optional[str] = None
    uncertainty_file_path: Optional[str] = None


@dataclass
class SceneMeta:
    fl_x: float
    fl_y: float
    cx: float
    cy: float
    w: int
    h: int
    frames: List[Frame] = field(default_factory=list)
    root: str = "."


def load_transforms(path: str) -> SceneMeta:
    """``transforms.json`` with the per-frame keys of reference ``utils/add_depth_file_path_to_transforms.py:36-50``."""
    with open(path) as f:
        data = json.load(f)
    frames = []
    for fr in data["frames"]:
        frames.append(Frame(fr["file_path"], np.array(fr["transform_matrix"], dtype=np.float64),
                            fr.get("depth_file_path"), fr.get("uncertainty_file_path")))
    if "camera_angle_x" in data and "fl_x" not in data:            # blender-style file: derive the intrinsics
        w, h = int(data.get("w", 800)), int(data.get("h", 800))
        fl = 0.5 * w / math.tan(0.5 * float(data["camera_angle_x"]))
        data = dict(data, fl_x=fl, fl_y=fl, cx=w / 2.0, cy=h / 2.0, w=w, h=h)
    return SceneMeta(float(data["fl_x"]), float(data["fl_y"]), float(data["cx"]), float(data["cy"]),
                     int(data["w"]), int(data["h"]), frames, os.path.dirname(os.path.abspath(path)))


def split_indices(num_images: int, train_split_fraction: float = 0.9) -> Tuple[np.ndarray, np.ndarray]:
    """Reference ``legacy/dataparser_tactile.py:199-214``: equally spaced training images incl. first and last."""
    num_train = math.ceil(num_images * train_split_fraction)
    i_all = np.arange(num_images)
    i_train = np.linspace(0, num_images - 1, num_train, dtype=int)
    i_eval = np.setdiff1d(i_all, i_train)
    return i_train, i_eval


def touch_cloud_split(num_images: int, train_split_fraction: float = 0.9) -> Tuple[np.ndarray, np.ndarray]:
    """The split variant of reference ``utils/create_point_cloud_from_touches.py:174-198`` (which views feed the touch
    seed cloud): ``linspace(0, n-1, ceil(n f) + 1)`` without its last value."""
    num_train = math.ceil(num_images * train_split_fraction)
    i_train = np.linspace(0, num_images - 1, num_train + 1, dtype=int)[:-1]
    return i_train, np.setdiff1d(np.arange(num_images), i_train)


def back_project_touch_points(depth_m: np.ndarray, color_rgb: np.ndarray, intrinsics, c2w_gl: np.ndarray):
    """Vectorised restatement of reference ``utils/create_point_cloud_from_touches.py:19-73``: pixels with depth != 0 ->
    world points ``R diag(1,-1,-1) [ (u-cx) Z / fx, (v-cy) Z / fy, Z ] + t`` (row-major pixel order) and colours / 255."""
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    depth_m = np.asarray(depth_m, dtype=np.float64)
    v, u = np.nonzero(depth_m != 0)
    Z = depth_m[v, u]
    P = np.stack([(u - cx) * Z / fx, (v - cy) * Z / fy, Z], 0)
    c2w = np.asarray(c2w_gl, dtype=np.float64)
    R = c2w[:3, :3] @ np.diag([1.0, -1.0, -1.0])
    pts = (R @ P + c2w[:3, 3:4]).T
    return pts, np.asarray(color_rgb)[v, u] / 255.0


def pose_scale_factor(c2w: np.ndarray, auto_scale_poses: bool = True, scale_factor: float = 1.0) -> float:
    """Reference ``legacy/dataparser_tactile.py:229-235``: ``1 / max|translation|`` (if auto) times ``scale_factor``.
    The same factor multiplies the metric depths, so poses and depth targets stay consistent."""
    s = 1.0
    if auto_scale_poses:
        s /= float(np.max(np.abs(c2w[:, :3, 3])))
    return s * scale_factor


def camera_from_c2w(c2w_gl: np.ndarray, meta: SceneMeta, znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """Operator camera (transposed view / full projection, tan of the half fov) from a camera-to-world matrix in
    OpenGL axes; the axis flip is reference ``utils/create_point_cloud_from_touches.py:64`` (``diag(1,-1,-1)``).
    The principal point offset (cx - w/2, cy - h/2) is returned by :meth:`TouchGSDataset.principal_offset`."""
    c2w = np.array(c2w_gl, dtype=np.float64)
    c2w[:3, 1:3] *= -1.0                                           # OpenGL (y up, z back) -> OpenCV (y down, z forward)
    w2c = np.linalg.inv(c2w)
    W, H = meta.w, meta.h
    tanx, tany = 0.5 * W / meta.fl_x, 0.5 * H / meta.fl_y
    P = np.zeros((4, 4))
    P[0, 0], P[1, 1] = 1.0 / tanx, 1.0 / tany
    P[2, 2], P[2, 3] = zfar / (zfar - znear), -(zfar * znear) / (zfar - znear)
    P[3, 2] = 1.0
    full = P @ w2c
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).float()
    return Camera(W, H, float(tanx), float(tany), t(w2c), t(full), torch.from_numpy(c2w[:3, 3].copy()).float())


# ------------------------------------------------------------------------------------ device decode
def decode_touch_maps(depth_u16, sigma_u16=None, depth_unit: float = DEPTH_UNIT_SCALE_FACTOR,
                      depth_loss_type: str = "DEPTH_UNCERTAINTY_WEIGHTED_LOSS", uncertainty_weight: float = 1.0,
                      device=None, want_weight: bool = True):
    """uint16 mm depth (+ uint16 sigma x 1000) -> (touch_depth fp32 [H,W], touch_weight fp32 [H,W] | None) on the GPU.
    ``depth_unit`` = 1e-3 * pose scale.  ``depth_loss_type`` / ``uncertainty_weight``: reference
    ``scripts/train_bunny_real.sh:52``, ``scripts/train_block_data.sh:50``."""
    if depth_loss_type not in WEIGHT_MODES:
        raise ValueError(f"depth_loss_type must be one of {list(WEIGHT_MODES)}")
    mode = WEIGHT_MODES[depth_loss_type]
    lib = L.load()

    def to_dev(a):
        if a is None:
            return None
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16)) if isinstance(a, np.ndarray) else a
        if t.dtype not in (torch.uint16, torch.int16):
            raise ValueError(f"depth / uncertainty maps must be uint16, got {t.dtype}")
        dev = torch.device(device) if device is not None else (t.device if t.device.type == "cuda" else torch.device("cuda"))
        return t.to(dev, non_blocking=True).contiguous()
    d, s = to_dev(depth_u16), to_dev(sigma_u16)
    if d.device.type != "cuda":
        raise RuntimeError("decode_touch_maps is CUDA-only (no CPU fallback)")
    if s is not None and s.shape != d.shape:
        raise ValueError("depth and uncertainty maps must have the same shape")
    with torch.cuda.device(d.device):
        target = torch.empty(d.shape, dtype=torch.float32, device=d.device)
        weight = torch.empty(d.shape, dtype=torch.float32, device=d.device) if (want_weight and (mode == 0 or s is not None)) else None
        import ctypes as C
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        L.check(lib.tgs_decode_touch_maps(p(d), p(s), d.numel(), float(depth_unit), float(uncertainty_weight), mode,
                                          p(target), p(weight), C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)),
                "tgs_decode_touch_maps")
    return target, weight


# ---------------------------------------------------------------------------------------- seeding
def seed_gaussians(points: np.ndarray, colors_255: np.ndarray, sh_degree: int = 3, scale: float = 1.0,
                   init_opacity: float = 0.1, max_points: Optional[int] = None, seed: int = 0):
    """Raw trainer parameters from the touch seed cloud (``points_touch.npy`` / ``points_colors.npy``, colours x 255:
    reference ``utils/create_point_cloud_from_touches.py:171,243-244``), the way the splat trainers of that era
    initialise from a point cloud: log-scale = log of the mean distance to the 3 nearest neighbours,
    identity rotations, opacity logit(0.1), SH DC = (colour - 0.5) / C0.
    Returns (means [M,3], shs [M,K,3], opacity_logit [M], scales_log [M,3], quats [M,4]) as CPU float32 tensors."""
    pts = np.asarray(points, dtype=np.float64) * float(scale)
    col = np.asarray(colors_255, dtype=np.float64) / 255.0
    if pts.ndim != 2 or pts.shape[1] != 3 or col.shape != pts.shape:
        raise ValueError(f"points / colours must both be [M,3], got {pts.shape} / {col.shape}")
    if max_points is not None and pts.shape[0] > max_points:
        sel = np.random.default_rng(seed).choice(pts.shape[0], max_points, replace=False)
        sel.sort()
        pts, col = pts[sel], col[sel]
    from scipy.spatial import cKDTree
    k = min(4, pts.shape[0])
    d, _ = cKDTree(pts).query(pts, k=k)
    nn = d[:, 1:].mean(axis=1) if k > 1 else np.full(pts.shape[0], 0.01)
    nn = np.clip(nn, 1e-7, None)
    M, K = pts.shape[0], (sh_degree + 1) ** 2
    shs = np.zeros((M, K, 3))
    shs[:, 0] = (col - 0.5) / SH_C0
    quats = np.zeros((M, 4))
    quats[:, 0] = 1.0
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    return (f32(pts), f32(shs), torch.full((M,), math.log(init_opacity / (1.0 - init_opacity))),
            f32(np.log(nn)[:, None].repeat(3, 1)), f32(quats))


# ---------------------------------------------------------------------------------------- dataset
class TouchGSDataset:
    """One scene: cameras, RGB images, touch depth targets and weights, seed cloud.

    ``data_dir`` holds ``transforms.json`` and the relative paths in it.  ``__getitem__`` returns a dict with the
    operator inputs of one training view, decoded on ``device``."""

    def __init__(self, data_dir: str, split: str = "train", train_split_fraction: float = 0.9, auto_scale_poses: bool = True,
                 scale_factor: float = 1.0, depth_unit_scale_factor: float = DEPTH_UNIT_SCALE_FACTOR,
                 depth_loss_type: str = "DEPTH_UNCERTAINTY_WEIGHTED_LOSS", uncertainty_weight: float = 1.0,
                 device="cuda", transforms: str = "transforms.json"):
        self.meta = load_transforms(os.path.join(data_dir, transforms))
        self.data_dir, self.device = data_dir, torch.device(device)
        n = len(self.meta.frames)
        if n == 0:
            raise ValueError("No image files found. You should check the file_paths in the transforms.json file "
                             "to make sure they are correct.")                      # reference legacy/dataparser_tactile.py:168-173
        i_train, i_eval = split_indices(n, train_split_fraction)
        if split == "train":
            self.indices = i_train
        elif split in ("val", "test"):
            self.indices = i_eval
        else:
            raise ValueError(f"Unknown dataparser split {split}")                   # reference legacy/dataparser_tactile.py:214
        c2w = np.stack([f.transform_matrix for f in self.meta.frames])
        self.scale = pose_scale_factor(c2w, auto_scale_poses, scale_factor)           # computed over ALL frames, then split
        self.depth_unit = depth_unit_scale_factor * self.scale
        self.depth_loss_type, self.uncertainty_weight = depth_loss_type, uncertainty_weight
        c2w = c2w.copy()
        c2w[:, :3, 3] *= self.scale
        self.c2w = c2w

    def __len__(self):
        return len(self.indices)

    @property
    def principal_offset(self) -> Tuple[float, float]:
        return self.meta.cx - 0.5 * self.meta.w, self.meta.cy - 0.5 * self.meta.h

    def _path(self, rel: str) -> str:
        return rel if os.path.isabs(rel) else os.path.join(self.data_dir, rel)

    def camera(self, i: int) -> Camera:
        return camera_from_c2w(self.c2w[int(self.indices[i])], self.meta)

    def __getitem__(self, i: int) -> Dict[str, object]:
        fr = self.meta.frames[int(self.indices[i])]
        out: Dict[str, object] = {"camera": self.camera(i), "file_path": fr.file_path}
        ip = self._path(fr.file_path)
        if os.path.exists(ip):
            img = read_png(ip)
            if img.ndim == 2:
                img = np.repeat(img[..., None], 3, -1)
            rgb = torch.from_numpy(np.ascontiguousarray(img[..., :3])).to(self.device)
            out["image"] = (rgb.permute(2, 0, 1).float() / 255.0).contiguous()          # [3,H,W] like the operator's output
        if fr.depth_file_path is not None:
            d = read_png(self._path(fr.depth_file_path))
            s = read_png(self._path(fr.uncertainty_file_path)) if fr.uncertainty_file_path is not None else None
            if d.dtype != np.uint16 or (s is not None and s.dtype != np.uint16):
                raise ValueError("depth / uncertainty maps must be 16-bit PNGs (millimetres / sigma x 1000)")
            out["touch_depth"], out["touch_weight"] = decode_touch_maps(
                d, s, self.depth_unit, self.depth_loss_type, self.uncertainty_weight, self.device)
        return out

    def seed_points(self, sh_degree: int = 3, max_points: Optional[int] = None):
        """Gaussians seeded from ``points_touch.npy`` / ``points_colors.npy`` (scaled like the poses)."""
        pp, pc = os.path.join(self.data_dir, "points_touch.npy"), os.path.join(self.data_dir, "points_colors.npy")
        if not (os.path.exists(pp) and os.path.exists(pc)):
            raise FileNotFoundError(f"{pp} / {pc}: run the reference's create_point_cloud_from_touches.py first")
        return seed_gaussians(np.load(pp), np.load(pc), sh_degree, self.scale, max_points=max_points)
