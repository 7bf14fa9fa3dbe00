"""Seeded synthetic scenes, cameras and touch-depth maps.

No dataset is reachable (reference submodule ``touch-gs-data`` is empty, reference
``.gitmodules:1-3``), so every test and bench input is generated here.  The touch
target / uncertainty encodings follow the reference's on-disk formats:

* depth target: millimetre-quantised, 0 = invalid (reference
  ``utils/fuse_touch_vision.py:372-388``, ``utils/read_touch_depths.py:48-56``);
* uncertainty sigma: touched pixels have tiny sigma, elsewhere the vision heuristic
  ``clip(0.05*depth, 0, 10) + 5`` (reference ``utils/fuse_touch_vision.py:310-313``);
* weight handed to the kernel = 1 / sigma.

All tensors are created on CPU with ``torch.Generator(seed)`` and moved by the caller.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch

SH_C0 = 0.28209479177387814

# BASELINE.json configs c1..c5: (N, W, H, sh_degree, log-scale range)
CONFIGS = {
    "c1": dict(N=1_000, W=128, H=128, sh_degree=0, smin=0.02, smax=0.2),
    "c2": dict(N=100_000, W=800, H=800, sh_degree=3, smin=0.004, smax=0.04),
    "c3": dict(N=1_000_000, W=1920, H=1080, sh_degree=3, smin=0.002, smax=0.02),
    "c5": dict(N=5_000_000, W=3840, H=2160, sh_degree=3, smin=0.002, smax=0.02),
    # Gaussians seeded from the reference's sample point cloud, cameras from its sample poses
    "fixture": dict(N=71_283, W=800, H=800, sh_degree=3, smin=0.0, smax=0.0, fixture_copies=1),
    "fixture1m": dict(N=71_283 * 14, W=1920, H=1080, sh_degree=3, smin=0.0, smax=0.0, fixture_copies=14),
}


class Scene(NamedTuple):
    means3D: torch.Tensor     # [N,3]
    scales: torch.Tensor      # [N,3] (already exponentiated)
    rotations: torch.Tensor   # [N,4] unit quaternions (w,x,y,z)
    opacities: torch.Tensor   # [N,1] in (0,1)
    shs: torch.Tensor         # [N,K,3]
    sh_degree: int


class Camera(NamedTuple):
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor  # [4,4] TRANSPOSED world->view
    projmatrix: torch.Tensor  # [4,4] TRANSPOSED full projection
    campos: torch.Tensor      # [3]


def make_scene(N: int, sh_degree: int = 3, smin: float = 0.002, smax: float = 0.02,
               seed: int = 0, sh_rest_std: float = 0.1) -> Scene:
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(N, 3, generator=g) * 2.0 - 1.0
    ls = torch.rand(N, 3, generator=g) * (math.log(smax) - math.log(smin)) + math.log(smin)
    scales = torch.exp(ls)
    q = torch.randn(N, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    opac = torch.sigmoid(torch.randn(N, 1, generator=g) * 2.0)
    K = (sh_degree + 1) ** 2
    shs = torch.randn(N, K, 3, generator=g) * sh_rest_std
    shs[:, 0, :] = (torch.rand(N, 3, generator=g) - 0.5) / SH_C0
    return Scene(means.contiguous(), scales.contiguous(), q.contiguous(), opac.contiguous(),
                 shs.contiguous(), sh_degree)


def look_at_camera(W: int, H: int, eye, target=(0.0, 0.0, 0.0), fovx_deg: float = 60.0,
                   znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """Pinhole camera, +z forward / +y down in view space (the convention the rasterizer's
    NDC->pixel map assumes).  Matrices are returned TRANSPOSED (row-vector convention)."""
    eye_t = torch.tensor(eye, dtype=torch.float64)
    tgt = torch.tensor(target, dtype=torch.float64)
    fwd = tgt - eye_t
    fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    if abs(float(fwd @ up)) > 0.999:
        up = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    Rm = torch.stack([right, down, fwd], 0)          # world->view rotation rows
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = Rm
    V[:3, 3] = -Rm @ eye_t
    tanx = math.tan(math.radians(fovx_deg) * 0.5)
    tany = tanx * H / W
    P = torch.zeros(4, 4, dtype=torch.float64)
    P[0, 0] = 1.0 / tanx
    P[1, 1] = 1.0 / tany
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    P[3, 2] = 1.0
    full = P @ V
    return Camera(W, H, tanx, tany, V.t().contiguous().float(), full.t().contiguous().float(),
                  eye_t.float())


def orbit_cameras(W: int, H: int, n: int = 8, radius: float = 3.0, seed: int = 0,
                  fovx_deg: float = 60.0):
    """n seeded orbit poses looking at the origin from distance ``radius``."""
    g = torch.Generator().manual_seed(seed + 12345)
    cams = []
    for i in range(n):
        az = 2.0 * math.pi * (i + float(torch.rand(1, generator=g))) / n
        el = math.radians(float(torch.rand(1, generator=g)) * 40.0 - 20.0)
        eye = (radius * math.cos(el) * math.sin(az), radius * math.sin(el), -radius * math.cos(el) * math.cos(az))
        cams.append(look_at_camera(W, H, eye, fovx_deg=fovx_deg))
    return cams


def make_touch_maps(rendered_depth: torch.Tensor, seed: int = 0, n_patches: int = 10,
                    patch_radius: int = 32, invalid_frac: float = 0.05,
                    touch_sigma: float = 0.005):
    """Build (target [H,W], weight [H,W]) from an expected-depth render of a perturbed scene.

    target: quantised to 1 mm, 0 where invalid (no coverage, or a seeded 5 % dropout).
    sigma : ``touch_sigma`` inside ``n_patches`` random discs ("touches"), elsewhere
            clip(0.05*depth,0,10)+5 ;  weight = 1/sigma."""
    d = rendered_depth.detach().float().cpu()
    H, W = d.shape
    g = torch.Generator().manual_seed(seed + 777)
    target = torch.round(torch.clamp_min(d, 0.0) * 1000.0).clamp(0, 65535) / 1000.0
    drop = torch.rand(H, W, generator=g) < invalid_frac
    target = torch.where(drop, torch.zeros_like(target), target)
    sigma = torch.clamp(0.05 * target, 0.0, 10.0) + 5.0
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    for _ in range(n_patches):
        cx = int(torch.randint(0, W, (1,), generator=g))
        cy = int(torch.randint(0, H, (1,), generator=g))
        disc = (xs - cx) ** 2 + (ys - cy) ** 2 <= patch_radius ** 2
        sigma = torch.where(disc, torch.full_like(sigma, touch_sigma), sigma)
    weight = 1.0 / sigma
    return target.contiguous(), weight.contiguous()


def perturbed(scene: Scene, std: float = 0.01, seed: int = 0) -> Scene:
    g = torch.Generator().manual_seed(seed + 4242)
    return scene._replace(means3D=(scene.means3D + torch.randn(scene.means3D.shape, generator=g) * std).contiguous())


def config_scene(name: str, seed: int = 0, N: Optional[int] = None):
    c = CONFIGS[name]
    n = c["N"] if N is None else N
    scene = make_scene(n, c["sh_degree"], c["smin"], c["smax"], seed)
    cams = orbit_cameras(c["W"], c["H"], 8, 3.0, seed)
    return scene, cams


# ------------------------------------------------------------------ realistic fixture
def _fixture_path(path: Optional[str] = None) -> str:
    import os
    return path or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                "fixture_scene.npz")


def fixture_scene(sh_degree: int = 3, copies: int = 1, seed: int = 0, path: Optional[str] = None) -> Scene:
    """Gaussians initialised from the reference's sample point cloud (``sample_pc_data/sparse.ply``, 71 283 coloured
    points; committed as ``tests/golden/fixture_scene.npz`` by ``tests/golden/make_fixture_scene.py``) the way the
    splat trainers of that era seed a scene: mean = point, isotropic scale = mean distance to the 3 nearest
    neighbours, opacity 0.1, SH DC = (rgb - 0.5)/C0, higher bands 0, random orientation.
    ``copies`` > 1 replicates every point with a jitter of its own neighbour distance (a denser cloud with the same
    spatial distribution: 14 copies ~ 1M Gaussians)."""
    import numpy as np
    z = np.load(_fixture_path(path))
    g = torch.Generator().manual_seed(seed)
    pts = torch.from_numpy(z["points"]).float()
    col = torch.from_numpy(z["colors"]).float() / 255.0
    knn = torch.from_numpy(z["knn_dist"]).float().clamp_min(1e-4)
    if copies > 1:
        pts = pts.repeat(copies, 1)
        col = col.repeat(copies, 1)
        knn = knn.repeat(copies)
        jit = torch.randn(pts.shape, generator=g) * knn[:, None]
        jit[: z["points"].shape[0]] = 0.0
        pts = pts + jit
        knn = knn / float(copies) ** (1.0 / 3.0)
    N = pts.shape[0]
    K = (sh_degree + 1) ** 2
    shs = torch.zeros(N, K, 3)
    shs[:, 0] = (col - 0.5) / SH_C0
    q = torch.randn(N, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    return Scene(pts.contiguous(), knn[:, None].repeat(1, 3).contiguous(), q.contiguous(),
                 torch.full((N, 1), 0.1), shs.contiguous(), sh_degree)


def fixture_cameras(W: int, H: int, n: int = 8, path: Optional[str] = None, znear: float = 0.01, zfar: float = 100.0):
    """The first ``n`` (of 100, evenly strided) poses of the reference's ``sample_blender_data/transforms_train.json``
    with its ``camera_angle_x``.  Blender / OpenGL camera-to-world (x right, y up, looking down -z) -> our view
    convention (+z forward, +y down) with the diag(1,-1,-1) flip of reference
    ``utils/create_point_cloud_from_touches.py:64``."""
    import numpy as np
    z = np.load(_fixture_path(path))
    poses = torch.from_numpy(z["poses_c2w"]).double()
    ang = float(z["camera_angle_x"])
    idx = torch.linspace(0, poses.shape[0] - 1, n).round().long()
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], dtype=torch.float64))
    tanx = math.tan(0.5 * ang)
    tany = tanx * H / W
    cams = []
    for i in idx.tolist():
        V = flip @ torch.linalg.inv(poses[i])
        P = torch.zeros(4, 4, dtype=torch.float64)
        P[0, 0], P[1, 1] = 1.0 / tanx, 1.0 / tany
        P[2, 2], P[2, 3], P[3, 2] = zfar / (zfar - znear), -(zfar * znear) / (zfar - znear), 1.0
        cams.append(Camera(W, H, tanx, tany, V.t().contiguous().float(), (P @ V).t().contiguous().float(),
                           poses[i][:3, 3].float()))
    return cams


def zbuffer_depth(points: torch.Tensor, cam: Camera) -> torch.Tensor:
    """Point-cloud -> depth image by z-buffering, the logic of reference
    ``data_preprocessing/vision/point_cloud/read_point_cloud.py:224-266`` (``project_points_with_colors``): pinhole
    projection with fx = W / (2 tan(fovx/2)), principal point at the image centre, ``int()`` truncation of (u, v),
    nearest depth wins, 0 where no point lands."""
    W, H = cam.image_width, cam.image_height
    V = cam.viewmatrix.t().double()
    pc = points.double() @ V[:3, :3].t() + V[:3, 3]
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    ok = pc[:, 2] > 0
    pc = pc[ok]
    u = torch.trunc(fx * pc[:, 0] / pc[:, 2] + W / 2.0).long()
    v = torch.trunc(fy * pc[:, 1] / pc[:, 2] + H / 2.0).long()
    inside = (u >= 0) & (u < W) & (v >= 0) & (v < H)
    flat = (v * W + u)[inside]
    depth = torch.full((H * W,), float("inf"), dtype=torch.float64)
    depth.scatter_reduce_(0, flat, pc[inside, 2], reduce="amin")
    depth = torch.where(torch.isinf(depth), torch.zeros_like(depth), depth)
    return depth.reshape(H, W).float()
