"""the dataset reader against files written by the REFERENCE's own code
(tests/golden/make_dataset_golden.py: add_depth_file_path_to_transforms.py as a subprocess, fuse_touch_vision.save(),
get_point_cloud_from_depth_and_color, get_train_eval_split_fraction; arrays read back with cv2 as the reference does)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import T, ROOT

D = T.dataset
SCENE = os.path.join(ROOT, "tests", "golden", "dataset_scene")
Z = np.load(os.path.join(ROOT, "tests", "golden", "dataset_reference.npz"))


def test_transforms_keys_added_by_the_reference_script():
    meta = D.load_transforms(os.path.join(SCENE, "transforms.json"))
    assert (meta.w, meta.h) == (48, 36) and len(meta.frames) == 7
    for i, fr in enumerate(meta.frames):
        name = fr.file_path.split("/")[-1]
        assert fr.depth_file_path == f"fused_depth/{name}"                        # <template>/<image file name>
        assert fr.uncertainty_file_path == f"fused_depth_uncertainty/{name}"
        assert fr.transform_matrix.shape == (4, 4)


def test_png_decoder_matches_what_the_reference_reads_back():
    for i in range(7):
        d = D.read_png(os.path.join(SCENE, "fused_depth", f"{i:04d}.png"))
        s = D.read_png(os.path.join(SCENE, "fused_depth_uncertainty", f"{i:04d}.png"))
        assert d.dtype == np.uint16 and np.array_equal(d, Z[f"depth_u16_{i}"])
        assert s.dtype == np.uint16 and np.array_equal(s, Z[f"sigma_u16_{i}"])
        img = D.read_png(os.path.join(SCENE, "images", f"{i:04d}.png"))
        assert np.array_equal(img, Z[f"image_rgb_{i}"])


def test_png_all_filter_types_round_trip(tmp_path):
    """cv2 / libpng may pick any of the five row filters: decode files written with each of them."""
    import struct
    import zlib
    rng = np.random.default_rng(0)
    img = (rng.random((9, 13)) * 65535).astype(np.uint16)
    be = img.astype(">u2").tobytes()
    bpp, stride = 2, 26

    def filt(ft, row, prev):
        out = bytearray(stride)
        for x in range(stride):
            a = row[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            if ft == 0: pred = 0
            elif ft == 1: pred = a
            elif ft == 2: pred = b
            elif ft == 3: pred = (a + b) >> 1
            else:
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[x] = (row[x] - pred) & 0xFF
        return bytes(out)
    raw, prev = b"", bytes(stride)
    for y in range(9):
        row = be[y * stride:(y + 1) * stride]
        ft = y % 5
        raw += bytes([ft]) + filt(ft, row, prev)
        prev = row
    chunk = lambda t, b: struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xFFFFFFFF)
    p = tmp_path / "f.png"
    p.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 13, 9, 16, 0, 0, 0, 0)) +
                  chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))
    assert np.array_equal(D.read_png(str(p)), img)
    q = tmp_path / "w.png"
    D.write_png_u16(str(q), img)
    assert np.array_equal(D.read_png(str(q)), img)
    import cv2
    assert np.array_equal(cv2.imread(str(q), cv2.IMREAD_ANYDEPTH), img)           # and cv2 reads what we write


def test_splits():
    tr, ev = D.split_indices(10, 0.9)                    # reference legacy/dataparser_tactile.py:199-214
    assert tr.tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 9] and ev.tolist() == [8]
    for key in Z.files:
        if key.startswith("cloud_split_train_"):         # outputs of the reference's get_train_eval_split_fraction
            n, f = key.split("_")[3:5]
            tr, ev = D.touch_cloud_split(int(n), int(f) / 100.0)
            assert np.array_equal(tr, Z[key]) and np.array_equal(ev, Z[key.replace("train", "eval")])
    assert any(k.startswith("cloud_split_train_") for k in Z.files)


def test_back_projection_matches_the_reference_and_our_camera_convention():
    meta = D.load_transforms(os.path.join(SCENE, "transforms.json"))
    intr = [meta.fl_x, meta.fl_y, meta.cx, meta.cy]
    for i in (0, 3):
        depth = Z[f"depth_u16_{i}"] / 1000
        pts, col = D.back_project_touch_points(depth, Z[f"image_rgb_{i}"], intr, meta.frames[i].transform_matrix)
        assert np.allclose(pts, Z[f"cloud_points_{i}"], rtol=0, atol=1e-12) and np.array_equal(col, Z[f"cloud_colors_{i}"])
        # the reference's world points, seen through OUR operator camera, land on their pixel with their depth
        cam = D.camera_from_c2w(meta.frames[i].transform_matrix, meta)
        P = torch.from_numpy(pts).float()
        hom = torch.cat([P, torch.ones(len(P), 1)], 1)
        view = hom @ cam.viewmatrix                      # transposed matrices: row vectors
        clip = hom @ cam.projmatrix
        ndc = clip[:, :2] / clip[:, 3:4]
        px = ((ndc[:, 0] + 1) * meta.w - 1) * 0.5 + (meta.cx - meta.w / 2)
        py = ((ndc[:, 1] + 1) * meta.h - 1) * 0.5 + (meta.cy - meta.h / 2)
        v, u = np.nonzero(depth != 0)
        # SURVEY A1 pixel mean ((ndc+1) W - 1)/2 puts pixel centres at integers + the half-pixel of the NDC convention
        assert float((px - (torch.from_numpy(u).float() - 0.5)).abs().max()) < 2e-3
        assert float((py - (torch.from_numpy(v).float() - 0.5)).abs().max()) < 2e-3
        assert float((view[:, 2] - torch.from_numpy(depth[v, u]).float()).abs().max()) < 1e-4
    allp = np.load(os.path.join(SCENE, "points_touch.npy"))
    assert allp.shape == (len(Z["cloud_points_0"]) + len(Z["cloud_points_3"]), 3)


def test_pose_scale_and_seed_points():
    meta = D.load_transforms(os.path.join(SCENE, "transforms.json"))
    c2w = np.stack([f.transform_matrix for f in meta.frames])
    s = D.pose_scale_factor(c2w, True, 1.0)
    assert math.isclose(s, 1.0 / np.abs(c2w[:, :3, 3]).max())       # reference legacy/dataparser_tactile.py:229-235
    assert D.pose_scale_factor(c2w, False, 2.5) == 2.5
    means, shs, op, sl, q = D.seed_gaussians(np.load(os.path.join(SCENE, "points_touch.npy")),
                                             np.load(os.path.join(SCENE, "points_colors.npy")), sh_degree=2, scale=s)
    M = means.shape[0]
    assert shs.shape == (M, 9, 3) and op.shape == (M,) and sl.shape == (M, 3) and q.shape == (M, 4)
    col = np.load(os.path.join(SCENE, "points_colors.npy")) / 255.0
    assert np.allclose(shs[:, 0].numpy() * T.synth.SH_C0 + 0.5, col, atol=1e-6)        # SH DC reproduces the colour
    assert float(shs[:, 1:].abs().max()) == 0 and torch.isfinite(sl).all() and float(q[:, 0].min()) == 1.0
    assert np.allclose(means.numpy(), np.load(os.path.join(SCENE, "points_touch.npy")) * s, atol=1e-6)


@pytest.mark.gpu
def test_dataset_decodes_on_the_gpu_and_feeds_the_trainer():
    assert torch.cuda.is_available()
    ds = D.TouchGSDataset(SCENE, split="train", train_split_fraction=0.9, uncertainty_weight=0.01, device="cuda")
    assert len(ds) == 7 and math.isclose(ds.depth_unit, 1e-3 * ds.scale)
    own0, _ = T._lib.launch_counts()
    item = ds[2]
    d16, s16 = Z["depth_u16_2"].astype(np.float32), Z["sigma_u16_2"].astype(np.float32)
    want_t = d16 * np.float32(ds.depth_unit)
    sg = s16 * np.float32(1e-3) * np.float32(0.01)
    want_w = np.where(sg > 0, 1.0 / np.where(sg > 0, sg, 1), 0).astype(np.float32)
    assert np.allclose(item["touch_depth"].cpu().numpy(), want_t, rtol=1e-6, atol=0)
    assert np.allclose(item["touch_weight"].cpu().numpy(), want_w, rtol=1e-6, atol=0)
    assert (item["touch_depth"] == 0).sum() == int((Z["depth_u16_2"] == 0).sum())      # 0 stays "invalid"
    assert item["image"].shape == (3, 36, 48) and T._lib.launch_counts()[0] > own0
    simple, w1 = D.decode_touch_maps(Z["depth_u16_2"], None, ds.depth_unit, "SIMPLE_LOSS", device="cuda")
    assert float((w1 - 1).abs().max()) == 0 and torch.equal(simple, item["touch_depth"])
    # seed the trainer from points_touch.npy / points_colors.npy and take a few steps on the scene's own views
    means, shs, op, sl, q = (t.cuda() for t in ds.seed_points(sh_degree=1))
    cfg = T.TrainConfig(sh_degree=1, depth_loss_mult=0.2, depth_loss_type="DEPTH_UNCERTAINTY_WEIGHTED_LOSS",
                        uncertainty_weight=1.0, refine_every=0, sh_degree_interval=0)
    tr = T.TouchGSTrainer(means, shs, op, sl, q, cfg)
    losses = []
    for it in range(6):
        b = ds[it % len(ds)]
        cam = b["camera"]
        rs = T.GaussianRasterizationSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                                             torch.zeros(3, device="cuda"), 1.0, cam.viewmatrix.cuda(), cam.projmatrix.cuda(),
                                             1, cam.campos.cuda(), False, False)
        losses.append(float(tr.train_step(rs, b["image"], b["touch_depth"], b["touch_weight"])))
    assert all(math.isfinite(l) for l in losses) and tr.last["num_rendered"] > 0
    hit = (tr.last["depth"][0] > 0)
    assert int(hit.sum()) > 100, "the seeded cloud must be visible from its own cameras"
