"""Generate tests/golden/dataset_scene/ + dataset_reference.npz by RUNNING THE REFERENCE's own code on a small seeded
scene (the on-disk formats either side of the hot path).  Only runs in the build container (the
reference tree is not on the GPU box); everything it writes is committed.

  * transforms.json is patched by the reference SCRIPT utils/add_depth_file_path_to_transforms.py (run as a subprocess);
  * the depth / uncertainty PNGs are written by the reference's save() (utils/fuse_touch_vision.py:372-388);
  * points_touch.npy / points_colors.npy come from the reference's get_point_cloud_from_depth_and_color
    (utils/create_point_cloud_from_touches.py:19-73) with its "* 255.0" of :171, saved like :243-244;
  * the expected decoded arrays are what the reference itself reads back: cv2.imread(..., IMREAD_ANYDEPTH) / 1000
    (utils/create_point_cloud_from_touches.py:131-132), and its train/eval split helper (:174-198).
matplotlib / mpl_toolkits / open3d are absent here and only used for visualisation: stubbed for the import.
"""
import json
import os
import shutil
import subprocess
import sys
import types

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
SCENE = os.path.join(HERE, "dataset_scene")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "open3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.path.insert(0, os.path.join(REF, "utils"))
    import fuse_touch_vision as fuse                    # noqa: E402
    import create_point_cloud_from_touches as cloud     # noqa: E402
    return fuse, cloud


if __name__ == "__main__":
    fuse, cloud = import_reference()
    shutil.rmtree(SCENE, ignore_errors=True)
    for d in ("images", "vision_aligned", "vision_aligned_baseline", "fused_depth", "fused_depth_uncertainty"):
        os.makedirs(os.path.join(SCENE, d))
    rng = np.random.default_rng(3)
    H, W, n = 36, 48, 7
    fl, cx, cy = 60.0, 23.25, 17.5
    src = json.load(open(os.path.join(REF, "data_preprocessing/vision/point_cloud/sample_blender_data/transforms_train.json")))
    frames = []
    for i in range(n):
        T = np.array(src["frames"][i * 9]["transform_matrix"])
        frames.append({"file_path": f"images/{i:04d}.png", "transform_matrix": T.tolist()})
    meta = {"fl_x": fl, "fl_y": fl * 1.01, "cx": cx, "cy": cy, "w": W, "h": H, "frames": frames}
    json.dump(meta, open(os.path.join(SCENE, "transforms.json"), "w"), indent=2)
    # ---- the reference script adds depth_file_path / uncertainty_file_path
    subprocess.run([sys.executable, os.path.join(REF, "utils/add_depth_file_path_to_transforms.py"), "--base_repo_path", SCENE,
                    "--filename", "transforms.json", "--depth_file_path_template", "fused_depth",
                    "--uncertainty_file_path_template", "fused_depth_uncertainty"], check=True)
    save = {}
    yy, xx = np.mgrid[0:H, 0:W]
    pts_all, col_all = [], []
    T_by_name, intr = cloud.transforms_utils.read_nerfstudio_transform_positions(os.path.join(SCENE, "transforms.json"),
                                                                                   return_full_transforms=True)
    for i in range(n):
        img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
        cv2.imwrite(os.path.join(SCENE, "images", f"{i:04d}.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        depth = 3.2 + 0.6 * np.sin(xx / 7.0 + i) * np.cos(yy / 5.0) + rng.normal(0, 0.002, (H, W))         # metres
        depth[rng.random((H, W)) < 0.1] = 0.0                                                           # invalid pixels
        sigma = np.where(rng.random((H, W)) < 0.2, np.abs(rng.normal(0.004, 0.002, (H, W))) + 0.001,
                         np.clip(0.05 * depth, 0, 10) + 5.0)
        sigma[0, :4] = 0.0
        # the reference's own writer: vision / baseline maps are not consumed by the trainer, fused + uncertainty are
        fuse.save(os.path.join(SCENE, "vision_aligned"), os.path.join(SCENE, "fused_depth"), f"{i:04d}", depth * 0.9, depth * 0.95,
                  depth, sigma)
        # what the reference reads back (create_point_cloud_from_touches.py:131-132)
        d_back = cv2.imread(os.path.join(SCENE, "fused_depth", f"{i:04d}.png"), cv2.IMREAD_ANYDEPTH)
        s_back = cv2.imread(os.path.join(SCENE, "fused_depth_uncertainty", f"{i:04d}.png"), cv2.IMREAD_ANYDEPTH)
        save[f"depth_u16_{i}"], save[f"sigma_u16_{i}"] = d_back, s_back
        save[f"image_rgb_{i}"] = img
        if i in (0, 3):                                   # seed cloud from two views with the reference's back-projection
            p, c = cloud.get_point_cloud_from_depth_and_color(d_back / 1000, img, intr, T_by_name[f"{i:04d}"])
            pts_all.append(p); col_all.append(c)
            save[f"cloud_points_{i}"], save[f"cloud_colors_{i}"] = p, c
    pts, col = np.concatenate(pts_all), np.concatenate(col_all) * 255.0        # ":171" colours x 255
    np.save(os.path.join(SCENE, "points_touch.npy"), pts)                      # ":243-244"
    np.save(os.path.join(SCENE, "points_colors.npy"), col)
    shutil.rmtree(os.path.join(SCENE, "vision_aligned"))
    shutil.rmtree(os.path.join(SCENE, "vision_aligned_baseline"))
    for nimg, frac in ((7, 0.9), (10, 0.9), (100, 0.9), (23, 0.5), (40, 0.8), (151, 0.9)):
        try:                                  # the helper's own assert fires when its linspace repeats an index
            tr, ev = cloud.get_train_eval_split_fraction(list(range(nimg)), frac)
        except AssertionError:
            print(f"reference split helper asserts for n={nimg}, fraction={frac}: not recorded")
            continue
        save[f"cloud_split_train_{nimg}_{int(frac * 100)}"] = tr
        save[f"cloud_split_eval_{nimg}_{int(frac * 100)}"] = ev
    np.savez_compressed(os.path.join(HERE, "dataset_reference.npz"), **save)
    print("wrote", SCENE, "and dataset_reference.npz;", len(pts), "seed points")
