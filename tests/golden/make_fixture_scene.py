"""Build tests/golden/fixture_scene.npz from the reference's OWN sample data (run in the build container only;
/root/reference does not exist on the GPU box):

* seed points + colours: reference data_preprocessing/vision/point_cloud/sample_pc_data/sparse.ply
  (71 283 coloured COLMAP points -- SURVEY.md §8c "usable fixtures", §8f N4);
* camera poses: reference data_preprocessing/vision/point_cloud/sample_blender_data/transforms_train.json
  (100 camera-to-world matrices, camera_angle_x = 0.6911).

The two files live in different world frames (the reference aligns them with
data_preprocessing/vision/colmap/compute_colmap_blender_transform.py from data we do not have), so the cloud is
normalised into the pose frame: all 100 optical axes meet in one point (the object the poses orbit at distances
0.35-0.75; found by least squares), so the cloud's centroid is moved there and its 95th-percentile radius set to
0.2 (it then fills the 39.6 degree field of view from the mean camera distance).  Also stored: each point's mean
distance to its 3 nearest neighbours (the initial Gaussian scale of the splat trainers of that era).

    python tests/golden/make_fixture_scene.py
"""
import json
import os

import numpy as np
from scipy.spatial import cKDTree

REF = "/root/reference/data_preprocessing/vision/point_cloud"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixture_scene.npz")


def read_ply(path):
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
        v = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return np.stack([v["x"], v["y"], v["z"]], -1).astype(np.float32), np.stack([v["r"], v["g"], v["b"]], -1)


def main():
    pts, col = read_ply(os.path.join(REF, "sample_pc_data", "sparse.ply"))
    tr = json.load(open(os.path.join(REF, "sample_blender_data", "transforms_train.json")))
    poses = np.array([f["transform_matrix"] for f in tr["frames"]], dtype=np.float32)
    o, d = poses[:, :3, 3].astype(np.float64), -poses[:, :3, 2].astype(np.float64)
    A, b = np.zeros((3, 3)), np.zeros(3)
    for oi, di in zip(o, d):                      # point closest to all optical axes
        M = np.eye(3) - np.outer(di, di)
        A += M
        b += M @ oi
    focus = np.linalg.solve(A, b)
    c = pts.mean(0)
    r95 = np.percentile(np.linalg.norm(pts - c, axis=1), 95)
    pts_n = ((pts - c) / r95 * 0.2 + focus).astype(np.float32)
    d, _ = cKDTree(pts_n).query(pts_n, k=4)
    knn = d[:, 1:].mean(1).astype(np.float32)
    np.savez_compressed(OUT, points=pts_n, colors=col, poses_c2w=poses, camera_angle_x=np.float32(tr["camera_angle_x"]),
                        knn_dist=knn, focus=focus.astype(np.float32))
    print(OUT, pts_n.shape, poses.shape, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
