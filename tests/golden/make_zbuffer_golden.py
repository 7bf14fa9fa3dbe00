"""Golden vector for synth.zbuffer_depth, produced by the REFERENCE'S OWN function
``project_points_with_colors`` (reference data_preprocessing/vision/point_cloud/read_point_cloud.py:224-266).
The module imports open3d (not installed), so only that function's source is extracted with ``ast`` and executed
with numpy.  Build container only (/root/reference does not exist on the GPU box):

    python tests/golden/make_zbuffer_golden.py
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/data_preprocessing/vision/point_cloud/read_point_cloud.py"


def reference_function():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "project_points_with_colors")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["project_points_with_colors"]


def main():
    import touchgs_b200 as T
    f = reference_function()
    sc = T.synth.fixture_scene(0)
    W, H = 160, 120
    out = {}
    for i, cam in enumerate(T.synth.fixture_cameras(W, H, 3)):
        V = cam.viewmatrix.t().double().numpy()                      # world -> camera, +z forward, +y down
        fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
        K = np.array([[fx, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1.0]])
        pts = sc.means3D.numpy()[::7]
        _, _, depth = f(pts, np.zeros((pts.shape[0], 3)), K, V, W, H)
        out[f"depth_{i}"] = depth.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "zbuffer_reference.npz"), **out)
    print({k: (v.shape, float((v > 0).mean())) for k, v in out.items()})


if __name__ == "__main__":
    main()
