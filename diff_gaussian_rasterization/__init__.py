"""Drop-in alias: the operator names a reference-era splat model imports
(``from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer``),
served by the B200-native implementation in ``touch-gs_b200/``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import touchgs_b200 as _t  # noqa: E402

GaussianRasterizationSettings = _t.GaussianRasterizationSettings
GaussianRasterizer = _t.GaussianRasterizer
rasterize_gaussians = _t.rasterize_gaussians
TouchOptions = _t.TouchOptions
__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "TouchOptions"]
