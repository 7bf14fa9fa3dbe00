"""touchgs_b200 -- B200-native differentiable Gaussian-splat rasterizer with fused touch-depth
supervision (the one hot path of Touch-GS training).

The directory is called ``touch-gs_b200`` (not importable by a plain ``import`` statement because of
the hyphen); use ``import touchgs_b200`` (top-level shim) or ``import diff_gaussian_rasterization``
(drop-in alias with the reference-era operator names).
"""
from . import _lib
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, TouchOptions,
                         rasterize_gaussians, _RasterizeGaussians)
from . import synth, sharding, inspect_state, touch_inputs, refstructure, train_step, dataset
from .train_step import TouchGSTrainer, TrainConfig, photometric_loss, adam_step, activate, densify

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "TouchOptions", "rasterize_gaussians",
           "synth", "sharding", "inspect_state", "touch_inputs", "dataset", "refstructure", "train_step", "TouchGSTrainer", "TrainConfig", "photometric_loss",
           "adam_step", "activate", "densify", "_lib"]
