"""Shared test helpers: scene/settings construction and comparison metrics."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import touchgs_b200 as T  # noqa: E402
import oracle as O  # noqa: E402

synth = T.synth


def oracle_settings(cam, sh_degree, bg=(0.0, 0.0, 0.0), mod=1.0):
    return O.OracleSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                            torch.tensor(bg, dtype=torch.float32), mod, cam.viewmatrix, cam.projmatrix,
                            sh_degree, cam.campos)


def cuda_settings(cam, sh_degree, dev, bg=(0.0, 0.0, 0.0), mod=1.0, debug=False):
    return T.GaussianRasterizationSettings(
        cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
        torch.tensor(bg, dtype=torch.float32, device=dev), mod, cam.viewmatrix.to(dev), cam.projmatrix.to(dev),
        sh_degree, cam.campos.to(dev), False, debug)


def rel_inf(a, b, eps=1e-12):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), eps))


def frac_outliers(a, b, rtol=1e-4, atol_rel=1e-4):
    """fraction of elements with |a-b| > rtol*|b| + atol_rel*max|b|"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    tol = rtol * b.abs() + atol_rel * max(float(b.abs().max()), 1e-30)
    return float(((a - b).abs() > tol).double().mean())


def assert_close_tensor(a, b, name, rel=1e-4, max_outlier_frac=0.0, outlier_rtol=1e-4):
    """The float-parity bar (north star: 1e-4 rel).  Per tensor: ||a-b||_inf / ||b||_inf <= rel, OR
    (for image-like tensors where a 1-ulp difference in exp() can flip an alpha >= 1/255 / T < 1e-4
    decision and move one pixel by ~1/255) at most `max_outlier_frac` of the elements off."""
    r = rel_inf(a, b)
    if r <= rel:
        return
    f = frac_outliers(a, b, outlier_rtol, rel)
    assert f <= max_outlier_frac, f"{name}: rel_inf={r:.3e} > {rel:g} and outlier fraction {f:.3e} > {max_outlier_frac:g}"
