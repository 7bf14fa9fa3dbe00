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


PARITY_LOG = []      # (test id, tensor name, rel_inf, outlier fraction or None, branch) -- printed by conftest's summary


def assert_close_tensor(a, b, name, rel=1e-4, max_outlier_frac=0.0, outlier_rtol=1e-4):
    """The float-parity bar (north star: 1e-4 rel).  Per tensor: ||a-b||_inf / ||b||_inf <= rel, OR
    (only where the call site grants an outlier budget: a 1-ulp difference between ex2.approx and expf can flip an
    alpha >= 1/255 / T < 1e-4 decision of a (pixel, splat) pair sitting exactly on the threshold, which moves one
    pixel by ~1/255 and the gradients of the splats behind it) at most `max_outlier_frac` of the elements off by
    more than outlier_rtol*|b| + rel*max|b|.  Every call is logged with the branch that passed; the pytest terminal
    summary prints the table (and writes gpurun_out/parity_report.json), so the slack actually used is visible."""
    import os
    r = rel_inf(a, b)
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    if r <= rel:
        PARITY_LOG.append(dict(test=test, tensor=name, rel_inf=r, outlier_frac=None, branch="rel_inf", rel=rel))
        return
    f = frac_outliers(a, b, outlier_rtol, rel)
    ok = f <= max_outlier_frac
    PARITY_LOG.append(dict(test=test, tensor=name, rel_inf=r, outlier_frac=f, branch="outlier_budget" if ok else "FAIL",
                           rel=rel, budget=max_outlier_frac))
    assert ok, f"{name}: rel_inf={r:.3e} > {rel:g} and outlier fraction {f:.3e} > {max_outlier_frac:g}"
