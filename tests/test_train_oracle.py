"""CPU checks of the train-step oracle (oracle/train_oracle.py) and of the host logic of the trainer
that needs no GPU.  Parity is UNPINNED by the reference for these rows (the trainer is in its empty submodule), so
the oracle is pinned against torch itself (torch.optim.Adam, autograd, closed forms)."""
import math

import pytest
import torch

from helpers import O, T

TO = O.train_oracle


def test_ssim_window_and_identity():
    w = TO.ssim_window()
    assert w.numel() == 11 and abs(float(w.sum()) - 1.0) < 1e-6 and torch.equal(w, w.flip(0))
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 40, 52, generator=g)
    # identical images: SSIM == 1 everywhere, L1 == 0  ->  loss 0, for any lambda
    assert abs(float(TO.photometric_loss(img, img.clone(), 0.2))) < 1e-6
    # lambda = 0 is the plain L1 mean
    other = torch.rand(3, 40, 52, generator=g)
    assert abs(float(TO.photometric_loss(img, other, 0.0)) - float((img - other).abs().mean())) < 1e-7


def test_photometric_loss_bands_add_up_and_gradcheck():
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(3, 45, 37, generator=g), torch.rand(3, 45, 37, generator=g)
    full = TO.photometric_loss(a, b, 0.2)
    parts = sum(TO.photometric_loss(a, b, 0.2, rows=r) for r in ((0, 16), (16, 32), (32, 45)))
    assert abs(float(full) - float(parts)) < 1e-6
    x = torch.rand(3, 14, 13, generator=g, dtype=torch.float64).requires_grad_(True)
    y = torch.rand(3, 14, 13, generator=g, dtype=torch.float64)
    # away from x == y the loss is smooth: finite differences agree with autograd in float64
    assert torch.autograd.gradcheck(lambda t: TO.photometric_loss(t, y, 0.2), (x,), eps=1e-6, atol=1e-6)


def test_adam_reference_closed_form_first_step():
    p, g = torch.tensor([1.0, -2.0, 3.0]), torch.tensor([0.5, -0.25, 1e-3])
    (q,), (m,), (v,) = TO.adam_reference([p], [g], [1e-2], 1)
    # first Adam step with bias correction moves every element by lr * sign(g)
    assert torch.allclose(q, p - 1e-2 * torch.sign(g), atol=1e-7)
    assert torch.allclose(m, 0.1 * g) and torch.allclose(v, 0.001 * g * g)


def test_densify_reference_invariants():
    g = torch.Generator().manual_seed(2)
    N, K = 500, 4
    means, shs = torch.randn(N, 3, generator=g), torch.randn(N, K, 3, generator=g)
    op = torch.randn(N, generator=g) * 2
    sl = torch.log(torch.rand(N, 3, generator=g) * 0.05 + 0.001)
    sl[:5] = math.log(0.9)                                       # huge -> culled
    q = torch.randn(N, 4, generator=g)
    acc = torch.rand(N, generator=g) * 6e-4
    vc = torch.randint(0, 3, (N,), generator=g).int()
    noise = torch.randn(N, 2, 3, generator=g)
    cfg = TO.DensifyConfig()
    out = TO.densify_reference(means, shs, op, sl, q, acc, vc, noise, cfg)
    cnt = out["count"]
    assert set(cnt.tolist()) <= {0, 1, 2} and int(cnt.sum()) == out["means"].shape[0]
    assert bool((cnt[:5] == 0).all())
    src, new = out["src"], out["is_new"]
    assert bool((src[1:] >= src[:-1]).all()), "outputs must stay in source-id order"
    kept = ~new
    assert torch.equal(out["means"][kept], means[src[kept]]) and torch.equal(out["shs"], shs[src])
    # split samples shrink by 1.6, duplicates keep their scale
    smax = sl.exp().max(-1).values
    is_split = new & (smax[src] > cfg.size_thresh)
    assert torch.allclose(out["scales_log"][is_split], torch.log(sl[src[is_split]].exp() / 1.6))
    no = TO.densify_reference(means, shs, op, sl, q, acc, vc, noise, cfg, allow_split_dup=False)
    assert int(no["count"].max()) <= 1
    st = TO.densify_stats_reference(torch.tensor([[3.0, 4.0, 0.0], [1.0, 0.0, 0.0]]), torch.tensor([2, 0], dtype=torch.int32),
                                    torch.zeros(2), torch.zeros(2, dtype=torch.int32), torch.zeros(2, dtype=torch.int32))
    assert st[0].tolist() == [5.0, 0.0] and st[1].tolist() == [1, 0] and st[2].tolist() == [2, 0]


def test_trainer_rejects_cpu_tensors_and_bad_loss_type(tgs_lib):
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA-only"):
        T.TouchGSTrainer(z(4, 3), z(4, 16, 3), z(4), z(4, 3), z(4, 4))
    with pytest.raises(ValueError, match="depth_loss_type"):
        T.TouchGSTrainer(z(4, 3), z(4, 16, 3), z(4), z(4, 3), z(4, 4), T.TrainConfig(depth_loss_type="NOPE"))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        T.photometric_loss(z(3, 8, 8), z(3, 8, 8))
    c = T.TrainConfig()
    assert c.adam_eps == 1e-15 and c.depth_loss_mult == 0.2 and c.max_steps == 30000
    # position learning rate: log-linear decay 1.6e-4 -> 1.6e-6; SH degree: +1 every 1000 steps up to sh_degree
    assert c.lr_means_at(0) == 1.6e-4 and abs(c.lr_means_at(15000) - 1.6e-5) < 1e-9 and abs(c.lr_means_at(10**6) - 1.6e-6) < 1e-12
    assert [c.active_sh_degree(s) for s in (1, 999, 1000, 2500, 99999)] == [0, 0, 1, 2, 3]
    assert T.TrainConfig(sh_degree_interval=0).active_sh_degree(1) == 3


def test_train_oracle_regression_fixture():
    """tests/golden/train_oracle_small.npz pins the train-step oracle against accidental change (a regression
    fixture generated by the oracle itself, not a reference-side vector)."""
    import os
    import numpy as np
    from importlib import import_module
    from helpers import ROOT
    mk = import_module("golden.make_train_golden")
    got = mk.compute()
    z = np.load(os.path.join(ROOT, "tests", "golden", "train_oracle_small.npz"))
    assert set(z.files) == set(got)
    for k in z.files:
        if z[k].dtype.kind in "iub":
            assert np.array_equal(z[k], got[k]), k
        else:
            assert np.allclose(z[k], got[k], rtol=1e-5, atol=1e-7), k
