"""Regression fixture of the train-step oracle (oracle/train_oracle.py): NOT a reference-side vector (the trainer is in
the reference's empty submodule; parity unpinned) -- it pins OUR oracle against accidental change.

    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

TO = O.train_oracle


def inputs():
    g = torch.Generator().manual_seed(1234)
    a = torch.rand(3, 37, 45, generator=g)
    b = (a + 0.15 * torch.randn(3, 37, 45, generator=g)).clamp(0, 1)
    N, K = 64, 4
    means, shs = torch.randn(N, 3, generator=g), torch.randn(N, K, 3, generator=g)
    op = torch.randn(N, generator=g) * 2
    sl = torch.log(torch.rand(N, 3, generator=g) * 0.03 + 0.001)
    sl[:3] = float(np.log(0.8))
    q = torch.randn(N, 4, generator=g)
    acc = torch.rand(N, generator=g) * 8e-4
    vc = torch.randint(0, 4, (N,), generator=g).int()
    noise = torch.randn(N, 2, 3, generator=g)
    p = torch.randn(50, generator=g)
    gr = torch.randn(50, generator=g) * 0.1
    return a, b, (means, shs, op, sl, q, acc, vc, noise), (p, gr)


def compute():
    a, b, dens, (p, gr) = inputs()
    x = a.clone().requires_grad_(True)
    loss = TO.photometric_loss(x, b, 0.2)
    loss.backward()
    d = TO.densify_reference(*dens, TO.DensifyConfig())
    (q,), (m,), (v,) = TO.adam_reference([p], [gr], [1e-2], 3)
    return dict(loss=loss.detach().numpy(), dloss=x.grad.numpy(), d_means=d["means"].numpy(), d_scales=d["scales_log"].numpy(),
                d_src=d["src"].numpy(), d_new=d["is_new"].numpy(), adam_p=q.numpy(), adam_m=m.numpy(), adam_v=v.numpy())


if __name__ == "__main__":
    out = compute()
    np.savez_compressed(os.path.join(HERE, "train_oracle_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})
