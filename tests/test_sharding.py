"""Multi-GPU host logic on CPU: band arithmetic and the one exchange step (all-reduce of the [N,10]
screen-space gradients) over a world_size-2 gloo group, using the oracle as the per-rank renderer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, T, synth, oracle_settings

sharding = T.sharding


def test_even_bands_cover_all_rows():
    for H in (16, 128, 1080, 2160, 17):
        Ty = sharding.tile_rows(H)
        for g in (1, 2, 4, 8):
            b = sharding.even_bands(H, g)
            assert b[0][0] == 0 and b[-1][1] == Ty
            assert all(b[i][1] == b[i + 1][0] for i in range(g - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    assert sharding.even_bands(1080, 2) == [(0, 34), (34, 68)]


def test_balanced_bands():
    work = [0, 0, 10, 10, 10, 10, 0, 0]
    b = sharding.balanced_bands(work, 2)
    assert b == [(0, 4), (4, 8)]
    b = sharding.balanced_bands([1] * 68, 8)
    assert b[0][0] == 0 and b[-1][1] == 68 and all(e > s for s, e in b)
    b = sharding.balanced_bands([5, 1, 1, 1], 4)
    assert b[0][0] == 0 and b[-1][1] == 4 and all(e >= s for s, e in b)
    assert sharding.band_pixel_rows((34, 68), 1080) == (544, 1080)


def test_balanced_bands_properties():
    """Property test (hypothesis): the bands are a contiguous partition of the tile rows, nobody is left without a row
    while rows remain, and no band carries more than its share plus one row's worth of work."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.one_of(st.just(0.0), st.floats(0.0, 1e6, allow_nan=False, allow_infinity=False),
                              st.integers(0, 5).map(float)), min_size=1, max_size=140),
           st.sampled_from([1, 2, 3, 4, 8]))
    def check(work, g):
        b = sharding.balanced_bands(work, g)
        Ty = len(work)
        assert len(b) == g and b[0][0] == 0 and b[-1][1] == Ty
        assert all(b[i][1] == b[i + 1][0] for i in range(g - 1)) and all(e >= s for s, e in b)
        if Ty >= g:
            assert all(e > s for s, e in b)
            total = sum(work)
            if total > 0:
                assert max(sum(work[s:e]) for s, e in b) <= total / g + max(work) + 1e-6 * total

    check()


def _leaf_pre(pre):
    """Detach the per-Gaussian screen-space quantities into leaves (depth is an ancestor of conic in
    the full graph, so *partial* derivatives -- what tgs_backward_render emits -- need a cut graph)."""
    leaves = {k: getattr(pre, k).detach().clone().requires_grad_(True) for k in ("xy", "conic", "opacity", "rgb", "depth")}
    return pre._replace(**leaves), leaves


def _screen_grads(leaves, img_loss):
    """d loss / d (xy, conic, opacity, rgb, depth) packed [N,10]."""
    gs = torch.autograd.grad(img_loss, [leaves[k] for k in ("xy", "conic", "opacity", "rgb", "depth")], allow_unused=True)
    N = leaves["xy"].shape[0]
    shapes = [(N, 2), (N, 3), (N, 1), (N, 3), (N, 1)]
    return torch.cat([(g.reshape(sh) if g is not None else torch.zeros(sh)) for g, sh in zip(gs, shapes)], 1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sc = synth.make_scene(300, 1, 0.04, 0.3, seed=21)
    cam = synth.look_at_camera(64, 64, (0.3, 0.1, -3.0))
    S = oracle_settings(cam, 1, (0.1, 0.1, 0.1))
    g = torch.Generator().manual_seed(5)
    grgb = torch.rand(3, 64, 64, generator=g)
    band = sharding.even_bands(64, world)[rank]
    m = sc.means3D.clone().requires_grad_(True)
    op, sh = sc.opacities.clone().requires_grad_(True), sc.shs.clone().requires_grad_(True)
    pre = O.preprocess(m, sc.scales, sc.rotations, op, sh, None, None, S, band)
    lpre, leaves = _leaf_pre(pre)
    img = O.render_tiles(lpre, O.bin_and_sort(lpre, S), S)
    y0, y1 = sharding.band_pixel_rows(band, 64)
    loss = (img.color[:, y0:y1] * grgb[:, y0:y1]).sum() + 0.3 * img.depth[y0:y1].sum()
    sg = _screen_grads(leaves, loss).detach().contiguous()
    sharding.all_reduce_screen_grads(sg)                 # the ONE collective of the path
    # replicated preprocess backward from the reduced screen grads
    L = ((pre.xy * sg[:, 0:2]).sum() + (pre.conic * sg[:, 2:5]).sum() + (pre.opacity * sg[:, 5]).sum()
         + (pre.rgb * sg[:, 6:9]).sum() + (pre.depth * sg[:, 9]).sum())
    L.backward()
    q.put((rank, sg, m.grad.clone(), pre.radii.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference
    sc = synth.make_scene(300, 1, 0.04, 0.3, seed=21)
    cam = synth.look_at_camera(64, 64, (0.3, 0.1, -3.0))
    S = oracle_settings(cam, 1, (0.1, 0.1, 0.1))
    g = torch.Generator().manual_seed(5)
    grgb = torch.rand(3, 64, 64, generator=g)
    m = sc.means3D.clone().requires_grad_(True)
    op, sh = sc.opacities.clone().requires_grad_(True), sc.shs.clone().requires_grad_(True)
    pre = O.preprocess(m, sc.scales, sc.rotations, op, sh, None, None, S)
    lpre, leaves = _leaf_pre(pre)
    img_l = O.render_tiles(lpre, O.bin_and_sort(lpre, S), S)
    sg = _screen_grads(leaves, (img_l.color * grgb).sum() + 0.3 * img_l.depth.sum()).detach()
    img = O.render_tiles(pre, O.bin_and_sort(pre, S), S)
    loss = (img.color * grgb).sum() + 0.3 * img.depth.sum()
    loss.backward()
    for rank, sg_r, gm, radii in res:
        assert torch.equal(radii, pre.radii)                       # radii are band-independent
        assert float((sg_r - sg).abs().max()) <= 1e-4 * float(sg.abs().max())
        assert float((gm - m.grad).abs().max()) <= 1e-4 * float(m.grad.abs().max())
    assert torch.equal(res[0][1], res[1][1])                       # replicas stay in sync


def test_halo_bands_cover_the_image_and_the_ssim_window():
    """Sharded train step: loss rows partition the image, rendered rows add >= 5 rows (the SSIM window
    radius) on every interior border, gradient rows == rendered rows."""
    from helpers import T
    S = T.sharding
    for H in (1080, 2160, 272, 117, 16):
        for world in (1, 2, 4, 8):
            if world > S.tile_rows(H):
                continue
            hb = S.halo_bands(H, world)
            assert len(hb) == world
            rows = []
            for r, (ext, loss, grad) in enumerate(hb):
                assert grad == S.band_pixel_rows(ext, H)
                assert grad[0] <= loss[0] <= loss[1] <= grad[1]
                if loss[1] > loss[0]:
                    if loss[0] > 0:
                        assert loss[0] - grad[0] >= 5
                    if loss[1] < H:
                        assert grad[1] - loss[1] >= 5
                rows.append(loss)
            assert rows[0][0] == 0 and rows[-1][1] == H
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1)), "loss rows must partition the image"
