"""Multi-GPU equivalence check (run with torchrun on a >= 2-GPU box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank also computes the UNSHARDED result on its own GPU and compares:
  1. operator: tile-row band + exchange (NCCL all-reduce and the fused P2P gather) == full-image gradients;
  2. trainer: band + one-tile halo, band-local L1+SSIM and touch loss, == the single-GPU train step
     (gradients of all five raw tensors, summed partial losses, refine statistics).
Prints one JSON line per rank-0 check; exits non-zero on failure."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import touchgs_b200 as T  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    H, W, deg, N = 272, 320, 2, 30000
    sc = T.synth.make_scene(N, deg, 0.005, 0.06, seed=3)
    cam = T.synth.look_at_camera(W, H, (0.4, 0.3, -3.0))
    rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                         cam.viewmatrix.to(dev), cam.projmatrix.to(dev), deg, cam.campos.to(dev), False, False)
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(3, H, W, generator=g).to(dev)
    P = [t.to(dev) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations)]
    with torch.no_grad():
        _, _, d, _, _ = T.GaussianRasterizer(rs)(P[0], None, P[1], shs=P[2], scales=P[3], rotations=P[4])
    tgt, wgt = T.synth.make_touch_maps(d[0].cpu() + 0.02, seed=0, n_patches=4, patch_radius=20)
    tgt, wgt = tgt.to(dev), wgt.to(dev)
    ok = True

    def op_grads(tile_rows=None, pg=None, peer=None):
        ins = [t.clone().requires_grad_(True) for t in P]
        color, radii, depth, alpha, _ = T.GaussianRasterizer(rs)(
            ins[0], None, ins[1], shs=ins[2], scales=ins[3], rotations=ins[4], touch_depth=tgt, touch_weight=wgt,
            depth_loss="l1", depth_loss_mult=0.2, tile_rows=tile_rows, process_group=pg, peer_exchange=peer)
        y0, y1 = (0, H) if tile_rows is None else T.sharding.band_pixel_rows(tile_rows, H)
        ((color[:, y0:y1] - gt[:, y0:y1]).abs().sum() / (3 * H * W)).backward()
        return [t.grad for t in ins]

    full = op_grads()
    bands = T.sharding.even_bands(H, world)
    res = {}
    got = op_grads(bands[rank], group)
    res["nccl"] = max(rel(a, b) for a, b in zip(got, full))
    peer = T.sharding.make_peer_exchange(group, N, dev)
    res["p2p_available"] = peer is not None
    if peer is not None:
        peer.bands = bands
        for it in range(3):                              # three rounds: exercises the double buffering
            got = op_grads(bands[rank], group, peer)
        res["p2p"] = max(rel(a, b) for a, b in zip(got, full))
        # replicas must be bit-identical: compare rank 0's gradient with everyone's
        ref0 = got[0].clone()
        dist.broadcast(ref0, 0)
        res["p2p_replicas_identical"] = bool(torch.equal(ref0, got[0]))
        ok &= res["p2p"] < 1e-4 and res["p2p_replicas_identical"]
    ok &= res["nccl"] < 1e-4

    # ---- trainer: sharded (band + halo) vs single GPU
    raw = [sc.means3D, sc.shs, torch.logit(sc.opacities.reshape(-1).clamp(1e-4, 1 - 1e-4)), torch.log(sc.scales), sc.rotations]
    raw = [t.to(dev) for t in raw]
    cfg = T.TrainConfig(sh_degree=deg, depth_loss_type="DEPTH_UNCERTAINTY_WEIGHTED_LOSS", refine_every=0, sh_degree_interval=0)
    single = T.TouchGSTrainer(*raw, cfg)
    l_full = single.train_step(rs, gt, tgt, wgt)
    for name, px in (("nccl", None), ("p2p", peer)):
        if name == "p2p" and peer is None:
            continue
        tr = T.TouchGSTrainer(*raw, cfg, process_group=group, peer_exchange=px)
        l_part = tr.train_step(rs, gt, tgt, wgt).clone()
        dist.all_reduce(l_part)
        e = max(rel(tr.last["grads"][k], single.last["grads"][k]) for k in single.last["grads"])
        res[f"trainer_{name}_grad"] = e
        res[f"trainer_{name}_loss"] = abs(float(l_part) - float(l_full)) / abs(float(l_full))
        res[f"trainer_{name}_stats"] = bool(torch.equal(tr.vis_count, single.vis_count)) and rel(tr.grad_accum, single.grad_accum) < 1e-4
        ok &= e < 1e-4 and res[f"trainer_{name}_loss"] < 1e-5 and res[f"trainer_{name}_stats"]
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "ok": bool(t.item() > 0), **res}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() > 0 else 1)


if __name__ == "__main__":
    main()
