#!/usr/bin/env python
"""bench.py -- forward+backward Gaussians/s of the Touch-GS rasterizer hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # our arm (libtgs.so, sm_100a)
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # CPU reference arm (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # tile-row sharded, one rank per GPU

A "step" = one forward + backward of the operator for one camera of the workload:
preprocess -> bin + sort -> per-tile compositing of RGB + expected depth -> backward with the touch
depth-L1 gradient fused in -> preprocess backward.  Workload = BASELINE config c3: 1M synthetic
Gaussians, 1920x1080, SH degree 3, fused tactile depth-L1.  With N > 1 ranks the image
is sharded by tile rows (bands balanced by measured instance counts) and the [N,10] screen-space gradients are
exchanged once per step: by default a P2P gather over NVLink fused into the preprocess-backward kernel
(--exchange p2p), or one NCCL all-reduce (--exchange nccl).

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for the meaning of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "fwd+bwd Gaussians/sec @1M splats/1080p"
UNIT = "Gaussians/s"
DEPTH_LOSS_MULT = 0.2       # reference scripts/train_block_data.sh:50 (--pipeline.model.depth-loss-mult 0.2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c5", "fixture", "fixture1m"],
                    help="fixture*: Gaussians seeded from the reference's sample point cloud + its sample camera poses "
                         "(tests/golden/fixture_scene.npz; SURVEY §8f N4), z-buffer depth of the cloud as the touch target")
    ap.add_argument("--num-gaussians", type=int, default=None, help="override N (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-refcuda", action="store_true", help="skip the reference-structure CUDA comparison arm")
    ap.add_argument("--cameras", type=int, default=8)
    ap.add_argument("--workload", default="rasterizer", choices=["rasterizer", "touch_inputs", "train_step"],
                    help="touch_inputs: roofline of the per-pixel touch/vision fusion kernel (SURVEY §8f N2); train_step: the "
                         "full Touch-GS train step (activations, rasterizer, L1+SSIM loss, fused touch depth-L1, Adam, refine "
                         "every --refine-every steps: SURVEY §8f N1 / BASELINE config c5); neither is the headline metric")
    ap.add_argument("--refine-every", type=int, default=100)
    ap.add_argument("--forward-only", action="store_true",
                    help="eval / render path (reference experiment_utils/run_eval.py:43-48 -> ns-eval, ns-render): forward only, "
                         "no grad, markVisible prefiltering; a secondary line, not the headline metric")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange of the [N,10] screen gradients: p2p = gather fused into the preprocess-backward "
                         "kernel over peer-mapped memory (NVLink); nccl = one all-reduce (also the fallback if p2p is unavailable)")
    ap.add_argument("--no-hints", action="store_true", help="synchronous sizing in every forward (no rendered_hint)")
    ap.add_argument("--defer-count", default="auto", choices=["auto", "on", "off"],
                    help="operator option defer_count (with the per-view hints): auto = on for N > 1")
    ap.add_argument("--even-bands", action="store_true", help="N > 1: equal tile-row bands instead of bands balanced by instance count")
    ap.add_argument("--no-measured-configs", action="store_true",
                    help="--impl reference: skip the full CPU runs of configs c1 / c2 (about a minute on 8 cores)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).
    The sampler process is started BEFORE warm-up (its NVML initialisation stalls the driver for tens
    of milliseconds, which must not land inside a timed window); samples are windowed by wall clock."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # one GPU: the default timed region is ~150 ms, several 25 ms samples land inside it.  Under torchrun every rank runs
    # its own sampler: keep the 100 ms period there so that 8 pollers do not compete for the driver
    PERIOD_MS = 25 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 100

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS), "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:      # wait until NVML is up
                time.sleep(0.02)
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        time.sleep(self.PERIOD_MS * 1.2e-3)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()

    def summary(self, t_begin, t_end, t_warm):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        window = "timed region"
        rows = [l for t, l in self.lines if t_begin <= t <= t_end + self.PERIOD_MS * 1.1e-3]
        if not rows:        # timed region shorter than the polling period: use everything since warm-up began
            rows = [l for t, l in self.lines if t_warm <= t <= t_end + self.PERIOD_MS * 1.1e-3]
            window = "warm-up + timed region (timed region shorter than the polling period)"
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- workload
def alg_bytes(N, I, P, T, K):
    """BASELINE.md §4 / SURVEY.md §8(d): algorithmic bytes of one fwd+bwd step."""
    return N * (12 * (11 + 3 * K) + 124) + I * 172 + P * 56 + T * 8


def make_workload(cfg, N, n_cams, dev, rank, band, seed=0):
    import touchgs_b200 as T
    synth = T.synth
    fixture = "fixture_copies" in cfg
    if fixture:
        scene = synth.fixture_scene(cfg["sh_degree"], cfg["fixture_copies"], seed)
        cams = synth.fixture_cameras(cfg["W"], cfg["H"], n_cams)
    else:
        scene = synth.make_scene(N, cfg["sh_degree"], cfg["smin"], cfg["smax"], seed)
        cams = synth.orbit_cameras(cfg["W"], cfg["H"], n_cams, 3.0, seed)
    H, W = cfg["H"], cfg["W"]
    params = {k: getattr(scene, k).to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    pert = synth.perturbed(scene, 0.01, seed)
    batches = []
    bg = torch.zeros(3, device=dev)
    g = torch.Generator().manual_seed(seed + 31)
    for ci, cam in enumerate(cams):
        rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix.to(dev),
                                             cam.projmatrix.to(dev), cfg["sh_degree"], cam.campos.to(dev), False, False)
        if fixture:
            # touch target = z-buffer depth image of the seed cloud (reference read_point_cloud.py:224-266), mm-quantised,
            # 0 = no point; weight = 1/sigma with the touch sigma on the hit pixels
            zb = synth.zbuffer_depth(scene.means3D[: 71_283], cam)
            target = torch.round(zb * 1000.0).clamp(0, 65535) / 1000.0
            weight = torch.where(target > 0, torch.full_like(target, 1.0 / 0.005), torch.zeros_like(target))
        else:
            with torch.no_grad():      # touch target = expected depth of the PERTURBED scene, rendered by us
                _, _, d, _, _ = T.GaussianRasterizer(rs)(pert.means3D.to(dev), None, pert.opacities.to(dev),
                                                         shs=pert.shs.to(dev), scales=pert.scales.to(dev),
                                                         rotations=pert.rotations.to(dev))
            target, weight = synth.make_touch_maps(d[0].cpu(), seed=seed + ci)
        gt = torch.rand(3, H, W, generator=g)
        host = dict(gt=gt.pin_memory(), target=target.pin_memory(), weight=weight.pin_memory(),
                    view=cam.viewmatrix.contiguous().pin_memory(), proj=cam.projmatrix.contiguous().pin_memory(),
                    campos=cam.campos.contiguous().pin_memory())
        devb = {k: v.to(dev) for k, v in host.items()}
        batches.append(dict(cam=cam, host=host, dev=devb))
    return scene, params, batches, bg


class Stepper:
    """One forward+backward through the public operator (GaussianRasterizer autograd Function)."""

    def __init__(self, cfg, params, bg, dev, band, group):
        import touchgs_b200 as T
        self.T, self.cfg, self.p, self.bg, self.dev = T, cfg, params, bg, dev
        self.band, self.group = band, group
        self.peer = None
        self.use_hints = True
        self.defer_count = False
        H = cfg["H"]
        self.y0, self.y1 = (0, H) if band is None else T.sharding.band_pixel_rows(band, H)
        self.inv = 1.0 / (3.0 * cfg["H"] * cfg["W"])
        # persistent device staging buffers for the end-to-end path
        self.stage = None
        self.hints = {}          # per training view: instance count of its previous render (+5 %), see _run
        self.rasterizers = {}

    def _run(self, cam, view, proj, campos, gt, target, weight):
        """`rendered_hint`: a trainer revisits the same views every epoch, so it passes the view's previous
        instance count (+5 %) and the operator sizes its binning buffers speculatively, hiding the
        forward's host sync; results are identical with or without the hint (the operator re-runs the
        binning exactly if the hint was too small)."""
        T, cfg, p = self.T, self.cfg, self.p
        key = id(cam)
        # the rasterizer module of a view is built once (a trainer keeps one per camera): the settings only hold
        # references to the camera tensors, whose CONTENTS the end-to-end path refreshes in place every step
        rkey = (key, view.data_ptr())
        ras = self.rasterizers.get(rkey)
        if ras is None:
            rs = T.GaussianRasterizationSettings(cfg["H"], cfg["W"], cam.tanfovx, cam.tanfovy, self.bg, 1.0, view, proj,
                                                 cfg["sh_degree"], campos, False, False)
            ras = self.rasterizers[rkey] = T.GaussianRasterizer(rs)
        for v in p.values():
            v.grad = None
        color, radii, depth, alpha, resid = ras(
            p["means3D"], None, p["opacities"], shs=p["shs"], scales=p["scales"], rotations=p["rotations"],
            touch_depth=target, touch_weight=weight, depth_loss="l1", depth_loss_mult=DEPTH_LOSS_MULT,
            depth_normalize=True, tile_rows=self.band, process_group=self.group, peer_exchange=self.peer,
            rendered_hint=self.hints.get(key, 0) if self.use_hints else 0, defer_count=self.defer_count)
        y0, y1 = self.y0, self.y1
        # mean |C - C*| (band-local part) through the library's fused photometric loss (lambda_dssim = 0: plain L1,
        # one kernel forward, one backward)
        rows = None if self.band is None else (y0, y1)
        loss = T.photometric_loss(color, gt, 0.0, rows, rows)
        loss.backward()
        # (read after backward: with a deferred count the number is a ticket that backward has redeemed by now)
        self.hints[key] = int(ras.last_num_rendered * 1.05) + 4096
        return loss

    def forward_only_step(self, b):
        """The eval / render path: markVisible, then the forward of the operator under no_grad."""
        T, cfg, p, d, cam = self.T, self.cfg, self.p, b["dev"], b["cam"]
        rs = T.GaussianRasterizationSettings(cfg["H"], cfg["W"], cam.tanfovx, cam.tanfovy, self.bg, 1.0, d["view"], d["proj"],
                                             cfg["sh_degree"], d["campos"], True, False)
        with torch.no_grad():
            ras = T.GaussianRasterizer(rs)
            vis = ras.markVisible(p["means3D"])
            color, radii, depth, alpha, _ = ras(p["means3D"], None, p["opacities"], shs=p["shs"], scales=p["scales"],
                                                rotations=p["rotations"], tile_rows=self.band,
                                                rendered_hint=self.hints.get(id(cam), 0) if self.use_hints else 0)
            self.hints[id(cam)] = int(ras.last_num_rendered * 1.05) + 4096
        return color, vis

    def device_step(self, b):
        if getattr(self, "forward_only", False):
            return self.forward_only_step(b)
        d = b["dev"]
        return self._run(b["cam"], d["view"], d["proj"], d["campos"], d["gt"], d["target"], d["weight"])

    def copy_rows(self):
        """Image rows of the per-step inputs this rank needs on its GPU: its own band (rasterizer workload) or its
        band plus the one-tile halo the SSIM window reaches into (train step); everything on a single GPU."""
        return self.y0, self.y1

    def _prefetch(self, b, slot):
        """H2D of one step's inputs from pinned host memory on the copy stream.  A rank of the tile-row shard copies
        only the image rows it consumes (ground truth, touch depth and weight of its band); cameras are copied whole."""
        r0, r1 = self.copy_rows()
        with torch.cuda.stream(self.copy_stream):
            for k, v in b["host"].items():
                if k in ("gt", "target", "weight") and (r0, r1) != (0, self.cfg["H"]):
                    # one CONTIGUOUS row range per channel: a strided [3, rows, W] slice would make torch stage the
                    # copy through a pageable temporary (synchronous, and a host memcpy on top)
                    dst = self.stage[slot][k]
                    if v.dim() == 3:
                        for c in range(v.shape[0]):
                            dst[c, r0:r1].copy_(v[c, r0:r1], non_blocking=True)
                    else:
                        dst[r0:r1].copy_(v[r0:r1], non_blocking=True)
                else:
                    self.stage[slot][k].copy_(v, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def h2d_bytes_rank(self, b):
        r0, r1 = self.copy_rows()
        H = self.cfg["H"]
        tot = 0
        for k, v in b["host"].items():
            n = v.numel() * v.element_size()
            tot += n * (r1 - r0) // H if k in ("gt", "target", "weight") else n
        return int(tot)

    def e2e_begin(self, batches):
        h = batches[0]["host"]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.stage = [{k: torch.zeros_like(v, device=self.dev) for k, v in h.items()} for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.batches, self.k, self.losses = batches, 0, []
        self._prefetch(batches[0], 0)

    def e2e_step(self, i):
        """End-to-end step i: its inputs (camera, ground-truth image, touch depth + weight) start in
        PINNED HOST memory and are copied H2D on a copy stream (issued one step ahead, double-buffered);
        the step's result (loss) is copied D2H into pinned memory and read by the host one step later."""
        n = len(self.batches)
        b, slot = self.batches[i % n], self.k & 1
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[slot])
        s = self.stage[slot]
        loss = self._run(b["cam"], s["view"], s["proj"], s["campos"], s["gt"], s["target"], s["weight"])
        self.loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)     # D2H of the result
        self.loss_ev[slot].record(cur)
        self.consumed[slot].record(cur)
        # next step's inputs: H2D into the other slot once the step that used it has finished
        nslot = slot ^ 1
        self.copy_stream.wait_event(self.consumed[nslot])
        self._prefetch(self.batches[(i + 1) % n], nslot)
        if self.k > 0:                       # host reads the PREVIOUS step's loss (already on its way)
            self.loss_ev[nslot].synchronize()
            self.losses.append(float(self.loss_host[nslot]))
        self.k += 1

    def e2e_finish(self):
        slot = (self.k - 1) & 1
        self.loss_ev[slot].synchronize()
        self.losses.append(float(self.loss_host[slot]))

    @staticmethod
    def h2d_bytes(b):
        return int(sum(v.numel() * v.element_size() for v in b["host"].values()))


class TrainStepper(Stepper):
    """One FULL train step through the public trainer API (TouchGSTrainer.train_step): activations -> rasterizer
    forward -> L1 + SSIM photometric loss -> backward with the fused touch depth-L1 gradient -> [all-reduce] ->
    activation backward -> refine statistics -> one-launch Adam; refine (densify / cull) every `refine_every` steps."""

    def __init__(self, cfg, params, bg, dev, band, group, refine_every):
        super().__init__(cfg, params, bg, dev, band, group)
        T = self.T
        raw = [params["means3D"].detach(), params["shs"].detach(),
               torch.logit(params["opacities"].detach().reshape(-1).clamp(1e-4, 1 - 1e-4)),
               torch.log(params["scales"].detach()), params["rotations"].detach()]
        tc = T.TrainConfig(sh_degree=cfg["sh_degree"], depth_loss_mult=DEPTH_LOSS_MULT,
                           depth_loss_type="DEPTH_UNCERTAINTY_WEIGHTED_LOSS", uncertainty_weight=1.0,
                           refine_every=refine_every, warmup_length=0, sh_degree_interval=0)
        self.trainer = T.TouchGSTrainer(*raw, tc, process_group=group)
        self.raw_n = int(raw[0].shape[0])
        self.n_history = [self.trainer.num_points]

    def copy_rows(self):
        if self.group is None:
            return 0, self.cfg["H"]
        return self.trainer._bands(self.cfg["H"])[2]          # band + halo rows

    def _run(self, cam, view, proj, campos, gt, target, weight):
        T, cfg = self.T, self.cfg
        rs = T.GaussianRasterizationSettings(cfg["H"], cfg["W"], cam.tanfovx, cam.tanfovy, self.bg, 1.0, view, proj,
                                             cfg["sh_degree"], campos, False, False)
        loss = self.trainer.train_step(rs, gt, target, weight, view_key=id(cam) if self.use_hints else None)
        if self.trainer.num_points != self.n_history[-1]:
            self.n_history.append(self.trainer.num_points)
        return loss


# ------------------------------------------------------------------------ CPU reference
class CpuReference:
    """The reference's CPU path for this hot path: the pure-PyTorch oracle (kind = "port"; the reference
    vendors no rasterizer to compile -- SURVEY.md §0).  Timed on a BOUNDED SAMPLE of the workload:
    preprocess fwd+bwd and binning on a 1/`gauss_div` slice of the Gaussians, compositing fwd+bwd
    (with the touch depth-L1 loss) on `n_tiles` tiles; each part is scaled to the full workload by its
    unit count (Gaussians resp. (pixel, splat) pairs) to estimate whole-step Gaussians/s."""

    def __init__(self, cfg, scene, cam, target, weight, n_tiles=48, gauss_div=10):
        import oracle as O
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        H, W = cfg["H"], cfg["W"]
        self.cfg, self.scene = cfg, scene
        self.S = O.OracleSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3), 1.0, cam.viewmatrix, cam.projmatrix,
                                  cfg["sh_degree"], cam.campos)
        self.target, self.weight = target, weight
        N = scene.means3D.shape[0]
        self.N = N
        with torch.no_grad():           # one-time, untimed: list lengths of the full workload
            pre = O.preprocess(scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs, None, None, self.S)
            self.bins = O.bin_and_sort(pre, self.S)
        self.pre = pre
        lens = (self.bins.ranges[:, 1] - self.bins.ranges[:, 0]).long()
        self.I = int(lens.sum())
        self.pairs_total = int(lens.sum()) * 256
        nz = torch.nonzero(lens > 0).flatten()
        if nz.numel() == 0:
            self.tiles = []
        else:
            idx = torch.linspace(0, nz.numel() - 1, min(n_tiles, nz.numel())).round().long()
            self.tiles = nz[idx].tolist()
        self.pairs_sample = int(lens[self.tiles].sum()) * 256 if self.tiles else 1
        self.nonempty_tiles = int((lens > 0).sum())
        self.ns = max(1, N // gauss_div)
        g = torch.Generator().manual_seed(5)
        self.gt = torch.rand(3, H, W, generator=g)
        self.sg = torch.randn(self.ns, 10, generator=g)
        self.sample = (f"oracle (pure PyTorch, {self.cores} threads): preprocess fwd+bwd + binning on {self.ns} of {N} "
                       f"Gaussians, compositing fwd+bwd with touch depth-L1 on {len(self.tiles)} of "
                       f"{int((lens > 0).sum())} non-empty tiles ({self.pairs_sample} of {self.pairs_total} "
                       f"(pixel,splat) pairs); parts scaled by unit count to the full step")

    def step(self):
        O, sc, S, ns = self.O, self.scene, self.S, self.ns
        t0 = time.perf_counter()
        ins = [t[:ns].clone().requires_grad_(True) for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)]
        pre_s = O.preprocess(ins[0], ins[1], ins[2], ins[3], ins[4], None, None, S)
        t1 = time.perf_counter()
        O.bin_and_sort(pre_s, S)
        t2 = time.perf_counter()
        L = ((pre_s.xy * self.sg[:, 0:2]).sum() + (pre_s.conic * self.sg[:, 2:5]).sum() + (pre_s.opacity * self.sg[:, 5]).sum()
             + (pre_s.rgb * self.sg[:, 6:9]).sum() + (pre_s.depth * self.sg[:, 9]).sum())
        L.backward()
        t3 = time.perf_counter()
        leaves = {k: getattr(self.pre, k).detach().clone().requires_grad_(True) for k in ("xy", "conic", "opacity", "rgb", "depth")}
        lpre = self.pre._replace(**leaves)
        img = O.render_tiles(lpre, self.bins, S, tiles=self.tiles)
        scale = O.loss_scale_from_target(self.target, DEPTH_LOSS_MULT)
        tl, _, _ = O.touch_loss(img.depth, img.alpha, self.target, self.weight, "l1", scale, True)
        loss = (img.color - self.gt).abs().mean() + tl
        loss.backward()
        t4 = time.perf_counter()
        per_gauss = (t1 - t0) + (t2 - t1) + (t3 - t2)
        render = t4 - t3
        est = per_gauss * (self.N / ns) + render * (self.pairs_total / self.pairs_sample)
        return est, (t4 - t0)


def _reference_touch_target(O, synth, scene, cam, S, tiles, seed=0):
    """The GPU arm's touch target (expected depth of the scene perturbed by N(0, 0.01), mm-quantised, touch patches,
    5 % invalid: SURVEY §8d) rebuilt on the CPU arm with the oracle.  With `tiles` it is rendered on those tiles only
    (0 = invalid elsewhere): the values on the sampled tiles are the GPU arm's."""
    pert = synth.perturbed(scene, 0.01, seed)
    with torch.no_grad():
        ppre = O.preprocess(pert.means3D, pert.scales, pert.rotations, pert.opacities, pert.shs, None, None, S)
        pimg = O.render_tiles(ppre, O.bin_and_sort(ppre, S), S, tiles=tiles)
        has = pimg.alpha > 0
        d = torch.where(has, pimg.depth / pimg.alpha.clamp_min(1e-30), torch.zeros_like(pimg.depth))
    return synth.make_touch_maps(d, seed=seed)


def _measure_full_cpu_config(name, naive=False, naive_row_step=1, runs=1):
    """One FULL forward + backward of the oracle on a BASELINE config (SURVEY §8d "Reference CPU path"): really run, wall
    clock.  naive = the per-pixel Python loop (config c1's "pure-PyTorch rasterize_gaussians"); with naive_row_step > 1
    only every naive_row_step-th pixel row is composited (bounded sample, reported as such)."""
    import touchgs_b200 as T
    import oracle as O
    cfg = T.synth.CONFIGS[name]
    H, W, deg, N = cfg["H"], cfg["W"], cfg["sh_degree"], cfg["N"]
    scene = T.synth.make_scene(N, deg, cfg["smin"], cfg["smax"], 0)
    cam = T.synth.orbit_cameras(W, H, 8, 3.0, 0)[0]
    S = O.OracleSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3), 1.0, cam.viewmatrix, cam.projmatrix, deg, cam.campos)
    target, weight = _reference_touch_target(O, T.synth, scene, cam, S, None)
    g = torch.Generator().manual_seed(31)
    gt = torch.rand(3, H, W, generator=g)
    rows = None if (not naive or naive_row_step <= 1) else list(range(0, H, naive_row_step))
    times, I = [], 0
    for _ in range(runs):
        ins = [t.clone().requires_grad_(True) for t in (scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs)]
        t0 = time.perf_counter()
        out = O.rasterize(ins[0], ins[3], S, shs=ins[4], scales=ins[1], rotations=ins[2], touch_depth=target, touch_weight=weight,
                          depth_loss="l1", depth_loss_mult=DEPTH_LOSS_MULT, naive=naive, naive_rows=rows)
        t1 = time.perf_counter()
        ((out.color - gt).abs().mean() + out.touch_loss).backward()
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
        I = int(out.bins.keys.numel())
    fwd = min(t[0] for t in times)
    bwd = min(t[1] for t in times)
    frac = 1.0 if rows is None else len(rows) / H
    rec = {"config": name, "path": "naive per-pixel Python loop" if naive else "vectorised oracle", "N": N, "image": f"{W}x{H}",
           "num_rendered": I, "forward_s": fwd, "backward_s": bwd, "runs": runs, "pixel_rows_composited": frac,
           "estimated": frac < 1.0}
    if frac < 1.0:
        # per-Gaussian work (preprocess, binning) is done in full; the compositing loop scales with the rows composited
        rec["note"] = (f"compositing loop run on every {naive_row_step}th pixel row ({len(rows)} of {H}); forward + backward time "
                       "scaled by 1 / fraction for the Gaussians/s figure (the full loop takes ~6 min on 8 cores)")
        rec["gaussians_per_s"] = N / ((fwd + bwd) / frac)
    else:
        rec["gaussians_per_s"] = N / (fwd + bwd)
    return rec


def run_reference(args, cfg, N):
    """`--impl reference`: rank 0 only, CPU, same config / metric / unit.  Each step = a bounded sample of the c3
    workload (the JSON line says `estimated: true` and gives the sampled fractions); the configs SURVEY §8d prescribes
    for the CPU path -- c1 naive loop, c1 and c2 vectorised -- are additionally RUN IN FULL once and reported under
    `measured_configs`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import touchgs_b200 as T
    import oracle as O
    t_start = time.perf_counter()
    scene = T.synth.make_scene(N, cfg["sh_degree"], cfg["smin"], cfg["smax"], 0)
    cam = T.synth.orbit_cameras(cfg["W"], cfg["H"], args.cameras, 3.0, 0)[0]
    H, W = cfg["H"], cfg["W"]
    ref = CpuReference(cfg, scene, cam, None, None)
    # same touch target as the GPU arm's camera 0, on the sampled tiles
    ref.target, ref.weight = _reference_touch_target(O, T.synth, scene, cam, ref.S, ref.tiles)
    for _ in range(args.warmup):
        ref.step()
    ests, walls = [], []
    for _ in range(args.steps):
        e, w = ref.step()
        ests.append(e); walls.append(w)
    est = sum(ests) / len(ests)
    val = N / est
    measured = {}
    if not args.no_measured_configs:
        measured["c1_vectorised"] = _measure_full_cpu_config("c1", runs=3)
        measured["c2_vectorised"] = _measure_full_cpu_config("c2", runs=1)
        measured["c1_naive"] = _measure_full_cpu_config("c1", naive=True, naive_row_step=8, runs=1)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "estimated": True,
           "sampled_fraction": {"gaussians": ref.ns / ref.N, "pixel_splat_pairs": ref.pairs_sample / max(ref.pairs_total, 1),
                                "tiles": len(ref.tiles) / max(ref.nonempty_tiles, 1)},
           "config": {"workload": f"{args.config}: {N} Gaussians, {W}x{H}, SH deg {cfg['sh_degree']}, fused touch depth-L1 "
                                  f"(mult {DEPTH_LOSS_MULT}), camera 0 of the GPU arm, same seeds and the same touch target",
                      "num_rendered": ref.I,
                      "note": "value / ms_per_step are the ESTIMATED full-step CPU time: every step really runs the bounded sample "
                              "described in cpu_baseline.sample and scales its parts by their unit counts; measured_configs holds "
                              "full, unscaled CPU runs of BASELINE configs c1 and c2",
                      "sample_wall_ms": 1e3 * sum(walls) / len(walls)},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample},
           "measured_configs": measured,
           "wall_s_total": None,
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out["wall_s_total"] = time.perf_counter() - t_start
    print(json.dumps(out), flush=True)


def run_touch_inputs(args):
    """Secondary line: the fp64 per-pixel fusion kernel on a dataset-sized batch (100 frames of 1280x720,
    the reference's native real-world size -- SURVEY A7), uint16 in / uint16 + fp32 out, 22 B per pixel."""
    import touchgs_b200 as T
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    n = 100 * 1280 * 720
    g = torch.Generator(device="cpu").manual_seed(0)
    mk = lambda hi: torch.randint(0, hi, (n,), generator=g, dtype=torch.int32).to(torch.uint16).to(dev)
    touch, vision, tsig = mk(3000), mk(6000), mk(60)
    fn = lambda: T.touch_inputs.fuse_touch_vision(touch, vision, tsig, 1.3, -0.2, 0.017, True, 1.0)
    for _ in range(max(args.warmup, 3)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peak, src = peaks()
    gbs = 22.0 * n / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "touch/vision fusion pixels/s (secondary; SURVEY 8f N2)", "value": n / (ms * 1e-3), "unit": "pixels/s",
                      "ms_per_step": ms, "steps": args.steps, "dtype": "f64 math, u16/f32 I/O", "data": "synthetic",
                      "config": {"workload": "100 frames x 1280x720 uint16 (touch depth, vision depth, touch sigma)",
                                 "note": "includes 6 torch.empty output allocations per call; 737 MB working set > L2"},
                      "roofline": {"bound": "hbm", "kernel": "k_fuse_touch_vision", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                   "frac": gbs / peak, "traffic": None, "peak_source": src,
                                   "algorithmic_bytes_per_launch": 22 * n}}), flush=True)


# --------------------------------------------------------------------------------- main
def main():
    args = parse()
    if args.workload == "touch_inputs":
        run_touch_inputs(args)
        return
    import touchgs_b200 as T
    cfg = dict(T.synth.CONFIGS[args.config])
    N = args.num_gaussians or cfg["N"]
    if args.impl == "reference":
        run_reference(args, cfg, N)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    T._lib.load()
    H, W, K = cfg["H"], cfg["W"], (cfg["sh_degree"] + 1) ** 2
    Tx, Ty = (W + 15) // 16, (H + 15) // 16
    band = None if world == 1 else T.sharding.even_bands(H, world)[rank]

    scene, params, batches, bg = make_workload(cfg, N, args.cameras, dev, rank, band)
    bands_all, band_policy = None, "single GPU"
    if world > 1:
        bands_all, band_policy = T.sharding.even_bands(H, world), "even tile rows"
        if not args.even_bands:
            # bands balanced by measured work (SURVEY §8e "balance by instance count, not rows"): instances per tile row,
            # summed over the workload's cameras, from one untimed full-image forward per camera (identical on every rank)
            row_work = torch.zeros(Ty, dtype=torch.float64)
            with torch.no_grad():
                for b in batches:
                    d = b["dev"]
                    rs0 = T.GaussianRasterizationSettings(H, W, b["cam"].tanfovx, b["cam"].tanfovy, bg, 1.0, d["view"], d["proj"],
                                                          cfg["sh_degree"], d["campos"], False, False)
                    st0 = T.inspect_state.forward_state(params["means3D"].detach(), params["opacities"].detach(), rs0,
                                                        shs=params["shs"].detach(), scales=params["scales"].detach(),
                                                        rotations=params["rotations"].detach())
                    rg = st0["ranges"].long()
                    row_work += (rg[:, 1] - rg[:, 0]).clamp_min(0).view(Ty, Tx).sum(1).double().cpu()
                    del st0
            bands_all = T.sharding.balanced_bands([float(v) + 1.0 for v in row_work], world)
            band_policy = "balanced by instances per tile row (measured over the workload's cameras)"
        band = tuple(bands_all[rank])
    train_mode = args.workload == "train_step"
    if train_mode:
        stepper = TrainStepper(cfg, params, bg, dev, band, group, args.refine_every)
        args.no_refcuda = True
        args.no_cpu_baseline = True
    else:
        stepper = Stepper(cfg, params, bg, dev, band, group)
    stepper.use_hints = not args.no_hints
    # deferred count: the forward never waits for num_rendered (checked when backward runs).  On by default for N > 1, where
    # a rank's GPU work per step is comparable to the host path and the forward's wait would leave the GPU idle
    stepper.defer_count = stepper.use_hints and (args.defer_count == "on" or (args.defer_count == "auto" and world > 1)) and not train_mode
    stepper.forward_only = bool(args.forward_only) and not train_mode
    if stepper.forward_only:
        args.no_e2e = args.no_refcuda = args.no_cpu_baseline = True
    exchange = "none"
    if world > 1:
        exchange = "nccl all-reduce"
        if args.exchange == "p2p":
            peer = T.sharding.make_peer_exchange(group, int(N * (3 if train_mode else 1)), dev)
            if peer is not None:
                peer.bands = [tuple(b) for b in bands_all]
                stepper.peer = peer
                if train_mode:
                    stepper.trainer.peer = peer
                exchange = "p2p gather fused into preprocess-backward (symmetric memory over NVLink), no all-reduce"
    # allocator priming (setup, not warm-up): every camera has its own instance count, so touch each
    # once so that torch's caching allocator owns blocks of every size before anything is timed
    for _ in range(2):        # second pass: with the per-view hints of the first (speculative sizing: other buffer sizes)
        for b in batches:
            stepper.device_step(b)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False, finalize=None):
        sampler = ClockSampler(local)
        sampler.start()
        t_warm = time.time()
        for i in range(warmup):
            fn(i)
        barrier()
        if profile:
            T._lib.profile_enable(True)
            T._lib.profile_read()
        own0, cub0 = T._lib.launch_counts()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]     # per-step spread (p10/p50/p90)
        t_begin = time.time()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
            marks[i].record()
        if finalize is not None:
            finalize()
        e1.record()
        barrier()
        t_end = time.time()
        sampler.stop()
        clocks = sampler.summary(t_begin, t_end, t_warm)
        ms = e0.elapsed_time(e1)
        per = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))
        timed.spread = {"p10": per[len(per) // 10], "p50": per[len(per) // 2], "p90": per[(9 * len(per)) // 10]} if per else None
        prof = None
        if profile:
            prof = T._lib.profile_read()
            T._lib.profile_enable(False)
        own1, cub1 = T._lib.launch_counts()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, prof, (own1 - own0, cub1 - cub0)

    steps, warmup = args.steps, max(args.warmup, 3)
    ms, clocks, prof, (own, cub) = timed(lambda i: stepper.device_step(batches[i % len(batches)]), steps, warmup,
                                         profile=True)
    value = N * steps / (ms * 1e-3)
    spread = getattr(timed, "spread", None)

    # measured I (num_rendered) of the cameras used: read from a state-inspecting forward (untimed)
    with torch.no_grad():
        st = T.inspect_state.forward_state(params["means3D"].detach(), params["opacities"].detach(),
                                           T.GaussianRasterizationSettings(H, W, batches[0]["cam"].tanfovx, batches[0]["cam"].tanfovy,
                                                                           bg, 1.0, batches[0]["dev"]["view"], batches[0]["dev"]["proj"],
                                                                           cfg["sh_degree"], batches[0]["dev"]["campos"], False, False),
                                           shs=params["shs"].detach(), scales=params["scales"].detach(),
                                           rotations=params["rotations"].detach(), opt=T.TouchOptions(tile_rows=band))
        I_cam0 = int(st["num_rendered"])
        n_vis = int((st["radii"] > 0).sum())
        del st

    e2e = None
    if not args.no_e2e:
        stepper.e2e_begin(batches)
        ms_e, clocks_e, _, _ = timed(stepper.e2e_step, steps, warmup, finalize=stepper.e2e_finish)
        h2d_total = stepper.h2d_bytes_rank(batches[0])
        if world > 1:                                   # whole-job bytes: every rank copies its own rows + the cameras
            tb = torch.tensor([float(h2d_total)], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tb)
            h2d_total = int(tb.item())
        e2e = {"value": N * steps / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e / steps,
               "binding": "torch C++ extension _C" if T._lib.load_ext() is not None else "ctypes",
               "losses_read": len(stepper.losses), "clocks": clocks_e,
               "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": 4 * world,
               "what": "GaussianRasterizer fwd + L1 photometric + fused touch depth-L1 bwd; per-step camera, GT image, "
                       "touch depth and weight (each rank: the image rows of its band) copied H2D from pinned host memory (copy stream, issued one step ahead, "
                       "double-buffered); loss copied D2H every step and read by the host one step later; Gaussian "
                       "parameters are resident training state"}

    # ---- second end-to-end figure: the ALL-HOST-POINTER C entry point (tgs_train_step_host): what a non-PyTorch
    # trainer calls.  Everything crosses PCIe every step: parameters up, images up, all gradients + loss down.
    e2e_host = None
    if world == 1 and not args.no_e2e and not train_mode and not stepper.forward_only:
        import ctypes as C
        Lb = T._lib
        lib = Lb.load()
        pin = lambda t: t.detach().cpu().contiguous().pin_memory()
        hp = {k: pin(v) for k, v in params.items()}
        hp["opacities"] = pin(params["opacities"].detach().reshape(-1))
        hg = dict(dmeans2D=torch.empty(N, 3).pin_memory(), dmeans3D=torch.empty(N, 3).pin_memory(), dopacity=torch.empty(N).pin_memory(),
                  dshs=torch.empty(N, K, 3).pin_memory(), dscales=torch.empty(N, 3).pin_memory(), drotations=torch.empty(N, 4).pin_memory())
        loss_h = torch.zeros(1).pin_memory()
        bgh = torch.zeros(3).pin_memory()
        nr = C.c_int64(0)
        P_ = lambda t: C.c_void_p(t.data_ptr())

        def host_step(i):
            b = batches[i % len(batches)]
            h, cam = b["host"], b["cam"]
            st = Lb.TgsSettings(image_width=W, image_height=H, tanfovx=float(cam.tanfovx), tanfovy=float(cam.tanfovy),
                                scale_modifier=1.0, sh_degree=cfg["sh_degree"], sh_coeffs=K, depth_normalize=1,
                                viewmatrix=h["view"].data_ptr(), projmatrix=h["proj"].data_ptr(), campos=h["campos"].data_ptr(),
                                bg=bgh.data_ptr())
            gs = Lb.TgsGaussians(N=N, means3D=hp["means3D"].data_ptr(), opacities=hp["opacities"].data_ptr(), shs=hp["shs"].data_ptr(),
                                 scales=hp["scales"].data_ptr(), rotations=hp["rotations"].data_ptr())
            gr = Lb.TgsGrads(**{k: v.data_ptr() for k, v in hg.items()})
            Lb.check(lib.tgs_train_step_host(C.byref(st), C.byref(gs), P_(h["gt"]), P_(h["target"]), P_(h["weight"]), Lb.LOSS_L1,
                                             DEPTH_LOSS_MULT, C.byref(gr), None, None, None, P_(loss_h), C.byref(nr),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "tgs_train_step_host")
        hsteps = max(3, min(steps, 10))
        for i in range(3):
            host_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(hsteps):
            host_step(3 + i)                       # synchronises the stream before it returns
        dt = (time.perf_counter() - t0) / hsteps
        up = sum(v.numel() * 4 for v in hp.values()) + Stepper.h2d_bytes(batches[0]) + 12
        down = sum(v.numel() * 4 for v in hg.values()) + 4
        e2e_host = {"value": N / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": hsteps, "h2d_bytes_per_step": int(up),
                    "d2h_bytes_per_step": int(down), "loss": float(loss_h[0]), "num_rendered": int(nr.value),
                    "what": "tgs_train_step_host (C ABI, every pointer a pinned HOST pointer): Gaussian parameters, camera, GT image, "
                            "touch depth + weight uploaded, forward + L1 photometric + fused touch depth-L1 backward, all parameter "
                            "gradients + loss downloaded, stream synchronised -- every step; wall clock; scratch from cudaMallocAsync"}

    # ---- "reference CUDA path beside it" (SURVEY §8d): the upstream-STRUCTURED kernels (csrc/refstructure.cu) on
    # the same device, same scene, same cameras, same loss (touch depth-L1 formed in PyTorch, not fused)
    refcuda = None
    if world == 1 and not args.no_refcuda:
        R = T.refstructure
        rsteps = max(3, min(steps, 16))

        def ref_step(i):
            b = batches[i % len(batches)]
            d, cam = b["dev"], b["cam"]
            rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, d["view"], d["proj"],
                                                 cfg["sh_degree"], d["campos"], False, False)
            for v in params.values():
                v.grad = None
            color, _, draw, alpha = R.rasterize_refstructure(params["means3D"], params["opacities"], params["shs"],
                                                             params["scales"], params["rotations"], rs)
            loss = (color - d["gt"]).abs().sum() * stepper.inv + R.touch_depth_loss_unfused(
                draw, alpha, d["target"], d["weight"], DEPTH_LOSS_MULT, "l1")
            loss.backward()

        ms_r, clocks_r, _, _ = timed(ref_step, rsteps, 3)
        refcuda = {"value": N * rsteps / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r / rsteps, "steps": rsteps,
                   "clocks": clocks_r,
                   "what": "SUBSTITUTE for the reference's own CUDA rasterizer (not in its tree): the same algorithm in "
                           "the upstream kernels' structure, written from the spec for sm_100a (csrc/refstructure.cu): "
                           "id-order scan + blocking count read, 64-bit keys, one 12-byte-pair radix sort, 16x16 "
                           "one-pixel-per-thread compositing without culling, 10 per-thread atomics per (pixel, splat) "
                           "in backward, touch depth-L1 in PyTorch outside the kernels"}
        for v in params.values():
            v.grad = None

    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the stage timers of the timed region
    peak, peak_src = peaks()
    stage_ms = {k: (v[0] / max(v[1], 1), v[1]) for k, v in prof.items()}
    cands = ("render_fwd", "render_bwd", "sort", "preprocess", "preprocess_bwd", "pack", "duplicate")
    if train_mode:
        cands += ("adam", "photo_fwd", "photo_bwd")
    dom = max(cands, key=lambda k: prof[k][0])
    P_band = W * (stepper.y1 - stepper.y0)
    per_launch_bytes = {
        "render_bwd": 84 * I_cam0 + 32 * P_band,          # id 4 + record 40 + grad accumulate 40 per instance; 32 B / pixel
        "render_fwd": 44 * I_cam0 + 24 * P_band,          # id 4 + record 40 per instance; 24 B / pixel written
        "sort": 24 * I_cam0, "pack": 8 * I_cam0 + 96 * I_cam0, "duplicate": 20 * N + 12 * I_cam0,
        "preprocess": N * (4 * (11 + 3 * K) + 48 + 8), "preprocess_bwd": N * (4 * (11 + 3 * K) * 2 + 48),
        # train step: Adam reads p,g,m,v and writes p,m,v = 28 B per parameter element; the loss kernels
        # read the two images (24 B/px) and write / read the 3 derivative maps (36 B/px) and write the gradient (12 B/px)
        "adam": 28 * N * (11 + 3 * K), "photo_fwd": 60 * P_band, "photo_bwd": 72 * P_band,
    }[dom]
    dom_ms = stage_ms[dom][0]
    traffic = None          # DRAM bytes of the dominant kernel from the committed ncu capture (c3, camera 0)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if args.config == "c3" and world == 1 and dom in tj:
            traffic = tj[dom]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    achieved = per_launch_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    T_band = Tx * (Ty if band is None else band[1] - band[0])
    step_bytes = alg_bytes(N, I_cam0, P_band, T_band, K)
    out = {
        "metric": METRIC if not train_mode else "full train step Gaussians/sec (BASELINE config c5 pipeline: rasterizer fwd+bwd + L1/SSIM + fused touch depth-L1 + Adam + refine)",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "ms_per_step_spread": spread, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {N} Gaussians, {W}x{H}, SH deg {cfg['sh_degree']}, fused touch depth-L1 "
                               f"(mult {DEPTH_LOSS_MULT}), {len(batches)} orbit cameras cycled",
                   "num_rendered_cam0": I_cam0, "visible_cam0": n_vis,
                   "parallelism": "single GPU" if world == 1 else f"tile-row shard x{world}; [N,10] fp32 screen-gradient exchange: {exchange}",
                   "bands": None if bands_all is None else {"policy": band_policy, "tile_rows": [list(b) for b in bands_all]},
                   "rendered_hint": "off (synchronous sizing)" if args.no_hints else ("per-view instance count of the previous visit +5% (speculative sizing; "
                                    + ("count deferred: checked when backward runs, an overflow would raise)" if stepper.defer_count else "exact re-run on overflow)")),
                   "l2_policy": "working set per step (params+grads 472 MB, instance records >250 MB) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "gpu_launches": own + cub,
        "gpu_launches_detail": {"own_kernels": own, "cub_calls": cub, "per_step": (own + cub) / steps},
        "stage_ms_per_launch": {k: round(v[0], 4) for k, v in stage_ms.items()},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": per_launch_bytes, "launch_ms": dom_ms},
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms / steps * 1e-3) / 1e9,
                          "frac": step_bytes / (ms / steps * 1e-3) / 1e9 / peak},
    }
    if stepper.forward_only:
        out["metric"] = "forward-only Gaussians/sec (eval / render path: markVisible + operator forward under no_grad)"
    if train_mode:
        out["config"]["workload"] = (f"train_step on {args.config}: {N} Gaussians initially, {W}x{H}, SH deg {cfg['sh_degree']}, "
                                     f"L1+SSIM (lambda 0.2) + fused touch depth-L1 (uncertainty-weighted, mult {DEPTH_LOSS_MULT}), "
                                     f"Adam over 5 tensors in one launch, refine every {args.refine_every} steps, "
                                     f"{len(batches)} orbit cameras cycled; value = initial N x steps / time")
        out["config"]["population_history"] = stepper.n_history
        out["config"]["num_rendered_cam0"] = I_cam0
    if e2e is not None:
        out["e2e"] = e2e
    if e2e_host is not None:
        out["e2e_host_buffers"] = e2e_host
    if refcuda is not None:
        out["reference_structure_cuda"] = refcuda
    if world == 1 and not args.no_cpu_baseline:
        b0 = batches[0]
        ref = CpuReference(cfg, scene, b0["cam"], b0["host"]["target"].clone(), b0["host"]["weight"].clone())
        est, wall = ref.step()
        out["cpu_baseline"] = {"value": N / est, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.sample,
                               "estimated_ms_per_step": est * 1e3, "sample_wall_s": wall}
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
