"""Tile-row sharding of one camera's image across the GPUs of a box.

Every rank holds a full replica of the Gaussians and renders a contiguous band of tile rows; the
backward produces *partial* screen-space gradients [N,10]; the chain-rule kernel gathers them straight from the
peers' buffers over NVLink (PeerScreenGrads), or they are summed with ONE all-reduce per step (NCCL on GPUs, gloo in
the CPU tests).  No other data-path collective exists.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import _lib

TILE = 16


def tile_rows(image_height: int) -> int:
    return (image_height + TILE - 1) // TILE


def even_bands(image_height: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal tile-row bands: rank r owns [Ty*r/g, Ty*(r+1)/g)."""
    Ty = tile_rows(image_height)
    return [((Ty * r) // world_size, (Ty * (r + 1)) // world_size) for r in range(world_size)]


def balanced_bands(row_work: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Bands balanced by per-tile-row work (e.g. instances per row measured on a previous step)
    instead of row count: greedy prefix split at the g-quantiles of the cumulative work.
    Every rank gets at least one row while rows remain."""
    Ty = len(row_work)
    total = float(sum(row_work))
    if total <= 0 or world_size <= 1:
        return [((Ty * r) // world_size, (Ty * (r + 1)) // world_size) for r in range(world_size)]
    bounds = [0]
    acc = 0.0
    row = 0
    for r in range(1, world_size):
        target = total * r / world_size
        while row < Ty and acc + row_work[row] * 0.5 < target:
            acc += row_work[row]
            row += 1
        row = max(row, bounds[-1] + (1 if bounds[-1] < Ty - (world_size - r) else 0))
        row = min(row, Ty - (world_size - r))
        row = max(row, bounds[-1])
        acc = float(sum(row_work[:row]))
        bounds.append(row)
    bounds.append(Ty)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def band_pixel_rows(band: Tuple[int, int], image_height: int) -> Tuple[int, int]:
    return min(band[0] * TILE, image_height), min(band[1] * TILE, image_height)


def halo_bands(image_height: int, world_size: int, halo_tiles: int = 1):
    """Bands of the sharded TRAIN STEP: per rank (tile rows to render incl. halo, loss pixel rows, gradient pixel
    rows).  The 11x11 SSIM window reaches 5 rows across a band border, so every rank renders `halo_tiles` extra tile
    rows on each side; its loss (and touch loss) pixels stay its own band; its image gradient covers band + halo."""
    Ty = tile_rows(image_height)
    own = even_bands(image_height, world_size)
    out = []
    for b0, b1 in own:
        ext = (max(0, b0 - halo_tiles), min(Ty, b1 + halo_tiles))
        out.append((ext, band_pixel_rows((b0, b1), image_height), band_pixel_rows(ext, image_height)))
    return out


def all_reduce_screen_grads(sgrad, group=None):
    """The single exchange step: sum the [N,10] partial screen-space gradients over ranks."""
    import torch.distributed as dist
    dist.all_reduce(sgrad, op=dist.ReduceOp.SUM, group=group)
    return sgrad


class PeerScreenGrads:
    """Peer-mapped (symmetric-memory) [N,10] screen-gradient buffers for the FUSED exchange of the tile-row shard:
    ``tgs_backward_preprocess_gather`` reads the partial sums straight out of every peer's buffer over NVLink, so
    there is no all-reduce and no reduction kernel between BACKWARD::render and the per-Gaussian chain rule.

    Two buffers are used alternately: rank A may still be reading rank B's buffer k while B already renders step
    k+1 into the other one; B cannot reach step k+2 (which overwrites buffer k) before the barrier of step k+1, and
    A only enters that barrier after its reads of step k have been enqueued ahead of it -- one barrier per step.

    ``bands[r]`` = tile rows rank r renders (including any halo); set by the caller before each step.
    torch's symmetric memory provides the allocation, the handle exchange and the barrier (plumbing); the gather
    itself is our kernel."""

    def __init__(self, group, max_gaussians: int, device, n_grad: int = 10):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("PeerScreenGrads supports up to 8 ranks of one NVSwitch box")
        self.capacity, self.n_grad = int(max_gaussians), n_grad
        self.bufs, self.handles, self.ptr_arrays = [], [], []
        name = self.group.group_name
        try:
            symm.enable_symm_mem_for_group(name)       # required by older torch, a deprecated no-op in newer ones
        except Exception:  # noqa: BLE001
            pass
        for _ in range(2):
            # rows + contributor bytes (TgsSettings.contrib_flags): the gather asks a peer only for the rows it wrote
            t = symm.empty(max(_lib.screen_grad_floats(self.capacity, True), 1), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, name)
            ptrs = [int(p) for p in h.buffer_ptrs]
            self.bufs.append(t)
            self.handles.append(h)
            self.ptr_arrays.append((C.c_void_p * self.world)(*ptrs))
        self.k = 0
        self.bands = None

    def acquire(self, N: int):
        """-> (this step's buffer of THIS rank: screen_grad_floats(N, True) floats = [N,10] rows + contributor bytes,
        ctypes array of all ranks' pointers, barrier handle)"""
        if N > self.capacity:
            raise ValueError(f"PeerScreenGrads capacity {self.capacity} < {N} Gaussians")
        i = self.k & 1
        self.k += 1
        return self.bufs[i][: _lib.screen_grad_floats(N, True)], self.ptr_arrays[i], self.handles[i]

    def band_array(self):
        import ctypes as C
        if self.bands is None or len(self.bands) != self.world:
            raise ValueError("PeerScreenGrads.bands must list the rendered tile rows of every rank")
        flat = [int(v) for b in self.bands for v in b]
        return (C.c_int32 * (2 * self.world))(*flat)


def make_peer_exchange(group, max_gaussians: int, device):
    """PeerScreenGrads if symmetric memory works on EVERY rank of the group (NVLink P2P between all ranks), else None
    (the caller then uses the NCCL all-reduce path).  The decision is collective: a MIN all-reduce of the per-rank
    success flag, so a failure on some ranks cannot leave the others waiting in the peer barrier while those call
    all_reduce.  Both paths are GPU paths of this library; there is no CPU fallback."""
    import torch
    import torch.distributed as dist
    peer, err = None, None
    try:
        peer = PeerScreenGrads(group, max_gaussians, device)
    except Exception as e:  # noqa: BLE001
        err = e
    grp = group if group is not None else dist.group.WORLD
    flag = torch.tensor([1 if peer is not None else 0], dtype=torch.int32,
                        device=device if dist.get_backend(grp) == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=grp)
    if int(flag.item()) == 1:
        return peer
    import warnings
    why = f"{type(err).__name__}: {err}" if err is not None else "a peer rank could not set it up"
    warnings.warn(f"symmetric-memory peer exchange unavailable ({why}); every rank uses the NCCL all-reduce")
    return None
