"""Tile-row sharding of one camera's image across the GPUs of a box.

Every rank holds a full replica of the Gaussians and renders a contiguous band of tile rows; the
backward produces *partial* screen-space gradients [N,10] which are summed with ONE all-reduce per
step (NCCL over NVLink on GPUs, gloo in the CPU tests).  No other data-path collective exists.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

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


def all_reduce_screen_grads(sgrad, group=None):
    """The single exchange step: sum the [N,10] partial screen-space gradients over ranks."""
    import torch.distributed as dist
    dist.all_reduce(sgrad, op=dist.ReduceOp.SUM, group=group)
    return sgrad
