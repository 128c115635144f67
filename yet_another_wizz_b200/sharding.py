"""
Multi-GPU partitioning of patch-pair work (replaces the dynamic master/worker queue of
`yaw.utils.parallel`, reference `src/yaw/utils/parallel.py:251-343`).

One process per GPU (`torch.distributed`, NCCL over NVLink on the B200 box, gloo in CPU
tests).  Work units are patch pairs; they are independent (the reference only scatters
their results, `src/yaw/correlation/measurements.py:354-364`), so the pairs are assigned
statically by longest-processing-time-first on a cost model, every rank counts its share
into a zero-initialised result tensor and ONE sum-reduce to rank 0 finishes the job.
Every element has exactly one non-zero contributor, so the reduction is exact and
deterministic for integers and floats alike.
"""

from __future__ import annotations

import heapq
import os

import numpy as np

__all__ = ["Shard", "assign_pairs_lpt", "assign_patches_contiguous", "current_shard", "pair_costs"]


def pair_costs(pair_i, pair_j, n1_per_patch, n2_per_patch, radii1=None) -> np.ndarray:
    """Predicted cost of each patch pair after pruning: diagonal pairs carry nearly all of the work (SURVEY.md
    section 3.1: ~95 %), neighbours only a boundary strip.  The executed pair tests of a pair scale with the rows
    of the second patch times the surface DENSITY of the first one (every tile of the second catalog meets the rows
    of the first inside its search box), so with the patch radii of the first catalog the cost is
    n1(i) / radius(i)^2 * n2(j); without them n1(i) * n2(j), which overrates large patches (measured on the C3
    benchmark, 8 ranks: most loaded rank 12 % above the mean with the product, 3 % with the density)."""
    pair_i = np.asarray(pair_i)
    pair_j = np.asarray(pair_j)
    w1 = np.asarray(n1_per_patch, dtype=np.float64)
    if radii1 is not None:
        r = np.asarray(radii1, dtype=np.float64)
        area = np.where(r > 0.0, r * r, np.inf)
        w1 = w1 / area
        w1 = w1 / max(float(w1.max()), 1e-300) * float(np.max(n1_per_patch))  # keep the scale of a row count
    cost = w1[pair_i] * np.asarray(n2_per_patch, dtype=np.float64)[pair_j]
    off = pair_i != pair_j
    cost[off] *= 0.05
    return cost + 1.0


def assign_pairs_lpt(costs: np.ndarray, world_size: int) -> list[np.ndarray]:
    """Longest-processing-time-first bin packing; returns the pair indices of every rank."""
    order = np.argsort(-np.asarray(costs), kind="stable")
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    owned: list[list[int]] = [[] for _ in range(world_size)]
    for k in order:
        load, r = heapq.heappop(heap)
        owned[r].append(int(k))
        heapq.heappush(heap, (load + float(costs[k]), r))
    return [np.array(sorted(o), dtype=np.int64) for o in owned]


def assign_patches_contiguous(patch_costs: np.ndarray, centers_xyz: np.ndarray, world_size: int) -> list[np.ndarray]:
    """Deal patches to ranks as spatially compact groups of about equal cost: patches are ordered along a
    Morton curve of their centres (longitude, z) and the order is cut where the running cost crosses
    k / world_size of the total.  A rank then needs few first-catalog patches beyond its own (only the
    neighbours across the cut), which keeps per-rank uploads and index builds ~1/world_size."""
    patch_costs = np.asarray(patch_costs, dtype=np.float64)
    xyz = np.asarray(centers_xyz, dtype=np.float64)
    lon = np.arctan2(xyz[:, 1], xyz[:, 0])
    # unwrap around the mean direction so a field that straddles RA = 0 stays contiguous
    mean_lon = np.arctan2(xyz[:, 1].sum(), xyz[:, 0].sum())
    lon = (lon - mean_lon + np.pi) % (2 * np.pi)
    def quant(v):
        span = v.max() - v.min()
        return np.zeros(len(v), dtype=np.uint64) if span <= 0 else ((v - v.min()) / span * 1023).astype(np.uint64)
    qx, qy = quant(lon), quant(xyz[:, 2])
    code = np.zeros(len(xyz), dtype=np.uint64)
    for bit in range(10):
        code |= ((qx >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2 * bit)
        code |= ((qy >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2 * bit + 1)
    order = np.argsort(code, kind="stable")
    csum = np.cumsum(patch_costs[order])
    total = csum[-1] if len(csum) else 0.0
    bounds = [0]
    for r in range(1, world_size):
        bounds.append(min(len(order), int(np.searchsorted(csum, total * r / world_size * (1 - 1e-12), side="left")) + 1))
    bounds.append(len(order))
    bounds = np.maximum.accumulate(bounds)
    return [np.sort(order[bounds[r]:bounds[r + 1]]) for r in range(world_size)]


class Shard:
    """Rank / world size of this process and the reduce used to combine results."""

    def __init__(self, rank: int = 0, world_size: int = 1, group=None) -> None:
        self.rank = rank
        self.world_size = world_size
        self.group = group

    @property
    def active(self) -> bool:
        return self.world_size > 1

    def reduce_to_root(self, array: np.ndarray, device: str | None = None) -> np.ndarray:
        """Sum `array` over the ranks.  Every element has exactly one non-zero contributor, so the sum is exact.
        The result is delivered to EVERY rank (an all-reduce: the tensors are a few hundred KB), so that
        `crosscorrelate` / `autocorrelate` return complete counts wherever they are called.  `device`: where the
        collective runs with the NCCL backend -- the engine's GPU (`cuda:<engine.device>`), not whatever
        `torch.cuda.current_device()` happens to be; ignored for CPU backends (gloo in the tests)."""
        if not self.active:
            return array
        import torch
        import torch.distributed as dist

        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dev = device or f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
        else:
            dev = "cpu"
        t = torch.from_numpy(np.ascontiguousarray(array)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


def current_shard() -> Shard:
    """The shard of this process if `torch.distributed` is initialised, else a single shard.
    torch is imported only when a launcher (torchrun) has set up the environment."""
    if "RANK" not in os.environ and "WORLD_SIZE" not in os.environ:
        return Shard()
    try:
        import torch.distributed as dist
    except Exception:
        return Shard()
    if dist.is_available() and dist.is_initialized():
        return Shard(dist.get_rank(), dist.get_world_size())
    return Shard()
