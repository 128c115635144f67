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

__all__ = ["Shard", "assign_pairs_lpt", "assign_patches_contiguous", "assign_patch_fractions", "split_rows", "current_shard",
           "pair_costs"]


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


def _morton_order(centers_xyz: np.ndarray) -> np.ndarray:
    """patches ordered along a Morton curve of their centres (longitude unwrapped around the mean direction, z)"""
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
    return np.argsort(code, kind="stable")


def assign_patches_contiguous(patch_costs: np.ndarray, centers_xyz: np.ndarray, world_size: int) -> list[np.ndarray]:
    """Deal patches to ranks as spatially compact groups of about equal cost: patches are ordered along a
    Morton curve of their centres (longitude, z) and the order is cut where the running cost crosses
    k / world_size of the total.  A rank then needs few first-catalog patches beyond its own (only the
    neighbours across the cut), which keeps per-rank uploads and index builds ~1/world_size."""
    patch_costs = np.asarray(patch_costs, dtype=np.float64)
    order = _morton_order(centers_xyz)
    csum = np.cumsum(patch_costs[order])
    total = csum[-1] if len(csum) else 0.0
    bounds = [0]
    for r in range(1, world_size):
        bounds.append(min(len(order), int(np.searchsorted(csum, total * r / world_size * (1 - 1e-12), side="left")) + 1))
    bounds.append(len(order))
    bounds = np.maximum.accumulate(bounds)
    return [np.sort(order[bounds[r]:bounds[r + 1]]) for r in range(world_size)]


def assign_patch_fractions(patch_costs: np.ndarray, centers_xyz: np.ndarray, world_size: int,
                           min_fraction: float = 0.03) -> list[list[tuple[int, float, float]]]:
    """The same compact groups, cut EXACTLY at k / world_size of the total cost: the patch a cut falls into is
    shared by the two ranks, each taking a fraction [f0, f1) of its second-catalog rows (work items are tiles of the
    second catalog, so any subset of a patch's rows is a valid share; `split_rows` makes the subsets compact).
    With 64 patches on 8 ranks whole patches leave the most loaded rank 12 % above the mean (C3 benchmark);
    fractions bring that to the accuracy of the cost model.  Slivers below `min_fraction` of a patch are not
    worth a second copy of its first-catalog neighbours: such a cut snaps to the patch boundary.
    Returns, per rank, a list of (patch, f0, f1)."""
    patch_costs = np.asarray(patch_costs, dtype=np.float64)
    order = _morton_order(centers_xyz)
    cost = patch_costs[order]
    total = float(cost.sum())
    hi = np.cumsum(cost)
    lo = hi - cost
    cuts = [0.0]
    for r in range(1, world_size):
        c = total * r / world_size
        k = int(np.searchsorted(hi, c, side="right"))  # the patch the cut falls into
        if k < len(order) and cost[k] > 0.0:
            f = (c - lo[k]) / cost[k]
            if f < min_fraction:
                c = lo[k]
            elif f > 1.0 - min_fraction:
                c = hi[k]
        cuts.append(max(c, cuts[-1]))
    cuts.append(total)
    shares: list[list[tuple[int, float, float]]] = []
    for r in range(world_size):
        a, b = cuts[r], cuts[r + 1]
        mine = []
        for k, p in enumerate(order):
            if cost[k] <= 0.0:
                if a <= lo[k] < b or (r == world_size - 1 and lo[k] >= b):
                    mine.append((int(p), 0.0, 1.0))  # empty patches travel with their position on the curve
                continue
            f0 = min(max((a - lo[k]) / cost[k], 0.0), 1.0)
            f1 = min(max((b - lo[k]) / cost[k], 0.0), 1.0)
            if f1 - f0 > 1e-12:
                mine.append((int(p), float(f0), float(f1)))
        shares.append(mine)
    return shares


def split_rows(xyz: np.ndarray, f0: float, f1: float) -> np.ndarray:
    """Indices of the rows of ONE patch that make up its fraction [f0, f1): the rows are ordered along the longer
    axis of the patch (longitude or latitude), so every share is a compact strip.  Deterministic: every rank
    derives the same strips."""
    n = len(xyz)
    if f0 <= 0.0 and f1 >= 1.0:
        return np.arange(n)
    xyz = np.asarray(xyz, dtype=np.float64)
    lat = np.arcsin(np.clip(xyz[:, 2], -1.0, 1.0))
    mean_lon = np.arctan2(xyz[:, 1].sum(), xyz[:, 0].sum())
    lon = (np.arctan2(xyz[:, 1], xyz[:, 0]) - mean_lon + np.pi) % (2 * np.pi)
    span_lon = (lon.max() - lon.min()) * np.cos(lat.mean()) if n else 0.0
    key = lon if span_lon >= (lat.max() - lat.min() if n else 0.0) else lat
    order = np.argsort(key, kind="stable")
    return np.sort(order[int(round(f0 * n)):int(round(f1 * n))])


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
