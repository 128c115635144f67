"""
ORACLE / CPU BASELINE -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Port of the reference's CPU implementation of the hot path, used by `bench.py` as the
`cpu_baseline` leg and as the `--impl reference` arm on the GPU box, where
`/root/reference` does not exist (`cpu_baseline.kind = "port"`).  It follows the
reference's algorithm exactly -- same third-party kernel, same work decomposition:

  * one `scipy.spatial.KDTree(xyz, leafsize=16)` per patch and z-bin
    (`AngularTree.__init__`, src/yaw/catalog/trees.py:215-246; `build_trees` :365-429);
  * per patch pair and z-bin one dual-tree `count_neighbors(r=chords, weights=(w1, w2),
    cumulative=len(bins) < 8)` (`AngularTree.count`, trees.py:303-362; loop of
    `process_patch_pair`, src/yaw/correlation/measurements.py:88-128);
  * a task farm over patch pairs with `multiprocessing.Pool.imap_unordered`
    (src/yaw/utils/parallel.py:318-343), diagonal pairs first (measurements.py:258-289).

Differences from the reference that do not change the arithmetic: trees are built on the
worker pool like the reference's, but kept in the worker processes' memory (fork-inherited)
instead of being pickled to disk and re-read per pair, and the scale->angle conversion is
evaluated once per z-bin instead of once per pair.  Both make this port slightly FASTER than
the reference, i.e. a conservative baseline.

Pinned against the live reference in `tests/test_vs_reference.py::test_cpu_port_counts_equal_reference`.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np
from scipy.spatial import KDTree

import oracle

_STATE: dict = {}


class PortTree:
    """`AngularTree` (trees.py:163-362) on plain arrays."""

    __slots__ = ("tree", "weights", "sum_weights", "num_records")

    def __init__(self, xyz: np.ndarray, weights: np.ndarray | None, leafsize: int = 16):
        self.num_records = len(xyz)
        self.weights = weights
        self.sum_weights = float(len(xyz)) if weights is None else float(weights.sum())
        self.tree = KDTree(xyz, leafsize=leafsize, copy_data=True) if len(xyz) else None

    def count(self, other: "PortTree", ang_limits, ang_bins, chords, weight_scale) -> np.ndarray:
        if self.tree is None or other.tree is None:
            return np.zeros(len(ang_limits))
        cumulative = len(ang_bins) < 8
        counts = self.tree.count_neighbors(
            other.tree, r=chords, weights=(self.weights, other.weights), cumulative=cumulative
        ).astype(np.float64)
        counts = np.diff(counts) if cumulative else counts[1:]
        if weight_scale is not None:
            ang_weights = oracle.logarithmic_mid(ang_bins) ** weight_scale
            counts *= ang_weights / ang_weights.sum()
        return oracle.get_counts_for_limits(counts, ang_bins, ang_limits)


def build_patch_trees(xyz, weights, zbin, n_bins: int | None):
    """one tree per z-bin (or a single tree) of one patch"""
    if n_bins is None:
        return PortTree(xyz, weights)
    trees = []
    for b in range(n_bins):
        m = zbin == b
        trees.append(PortTree(xyz[m], None if weights is None else weights[m]))
    return tuple(trees)


def _process_patch_pair(pair):
    """`process_patch_pair`, measurements.py:88-128"""
    i, j = pair
    trees1, trees2 = _STATE["trees1"][i], _STATE["trees2"][j]
    plan = _STATE["plan"]
    n_bins = len(plan)
    counts = np.empty((len(plan[0][0]), n_bins))
    for b in range(n_bins):
        t1 = trees1[b]
        t2 = trees2[b] if isinstance(trees2, tuple) else trees2
        ang_limits, ang_bins, chords = plan[b]
        counts[:, b] = t1.count(t2, ang_limits, ang_bins, chords, _STATE["rweight"])
    return i, j, counts


def make_plan(ang_min, ang_max, rweight, resolution):
    plan = []
    for lo, hi in zip(np.atleast_2d(ang_min), np.atleast_2d(ang_max)):
        lim = oracle.parse_ang_limits(lo, hi)
        bins = oracle.get_ang_bins(lim, rweight, resolution)
        plan.append((lim, bins, oracle.angle_to_chord(bins)))
    return plan


def count_pairs(trees1, trees2, pairs, ang_min, ang_max, *, rweight=None, resolution=None, workers: int = 1):
    """Task farm over patch pairs; returns `{(i, j): counts[n_scales, n_bins]}` and the seconds spent."""
    _STATE.update(trees1=trees1, trees2=trees2, plan=make_plan(ang_min, ang_max, rweight, resolution),
                  rweight=rweight)
    pairs = sorted(pairs, key=lambda p: p[0] != p[1])  # diagonal (slowest) jobs first
    t0 = time.perf_counter()
    out = {}
    if workers <= 1:
        for res in map(_process_patch_pair, pairs):
            out[(res[0], res[1])] = res[2]
    else:
        with mp.get_context("fork").Pool(workers) as pool:  # trees are inherited by the forked workers
            for res in pool.imap_unordered(_process_patch_pair, pairs):
                out[(res[0], res[1])] = res[2]
    return out, time.perf_counter() - t0


def _build_one(args):
    xyz, weights, zbin, n_bins = args
    return build_patch_trees(xyz, weights, zbin, n_bins)


def build_catalog_trees(patch_rows, n_bins: int | None, workers: int = 1):
    """`Catalog.build_trees` (catalog.py:1406-1460): `patch_rows[p] = (xyz, weights, zbin)`, one task per
    patch on the worker pool (`parallel.iter_unordered(BinnedTrees.build, patches)`, catalog.py:1450-1456).
    The reference's workers pickle their trees to disk (trees.py:529-543); here they travel back to the parent
    pickled through the pool's pipe, so that the forked count workers inherit them."""
    t0 = time.perf_counter()
    args = [(x, w, z, n_bins) for (x, w, z) in patch_rows]
    if workers <= 1 or len(args) <= 1:
        trees = [_build_one(a) for a in args]
    else:
        with mp.get_context("fork").Pool(min(workers, len(args))) as pool:
            trees = pool.map(_build_one, args, chunksize=1)
    return trees, time.perf_counter() - t0


def physical_cores() -> int:
    """cores this process may use (the reference caps its pool at lscpu 'Core(s) per socket',
    src/yaw/utils/parallel.py:53-85; we report what is actually used)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
