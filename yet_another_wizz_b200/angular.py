"""
Angular-bin bookkeeping on the host (mirrors the helper functions of
`yaw.catalog.trees`, reference `src/yaw/catalog/trees.py:46-160`).  The GPU
returns the raw sub-bin histogram; everything here is O(n_edges) numpy.
"""

from __future__ import annotations

import math

import numpy as np

from .coordinates import AngularDistances

__all__ = [
    "parse_ang_limits", "get_ang_bins", "logarithmic_mid", "get_counts_for_limits", "squared_chord_edges",
    "AngularBinPlan",
]


def parse_ang_limits(ang_min, ang_max) -> np.ndarray:
    """`(n_scales, 2)` array of validated lower / upper angular limits in radian."""
    lo = np.atleast_1d(ang_min).astype(np.float64)
    hi = np.atleast_1d(ang_max).astype(np.float64)
    if lo.ndim != 1 or hi.ndim != 1:
        raise ValueError("'ang_min' and 'ang_max' must be 1-dim")
    if len(lo) != len(hi):
        raise ValueError("length of 'ang_min' and 'ang_max' does not match")
    if np.any(lo >= hi):
        raise ValueError("'ang_min' < 'ang_max' not satisfied")
    limits = np.column_stack((lo, hi))
    if np.any(limits < 0.0) or np.any(limits > np.pi):
        raise ValueError("'ang_min' and 'ang_max' not in range [0.0, pi]")
    return limits


def get_ang_bins(ang_range: np.ndarray, weight_scale, weight_res) -> np.ndarray:
    """Sorted unique bin edges: every scale limit plus, with r-weights, `weight_res + 1`
    log-spaced edges over the union range.  Goes through log10 -> unique -> 10** exactly like
    the reference so the edges are bit-identical."""
    with np.errstate(divide="ignore"):
        log_range = np.log10(ang_range)
    if weight_scale is not None:
        log_bins = np.linspace(log_range.min(), log_range.max(), weight_res + 1)
        log_bins = np.concatenate([log_bins, log_range.flatten()])
    else:
        log_bins = log_range.flatten()
    return 10.0 ** np.sort(np.unique(log_bins))


def logarithmic_mid(edges: np.ndarray) -> np.ndarray:
    log_edges = np.log10(edges)
    return 10.0 ** ((log_edges[:-1] + log_edges[1:]) / 2.0)


def get_counts_for_limits(counts: np.ndarray, ang_bins: np.ndarray, ang_limits: np.ndarray) -> np.ndarray:
    """Sum the sub-bins between the edges nearest to each pair of limits."""
    out = np.empty(len(ang_limits), dtype=counts.dtype)
    for i, (lo, hi) in enumerate(ang_limits):
        out[i] = counts[np.argmin(np.abs(ang_bins - lo)) : np.argmin(np.abs(ang_bins - hi))].sum()
    return out


def squared_chord_edges(ang_bins: np.ndarray) -> np.ndarray:
    """Thresholds on the squared chord length.  scipy squares the radii with libm `pow`,
    which differs from `r * r` by one ulp for ~0.08 % of values, hence
    `math.pow` per element and not `numpy ** 2`."""
    chords = AngularDistances(ang_bins).to_3d()
    return np.array([math.pow(float(r), 2.0) for r in chords], dtype=np.float64)


class AngularBinPlan:
    """Edges, thresholds and post-processing weights of every z-bin for one configuration.

    `r2[b]` goes to the GPU; `finish(hist)` turns the returned `(..., n_bins, n_sub)`
    histogram into `(n_scales, ..., n_bins)` counts exactly as `AngularTree.count` does after
    the tree query (`src/yaw/catalog/trees.py:356-362`).
    """

    def __init__(self, ang_min: np.ndarray, ang_max: np.ndarray, rweight, resolution) -> None:
        ang_min = np.atleast_2d(np.asarray(ang_min, dtype=np.float64))
        ang_max = np.atleast_2d(np.asarray(ang_max, dtype=np.float64))
        self.n_bins, self.n_scales = ang_min.shape
        self.rweight = rweight
        self.limits = [parse_ang_limits(ang_min[b], ang_max[b]) for b in range(self.n_bins)]
        self.ang_bins = [get_ang_bins(lim, rweight, resolution) for lim in self.limits]
        # np.unique may merge coinciding edges in some z-bins only: pad the threshold table with a
        # repeated last value (an empty sub-bin); the true edges are kept for the post-processing
        self.n_edges = max(len(e) for e in self.ang_bins)
        self.r2 = np.empty((self.n_bins, self.n_edges))
        for b, edges in enumerate(self.ang_bins):
            r2 = squared_chord_edges(edges)
            self.r2[b, : len(r2)] = r2
            self.r2[b, len(r2) :] = r2[-1]

    def finish(self, hist: np.ndarray) -> np.ndarray:
        """`hist[..., b, s]` -> `counts[scale, ..., b]` (float64)"""
        hist = np.asarray(hist)
        lead = hist.shape[:-2]
        out = np.zeros((self.n_scales, *lead, self.n_bins), dtype=np.float64)
        flat = hist.reshape(-1, self.n_bins, hist.shape[-1]).astype(np.float64)
        out_flat = out.reshape(self.n_scales, -1, self.n_bins)
        for b in range(self.n_bins):
            edges = self.ang_bins[b]
            counts = flat[:, b, : len(edges) - 1]
            if self.rweight is not None:
                with np.errstate(divide="ignore", invalid="ignore"):
                    ang_weights = logarithmic_mid(edges) ** self.rweight
                ang_weights = np.where(np.isfinite(ang_weights), ang_weights, 0.0)
                counts = counts * (ang_weights / ang_weights.sum())
            for s, (lo, hi) in enumerate(self.limits[b]):
                i0 = np.argmin(np.abs(edges - lo))
                i1 = np.argmin(np.abs(edges - hi))
                out_flat[s, :, b] = counts[:, i0:i1].sum(axis=1)
        return out
