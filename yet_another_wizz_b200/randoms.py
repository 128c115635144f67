"""
Uniform random points in an RA/Dec window -- the synthetic-input generator of the
benchmarks.  Mirrors `yaw.randoms.BoxRandoms`
(reference `src/yaw/randoms.py:85-96, 109-130, 187-259`): uniform in RA x sin(Dec),
seeded through `SeedSequence(seed).spawn(1)[0]`, optional redshifts / weights
resampled from a pool.  HealPix randoms are out of scope.
"""

from __future__ import annotations

import numpy as np

__all__ = ["BoxRandoms"]


class BoxRandoms:
    def __init__(self, ra_min: float, ra_max: float, dec_min: float, dec_max: float, *, weights=None,
                 redshifts=None, seed: int = 12345) -> None:
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.redshifts = None if redshifts is None else np.asarray(redshifts, dtype=np.float64)
        if self.weights is not None and self.redshifts is not None and len(self.weights) != len(self.redshifts):
            raise ValueError("number of 'weights' and 'redshifts' to draw from does not match")
        self.x_min, self.y_min = np.deg2rad(ra_min), np.sin(np.deg2rad(dec_min))
        self.x_max, self.y_max = np.deg2rad(ra_max), np.sin(np.deg2rad(dec_max))
        self.reseed(seed)

    def reseed(self, seed: int | None = None) -> None:
        if seed is not None:
            self.seed = int(seed)
        self.rng = np.random.default_rng(np.random.SeedSequence(self.seed).spawn(1)[0])

    @property
    def data_size(self) -> int:
        if self.weights is None and self.redshifts is None:
            return -1
        return len(self.redshifts if self.weights is None else self.weights)

    def __call__(self, probe_size: int) -> dict:
        """Draw `probe_size` points; returns a dict with `ra`, `dec` (radian) and the
        optional `weights` / `redshifts` (same draw order as the reference:
        RA, then sin(Dec), then the attribute indices)."""
        x = self.rng.uniform(self.x_min, self.x_max, probe_size)
        y = self.rng.uniform(self.y_min, self.y_max, probe_size)
        out = dict(ra=x, dec=np.arcsin(y))
        if self.data_size != -1:
            idx = self.rng.integers(0, self.data_size, size=probe_size)
            if self.weights is not None:
                out["weights"] = self.weights[idx]
            if self.redshifts is not None:
                out["redshifts"] = self.redshifts[idx]
        return out
