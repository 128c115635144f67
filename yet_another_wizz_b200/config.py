"""
Read-only configuration objects with the attribute surface the pair-count path
consumes (reference `src/yaw/config/classes.py:55-252` ScalesConfig, `:253-599`
BinningConfig, `:599-860` Configuration).  YAML round-trips and the parameter
specification machinery of the reference are out of scope (SURVEY.md section 2,
row 7); a real `yaw.Configuration` can be passed to `crosscorrelate` /
`autocorrelate` instead of this class -- both are consumed through the same
attributes: `scales.{scales, rweight, resolution, num_scales}`,
`binning.{binning, edges, closed, zmin}`, `cosmology`, `max_workers`.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .binning import Binning
from .cosmology import Scales, get_default_cosmology, new_scales

__all__ = ["BinningConfig", "Configuration", "ConfigError", "ScalesConfig"]


class ConfigError(Exception):
    pass


@dataclass(frozen=True)
class ScalesConfig:
    scales: Scales
    rweight: float | None = None
    resolution: int | None = None

    @property
    def rmin(self):
        return self.scales.scale_min.squeeze().tolist()

    @property
    def rmax(self):
        return self.scales.scale_max.squeeze().tolist()

    @property
    def unit(self) -> str:
        return str(self.scales.unit)

    @property
    def num_scales(self) -> int:
        return self.scales.num_scales

    @classmethod
    def create(cls, *, rmin, rmax, unit: str = "kpc", rweight=None, resolution=None) -> "ScalesConfig":
        try:
            scales = new_scales(rmin, rmax, unit=unit)
        except Exception as err:
            raise ConfigError(str(err)) from err
        return cls(
            scales,
            None if rweight is None else float(rweight),
            None if resolution is None else int(resolution),
        )


def _make_edges(zmin: float, zmax: float, num_bins: int, method: str, cosmology) -> np.ndarray:
    # RedshiftBinningFactory, src/yaw/cosmology.py:288-342
    if method == "linear":
        return np.linspace(zmin, zmax, num_bins + 1)
    if method == "logspace":
        log_min, log_max = np.log([1.0 + zmin, 1.0 + zmax])
        return np.logspace(log_min, log_max, num_bins + 1, base=np.e) - 1.0
    if method == "comoving":
        from scipy.optimize import brentq

        cmin, cmax = (float(np.asarray(cosmology.comoving_distance(z))) for z in (zmin, zmax))
        targets = np.linspace(cmin, cmax, num_bins + 1)
        edges = [brentq(lambda z, t=t: float(np.asarray(cosmology.comoving_distance(z))) - t, zmin, zmax)
                 for t in targets[1:-1]]
        return np.array([zmin, *edges, zmax])
    raise ConfigError(f"invalid binning method '{method}'")


@dataclass(frozen=True)
class BinningConfig:
    binning: Binning
    method: str = "linear"

    @property
    def edges(self) -> list:
        return self.binning.edges.tolist()

    @property
    def zmin(self) -> float:
        return float(self.binning.edges[0])

    @property
    def zmax(self) -> float:
        return float(self.binning.edges[-1])

    @property
    def num_bins(self) -> int:
        return len(self.binning)

    @property
    def closed(self) -> str:
        return str(self.binning.closed)

    @property
    def is_custom(self) -> bool:
        return self.method == "custom"

    @classmethod
    def create(cls, *, zmin=None, zmax=None, num_bins: int = 30, method: str = "linear", edges=None,
               closed: str = "right", cosmology=None) -> "BinningConfig":
        try:
            if edges is not None:
                return cls(Binning(edges, closed=closed), "custom")
            if zmin is None or zmax is None:
                raise ConfigError("either 'edges' or 'zmin' and 'zmax' are required")
            cosmology = cosmology or get_default_cosmology()
            return cls(Binning(_make_edges(float(zmin), float(zmax), int(num_bins), str(method), cosmology),
                               closed=closed), str(method))
        except ConfigError:
            raise
        except Exception as err:
            raise ConfigError(str(err)) from err


@dataclass(frozen=True)
class Configuration:
    scales: ScalesConfig
    binning: BinningConfig
    cosmology: object = None
    max_workers: int | None = None

    def __post_init__(self):
        if self.cosmology is None:
            object.__setattr__(self, "cosmology", get_default_cosmology())

    @classmethod
    def create(cls, *, rmin, rmax, unit: str = "kpc", rweight=None, resolution=None, zmin=None, zmax=None,
               num_bins: int = 30, method: str = "linear", edges=None, closed: str = "right",
               cosmology=None, max_workers=None) -> "Configuration":
        cosmology = cosmology or get_default_cosmology()
        if isinstance(cosmology, str):
            if cosmology != "Planck15":
                raise ConfigError(f"unknown cosmology '{cosmology}' (only 'Planck15' is built in)")
            cosmology = get_default_cosmology()
        scales = ScalesConfig.create(rmin=rmin, rmax=rmax, unit=unit, rweight=rweight, resolution=resolution)
        binning = BinningConfig.create(zmin=zmin, zmax=zmax, num_bins=num_bins, method=method, edges=edges,
                                       closed=closed, cosmology=cosmology)
        return cls(scales, binning, cosmology, max_workers)
