"""
Scale -> angle conversion on the host.

Mirrors `yaw.cosmology.Scales.get_angle_radian` and its three flavours
(reference `src/yaw/cosmology.py:158-175, 223-233, 250-259, 276-285`).  The
reference evaluates distances with astropy, which is not installed in this
image; `Planck15` below restates astropy's Planck15 (flat LambdaCDM, massive
neutrinos by the Komatsu et al. 2011 fit) and is used only when the caller does
not bring a cosmology of their own.  Any object with
`angular_diameter_distance(z)` / `comoving_distance(z)` in Mpc (astropy `FLRW`
instances included) is accepted, so with the real astropy installed the GPU
receives exactly the thresholds the reference would compute.
"""

from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

__all__ = ["CustomCosmology", "Planck15", "Scales", "get_default_cosmology", "new_scales"]


class CustomCosmology(ABC):
    """Interface for user supplied cosmologies (reference `src/yaw/cosmology.py:37-72`)."""

    @abstractmethod
    def comoving_distance(self, z):
        """comoving distance in Mpc"""

    @abstractmethod
    def angular_diameter_distance(self, z):
        """angular diameter distance in Mpc"""


class _FlatLambdaCDMNu(CustomCosmology):
    def __init__(self, name, H0, Om0, Tcmb0, Neff, m_nu) -> None:
        self.name = name
        self.H0, self.Om0, self.Tcmb0, self.Neff = H0, Om0, Tcmb0, Neff
        self.m_nu = np.asarray(m_nu, dtype=np.float64)
        h = H0 / 100.0
        self.Ogamma0 = 2.4728e-5 * (Tcmb0 / 2.7255) ** 4 / h**2
        # neutrino mass over temperature, m / (k_B * T_nu0)
        self._nu_y = self.m_nu / (8.617333262e-5 * 0.71377 * Tcmb0)
        self.Onu0 = self.Ogamma0 * self._nu_relative_density(0.0)
        self.Ode0 = 1.0 - Om0 - self.Ogamma0 - self.Onu0

    def _nu_relative_density(self, z: float) -> float:
        prefac, p, invp, k = 0.22710731766, 1.83, 0.54644808743, 0.3173
        per_species = (1.0 + (k * self._nu_y / (1.0 + z)) ** p) ** invp
        return prefac * self.Neff / 3.0 * per_species.sum()

    def inv_efunc(self, z: float) -> float:
        zp1 = 1.0 + z
        o_r = self.Ogamma0 * (1.0 + self._nu_relative_density(z))
        return 1.0 / np.sqrt(zp1**3 * (o_r * zp1 + self.Om0) + self.Ode0)

    def comoving_distance(self, z):
        from scipy.integrate import quad

        d_h = 299792.458 / self.H0
        zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
        out = np.array([d_h * quad(self.inv_efunc, 0.0, zi)[0] for zi in zs])
        return out if np.ndim(z) else float(out[0])

    def angular_diameter_distance(self, z):
        return self.comoving_distance(z) / (1.0 + np.asarray(z, dtype=np.float64))


Planck15 = _FlatLambdaCDMNu("Planck15", H0=67.74, Om0=0.3075, Tcmb0=2.7255, Neff=3.046, m_nu=(0.0, 0.0, 0.06))


def get_default_cosmology():
    return Planck15


def _distance_value(dist):
    return dist.value if hasattr(dist, "value") and hasattr(dist, "unit") else dist


ANGULAR_UNITS = ("rad", "deg", "arcmin", "arcsec")
PHYSICAL_UNITS = ("kpc", "Mpc")
COMOVING_UNITS = ("kpc_h", "Mpc_h")


class Scales:
    """Lower / upper correlation scale limits in one unit; `get_angle_radian(z)` converts them."""

    def __init__(self, scale_min, scale_max, *, unit: str = "kpc") -> None:
        unit = str(unit)
        if unit not in ANGULAR_UNITS + PHYSICAL_UNITS + COMOVING_UNITS:
            raise ValueError(f"'{unit}' is not a valid separation unit")
        scale_min = np.atleast_1d(np.asarray(scale_min, dtype=np.float64))
        scale_max = np.atleast_1d(np.asarray(scale_max, dtype=np.float64))
        if scale_min.ndim != 1 or scale_max.ndim != 1:
            raise ValueError("min/max scales must be scalars or one-dimensional arrays")
        if len(scale_min) != len(scale_max):
            raise ValueError("number of elements in min and max scales does not match")
        if np.any((scale_max - scale_min) <= 0.0):
            raise ValueError("all min scales must be smaller than corresponding max scales")
        self.unit = unit
        self.scale_min = scale_min
        self.scale_max = scale_max

    def __repr__(self) -> str:
        return f"Scales(min={self.scale_min.tolist()}, max={self.scale_max.tolist()}, unit='{self.unit}')"

    @property
    def num_scales(self) -> int:
        return len(self.scale_min)

    def _compute_angle(self, scales: np.ndarray, redshift: float, cosmology) -> np.ndarray:
        if self.unit in ANGULAR_UNITS:
            if self.unit == "rad":
                return scales
            if self.unit == "arcsec":
                scales = scales / 3600.0
            elif self.unit == "arcmin":
                scales = scales / 60.0
            return np.deg2rad(scales)
        if self.unit in ("kpc", "kpc_h"):
            scales = scales / 1000.0
        if self.unit in PHYSICAL_UNITS:
            dist = cosmology.angular_diameter_distance(redshift)
        else:
            dist = cosmology.comoving_distance(redshift)
        return scales / _distance_value(dist)

    def get_angle_radian(self, redshift: float, cosmology=None):
        cosmology = cosmology or get_default_cosmology()
        return (
            self._compute_angle(self.scale_min, redshift, cosmology),
            self._compute_angle(self.scale_max, redshift, cosmology),
        )


def new_scales(scale_min, scale_max, *, unit: str = "kpc") -> Scales:
    return Scales(scale_min, scale_max, unit=unit)
