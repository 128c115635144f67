"""
TEST INFRASTRUCTURE ONLY -- not part of the product path.

Makes the *unmodified* reference (`/root/reference/src/yaw`, v3.1.1) importable
in this container, where astropy / h5py / treecorr / strenum and the generated
`yaw/_version.py` are absent.  Only used to

  * validate the oracle restatement (`oracle/oracle.py`, `oracle/cpu_port.py`),
  * generate the golden vectors under `tests/golden/` (`tests/golden/make_golden.py`).

Nothing under `tests/ -m gpu`, `__graft_entry__.smoke()` or `bench.py` imports
this module: `/root/reference` does not exist on the GPU box.

The five stubs are installed into `sys.modules` *before* `import yaw`:
  1. `yaw._version`        (setuptools_scm artefact, `src/yaw/__init__.py:6`)
  2. `strenum.StrEnum`     (case-preserving; `src/yaw/options.py:11`)
  3. `h5py`                (placeholder File/Group; HDF5 I/O unavailable)
  4. `treecorr`            (placeholder; pass `patch_centers=` explicitly)
  5. `astropy{,.units,.cosmology,.io.fits}` with a Planck15 restatement
     (flat LCDM, massive-neutrino Komatsu fit) good to ~1e-7 relative.

Parity between engine and oracle does not depend on the accuracy of (5): both
sides receive their angular thresholds from the same host function.
"""

from __future__ import annotations

import enum
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("YAW_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "yaw"))


# --------------------------------------------------------------------------- #
# Planck15 restatement (astropy.cosmology.Planck15 parameters)
# --------------------------------------------------------------------------- #
class _Planck15Like:
    """Flat LambdaCDM with massive neutrinos, restating astropy's Planck15."""

    name = "Planck15"
    H0 = 67.74
    Om0 = 0.3075
    Tcmb0 = 2.7255
    Neff = 3.046
    m_nu = (0.0, 0.0, 0.06)

    def __init__(self) -> None:
        h = self.H0 / 100.0
        # photon density: Ogamma h^2 = 2.4728e-5 (T/2.7255)^4  [4 sigma T^4 / c^3 / rho_crit]
        self.Ogamma0 = 2.472_8e-5 * (self.Tcmb0 / 2.7255) ** 4 / h**2
        self._nu_y = np.array(self.m_nu) / (8.617333262e-5 * 0.71377 * self.Tcmb0)
        self.Onu0 = self.Ogamma0 * self._nu_rel(0.0)
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0

    def _nu_rel(self, z: float) -> float:
        # Komatsu et al. 2011 fitting formula used by astropy (nu_relative_density)
        prefac = 0.22710731766  # 7/8 (4/11)^(4/3)
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr = self._nu_y / (1.0 + z)
        rel_mass_per = (1.0 + (k * curr) ** p) ** invp
        return prefac * self.Neff / 3.0 * rel_mass_per.sum()

    def _inv_efunc(self, z: float) -> float:
        zp1 = 1.0 + z
        Or = self.Ogamma0 * (1.0 + self._nu_rel(z))
        return 1.0 / np.sqrt(zp1**3 * (Or * zp1 + self.Om0) + self.Ode0)

    def comoving_distance(self, z):
        from scipy.integrate import quad

        dh = 299792.458 / self.H0
        zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
        out = np.array([dh * quad(self._inv_efunc, 0.0, zi)[0] for zi in zs])
        return out if np.ndim(z) else float(out[0])

    def angular_diameter_distance(self, z):
        return self.comoving_distance(z) / (1.0 + np.asarray(z))


class FLRW:
    """stand-in for `astropy.cosmology.FLRW` (module level so that instances pickle: the reference ships its
    `Configuration` to `multiprocessing` workers)"""


class _P15(_Planck15Like, FLRW):
    pass


class StrEnum(str, enum.Enum):
    """`strenum.StrEnum`: case preserving, unlike `enum.StrEnum`"""

    def __str__(self) -> str:
        return self.value

    @staticmethod
    def _generate_next_value_(name, start, count, last_values):
        return name


def _install_stubs() -> None:
    # 1. yaw._version
    ver = types.ModuleType("yaw._version")
    ver.__version__ = "3.1.1"
    ver.__version_tuple__ = (3, 1, 1)
    sys.modules.setdefault("yaw._version", ver)

    # 2. strenum (case preserving, unlike enum.StrEnum)
    strenum = types.ModuleType("strenum")
    strenum.StrEnum = StrEnum
    sys.modules.setdefault("strenum", strenum)

    # 3. h5py placeholder
    h5py = types.ModuleType("h5py")

    class _NoHdf:
        def __init__(self, *a, **k):
            raise RuntimeError("h5py is not available in this container")

    h5py.File = _NoHdf
    h5py.Group = _NoHdf
    sys.modules.setdefault("h5py", h5py)

    # 4. treecorr placeholder
    treecorr = types.ModuleType("treecorr")

    class _NoTreecorr:
        def __init__(self, *a, **k):
            raise RuntimeError("treecorr is not available; pass patch_centers=")

    treecorr.Catalog = _NoTreecorr
    sys.modules.setdefault("treecorr", treecorr)

    # 5. astropy
    astropy = types.ModuleType("astropy")
    units = types.ModuleType("astropy.units")

    class Quantity:  # never instantiated by the stub cosmology
        pass

    units.Quantity = Quantity
    units.Mpc = 1.0
    cosmology = types.ModuleType("astropy.cosmology")

    planck15 = _P15()
    cosmology.FLRW = FLRW
    cosmology.Planck15 = planck15
    cosmology.available = ("Planck15",)
    cosmology.cosmology_equal = lambda a, b: a is b or getattr(a, "name", 0) == getattr(
        b, "name", 1
    )

    def z_at_value(*a, **k):
        raise RuntimeError("z_at_value is not available in the astropy stub")

    cosmology.z_at_value = z_at_value
    cosmology.__dict__["Planck15"] = planck15
    io = types.ModuleType("astropy.io")
    fits = types.ModuleType("astropy.io.fits")
    io.fits = fits
    astropy.units = units
    astropy.cosmology = cosmology
    astropy.io = io
    for name, mod in (
        ("astropy", astropy),
        ("astropy.units", units),
        ("astropy.cosmology", cosmology),
        ("astropy.io", io),
        ("astropy.io.fits", fits),
    ):
        sys.modules.setdefault(name, mod)


def import_reference():
    """Return the imported, unmodified reference package `yaw`."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    os.environ.setdefault("YAW_NUM_THREADS", "1")  # fixed at import (parallel.py:137-142)
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import yaw  # noqa: E402

    return yaw
