"""
Angular coordinates / distances on the host, mirroring the small part of
`yaw.coordinates` the pair-count path uses (reference
`src/yaw/coordinates.py:134-147, 162-205, 245-277`).  Kept in float64 numpy so the
GPU receives exactly the doubles the reference's trees are built from.
"""

from __future__ import annotations

import numpy as np

__all__ = ["AngularCoordinates", "AngularDistances"]


def _sgn(val):
    # positive numbers and 0 -> +1, negative -> -1 (src/yaw/coordinates.py:31-34)
    return np.where(val == 0, 1.0, np.sign(val))


class AngularCoordinates:
    """`(N, 2)` array of (ra, dec) in radian."""

    __slots__ = ("data",)

    def __init__(self, data) -> None:
        self.data = np.atleast_2d(data).astype(np.float64, copy=False)
        if self.data.ndim != 2 or self.data.shape[1] != 2:
            raise ValueError("invalid dimensions, expected 2-dim array with shape (N, 2)")

    @classmethod
    def from_coords(cls, coords):
        return cls(np.concatenate([c.data for c in coords]))

    @classmethod
    def from_3d(cls, xyz):
        x, y, z = np.transpose(np.atleast_2d(xyz))
        r_d2 = np.sqrt(x * x + y * y)
        r_d3 = np.sqrt(x * x + y * y + z * z)
        x_normed = np.ones_like(x)
        np.divide(x, r_d2, where=r_d2 > 0.0, out=x_normed)
        ra = np.arccos(x_normed) * _sgn(y) % (2.0 * np.pi)
        dec = np.arcsin(z / r_d3)
        return cls(np.column_stack([ra, dec]))

    def to_3d(self) -> np.ndarray:
        cos_dec = np.cos(self.dec)
        return np.column_stack([np.cos(self.ra) * cos_dec, np.sin(self.ra) * cos_dec, np.sin(self.dec)])

    @property
    def ra(self) -> np.ndarray:
        return self.data[:, 0]

    @property
    def dec(self) -> np.ndarray:
        return self.data[:, 1]

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, item):
        return type(self)(self.data[item])

    def __iter__(self):
        for row in self.data:
            yield type(self)(row)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.data!r})"

    def copy(self):
        return type(self)(self.data.copy())

    def tolist(self):
        return self.data.tolist()

    def mean(self, weights=None):
        return type(self).from_3d(np.average(self.to_3d(), weights=weights, axis=0))

    def distance(self, other: "AngularCoordinates") -> "AngularDistances":
        if not isinstance(other, type(self)):
            raise TypeError(f"cannot compute distance with type {type(other)}")
        diff_sq = (self.to_3d() - other.to_3d()) ** 2
        return AngularDistances.from_3d(np.sqrt(diff_sq.sum(axis=1)))


class AngularDistances:
    """1-dim array of angular separations in radian."""

    __slots__ = ("data",)

    def __init__(self, data) -> None:
        self.data = np.atleast_1d(data).astype(np.float64, copy=False)

    @classmethod
    def from_3d(cls, dists):
        if np.any(np.asarray(dists) > 2.0):
            raise ValueError("distance exceeds size of unit sphere")
        return cls(2.0 * np.arcsin(np.asarray(dists) / 2.0))

    def to_3d(self) -> np.ndarray:
        """chord length on the unit sphere, `2 sin(theta / 2)`"""
        return 2.0 * np.sin(self.data / 2.0)

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, item):
        return type(self)(self.data[item])

    def __iter__(self):
        for val in self.data:
            yield type(self)(val)

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.data!r})"

    def __lt__(self, other):
        return self.data < other.data

    def __add__(self, other):
        return type(self)(self.data + other.data)

    def __sub__(self, other):
        return type(self)(self.data - other.data)

    def min(self):
        return type(self)(self.data.min())

    def max(self):
        return type(self)(self.data.max())

    def tolist(self):
        return self.data.tolist()
