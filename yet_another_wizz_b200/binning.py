"""Redshift binning container (mirrors `yaw.Binning`, reference `src/yaw/binning.py:51-145`)."""

from __future__ import annotations

import numpy as np

__all__ = ["Binning"]


class Binning:
    """Contiguous, monotonically increasing bin edges; `closed` is "right" (default) or "left"."""

    __slots__ = ("edges", "closed")

    def __init__(self, edges, closed: str = "right") -> None:
        edges = np.asarray(edges, dtype=np.float64)
        if edges.ndim != 1 or len(edges) < 2:
            raise ValueError("bin edges must be one-dimensionals with length > 2")
        if np.any(np.diff(edges) <= 0.0):
            raise ValueError("bin edges must increase monotonically")
        closed = str(closed)
        if closed not in ("left", "right"):
            raise ValueError(f"'{closed}' is not a valid Closed")
        self.edges = edges
        self.closed = closed

    def __len__(self) -> int:
        return len(self.edges) - 1

    def __eq__(self, other) -> bool:
        if not isinstance(other, type(self)):
            return NotImplemented
        return np.array_equal(self.edges, other.edges) and self.closed == other.closed

    def __repr__(self) -> str:
        lb, rb = "[)" if self.closed == "left" else "(]"
        return f"{len(self)} bins @ {lb}{self.edges[0]:.3f}...{self.edges[-1]:.3f}{rb}"

    @property
    def mids(self) -> np.ndarray:
        return (self.edges[:-1] + self.edges[1:]) / 2.0

    @property
    def left(self) -> np.ndarray:
        return self.edges[:-1]

    @property
    def right(self) -> np.ndarray:
        return self.edges[1:]

    @property
    def dz(self) -> np.ndarray:
        return np.diff(self.edges)

    def copy(self) -> "Binning":
        return Binning(self.edges.copy(), closed=self.closed)

    def digitize(self, redshifts: np.ndarray) -> np.ndarray:
        """0-based z-bin index, -1 / len(self) for rows outside the binning
        (`np.digitize(..., right=closed=="right")`, `src/yaw/catalog/trees.py:408-414`)."""
        return np.digitize(redshifts, self.edges, right=(self.closed == "right")).astype(np.int32) - 1

    # HDF5 layout identical to the reference (src/yaw/binning.py:84-93)
    def to_hdf(self, dest) -> None:
        from .paircounts import HDF_COMPRESSION, write_version_tag

        write_version_tag(dest)
        dest.create_dataset("closed", data=str(self.closed))
        dest.create_dataset("edges", data=self.edges, **HDF_COMPRESSION)

    @classmethod
    def from_hdf(cls, source) -> "Binning":
        edges = source["edges"][:]
        closed = source["closed"][()].decode("utf-8")
        return cls(edges, closed=closed)
