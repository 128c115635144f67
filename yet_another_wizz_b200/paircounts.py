"""
Result containers filled by the engine.  Same attribute names, array shapes and
HDF5 layout as the reference's `PatchedSumWeights`, `PatchedCounts` and
`NormalisedCounts` (`src/yaw/correlation/paircounts.py:144-289, 291-457, 568-616`):

    /<counts>/{binning/{closed,edges}, auto, num_patches, patch_pairs (N x 2), binned_counts (N x n_bins)}
    /<sum_weights>/{binning, auto, sum_weights1, sum_weights2 (n_bins x n_patch)}

Resampling beyond the plain patch sums and the estimators stay on the
reference's host code (see `to_reference()` in `corrfunc.py`).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .binning import Binning

__all__ = ["NormalisedCounts", "NormalisedScalarCounts", "PatchedCounts", "PatchedSumWeights", "SampledPatchSum"]

HDF_COMPRESSION = dict(fletcher32=True, compression="gzip", shuffle=True)  # src/yaw/utils/misc.py:36
FORMAT_VERSION = "3.1.1"  # version tag of the reference whose layout is written


def write_version_tag(dest) -> None:
    if "version" not in dest:
        dest.create_dataset("version", data=FORMAT_VERSION)


@dataclass
class SampledPatchSum:
    """Sum over all patch pairs and its leave-one-patch-out jackknife samples."""

    binning: Binning
    data: np.ndarray     # (n_bins,)
    samples: np.ndarray  # (n_patch, n_bins)


def _sample_patch_sum(binning: Binning, array: np.ndarray, engine=None) -> SampledPatchSum:
    """`engine`: run the sums on the device over the non-zero patch pairs (`Engine.sample_patch_sum`, yawb_jackknife)"""
    if engine is not None:
        ids1, ids2 = np.nonzero(np.any(array, axis=0))
        total, samples = engine.sample_patch_sum(np.moveaxis(array[:, ids1, ids2], 0, -1), ids1, ids2, array.shape[1])
        return SampledPatchSum(binning, total, samples)
    # jackknife by subtraction: total - row_i - column_i + diagonal_i
    total = np.einsum("bij->b", array)
    samples = total[None, :] - np.einsum("bij->jb", array) - np.einsum("bij->ib", array) + np.einsum("bii->ib", array)
    return SampledPatchSum(binning, total, samples)


class PatchedSumWeights:
    __slots__ = ("binning", "auto", "sum_weights1", "sum_weights2")

    def __init__(self, binning: Binning, sum_weights1: np.ndarray, sum_weights2: np.ndarray, *, auto: bool) -> None:
        if sum_weights1.ndim != 2 or sum_weights2.ndim != 2:
            raise ValueError("'sum_weights1/2' must be two-dimensional")
        if sum_weights1.shape != sum_weights2.shape:
            raise ValueError("'sum_weights1' and 'sum_weights2' must have the same shape")
        if sum_weights1.shape[0] != len(binning):
            raise ValueError("first dimension of 'sum_weights1/2' must match 'binning'")
        self.binning = binning
        self.auto = bool(auto)
        self.sum_weights1 = sum_weights1.astype(np.float64)
        self.sum_weights2 = sum_weights2.astype(np.float64)

    @property
    def num_bins(self) -> int:
        return len(self.binning)

    @property
    def num_patches(self) -> int:
        return self.sum_weights1.shape[1]

    def __eq__(self, other) -> bool:
        if not isinstance(other, type(self)):
            return NotImplemented
        return (self.binning == other.binning and np.array_equal(self.sum_weights1, other.sum_weights1)
                and np.array_equal(self.sum_weights2, other.sum_weights2) and self.auto == other.auto)

    def get_array(self) -> np.ndarray:
        """`(n_bins, n_patch, n_patch)` products; auto: upper triangle with halved diagonal."""
        array = np.einsum("bi,bj->bij", self.sum_weights1, self.sum_weights2)
        if self.auto:
            array = np.triu(array)
            np.einsum("bii->bi", array)[:] *= 0.5
        return array

    def sample_patch_sum(self, engine=None) -> SampledPatchSum:
        return _sample_patch_sum(self.binning, self.get_array(), engine)

    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        self.binning.to_hdf(dest.create_group("binning"))
        dest.create_dataset("auto", data=self.auto)
        dest.create_dataset("sum_weights1", data=self.sum_weights1, **HDF_COMPRESSION)
        dest.create_dataset("sum_weights2", data=self.sum_weights2, **HDF_COMPRESSION)

    @classmethod
    def from_hdf(cls, source) -> "PatchedSumWeights":
        return cls(Binning.from_hdf(source["binning"]), source["sum_weights1"][:], source["sum_weights2"][:],
                   auto=bool(source["auto"][()]))


class PatchedCounts:
    __slots__ = ("binning", "counts", "auto")

    def __init__(self, binning: Binning, counts: np.ndarray, *, auto: bool) -> None:
        if counts.ndim != 3:
            raise ValueError("'counts' must be three-dimensional")
        if counts.shape[0] != len(binning):
            raise ValueError("first dimension of 'counts' must match 'binning'")
        if counts.shape[1] != counts.shape[2]:
            raise ValueError("'counts' must have shape (num_bins, num_patches, num_patches)")
        self.binning = binning
        self.auto = bool(auto)
        self.counts = counts.astype(np.float64)

    @classmethod
    def zeros(cls, binning: Binning, num_patches: int, *, auto: bool) -> "PatchedCounts":
        return cls(binning, np.zeros((len(binning), num_patches, num_patches)), auto=auto)

    @property
    def num_bins(self) -> int:
        return len(self.binning)

    @property
    def num_patches(self) -> int:
        return self.counts.shape[1]

    def __eq__(self, other) -> bool:
        if not isinstance(other, type(self)):
            return NotImplemented
        return self.binning == other.binning and np.array_equal(self.counts, other.counts) and self.auto == other.auto

    def get_array(self) -> np.ndarray:
        return self.counts

    def set_patch_pair(self, patch_id1: int, patch_id2: int, counts_binned: np.ndarray) -> None:
        self.counts[:, patch_id1, patch_id2] = counts_binned

    def sample_patch_sum(self, engine=None) -> SampledPatchSum:
        return _sample_patch_sum(self.binning, self.get_array(), engine)

    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        self.binning.to_hdf(dest.create_group("binning"))
        dest.create_dataset("auto", data=self.auto)
        dest.create_dataset("num_patches", data=self.num_patches)
        ids1, ids2 = np.nonzero(np.any(self.counts, axis=0))  # only patch pairs with any non-zero bin
        dest.create_dataset("patch_pairs", data=np.column_stack([ids1, ids2]), **HDF_COMPRESSION)
        dest.create_dataset("binned_counts", data=np.moveaxis(self.counts[:, ids1, ids2], 0, -1), **HDF_COMPRESSION)

    @classmethod
    def from_hdf(cls, source) -> "PatchedCounts":
        new = cls.zeros(Binning.from_hdf(source["binning"]), int(source["num_patches"][()]),
                        auto=bool(source["auto"][()]))
        for (id1, id2), counts in zip(source["patch_pairs"][:], source["binned_counts"][:]):
            new.set_patch_pair(id1, id2, counts)
        return new


class NormalisedCounts:
    """Raw pair counts + the sums of weights that normalise them."""

    __slots__ = ("_counts", "_weights")

    def __init__(self, counts: PatchedCounts, sum_weights: PatchedSumWeights) -> None:
        if counts.num_patches != sum_weights.num_patches:
            raise ValueError("number of patches of counts- and weights-container does not match")
        if counts.num_bins != sum_weights.num_bins:
            raise ValueError("number of bins of counts- and weights-container does not match")
        self._counts = counts
        self._weights = sum_weights

    @property
    def counts(self) -> PatchedCounts:
        return self._counts

    @property
    def sum_weights(self) -> PatchedSumWeights:
        return self._weights

    @property
    def binning(self) -> Binning:
        return self._counts.binning

    @property
    def auto(self) -> bool:
        return self._counts.auto

    @property
    def num_bins(self) -> int:
        return self._counts.num_bins

    @property
    def num_patches(self) -> int:
        return self._counts.num_patches

    def __eq__(self, other) -> bool:
        if type(self) is not type(other):
            return NotImplemented
        return self._counts == other._counts and self._weights == other._weights

    def get_array(self) -> np.ndarray:
        norm = self._weights.sample_patch_sum().data
        return self._counts.get_array() / norm[:, np.newaxis, np.newaxis]

    def sample_patch_sum(self, engine=None) -> SampledPatchSum:
        c, w = self._counts.sample_patch_sum(engine), self._weights.sample_patch_sum(engine)
        return SampledPatchSum(self.binning, c.data / w.data, c.samples / w.samples)

    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        self._counts.to_hdf(dest.create_group("counts"))
        self._weights.to_hdf(dest.create_group("sum_weights"))

    @classmethod
    def from_hdf(cls, source) -> "NormalisedCounts":
        return cls(PatchedCounts.from_hdf(source["counts"]), PatchedSumWeights.from_hdf(source["sum_weights"]))


class NormalisedScalarCounts:
    """Pair counts weighted by a scalar field, normalised by the plain pair counts of the same patch pairs
    (mirrors `yaw.correlation.paircounts.NormalisedScalarCounts`, reference `paircounts.py:619-666`;
    HDF5 groups `kappa_counts` / `number_counts`)."""

    __slots__ = ("_counts", "_weights")

    def __init__(self, kappa_counts: PatchedCounts, number_counts: PatchedCounts) -> None:
        if kappa_counts.num_patches != number_counts.num_patches:
            raise ValueError("number of patches of counts- and weights-container does not match")
        if kappa_counts.num_bins != number_counts.num_bins:
            raise ValueError("number of bins of counts- and weights-container does not match")
        self._counts = kappa_counts
        self._weights = number_counts

    kappa_counts = property(lambda self: self._counts)
    number_counts = property(lambda self: self._weights)
    binning = property(lambda self: self._counts.binning)
    auto = property(lambda self: self._counts.auto)
    num_bins = property(lambda self: self._counts.num_bins)
    num_patches = property(lambda self: self._counts.num_patches)

    def __eq__(self, other) -> bool:
        if type(self) is not type(other):
            return NotImplemented
        return self._counts == other._counts and self._weights == other._weights

    def sample_patch_sum(self, engine=None) -> SampledPatchSum:
        c, w = self._counts.sample_patch_sum(engine), self._weights.sample_patch_sum(engine)
        return SampledPatchSum(self.binning, c.data / w.data, c.samples / w.samples)

    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        self._counts.to_hdf(dest.create_group("kappa_counts"))
        self._weights.to_hdf(dest.create_group("number_counts"))

    @classmethod
    def from_hdf(cls, source) -> "NormalisedScalarCounts":
        return cls(PatchedCounts.from_hdf(source["kappa_counts"]), PatchedCounts.from_hdf(source["number_counts"]))
