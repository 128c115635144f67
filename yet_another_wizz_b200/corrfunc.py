"""
`CorrFunc(dd, dr, rd, rr)` container returned by `crosscorrelate` / `autocorrelate`
(mirrors `yaw.CorrFunc`, reference `src/yaw/correlation/corrfunc.py:280-427`; HDF5
group names `data_data / data_random / random_data / random_random`, `:323-325`).

Estimators, jackknife covariance and n(z) stay on the reference's host code: call
`to_reference()` to obtain a genuine `yaw.CorrFunc` holding the same arrays.
"""

from __future__ import annotations

from .paircounts import NormalisedCounts, NormalisedScalarCounts, write_version_tag

__all__ = ["CorrFunc", "EstimatorError", "ScalarCorrFunc"]

_COUNTS_NAME = dict(dd="data_data", dr="data_random", rd="random_data", rr="random_random")


class EstimatorError(Exception):
    pass


class CorrFunc:
    __slots__ = ("_counts_dict",)

    def __init__(self, dd: NormalisedCounts, dr: NormalisedCounts | None = None,
                 rd: NormalisedCounts | None = None, rr: NormalisedCounts | None = None) -> None:
        if type(dd) is not NormalisedCounts:
            raise TypeError(f"pair counts must be of type {NormalisedCounts}")
        if dr is None and rd is None and rr is None:
            raise EstimatorError("missing at least one additional pair count")
        self._counts_dict = dict(dd=dd)
        for kind, count in dict(dr=dr, rd=rd, rr=rr).items():
            if count is None:
                continue
            if count.num_patches != dd.num_patches or count.binning != dd.binning:
                raise ValueError(f"pair counts '{kind}' and 'dd' are not compatible")
            self._counts_dict[kind] = count

    dd = property(lambda self: self._counts_dict["dd"])
    dr = property(lambda self: self._counts_dict.get("dr"))
    rd = property(lambda self: self._counts_dict.get("rd"))
    rr = property(lambda self: self._counts_dict.get("rr"))

    @property
    def binning(self):
        return self.dd.binning

    @property
    def auto(self) -> bool:
        return self.dd.auto

    @property
    def num_patches(self) -> int:
        return self.dd.num_patches

    def __repr__(self) -> str:
        return (f"CorrFunc(counts={'|'.join(self._counts_dict)}, auto={self.auto}, "
                f"binning={self.binning}, num_patches={self.num_patches})")

    def __eq__(self, other) -> bool:
        if type(self) is not type(other):
            return NotImplemented
        keys = set(self._counts_dict) | set(other._counts_dict)
        return all(self._counts_dict.get(k) == other._counts_dict.get(k) for k in keys)

    def to_dict(self) -> dict:
        return self._counts_dict.copy()

    # ---- HDF5, same layout as the reference -------------------------------------------------------
    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        dest.create_dataset("kind", data="CorrFunc")
        for kind, count in self._counts_dict.items():
            count.to_hdf(dest.create_group(_COUNTS_NAME[kind]))

    @classmethod
    def from_hdf(cls, source) -> "CorrFunc":
        kwargs = {kind: NormalisedCounts.from_hdf(source[name]) for kind, name in _COUNTS_NAME.items()
                  if name in source}
        return cls(**kwargs)

    def to_file(self, path) -> None:
        import h5py  # optional dependency, absent in the build image

        with h5py.File(str(path), mode="w") as f:
            self.to_hdf(f)

    @classmethod
    def from_file(cls, path) -> "CorrFunc":
        import h5py

        with h5py.File(str(path)) as f:
            return cls.from_hdf(f)

    # ---- hand-over to the reference's host code -------------------------------------------------------
    def to_reference(self):
        """The same pair counts as a genuine `yaw.CorrFunc` (needs `yaw` importable), so
        `sample()`, the Davis-Peebles / Landy-Szalay estimators and `RedshiftData` run unchanged."""
        import yaw
        from yaw.correlation import paircounts as ref_pc

        def convert(nc: NormalisedCounts):
            binning = yaw.Binning(nc.binning.edges, closed=nc.binning.closed)
            counts = ref_pc.PatchedCounts(binning, nc.counts.counts, auto=nc.auto)
            sumw = ref_pc.PatchedSumWeights(binning, nc.sum_weights.sum_weights1, nc.sum_weights.sum_weights2,
                                            auto=nc.auto)
            return ref_pc.NormalisedCounts(counts, sumw)

        return yaw.CorrFunc(**{kind: convert(nc) for kind, nc in self._counts_dict.items()})


class ScalarCorrFunc:
    """`ScalarCorrFunc(dd, dr)` returned by `crosscorrelate_scalar` / `autocorrelate_scalar` (mirrors
    `yaw.ScalarCorrFunc`, reference `src/yaw/correlation/corrfunc.py:352-400`; HDF5 groups `data_data`,
    `data_random`)."""

    __slots__ = ("_counts_dict",)

    def __init__(self, dd: NormalisedScalarCounts, dr: NormalisedScalarCounts | None = None) -> None:
        if type(dd) is not NormalisedScalarCounts:
            raise TypeError(f"pair counts must be of type {NormalisedScalarCounts}")
        self._counts_dict = dict(dd=dd)
        if dr is not None:
            if dr.num_patches != dd.num_patches or dr.binning != dd.binning:
                raise ValueError("pair counts 'dr' and 'dd' are not compatible")
            self._counts_dict["dr"] = dr

    dd = property(lambda self: self._counts_dict["dd"])
    dr = property(lambda self: self._counts_dict.get("dr"))
    binning = property(lambda self: self.dd.binning)
    auto = property(lambda self: self.dd.auto)
    num_patches = property(lambda self: self.dd.num_patches)

    def __repr__(self) -> str:
        return (f"ScalarCorrFunc(counts={'|'.join(self._counts_dict)}, auto={self.auto}, "
                f"binning={self.binning}, num_patches={self.num_patches})")

    def to_dict(self) -> dict:
        return self._counts_dict.copy()

    def to_hdf(self, dest) -> None:
        write_version_tag(dest)
        dest.create_dataset("kind", data="ScalarCorrFunc")
        for kind, count in self._counts_dict.items():
            count.to_hdf(dest.create_group(dict(dd="data_data", dr="data_random")[kind]))
