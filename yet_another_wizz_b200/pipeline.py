"""
Pipelined schedule of the four pair counts of a cross-correlation for HOST-resident inputs.

With the rows resident in HBM a C3 pass takes ~10 ms, but pushing its 788 MB through PCIe takes ~16 ms,
so an end-to-end call is bound by the copy and by whatever is left to do once the last byte has arrived.
The schedule keeps that tail short:

* every copy is enqueued before the first count, smaller catalogs first (reference sample, unknown sample,
  then the randoms): the cheap DD count gets the GPU going a few ms into the transfer, and every count
  that does not need the last catalog has finished when it arrives;
* counts are issued in the order in which their inputs become complete; each builds the indexes it needs
  on first use (the first catalog's before the host waits for the copies of the second one);
* optionally (`groups` > 1) the unbinned catalogs (unknown sample, its randoms; "second role", register
  tiles) travel in slices of whole patches, each an independent device catalog; a patch pair (i, j) only
  ever needs the slice that holds patch j, so the counts against a slice can run while the next slice is
  still on the bus.  Pays only when a slice's work dwarfs its fixed costs (not at C3: DESIGN.md section 7).

Every upload is enqueued before the first count (`yawb_upload_catalog` is asynchronous), the slices are
views of the caller's buffers (rows of a patch are contiguous), and the per-slice results are scattered
into the full `(n_pairs, n_bins, n_sub)` arrays -- the result is identical to four whole-catalog calls.
This replaces the reference's task farm over patch pairs (`measurements.py:344-350`) on this path.
"""

from __future__ import annotations

import numpy as np

__all__ = ["COUNT_TYPES", "count_cross_pipelined", "split_patch_groups", "upload_patch_slices"]

# count type -> (first-role catalog, second-role catalog), `crosscorrelate` naming (measurements.py:623-626)
COUNT_TYPES = dict(DD=("ref", "unk"), DR=("ref", "unk_rand"), RD=("ref_rand", "unk"), RR=("ref_rand", "unk_rand"))


def split_patch_groups(patch_off: np.ndarray, n_groups: int) -> list[tuple[int, int]]:
    """Contiguous patch-id ranges `[lo, hi)` holding about the same number of rows each."""
    patch_off = np.asarray(patch_off, dtype=np.int64)
    n_patch, n_rows = len(patch_off) - 1, int(patch_off[-1])
    n_groups = max(1, min(int(n_groups), n_patch))
    if n_rows == 0 or n_groups == 1:
        return [(0, n_patch)]
    cuts = np.searchsorted(patch_off, n_rows * np.arange(1, n_groups) / n_groups, side="left")
    bounds = np.unique(np.concatenate([[0], np.clip(cuts, 0, n_patch), [n_patch]]))
    # drop empty ranges by merging them into their neighbour; the ranges always cover every patch
    keep = [int(bounds[0])]
    for a, b in zip(bounds[:-1], bounds[1:]):
        if patch_off[b] > patch_off[keep[-1]]:
            keep.append(int(b))
    keep[-1] = n_patch
    return list(zip(keep[:-1], keep[1:]))


def upload_patch_slices(engine, arrays: dict, n_groups: int) -> list[tuple[object, int, int]]:
    """Upload one catalog as independent device catalogs of whole patches.  `arrays` holds the keyword
    arguments of `Engine.upload_catalog` (xyz, patch_off, weights, zbin, n_bins); returns
    `(device catalog, first patch, one past the last patch)` per slice.  Patches outside a slice are
    empty in its device catalog, so patch ids keep their meaning."""
    off = np.asarray(arrays["patch_off"], dtype=np.int64)
    out = []
    for lo, hi in split_patch_groups(off, n_groups):
        a, b = int(off[lo]), int(off[hi])
        sub_off = np.clip(off, a, b) - a
        dev = engine.upload_catalog(
            arrays["xyz"][a:b], sub_off,
            weights=None if arrays.get("weights") is None else arrays["weights"][a:b],
            zbin=None if arrays.get("zbin") is None else arrays["zbin"][a:b],
            n_bins=arrays.get("n_bins", 1),
        )
        out.append((dev, lo, hi))
    return out


def count_cross_pipelined(engine, host: dict, pair_i: np.ndarray, pair_j: np.ndarray, r2: np.ndarray, *,
                          groups: int = 4, fuse: bool = True):
    """DD / DR / RD / RR (as far as the catalogs in `host` allow) between the patches listed in
    `(pair_i, pair_j)`.  `host[name]` = upload arguments of catalog `name` in ("ref", "ref_rand", "unk",
    "unk_rand") or None.  Returns `(counts_i64, sums_f64, stats, devices)`: per count type the
    `(n_pairs, n_bins, n_sub)` arrays, the merged kernel statistics, and the device catalogs (whole
    first-role catalogs and the slices of the second-role ones) for the caller to query and free.

    With `fuse` (default) and two first-role catalogs of the same kind (both weighted or both unweighted) the
    reference sample and its randoms are counted against each second-role catalog in ONE pass
    (`Engine.count2`: DD + RD, then DR + RR); the statistics of such a pass are reported under the joined tag,
    e.g. `stats["DD+RD"]`."""
    pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
    pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
    def rows(k):
        return int(host[k]["patch_off"][-1])

    # smaller catalogs first (ties: data before randoms): the cheap pair (reference x unknown) gets the GPU
    # going a few ms into the transfer, and every count that does not need the last catalog is done by the
    # time it arrives; what is left then is its index and the counts against it
    first = sorted((k for k in ("ref", "ref_rand") if host.get(k) is not None), key=rows)
    second = sorted((k for k in ("unk", "unk_rand") if host.get(k) is not None), key=rows)
    fused = fuse and len(first) == 2 and (host[first[0]].get("weights") is None) == (host[first[1]].get("weights") is None)
    if fused:  # both first-role catalogs are needed by the first count: F0, F1, S0, S1
        order = first + second
    else:  # upload order: F0, S0, F1, S1
        order = []
        for f, s2 in zip(first, second):
            order += [f, s2]
        order += first[len(second):] + second[len(first):]

    # enqueue every copy up front
    devices: dict[str, list] = {}
    arrival: dict[tuple[str, int], int] = {}
    for k in order:
        h = host[k]
        if k in first:
            devices[k] = [(engine.upload_catalog(h["xyz"], h["patch_off"], weights=h.get("weights"), zbin=h.get("zbin"),
                                                 n_bins=h.get("n_bins", 1)), 0, len(h["patch_off"]) - 1)]
        else:
            devices[k] = upload_patch_slices(engine, h, groups)
        for n in range(len(devices[k])):
            arrival[(k, n)] = len(arrival)

    n_bins = max(host[k].get("n_bins", 1) for k in first) if first else 1
    shape = (len(pair_i), n_bins, r2.shape[1] - 1)
    counts, sums, stats = {}, {}, {}
    def tag_of(k1, k2):
        return next(t for t, (a, b) in COUNT_TYPES.items() if (a, b) == (k1, k2))

    def ensure(tag):
        if tag not in counts:
            counts[tag] = np.zeros(shape, dtype=np.int64)
            sums[tag] = np.zeros(shape, dtype=np.float64)

    def add_stats(tag, st):
        acc = stats.setdefault(tag, {})
        for key, val in st.items():
            acc[key] = acc.get(key, 0) + val

    if fused:
        # one pass per second-role slice: (F0, F1) x S, in arrival order of the slices
        for k2 in second:
            tags = [tag_of(k1, k2) for k1 in first]
            for n, (dev2, lo, hi) in enumerate(devices[k2]):
                for tag in tags:
                    ensure(tag)
                sel = np.flatnonzero((pair_j >= lo) & (pair_j < hi))
                if len(sel) == 0:
                    continue
                (ia, fa), (ib, fb), st = engine.count2(devices[first[0]][0][0], devices[first[1]][0][0], dev2,
                                                       pair_i[sel], pair_j[sel], r2)
                counts[tags[0]][sel], sums[tags[0]][sel] = ia, fa
                counts[tags[1]][sel], sums[tags[1]][sel] = ib, fb
                add_stats("+".join(tags), st)
        return counts, sums, stats, devices
    # counts in the order in which their inputs are complete
    jobs = sorted(((max(arrival[(k1, 0)], arrival[(k2, n)]), arrival[(k2, n)], k1, k2, n)
                   for k1 in first for k2 in second for n in range(len(devices[k2]))))
    for _, _, k1, k2, n in jobs:
        tag = tag_of(k1, k2)
        ensure(tag)
        stats.setdefault(tag, {})
        dev2, lo, hi = devices[k2][n]
        sel = np.flatnonzero((pair_j >= lo) & (pair_j < hi))
        if len(sel) == 0:
            continue
        ci, cf, st = engine.count(devices[k1][0][0], dev2, pair_i[sel], pair_j[sel], r2)
        counts[tag][sel], sums[tag][sel] = ci, cf
        add_stats(tag, st)
    return counts, sums, stats, devices
