"""
Thin object wrappers over the C ABI: `Engine` (one per process / GPU) and
`DeviceCatalog` (one patch-partitioned catalog resident in HBM).

These are the B200 counterparts of the reference's worker pool
(`src/yaw/utils/parallel.py:318-343`) and of its per-patch `BinnedTrees` cache
(`src/yaw/catalog/trees.py:432-601`).  All arithmetic happens in libyawb.so.
"""

from __future__ import annotations

import ctypes
import os
import weakref
from ctypes import byref, c_double, c_int64, c_void_p

import numpy as np

from . import _lib


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(c_void_p)


class DeviceCatalog:
    """Handle of a catalog uploaded with `Engine.upload_catalog`."""

    def __init__(self, engine: "Engine", handle: c_void_p, n_patch: int, n_bins: int, binned: bool, weighted: bool):
        self.engine = engine
        self._h = handle
        self.n_patch = n_patch
        self.n_bins = n_bins
        self.binned = binned
        self.weighted = weighted

    def sum_weights(self) -> np.ndarray:
        """`(n_bins, n_patch)` sum of weights (row counts if unweighted)."""
        out = np.empty((self.n_bins, self.n_patch), dtype=np.float64)
        _lib.check(self.engine.lib.yawb_sum_weights(self._h, _ptr(out)))
        return out

    def patch_metadata(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """`(center_xyz (n_patch, 3), radius (n_patch,) in radian, num_records (n_patch,))` computed on the device
        from the uploaded rows: the quantities of the reference's `Metadata.compute`
        (`src/yaw/catalog/patch.py:104-147`; centres agree with numpy's to ~1e-15, the sums run in another order)."""
        center = np.empty((self.n_patch, 3), dtype=np.float64)
        chord = np.empty(self.n_patch, dtype=np.float64)
        num = np.empty(self.n_patch, dtype=np.int64)
        _lib.check(self.engine.lib.yawb_patch_metadata(self._h, _ptr(center), _ptr(chord), _ptr(num)))
        return center, 2.0 * np.arcsin(np.minimum(chord / 2.0, 1.0)), num

    def info(self) -> tuple[int, int]:
        n, nbytes = c_int64(), c_int64()
        _lib.check(self.engine.lib.yawb_catalog_info(self._h, byref(n), byref(nbytes)))
        return n.value, nbytes.value

    def build_index(self, role: int) -> float:
        ms = c_double()
        _lib.check(self.engine.lib.yawb_build_index(self._h, role, byref(ms)))
        return ms.value

    def drop_index(self) -> None:
        _lib.check(self.engine.lib.yawb_drop_index(self._h))

    def free(self) -> None:
        # a closed engine has already released every device block of its catalogs (and its streams are gone)
        if self._h is not None and self.engine._h is not None:
            self.engine.lib.yawb_free_catalog(self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One context on one CUDA device.  Raises if the device or library is missing."""

    def __init__(self, device: int = 0, *, staging: bool | None = None, cache_catalogs: bool | None = None,
                 cache_bytes: int | None = None):
        """`cache_catalogs` (default on; YAWB_CACHE=0 turns it off): device catalogs uploaded by the measurement
        functions stay resident, with their indexes, keyed by (catalog, z-binning) -- the counterpart of the
        reference's on-disk tree cache (`BinnedTrees.build` skips the rebuild while the `binning` file matches,
        `src/yaw/catalog/trees.py:515-526`).  A loop of `crosscorrelate` calls over tomographic bins then uploads and
        indexes only the catalogs that changed.  Least recently used entries are dropped beyond `cache_bytes` of
        device memory (default 64 GB, YAWB_CACHE_GB).

        `staging=True` (or YAWB_STAGING=1): measurement calls prepare large catalogs straight into a
        cache of page-locked buffers, so their copies are asynchronous and run at full PCIe rate.  Worth it
        for loops of many calls (C3: 225 -> 118 ms per call), not for a single one: page-locking the cache
        costs ~1.5 s per GB once."""
        self.staging = bool(int(os.environ.get("YAWB_STAGING", "0"))) if staging is None else bool(staging)
        self.staging_min_rows = 1_000_000  # smaller catalogs are not worth a page-locked detour
        # YAWB_DEVICE_DIGITIZE=1: measurement calls ship raw redshifts and the device assigns the z-bins
        # (`yawb_upload_catalog_z`); default: byte-sized ids assigned on the host (1/8 of the PCIe traffic)
        self.device_digitize = bool(int(os.environ.get("YAWB_DEVICE_DIGITIZE", "0")))
        self.lib = _lib.load()
        h = c_void_p()
        _lib.check(self.lib.yawb_create(int(device), byref(h)))
        self._h = h
        self.device = int(device)
        self._catalogs = weakref.WeakSet()  # live device catalogs: freed before the context goes away
        self.cache_catalogs = (os.environ.get("YAWB_CACHE", "1") != "0") if cache_catalogs is None else bool(cache_catalogs)
        self.cache_bytes = int(float(os.environ.get("YAWB_CACHE_GB", "64")) * 2**30) if cache_bytes is None else int(cache_bytes)
        self._cat_cache: dict = {}  # (identity, kappa) -> [signature of the binning, DeviceCatalog, bytes, weakref | None]
        self._cat_cache_tick = 0
        self._pinned: list[c_void_p] = []
        self._stage_free: list[tuple[c_void_p, int]] = []
        self._stage_used: list[tuple[c_void_p, int]] = []

    @property
    def num_sms(self) -> int:
        return self.lib.yawb_device_sms(self._h)

    def upload_catalog(
        self,
        xyz: np.ndarray,
        patch_off: np.ndarray,
        *,
        weights: np.ndarray | None = None,
        zbin: np.ndarray | None = None,
        n_bins: int = 1,
        redshifts: np.ndarray | None = None,
        edges: np.ndarray | None = None,
        closed: str = "right",
    ) -> DeviceCatalog:
        """`zbin` (+ `n_bins`): z-bin ids assigned on the host; or `redshifts` + `edges` (+ `closed`): the raw
        redshifts travel and the device assigns the bins with np.digitize's comparisons (`yawb_upload_catalog_z`,
        reference `src/yaw/catalog/trees.py:408-414`)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        patch_off = np.ascontiguousarray(patch_off, dtype=np.int64)
        n = len(xyz)
        if patch_off[-1] != n:
            raise ValueError("patch_off[-1] must equal the number of rows")
        if weights is not None:
            weights = np.ascontiguousarray(weights, dtype=np.float64)
            if len(weights) != n:
                raise ValueError("shape of 'xyz' and 'weights' does not match")
        upload = self.lib.yawb_upload_catalog
        if zbin is not None:
            if np.asarray(zbin).dtype == np.uint8:  # byte-sized ids (255 = dropped row): a quarter of the traffic
                zbin = np.ascontiguousarray(zbin)
                upload = self.lib.yawb_upload_catalog_u8
            else:
                zbin = np.ascontiguousarray(zbin, dtype=np.int32)
            if len(zbin) != n:
                raise ValueError("shape of 'xyz' and 'zbin' does not match")
        h = c_void_p()
        if redshifts is not None:
            if zbin is not None or edges is None:
                raise ValueError("pass either 'zbin' or 'redshifts' with 'edges'")
            redshifts = np.ascontiguousarray(redshifts, dtype=np.float64)
            edges = np.ascontiguousarray(edges, dtype=np.float64)
            if len(redshifts) != n:
                raise ValueError("shape of 'xyz' and 'redshifts' does not match")
            n_bins = len(edges) - 1
            _lib.check(
                self.lib.yawb_upload_catalog_z(
                    self._h, _ptr(xyz), _ptr(weights), _ptr(redshifts), _ptr(edges), int(str(closed) == "right"),
                    _ptr(patch_off), len(patch_off) - 1, int(n_bins), byref(h),
                )
            )
        else:
            _lib.check(
                upload(
                    self._h, _ptr(xyz), _ptr(weights), _ptr(zbin), _ptr(patch_off),
                    len(patch_off) - 1, int(n_bins), byref(h),
                )
            )
        binned = zbin is not None or redshifts is not None
        cat = DeviceCatalog(self, h, len(patch_off) - 1, int(n_bins) if binned else 1, binned, weights is not None)
        # the upload is asynchronous: keep the host buffers alive as long as the handle
        cat._keepalive = (xyz, weights, zbin, redshifts, patch_off)
        self._catalogs.add(cat)
        return cat

    def sample_patch_sum(self, values: np.ndarray, pair_i: np.ndarray, pair_j: np.ndarray,
                         n_patch: int) -> tuple[np.ndarray, np.ndarray]:
        """Sum over the linked patch pairs and its leave-one-patch-out jackknife samples on the device
        (`yawb_jackknife`): `values` is `(n_pairs, n_bins)`; returns `(total (n_bins,), samples (n_patch, n_bins))`,
        the `data` / `samples` of the reference's `SampledPatchSum` (`src/yaw/correlation/paircounts.py:113-141`)."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
        pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
        n_pairs, n_bins = values.shape
        total = np.empty(n_bins, dtype=np.float64)
        samples = np.empty((int(n_patch), n_bins), dtype=np.float64)
        _lib.check(self.lib.yawb_jackknife(self._h, _ptr(values), _ptr(pair_i), _ptr(pair_j), int(n_pairs), int(n_patch),
                                           int(n_bins), _ptr(total), _ptr(samples)))
        return total, samples

    def count(
        self,
        cat1: DeviceCatalog,
        cat2: DeviceCatalog,
        pair_i: np.ndarray,
        pair_j: np.ndarray,
        r2_edges: np.ndarray,
        *,
        exact: bool = False,
    ) -> tuple[np.ndarray, np.ndarray, dict]:
        """
        Returns `(counts_i64, sums_f64, stats)`, both arrays shaped
        `(n_pairs, n_bins, n_edges - 1)`.
        """
        pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
        pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
        r2_edges = np.ascontiguousarray(r2_edges, dtype=np.float64)
        n_bins = cat1.n_bins
        if r2_edges.ndim == 1:
            r2_edges = np.ascontiguousarray(np.broadcast_to(r2_edges, (n_bins, len(r2_edges))))
        if r2_edges.shape[0] != n_bins:
            raise ValueError(f"r2_edges must have shape (n_bins={n_bins}, n_edges)")
        n_edges = r2_edges.shape[1]
        n_pairs = len(pair_i)
        out_i = np.zeros((n_pairs, n_bins, n_edges - 1), dtype=np.int64)
        out_f = np.zeros((n_pairs, n_bins, n_edges - 1), dtype=np.float64)
        stats = _lib.YawbStats()
        flags = _lib.FLAG_EXACT_BRUTEFORCE if exact else 0
        _lib.check(
            self.lib.yawb_count(
                self._h, cat1._h, cat2._h, _ptr(pair_i), _ptr(pair_j), n_pairs, _ptr(r2_edges), n_edges,
                flags, _ptr(out_f), _ptr(out_i), byref(stats),
            )
        )
        return out_i, out_f, stats.as_dict()

    def count2(
        self,
        cat1a: DeviceCatalog,
        cat1b: DeviceCatalog,
        cat2: DeviceCatalog,
        pair_i: np.ndarray,
        pair_j: np.ndarray,
        r2_edges: np.ndarray,
        *,
        out_device: tuple[int, int, int, int] | None = None,
    ) -> tuple[tuple[np.ndarray, np.ndarray], tuple[np.ndarray, np.ndarray], dict]:
        """`count(cat1a, cat2)` and `count(cat1b, cat2)` in one pass over a fused index of the two first
        catalogs (`yawb_count2`): returns `((counts_a, sums_a), (counts_b, sums_b), stats)`.

        `out_device = (sums_a, counts_a, sums_b, counts_b)`: raw device pointers (0 = not wanted) of
        caller-owned buffers shaped `(n_pairs, n_bins, n_edges - 1)`; the results stay on the device
        (`YAWB_FLAG_OUT_DEVICE`, e.g. for an NCCL reduce) and the arrays returned are `None`."""
        pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
        pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
        r2_edges = np.ascontiguousarray(r2_edges, dtype=np.float64)
        n_bins = cat1a.n_bins
        if r2_edges.ndim == 1:
            r2_edges = np.ascontiguousarray(np.broadcast_to(r2_edges, (n_bins, len(r2_edges))))
        if r2_edges.shape[0] != n_bins:
            raise ValueError(f"r2_edges must have shape (n_bins={n_bins}, n_edges)")
        n_edges = r2_edges.shape[1]
        shape = (len(pair_i), n_bins, n_edges - 1)
        stats = _lib.YawbStats()
        if out_device is not None:
            ptrs = [c_void_p(p) if p else None for p in out_device]
            _lib.check(
                self.lib.yawb_count2(
                    self._h, cat1a._h, cat1b._h, cat2._h, _ptr(pair_i), _ptr(pair_j), len(pair_i), _ptr(r2_edges), n_edges,
                    _lib.FLAG_OUT_DEVICE, *ptrs, byref(stats),
                )
            )
            return (None, None), (None, None), stats.as_dict()
        out = [(np.zeros(shape, dtype=np.int64), np.zeros(shape, dtype=np.float64)) for _ in range(2)]
        _lib.check(
            self.lib.yawb_count2(
                self._h, cat1a._h, cat1b._h, cat2._h, _ptr(pair_i), _ptr(pair_j), len(pair_i), _ptr(r2_edges), n_edges, 0,
                _ptr(out[0][1]), _ptr(out[0][0]), _ptr(out[1][1]), _ptr(out[1][0]), byref(stats),
            )
        )
        return out[0], out[1], stats.as_dict()

    def count4(
        self,
        cat1a: DeviceCatalog,
        cat1b: DeviceCatalog,
        cat2a: DeviceCatalog,
        cat2b: DeviceCatalog,
        pair_i: np.ndarray,
        pair_j: np.ndarray,
        r2_edges: np.ndarray,
        *,
        out_device: tuple | None = None,
    ):
        """The four counts of a cross-correlation in ONE launch (`yawb_count4`): `count(cat1a, cat2a)`,
        `count(cat1b, cat2a)`, `count(cat1a, cat2b)`, `count(cat1b, cat2b)` -- DD, RD, DR, RR for (reference, its
        randoms, unknown, its randoms), `src/yaw/correlation/measurements.py:623-626`.  Returns
        `([(counts, sums)] * 4, stats)` in that order.

        `out_device = (sums_0, counts_0, ..., sums_3, counts_3)`: raw device pointers (0 = not wanted); the results
        stay on the device and the arrays returned are `None`."""
        pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
        pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
        r2_edges = np.ascontiguousarray(r2_edges, dtype=np.float64)
        n_bins = cat1a.n_bins
        if r2_edges.ndim == 1:
            r2_edges = np.ascontiguousarray(np.broadcast_to(r2_edges, (n_bins, len(r2_edges))))
        if r2_edges.shape[0] != n_bins:
            raise ValueError(f"r2_edges must have shape (n_bins={n_bins}, n_edges)")
        n_edges = r2_edges.shape[1]
        shape = (len(pair_i), n_bins, n_edges - 1)
        stats = _lib.YawbStats()
        tab_f, tab_i = (c_void_p * 4)(), (c_void_p * 4)()
        out = None
        if out_device is not None:
            for t in range(4):
                tab_f[t] = out_device[2 * t] or None
                tab_i[t] = out_device[2 * t + 1] or None
            flags = _lib.FLAG_OUT_DEVICE
        else:
            out = [(np.zeros(shape, dtype=np.int64), np.zeros(shape, dtype=np.float64)) for _ in range(4)]
            for t in range(4):
                tab_i[t] = out[t][0].ctypes.data
                tab_f[t] = out[t][1].ctypes.data
            flags = 0
        _lib.check(
            self.lib.yawb_count4(
                self._h, cat1a._h, cat1b._h, cat2a._h, cat2b._h, _ptr(pair_i), _ptr(pair_j), len(pair_i), _ptr(r2_edges),
                n_edges, flags, tab_f, tab_i, byref(stats),
            )
        )
        if out is None:
            return [(None, None)] * 4, stats.as_dict()
        return out, stats.as_dict()

    def count_into_device(self, cat1, cat2, pair_i, pair_j, r2_edges, out_f64_ptr: int, out_i64_ptr: int):
        """Variant writing into caller-owned device buffers (raw pointers)."""
        pair_i = np.ascontiguousarray(pair_i, dtype=np.int32)
        pair_j = np.ascontiguousarray(pair_j, dtype=np.int32)
        r2_edges = np.ascontiguousarray(r2_edges, dtype=np.float64)
        stats = _lib.YawbStats()
        _lib.check(
            self.lib.yawb_count(
                self._h, cat1._h, cat2._h, _ptr(pair_i), _ptr(pair_j), len(pair_i), _ptr(r2_edges),
                r2_edges.shape[1], _lib.FLAG_OUT_DEVICE,
                c_void_p(out_f64_ptr) if out_f64_ptr else None,
                c_void_p(out_i64_ptr) if out_i64_ptr else None, byref(stats),
            )
        )
        return stats.as_dict()

    def assign_patches(self, xyz: np.ndarray, centers_xyz: np.ndarray) -> np.ndarray:
        """Index of the nearest patch centre of every row (`scipy.cluster.vq.vq` arithmetic on the device;
        replaces `assign_patch_centers`, reference `src/yaw/catalog/catalog.py:229-249`)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        centers_xyz = np.ascontiguousarray(centers_xyz, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(xyz), dtype=np.int32)
        _lib.check(self.lib.yawb_assign_patches(self._h, _ptr(xyz), len(xyz), _ptr(centers_xyz), len(centers_xyz),
                                                _ptr(out)))
        return out

    # ---- resident catalogs across measurement calls -------------------------------------------------
    @staticmethod
    def _cache_identity(catalog):
        # the catalog OBJECT (an entry lives exactly as long as it does); a cache directory is only a label for
        # this package's in-memory catalogs, and on disk it can be overwritten, so paths are not identities
        return ("id", id(catalog))

    def cache_lookup(self, catalog, signature, kappa: bool):
        """Device catalog kept for `catalog` with this z-binning, or None.  An entry built for another binning
        is dropped (the reference rebuilds its trees in that case, `trees.py:519-526`)."""
        if not self.cache_catalogs:
            return None
        key = (self._cache_identity(catalog), bool(kappa))
        entry = self._cat_cache.get(key)
        if entry is None:
            return None
        if entry[0] != signature or entry[1]._h is None:
            self._cache_drop(key)
            return None
        self._cat_cache_tick += 1
        entry[4] = self._cat_cache_tick
        return entry[1]

    def cache_store(self, catalog, signature, kappa: bool, dev: DeviceCatalog) -> bool:
        if not self.cache_catalogs:
            return False
        ident = self._cache_identity(catalog)
        key = (ident, bool(kappa))
        try:
            ref = weakref.ref(catalog, lambda _r, k=key: self._cache_drop(k))
        except TypeError:
            return False
        self._cache_drop(key)
        nbytes = max(dev.info()[1], 0) * 3  # raw rows now, the indexes built later are about twice that
        self._cat_cache_tick += 1
        self._cat_cache[key] = [signature, dev, nbytes, ref, self._cat_cache_tick]
        total = sum(e[2] for e in self._cat_cache.values())
        while total > self.cache_bytes and len(self._cat_cache) > 1:  # least recently used first, never the new one
            victim = min((k for k in self._cat_cache if k != key), key=lambda k: self._cat_cache[k][4])
            total -= self._cat_cache[victim][2]
            self._cache_drop(victim)
        return True

    def _cache_drop(self, key) -> None:
        entry = self._cat_cache.pop(key, None)
        if entry is not None:
            try:
                entry[1].free()
            except Exception:
                pass

    def cache_clear(self) -> None:
        for key in list(self._cat_cache):
            self._cache_drop(key)

    def timer_start(self) -> None:
        _lib.check(self.lib.yawb_timer_start(self._h))

    def timer_stop(self) -> float:
        """milliseconds of device time since `timer_start` (CUDA events on the engine's stream)"""
        ms = c_double()
        _lib.check(self.lib.yawb_timer_stop(self._h, byref(ms)))
        return ms.value

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array backed by page-locked host memory (freed with the engine)"""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        ptr = c_void_p()
        _lib.check(self.lib.yawb_host_alloc(byref(ptr), nbytes))
        self._pinned.append(ptr)
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # ---- staging buffers for repeated measurement calls ----------------------------------------------
    def staging_empty(self, shape, dtype) -> np.ndarray:
        """Page-locked array from the engine's staging cache (new blocks are allocated on demand and kept;
        `staging_release` makes every block handed out so far available again).  Page-locking costs a few
        hundred ms per GB once; afterwards a measurement call prepares its catalogs straight into pinned
        memory and the copies run asynchronously at full PCIe rate."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        best = None
        for k, (ptr, size) in enumerate(self._stage_free):
            if nbytes <= size <= nbytes + nbytes // 2 + (1 << 20) and (best is None or size < self._stage_free[best][1]):
                best = k
        if best is not None:
            ptr, size = self._stage_free.pop(best)
        else:
            size = (nbytes + (1 << 20) - 1) & ~((1 << 20) - 1)
            ptr = c_void_p()
            _lib.check(self.lib.yawb_host_alloc(byref(ptr), size))
            self._pinned.append(ptr)
        self._stage_used.append((ptr, size))
        buf = (ctypes.c_char * size).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def staging_release(self) -> None:
        """Hand every staging block back to the cache; call when the uploads that used them are complete."""
        self._stage_free.extend(self._stage_used)
        self._stage_used.clear()

    def sync(self) -> None:
        _lib.check(self.lib.yawb_sync(self._h))

    def close(self) -> None:
        if self._h is not None:
            self.cache_clear()
            for cat in list(self._catalogs):
                cat.free()
            for ptr in self._pinned:
                self.lib.yawb_host_free(ptr)
            self._pinned.clear()
            self._stage_free.clear()
            self._stage_used.clear()
            self.lib.yawb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
