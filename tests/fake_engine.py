"""
TEST DOUBLE (tests only): an object with the interface of
`yet_another_wizz_b200.Engine` whose `count` is answered by the oracle.  It lets the
CPU suite exercise the *host* logic of `measurements.py` (linkage, pair lists, z-bin
digitisation, r-weight post-processing, scatter, auto halving, multi-rank sharding)
against the golden vectors without a GPU.  The product never uses it.
"""

from __future__ import annotations

import numpy as np

import oracle


class _FakeDeviceCatalog:
    def __init__(self, xyz, patch_off, weights, zbin, n_bins):
        self.xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.patch_off = np.asarray(patch_off)
        self.w = weights
        self.zbin = zbin
        self.binned = zbin is not None
        self.n_bins = n_bins if self.binned else 1
        self.n_patch = len(patch_off) - 1
        self.weighted = weights is not None
        self.patch = np.repeat(np.arange(self.n_patch), np.diff(self.patch_off))

    def select(self, patch, b):
        m = self.patch == patch
        if self.binned:
            m &= self.zbin == b
        return self.xyz[m], (None if self.w is None else self.w[m])

    def sum_weights(self):
        out = np.zeros((self.n_bins, self.n_patch))
        for p in range(self.n_patch):
            for b in range(self.n_bins):
                xyz, w = self.select(p, b)
                out[b, p] = len(xyz) if w is None else w.sum()
        return out

    def free(self):
        pass


class OracleEngine:
    def __init__(self):
        self.calls = 0

    def upload_catalog(self, xyz, patch_off, *, weights=None, zbin=None, n_bins=1):
        return _FakeDeviceCatalog(xyz, patch_off, weights, zbin, n_bins)

    def count(self, cat1, cat2, pair_i, pair_j, r2_edges, *, exact=False):
        self.calls += 1
        n_bins = cat1.n_bins
        nsub = r2_edges.shape[1] - 1
        out_i = np.zeros((len(pair_i), n_bins, nsub), dtype=np.int64)
        out_f = np.zeros((len(pair_i), n_bins, nsub), dtype=np.float64)
        for k, (i, j) in enumerate(zip(pair_i, pair_j)):
            for b in range(n_bins):
                xyz1, w1 = cat1.select(i, b)
                xyz2, w2 = cat2.select(j, b)
                h = oracle.pair_histogram(xyz1, xyz2, w1, w2, r2_edges[b])
                if h.dtype == np.int64:
                    out_i[k, b] = h
                    out_f[k, b] = h
                else:
                    out_f[k, b] = h
        return out_i, out_f, dict(kernel_ms=0.0, pair_tests=0, pair_tests_naive=0, rechecks=0, launches=0)

    def count2(self, cat1a, cat1b, cat2, pair_i, pair_j, r2_edges):
        ia, fa, st = self.count(cat1a, cat2, pair_i, pair_j, r2_edges)
        ib, fb, _ = self.count(cat1b, cat2, pair_i, pair_j, r2_edges)
        return (ia, fa), (ib, fb), st
