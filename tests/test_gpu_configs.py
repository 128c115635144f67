"""
GPU parity at the BASELINE.json configurations (scaled where the FP64 all-pairs cross-check would take
minutes): the production kernel (sky-cell pruning, FP32 test, FP64 recheck) against the unpruned FP64
kernel through the public API.  Integer counts bit-exact, weighted sums rtol 1e-12.

Size-independent properties checked alongside: DD == DR when the random catalog is the data catalog,
auto counts are symmetric-consistent (pairs counted once), sums of weights equal plain numpy sums.
"""

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu

BOX = (0.0, 40.0, -12.5, 12.5)


class ExactEngine:
    """routes every count through the FP64 all-pairs kernel (validation flag of the C ABI)"""

    def __init__(self, engine):
        self._e = engine

    def upload_catalog(self, *a, **k):
        return self._e.upload_catalog(*a, **k)

    def count(self, c1, c2, pi, pj, r2, **k):
        return self._e.count(c1, c2, pi, pj, r2, exact=True)


@pytest.fixture(scope="module")
def engine():
    from yet_another_wizz_b200 import Engine

    eng = Engine(0)
    yield eng
    eng.close()


def grid_centers(nx, ny, box=BOX):
    import yet_another_wizz_b200 as yb

    ras = box[0] + (np.arange(nx) + 0.5) * (box[1] - box[0]) / nx
    decs = box[2] + (np.arange(ny) + 0.5) * (box[3] - box[2]) / ny
    return yb.AngularCoordinates(np.deg2rad([[r, d] for d in decs for r in ras]))


def make(name, n, centers, *, zpool=None, wpool=None, seed=1, box=BOX):
    import yet_another_wizz_b200 as yb

    gen = yb.BoxRandoms(*box, redshifts=zpool, weights=wpool, seed=seed)
    return yb.Catalog.from_random(name, gen, n, patch_centers=centers)


def compare(fast, exact, kinds, exact_ints):
    for cf, ce in zip(fast, exact):
        for kind in kinds:
            a, b = getattr(cf, kind), getattr(ce, kind)
            if exact_ints:
                assert_array_equal(a.counts.counts, b.counts.counts)
            else:
                assert_allclose(a.counts.counts, b.counts.counts, rtol=1e-12, atol=0)
            assert b.counts.counts.sum() > 0
            assert_allclose(a.sum_weights.sum_weights1, b.sum_weights.sum_weights1, rtol=1e-13)
            assert_allclose(a.sum_weights.sum_weights2, b.sum_weights.sum_weights2, rtol=1e-13)


def test_c1_full_size_crosscorrelate(engine):
    """configs[0]: 1e5 ref x 1e5 unknown + 1e6 randoms per side, 16 patches, 100-1000 kpc, 10 z-bins"""
    import yet_another_wizz_b200 as yb

    centers = grid_centers(4, 4)
    pool = np.random.default_rng(7).uniform(0.1, 1.0, 1_000_000)
    ref = make("ref", 100_000, centers, zpool=pool, seed=1)
    unk = make("unk", 100_000, centers, seed=2)
    ref_rand = make("ref_rand", 1_000_000, centers, zpool=pool, seed=3)
    unk_rand = make("unk_rand", 1_000_000, centers, seed=4)
    config = yb.Configuration.create(rmin=100, rmax=1000, zmin=0.1, zmax=1.0, num_bins=10)
    fast = yb.crosscorrelate(config, ref, unk, ref_rand=ref_rand, unk_rand=unk_rand, engine=engine)
    exact = yb.crosscorrelate(config, ref, unk, ref_rand=ref_rand, unk_rand=unk_rand, engine=ExactEngine(engine))
    compare(fast, exact, ("dd", "dr", "rd", "rr"), exact_ints=True)
    # totals of SURVEY.md section 6 (reference probe of the same seeds): DD 9.65e4, RR 9.64e6
    assert abs(fast[0].dd.counts.counts.sum() / 9.65e4 - 1) < 0.01
    assert abs(fast[0].rr.counts.counts.sum() / 9.64e6 - 1) < 0.01
    # sums of weights = row counts inside the binning
    z = np.concatenate([ref[p].load_data()["redshifts"] for p in ref])
    assert fast[0].dd.sum_weights.sum_weights1.sum() == ((z > 0.1) & (z <= 1.0)).sum()
    assert fast[0].dd.sum_weights.sum_weights2.sum() == 10 * 100_000


def test_c2_like_autocorrelate_rweight(engine):
    """configs[1] scaled: autocorrelate with rweight=-1 and resolution=50 log sub-bins, 32 patches"""
    import yet_another_wizz_b200 as yb

    box = (0.0, 16.0, -4.0, 4.0)
    centers = grid_centers(8, 4, box)
    pool = np.random.default_rng(7).uniform(0.1, 1.0, 100_000)
    wpool = np.random.default_rng(8).uniform(0.5, 1.5, 100_000)
    data = make("data", 150_000, centers, zpool=pool, wpool=wpool, seed=1, box=box)
    rand = make("rand", 150_000, centers, zpool=pool, seed=3, box=box)
    config = yb.Configuration.create(rmin=100, rmax=1000, rweight=-1.0, resolution=50, zmin=0.1, zmax=1.0, num_bins=10)
    fast = yb.autocorrelate(config, data, rand, engine=engine)
    exact = yb.autocorrelate(config, data, rand, engine=ExactEngine(engine))
    compare(fast, exact, ("dd", "dr", "rr"), exact_ints=False)
    # RR is unweighted: r-weighted sums of integers must agree far below the 1e-12 bar
    assert_allclose(fast[0].rr.counts.counts, exact[0].rr.counts.counts, rtol=1e-14, atol=0)
    # auto counts: only the upper triangle is filled
    assert np.all(np.tril(fast[0].dd.counts.counts.sum(axis=0), -1) == 0)


def test_c4_like_multiscale_crosscorrelate(engine):
    """configs[3] scaled: three (rmin, rmax) scale pairs counted in one pass, 30 z-bins"""
    import yet_another_wizz_b200 as yb

    box = (0.0, 10.0, -5.0, 5.0)
    centers = grid_centers(4, 4, box)
    pool = np.random.default_rng(7).uniform(0.07, 1.42, 100_000)
    ref = make("ref", 60_000, centers, zpool=pool, seed=1, box=box)
    unk = make("unk", 400_000, centers, seed=2, box=box)
    config = yb.Configuration.create(rmin=[100, 300, 500], rmax=[1000, 1500, 2000], zmin=0.07, zmax=1.42, num_bins=30)
    fast = yb.crosscorrelate(config, ref, unk, unk_rand=unk.__class__(dict(unk.items()), "unk_copy"), engine=engine)
    exact = yb.crosscorrelate(config, ref, unk, unk_rand=unk.__class__(dict(unk.items()), "unk_copy"),
                              engine=ExactEngine(engine))
    assert len(fast) == 3
    compare(fast, exact, ("dd", "dr"), exact_ints=True)
    for cf in fast:  # the "random" catalog is the data catalog: DR must equal DD exactly
        assert_array_equal(cf.dd.counts.counts, cf.dr.counts.counts)
    # nested scales: 100-1000 kpc counts < 300-1500 < 500-2000 for a uniform field
    sums = [cf.dd.counts.counts.sum() for cf in fast]
    assert sums[0] < sums[1] < sums[2]


def test_many_bins_and_empty_patches(engine):
    """50 z-bins (configs[4] binning), patches with no rows and z-bins with no rows"""
    import yet_another_wizz_b200 as yb

    box = (0.0, 6.0, -3.0, 3.0)
    centers = grid_centers(3, 3, box)
    rng = np.random.default_rng(5)
    n1, n2 = 40_000, 120_000
    ra1 = rng.uniform(0.0, 4.0, n1)  # right third of the box stays empty in catalog 1
    dec1 = np.rad2deg(np.arcsin(rng.uniform(np.sin(np.deg2rad(-3)), np.sin(np.deg2rad(3)), n1)))
    z1 = rng.uniform(0.1, 0.6, n1)   # upper z-bins stay empty
    ra2 = rng.uniform(0.0, 6.0, n2)
    dec2 = np.rad2deg(np.arcsin(rng.uniform(np.sin(np.deg2rad(-3)), np.sin(np.deg2rad(3)), n2)))
    # keep the patch count consistent: a handful of far-right points so every patch id exists
    ra1[:9] = np.rad2deg(centers.ra)
    dec1[:9] = np.rad2deg(centers.dec)
    ref = yb.Catalog.from_arrays(ra1, dec1, patch_centers=centers, redshifts=z1)
    unk = yb.Catalog.from_arrays(ra2, dec2, patch_centers=centers)
    config = yb.Configuration.create(rmin=100, rmax=1000, zmin=0.07, zmax=1.42, num_bins=50)
    kw = dict(unk_rand=yb.Catalog.from_arrays(ra2, dec2, patch_centers=centers))
    try:
        fast = yb.crosscorrelate(config, ref, unk, engine=engine, **kw)
    except yb.InconsistentPatchesError:
        pytest.skip("sparse patch centres drifted too far for the reference's consistency check")
    exact = yb.crosscorrelate(config, ref, unk, engine=ExactEngine(engine), **kw)
    compare(fast, exact, ("dd", "dr"), exact_ints=True)
    assert fast[0].dd.counts.counts[30:].sum() == 0  # bins above z = 0.6 hold nothing


def test_many_bins_times_many_subbins(engine):
    """50 z-bins x 53 r-weight sub-bins (configs[4] binning with rweight): the per-warp accumulators no longer
    fit four warps per CTA, the launcher falls back to smaller CTAs; weighted first catalog"""
    import yet_another_wizz_b200 as yb

    box = (0.0, 4.0, -2.0, 2.0)
    centers = grid_centers(2, 2, box)
    pool = np.random.default_rng(7).uniform(0.07, 1.42, 50_000)
    wpool = np.random.default_rng(9).uniform(0.5, 1.5, 50_000)
    ref = make("ref", 30_000, centers, zpool=pool, wpool=wpool, seed=1, box=box)
    unk = make("unk", 120_000, centers, seed=2, box=box)
    unk2 = make("unk2", 60_000, centers, seed=4, box=box)
    config = yb.Configuration.create(rmin=[100, 200], rmax=[1000, 3000], rweight=-1.0, resolution=50,
                                     zmin=0.07, zmax=1.42, num_bins=50)
    fast = yb.crosscorrelate(config, ref, unk, unk_rand=unk2, engine=engine)
    exact = yb.crosscorrelate(config, ref, unk, unk_rand=unk2, engine=ExactEngine(engine))
    compare(fast, exact, ("dd", "dr"), exact_ints=False)


def test_irregular_patch_footprints(engine):
    """ring- and L-shaped patches: the patch box is mostly empty, so consecutive rows along the Hilbert curve
    can be far apart; oversized tiles are cut (index straggler guard) and counts stay exact"""
    import oracle

    rng = np.random.default_rng(11)
    n = 120_000
    # patch 0: annulus around (ra, dec) = (0.05, 0); patch 1: an L-shaped region next to it
    r = np.sqrt(rng.uniform(0.012**2, 0.02**2, n // 2))
    phi = rng.uniform(0, 2 * np.pi, n // 2)
    ra0, dec0 = 0.05 + r * np.cos(phi), r * np.sin(phi)
    u, v = rng.uniform(0, 0.04, n), rng.uniform(0, 0.04, n)
    keep = (u < 0.008) | (v < 0.008)
    ra1, dec1 = 0.08 + u[keep][: n // 2], -0.02 + v[keep][: n // 2]
    ra = np.concatenate([ra0, ra1]); dec = np.concatenate([dec0, dec1])
    patch_off = np.array([0, len(ra0), len(ra)])
    xyz = oracle.radec_to_xyz(ra, dec)
    zbin = rng.integers(0, 4, len(ra)).astype(np.int32)
    # second catalog: same footprints, different points
    xyz2 = oracle.radec_to_xyz(ra + rng.normal(0, 1e-4, len(ra)), dec + rng.normal(0, 1e-4, len(ra)))
    d1 = engine.upload_catalog(xyz, patch_off, zbin=zbin, n_bins=4)
    d2 = engine.upload_catalog(xyz2, patch_off)
    r2 = np.tile(oracle.chord_sq_edges(np.array([1e-4, 8e-4])), (4, 1))
    pi, pj = np.array([0, 0, 1, 1]), np.array([0, 1, 0, 1])
    fi, _, fs = engine.count(d1, d2, pi, pj, r2)
    ei, _, es = engine.count(d1, d2, pi, pj, r2, exact=True)
    assert_array_equal(fi, ei)
    assert ei.sum() > 1e5
    assert fs["pair_tests"] < 0.02 * es["pair_tests"]  # pruning still effective on hollow boxes
    d1.free(); d2.free()


def test_duplicates_and_dense_clumps(engine):
    """collisions: many rows at exactly the same position (d2 == 0 must never be counted when the lower
    edge is 0 <, and is counted in the first sub-bin otherwise), a clump much denser than the rest of the
    patch (one sky cell holds thousands of rows), and single-row / empty z-bins"""
    import oracle

    rng = np.random.default_rng(23)
    base_ra, base_dec = rng.uniform(0.0, 0.03, 4000), rng.uniform(-0.01, 0.01, 4000)
    clump_ra, clump_dec = 0.011 + rng.normal(0, 2e-5, 6000), 0.002 + rng.normal(0, 2e-5, 6000)
    dup_ra, dup_dec = np.full(3000, 0.02), np.full(3000, -0.004)  # 3000 identical points
    ra = np.concatenate([base_ra, clump_ra, dup_ra]); dec = np.concatenate([base_dec, clump_dec, dup_dec])
    order = rng.permutation(len(ra))
    ra, dec = ra[order], dec[order]
    xyz = oracle.radec_to_xyz(ra, dec)
    patch_off = np.array([0, len(ra) // 2, len(ra)])
    zbin = rng.integers(0, 3, len(ra)).astype(np.int32)
    zbin[:5] = 5  # out of range: dropped
    zbin[zbin == 2] = 1  # z-bin 2 is empty
    zbin[7] = 2  # ... except for a single row
    d1 = engine.upload_catalog(xyz, patch_off, zbin=zbin, n_bins=3)
    d2 = engine.upload_catalog(xyz, patch_off)
    pi, pj = np.array([0, 0, 1, 1]), np.array([0, 1, 0, 1])
    for edges in (np.array([1e-5, 3e-4]), np.array([1e-6, 1e-5, 1e-4, 1e-3]), np.array([3e-5, 2e-3])):
        r2 = np.tile(oracle.chord_sq_edges(edges), (3, 1))
        fi, _, fs = engine.count(d1, d2, pi, pj, r2)
        ei, _, _ = engine.count(d1, d2, pi, pj, r2, exact=True)
        assert_array_equal(fi, ei)
        assert ei.sum() > 1e6
    # cross-check one case against the numpy oracle on a sub-sample of rows (self pairs at d2 == 0 excluded
    # by the open lower edge)
    sel = np.flatnonzero((zbin[: patch_off[1]] == 0))[:1500]
    a = xyz[: patch_off[1]][sel]
    sub_off = np.array([0, len(a)])
    s1 = engine.upload_catalog(a, sub_off, zbin=np.zeros(len(a), dtype=np.int32), n_bins=1)
    s2 = engine.upload_catalog(a, sub_off)
    r2 = oracle.chord_sq_edges(np.array([1e-6, 1e-4, 1e-3]))[None, :]
    fi, _, _ = engine.count(s1, s2, np.array([0]), np.array([0]), r2)
    want = oracle.pair_histogram(a, a, None, None, r2[0])
    assert_array_equal(fi[0, 0], want)
    for d in (d1, d2, s1, s2):
        d.free()
