"""
CPU tests of the host side (`yet_another_wizz_b200.measurements` and friends) with the
oracle-backed test double standing in for the GPU engine: the reference has no unit tests
for `measurements.py` / `paircounts.py`, so these pin the patch-pair loop,
linkage, auto halving and sum-of-weights bookkeeping against the goldens of the unmodified
reference.
"""

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_cases
import golden_io
from fake_engine import OracleEngine

import yet_another_wizz_b200 as yb
from yet_another_wizz_b200 import angular
from yet_another_wizz_b200.measurements import PatchLinkage


def test_angles_match_reference_cosmology():
    # scale -> angle with the built-in Planck15 restatement vs. the reference's get_angle_radian
    for name in ("cross_unweighted", "cross_weighted_multiscale", "auto_rweight_polewrap"):
        g = golden_io.load(name)
        config = golden_cases.config_from_golden(g)
        from yet_another_wizz_b200.measurements import _angles_per_bin, get_max_angle

        amin, amax = _angles_per_bin(config)
        assert_allclose(amin, g["ang_min"], rtol=1e-9)
        assert_allclose(amax, g["ang_max"], rtol=1e-9)
        assert_allclose(get_max_angle(config).data, g["max_angle"], rtol=1e-9)


@pytest.mark.parametrize("name", ["cross_unweighted", "auto_rweight_polewrap"])
def test_linkage_matches_reference(name):
    g = golden_io.load(name)
    config = golden_cases.config_from_golden(g)
    keys = ("ref", "unk", "ref_rand", "unk_rand") if name.startswith("cross") else ("data", "rand")
    cats = [golden_cases.catalog_from_golden(g, k) for k in keys]
    links = PatchLinkage.from_catalogs(config, *cats)
    assert links.patch_links == golden_io.links_of(g)
    # meta data equals the reference's (centres given, radii recomputed)
    for k, cat in zip(keys, cats):
        assert_allclose(cat.get_radii().data, g[f"{k}_radii"], rtol=1e-12)
        assert_allclose(cat.get_centers().data, g[f"{k}_centers"], rtol=1e-15)
    # pair iteration visits every link once (cross) / the upper triangle (auto)
    cross = set(links.iter_patch_id_pairs(auto=False))
    assert cross == {(i, j) for i, js in links.patch_links.items() for j in js}
    auto = list(links.iter_patch_id_pairs(auto=True))
    assert len(auto) == len(set(auto)) and all(j >= i for i, j in auto)
    assert set(auto) == {(i, j) for (i, j) in cross if j >= i}


@pytest.mark.parametrize("name", ["cross_unweighted", "cross_weighted_multiscale"])
def test_crosscorrelate_host_logic(name):
    g = golden_io.load(name)
    corrs = golden_cases.run_cross(g, OracleEngine())
    assert len(corrs) == g["ang_min"].shape[1]
    golden_cases.check_corrfunc(g, "cross", corrs, ("dd", "dr", "rd", "rr"), exact=name == "cross_unweighted")


@pytest.mark.parametrize("name", ["auto_unweighted", "auto_rweight_polewrap"])
def test_autocorrelate_host_logic(name):
    g = golden_io.load(name)
    corrs = golden_cases.run_auto(g, OracleEngine())
    golden_cases.check_corrfunc(g, "auto", corrs, ("dd", "dr", "rr"), exact=name == "auto_unweighted")
    assert corrs[0].rd is None and corrs[0].auto


@pytest.mark.parametrize("name", ["scalar_weighted", "scalar_unweighted"])
def test_scalar_modes_host_logic(name):
    """crosscorrelate_scalar (with / without unknown randoms) and autocorrelate_scalar against the
    reference's goldens: kappa x weight on the "k" side, NN companion counts, mean-kappa DR."""
    g = golden_io.load(name)
    results = golden_cases.run_scalar(g, OracleEngine())
    golden_cases.check_scalar(g, results, exact_numbers=name == "scalar_unweighted")
    assert results["cross"][0].dr is not None and results["auto"][0].dr is None


def test_bundled_example_host_logic():
    """per-patch-pair counts of the reference's bundled 2dFLenS example (11 patches, weighted)"""
    g = golden_io.load("example_2dflens")
    cross, auto = golden_cases.run_example(g, OracleEngine())
    golden_cases.check_example(g, cross, auto)


def test_scalar_modes_need_kappa():
    g = golden_io.load("cross_unweighted")
    config = golden_cases.config_from_golden(g)
    ref, unk = golden_cases.catalog_from_golden(g, "ref"), golden_cases.catalog_from_golden(g, "unk")
    with pytest.raises(ValueError, match="kappa"):
        yb.crosscorrelate_scalar(config, ref, unk, engine=OracleEngine())
    with pytest.raises(ValueError, match="kappa"):
        yb.autocorrelate_scalar(config, ref, engine=OracleEngine())


def test_error_behaviour():
    g = golden_io.load("cross_unweighted")
    config = golden_cases.config_from_golden(g)
    ref = golden_cases.catalog_from_golden(g, "ref")
    unk = golden_cases.catalog_from_golden(g, "unk")
    with pytest.raises(ValueError, match="at least one random dataset must be provided"):
        yb.crosscorrelate(config, ref, unk, engine=OracleEngine())
    with pytest.raises(ValueError, match="separate cache directory"):
        yb.crosscorrelate(config, ref, ref, unk_rand=unk, engine=OracleEngine())
    # unknown catalog without redshifts cannot be z-binned (trees.py:397-398)
    with pytest.raises(ValueError, match="redshifts"):
        yb.autocorrelate(config, unk, golden_cases.catalog_from_golden(g, "unk_rand"), engine=OracleEngine())
    # misaligned patches
    c = golden_io.catalog_arrays(g, "unk")
    shifted = yb.Catalog.from_arrays(c["ra"] + 0.02, c["dec"], patch_ids=c["patch"], degrees=False)
    with pytest.raises(yb.InconsistentPatchesError):
        yb.crosscorrelate(config, ref, unk, unk_rand=shifted, engine=OracleEngine())
    fewer = yb.Catalog.from_arrays(c["ra"][c["patch"] < 3], c["dec"][c["patch"] < 3],
                                   patch_ids=c["patch"][c["patch"] < 3], degrees=False)
    with pytest.raises(yb.InconsistentPatchesError, match="patch IDs do not match"):
        yb.crosscorrelate(config, ref, unk, unk_rand=fewer, engine=OracleEngine())
    # rweight without resolution raises like the reference (np.linspace(..., None + 1), trees.py:110)
    bad = yb.Configuration.create(rmin=100, rmax=1000, rweight=-1.0, zmin=0.1, zmax=1.0, num_bins=3)
    with pytest.raises(TypeError):
        yb.crosscorrelate(bad, ref, unk, unk_rand=golden_cases.catalog_from_golden(g, "unk_rand"),
                          engine=OracleEngine())


def test_angular_helpers():
    # restated from reference tests/catalog/test_trees.py:14-131
    assert_array_equal(angular.parse_ang_limits([0.0, 1.0], [1.0, np.pi]), [[0.0, 1.0], [1.0, np.pi]])
    for bad in (([1.0], [0.5]), ([-1.0], [1.0]), ([1.0], [np.pi + 1e-9]), ([[1.0]], [2.0]), ([1.0, 2.0], [3.0])):
        with pytest.raises(ValueError):
            angular.parse_ang_limits(*bad)
    assert_allclose(angular.get_ang_bins(np.array([[0.01, 1.0]]), -1.0, 2), [0.01, 0.1, 1.0])
    assert_allclose(angular.logarithmic_mid(np.array([0.01, 1.0, 100.0])), [0.1, 10.0])
    got = angular.get_counts_for_limits(np.array([1.0, 2.0, 3.0, 4.0]), np.array([1.0, 2.0, 3.0, 4.0, 5.0]),
                                        np.array([[1.0, 5.0], [2.0, 4.0]]))
    assert_array_equal(got, [10.0, 5.0])
    # pow(r, 2.0), not r * r
    import math
    e = np.array([1e-3, 2.7e-3, 0.4])
    assert_array_equal(angular.squared_chord_edges(e), [math.pow(2.0 * math.sin(x / 2.0), 2.0) for x in e])


def test_boxrandoms_and_catalog_assignment():
    gen = yb.BoxRandoms(0, 40, -12.5, 12.5, redshifts=np.linspace(0.1, 1.0, 1000), seed=3)
    centers = yb.AngularCoordinates(np.deg2rad([[10.0, -6.0], [30.0, -6.0], [10.0, 6.0], [30.0, 6.0]]))
    cat = yb.Catalog.from_random("mem", gen, 20000, patch_centers=centers)
    assert cat.num_patches == 4 and cat.has_redshifts and not cat.has_weights
    assert sum(cat.get_num_records()) == 20000
    for pid in cat:
        rows = cat[pid].load_data()
        xyz = yb.AngularCoordinates(np.column_stack([rows["ra"], rows["dec"]])).to_3d()
        d = ((xyz[:, None, :] - centers.to_3d()[None]) ** 2).sum(axis=2)
        assert np.all(np.argmin(d, axis=1) == pid)  # nearest-centre assignment
        assert np.rad2deg(rows["ra"]).min() >= 0 and np.rad2deg(rows["ra"]).max() <= 40
    # same seed -> same points
    a = yb.BoxRandoms(0, 40, -12.5, 12.5, seed=5)(100)
    b = yb.BoxRandoms(0, 40, -12.5, 12.5, seed=5)(100)
    assert_array_equal(a["ra"], b["ra"])


def test_pipelined_schedule_matches_whole_catalog_counts():
    """`pipeline.count_cross_pipelined`: slicing the unbinned catalogs by patch groups changes the schedule,
    never the result; the slices cover every patch, also with empty patches"""
    from yet_another_wizz_b200 import pipeline
    from yet_another_wizz_b200.measurements import _as_binning, prepare_catalog_arrays

    g = golden_io.load("cross_unweighted")
    config = golden_cases.config_from_golden(g)
    cats = {k: golden_cases.catalog_from_golden(g, k) for k in ("ref", "unk", "ref_rand", "unk_rand")}
    binning = _as_binning(config)
    host = {k: prepare_catalog_arrays(c, binning if k in ("ref", "ref_rand") else None) for k, c in cats.items()}
    links = PatchLinkage.from_catalogs(config, *cats.values(), engine=OracleEngine())
    pair_i, pair_j = links.get_patch_id_pairs(auto=False)
    r2 = links._get_plan().r2

    whole = pipeline.count_cross_pipelined(OracleEngine(), host, pair_i, pair_j, r2, groups=1)
    sliced = pipeline.count_cross_pipelined(OracleEngine(), host, pair_i, pair_j, r2, groups=3)
    assert set(whole[0]) == {"DD", "DR", "RD", "RR"}
    assert len(sliced[3]["unk"]) == 3 and len(whole[3]["unk"]) == 1
    for tag in whole[0]:
        assert_array_equal(sliced[0][tag], whole[0][tag])
        assert whole[0][tag].sum() > 0
    # the DD counts are the golden ones
    plan = links._get_plan()
    dd = plan.finish(whole[0]["DD"])[0]
    want = g["cross_dd_counts_s0"][:, pair_i, pair_j].T
    assert_array_equal(dd, want)

    for off, n, expect in (([0, 0, 0, 5, 5], 3, [(0, 4)]), ([0, 4, 4, 4, 9, 9, 12], 3, [(0, 1), (1, 4), (4, 6)]),
                           ([0, 0, 0], 4, [(0, 2)])):
        ranges = pipeline.split_patch_groups(np.array(off), n)
        assert ranges == expect
        assert ranges[0][0] == 0 and ranges[-1][1] == len(off) - 1
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))


def test_angle_and_pair_list_caches_return_fresh_equal_values():
    """the per-configuration caches of the host path (scale -> angle conversion per z-bin, pair list of a linkage)
    return what an uncached evaluation returns, do not call the cosmology again, and hand out copies"""
    from yet_another_wizz_b200 import measurements

    class CountingCosmology:
        def __init__(self, inner):
            self.inner, self.calls = inner, 0

        def comoving_distance(self, z):
            self.calls += 1
            return self.inner.comoving_distance(z)

        def angular_diameter_distance(self, z):
            self.calls += 1
            return self.inner.angular_diameter_distance(z)

    cosmo = CountingCosmology(yb.cosmology.get_default_cosmology())
    config = yb.Configuration.create(rmin=100, rmax=1000, zmin=0.1, zmax=1.0, num_bins=7, cosmology=cosmo)
    a0, b0 = measurements._angles_per_bin(config)
    calls = cosmo.calls
    assert calls > 0
    a1, b1 = measurements._angles_per_bin(config)
    assert cosmo.calls == calls  # answered from the cache
    assert_array_equal(a0, a1)
    assert_array_equal(b0, b1)
    a1[:] = -1.0  # a copy: the cache is not poisoned
    a2, _ = measurements._angles_per_bin(config)
    assert_array_equal(a0, a2)
    # another binning of the same cosmology is a different entry
    config2 = yb.Configuration.create(rmin=100, rmax=1000, zmin=0.1, zmax=1.0, num_bins=5, cosmology=cosmo)
    a3, _ = measurements._angles_per_bin(config2)
    assert len(a3) == 5 and cosmo.calls > calls

    links = PatchLinkage(config, {0: {0, 1}, 1: {0, 1, 2}, 2: {1, 2}})
    for auto in (False, True):
        want = list(links.iter_patch_id_pairs(auto=auto))
        for _ in range(2):  # the second answer comes from the cache
            pi, pj = links.get_patch_id_pairs(auto=auto)
            assert list(zip(pi.tolist(), pj.tolist())) == want
            pi[:] = 99  # copies
