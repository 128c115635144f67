"""
GPU parity tests through the public API (`crosscorrelate` / `autocorrelate` ->
C ABI -> CUDA kernels) against the golden vectors produced by the unmodified
reference (`tests/golden/make_golden.py`).

Bar: unweighted counts bit-exact (`assert_array_equal`), weighted / r-weighted
sums within 1e-12 relative.
"""

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import golden_cases
import golden_io

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from yet_another_wizz_b200 import Engine

    eng = Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", ["cross_unweighted", "cross_weighted_multiscale"])
def test_crosscorrelate_golden(engine, name):
    g = golden_io.load(name)
    corrs = golden_cases.run_cross(g, engine)
    golden_cases.check_corrfunc(g, "cross", corrs, ("dd", "dr", "rd", "rr"), exact=name == "cross_unweighted")


@pytest.mark.parametrize("name", ["auto_unweighted", "auto_rweight_polewrap"])
def test_autocorrelate_golden(engine, name):
    g = golden_io.load(name)
    corrs = golden_cases.run_auto(g, engine)
    golden_cases.check_corrfunc(g, "auto", corrs, ("dd", "dr", "rr"), exact=name == "auto_unweighted")


@pytest.mark.parametrize("name", ["scalar_weighted", "scalar_unweighted"])
def test_scalar_modes_golden(engine, name):
    g = golden_io.load(name)
    results = golden_cases.run_scalar(g, engine)
    golden_cases.check_scalar(g, results, exact_numbers=name == "scalar_unweighted")


def test_bundled_example_golden(engine):
    """the reference's own example data set (SURVEY.md section 8c ii): counts that reproduce its
    golden n(z) files, see tests/golden/make_golden.py::case_example"""
    g = golden_io.load("example_2dflens")
    cross, auto = golden_cases.run_example(g, engine)
    golden_cases.check_example(g, cross, auto)


def test_pipelined_schedule_gpu(engine):
    """asynchronous uploads in patch slices + counts in arrival order: same integers as whole catalogs"""
    from yet_another_wizz_b200 import pipeline
    from yet_another_wizz_b200.measurements import PatchLinkage, _as_binning, prepare_catalog_arrays

    g = golden_io.load("cross_unweighted")
    config = golden_cases.config_from_golden(g)
    cats = {k: golden_cases.catalog_from_golden(g, k) for k in ("ref", "unk", "ref_rand", "unk_rand")}
    binning = _as_binning(config)
    host = {k: prepare_catalog_arrays(c, binning if k in ("ref", "ref_rand") else None) for k, c in cats.items()}
    links = PatchLinkage.from_catalogs(config, *cats.values(), engine=engine)
    pair_i, pair_j = links.get_patch_id_pairs(auto=False)
    plan = links._get_plan()
    for groups in (1, 2, 6):
        counts, _, stats, devs = pipeline.count_cross_pipelined(engine, host, pair_i, pair_j, plan.r2, groups=groups)
        for tag, kind in (("DD", "dd"), ("DR", "dr"), ("RD", "rd"), ("RR", "rr")):
            got = plan.finish(counts[tag])[0]
            assert np.array_equal(got, g[f"cross_{kind}_counts_s0"][:, pair_i, pair_j].T), (groups, tag)
        assert all(st["pair_tests"] > 0 for st in stats.values())
        assert set("+".join(stats).split("+")) == {"DD", "DR", "RD", "RR"}  # fused passes report under joined tags
        for lst in devs.values():
            for dev, _, _ in lst:
                dev.free()


def test_staging_cache_engine():
    """`Engine(staging=True)`: catalogs prepared into the page-locked cache, blocks recycled between calls"""
    from yet_another_wizz_b200 import Engine

    eng = Engine(0, staging=True)
    eng.staging_min_rows = 0
    try:
        g = golden_io.load("cross_unweighted")
        for _ in range(3):
            corrs = golden_cases.run_cross(g, eng)
            golden_cases.check_corrfunc(g, "cross", corrs, ("dd", "dr", "rd", "rr"), exact=True)
        assert len(eng._stage_used) == 0 and len(eng._stage_free) > 0
        n_blocks = len(eng._stage_free)
        golden_cases.run_cross(g, eng)
        assert len(eng._stage_free) == n_blocks  # nothing new was page-locked
    finally:
        eng.close()


def test_crosscorrelate_with_device_digitize(engine):
    """z-bins assigned on the device (`Engine.device_digitize`, yawb_upload_catalog_z) give the reference's counts"""
    g = golden_io.load("cross_unweighted")
    engine.device_digitize = True
    try:
        corrs = golden_cases.run_cross(g, engine)
    finally:
        engine.device_digitize = False
    golden_cases.check_corrfunc(g, "cross", corrs, ("dd", "dr", "rd", "rr"), exact=True)


def test_default_engine_and_stats():
    import yet_another_wizz_b200 as yb
    from yet_another_wizz_b200 import measurements

    g = golden_io.load("cross_unweighted")
    corrs = golden_cases.run_cross(g, None)  # process-wide engine on cuda:LOCAL_RANK
    golden_cases.check_corrfunc(g, "cross", corrs, ("dd", "rr"), exact=True)
    stats = measurements.last_stats()
    assert set(stats) == {"DD+RD", "DR+RR"}  # the reference sample and its randoms are counted in one pass
    for s in stats.values():
        assert s["launches"] >= 1 and s["pair_tests"] > 0
        assert s["pair_tests"] < s["pair_tests_naive"]  # sky-cell pruning is active


def test_medium_synthetic_vs_oracle(engine):
    """C1-like geometry at a size the oracle's C brute force finishes in seconds:
    BoxRandoms catalogs, 16 patches, 10 z-bins, 100-1000 kpc, unweighted -> bit-exact."""
    import oracle
    import yet_another_wizz_b200 as yb

    n_ref, n_unk = 20000, 60000
    pool = np.random.default_rng(7).uniform(0.1, 1.0, 100000)
    ras = (np.arange(4) + 0.5) * 10.0 / 4
    decs = -2.5 + (np.arange(4) + 0.5) * 5.0 / 4
    centers = yb.AngularCoordinates(np.deg2rad([[r, d] for d in decs for r in ras]))
    ref = yb.Catalog.from_random("ref", yb.BoxRandoms(0, 10, -2.5, 2.5, redshifts=pool, seed=1), n_ref,
                                 patch_centers=centers)
    unk = yb.Catalog.from_random("unk", yb.BoxRandoms(0, 10, -2.5, 2.5, seed=2), n_unk, patch_centers=centers)
    config = yb.Configuration.create(rmin=100, rmax=1000, zmin=0.1, zmax=1.0, num_bins=10)
    (corr,) = yb.crosscorrelate(config, ref, unk, unk_rand=unk.__class__(dict(unk.items()), "unk2"), engine=engine)

    links = yb.PatchLinkage.from_catalogs(config, ref, unk)
    from yet_another_wizz_b200.measurements import _angles_per_bin

    amin, amax = _angles_per_bin(config)
    p1 = [oracle.OraclePatch(*(ref[p].load_data()[f] for f in ("ra", "dec")), None, ref[p].load_data()["redshifts"])
          for p in ref]
    p2 = [oracle.OraclePatch(*(unk[p].load_data()[f] for f in ("ra", "dec"))) for p in unk]
    sw1, sw2, counts = oracle.count_pairs(p1, p2, links.patch_links, zedges=np.array(config.binning.edges),
                                          closed="right", ang_min=amin, ang_max=amax)
    assert_array_equal(corr.dd.counts.counts, counts[0])
    assert_array_equal(corr.dr.counts.counts, counts[0])
    assert_array_equal(corr.dd.sum_weights.sum_weights1, sw1)
    assert_array_equal(corr.dd.sum_weights.sum_weights2, sw2)
    assert counts.sum() > 1e4


def test_resident_catalogs_across_calls():
    """the engine keeps device catalogs and their indexes between measurement calls (the reference's tree cache,
    `trees.py:515-526`): same numbers, no re-indexing on the second call, a rebuild when the z-binning changes"""
    import yet_another_wizz_b200 as yb
    from yet_another_wizz_b200 import measurements

    g = golden_io.load("cross_unweighted")
    config = golden_cases.config_from_golden(g)
    cats = {k: golden_cases.catalog_from_golden(g, k) for k in ("ref", "unk", "ref_rand", "unk_rand")}
    eng = yb.Engine(0, cache_catalogs=True)
    try:
        kw = dict(ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"], engine=eng)
        first = yb.crosscorrelate(config, cats["ref"], cats["unk"], **kw)
        assert sum(s["index_ms"] for s in measurements.last_stats().values()) > 0
        assert len(eng._cat_cache) == 4
        second = yb.crosscorrelate(config, cats["ref"], cats["unk"], **kw)
        assert sum(s["index_ms"] for s in measurements.last_stats().values()) == 0  # nothing was rebuilt
        golden_cases.check_corrfunc(g, "cross", second, ("dd", "dr", "rd", "rr"), exact=True)
        for a, b in zip(first, second):
            assert a == b
        # another z-binning: the binned catalogs are rebuilt (their cache entries replaced), the unbinned ones are reused
        cfg = golden_io.config_of(g)
        edges = np.asarray(cfg["zedges"])
        config2 = yb.Configuration.create(rmin=cfg["rmin"], rmax=cfg["rmax"], edges=edges[: len(edges) // 2 + 1], closed=cfg["closed"])
        third = yb.crosscorrelate(config2, cats["ref"], cats["unk"], **kw)
        assert third[0].dd.counts.num_bins == len(edges) // 2
        assert len(eng._cat_cache) == 4
        ref_only = yb.Engine(0, cache_catalogs=False)
        try:
            want = yb.crosscorrelate(config2, cats["ref"], cats["unk"], ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"], engine=ref_only)
        finally:
            ref_only.close()
        for a, b in zip(third, want):
            assert a == b
        eng.cache_clear()
        assert len(eng._cat_cache) == 0
        fourth = yb.crosscorrelate(config, cats["ref"], cats["unk"], **kw)  # everything uploaded and indexed again
        assert sum(s["index_ms"] for s in measurements.last_stats().values()) > 0
        for a, b in zip(first, fourth):
            assert a == b
    finally:
        eng.close()
