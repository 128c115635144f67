"""
Checks against the LIVE unmodified reference (`/root/reference`, imported through
`oracle/refshim.py`).  Only runnable in the build container; skipped elsewhere.
"""

import os
import shutil
import tempfile

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import refshim
from fake_engine import OracleEngine

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not refshim.reference_available(), reason="reference checkout not present"),
]


@pytest.fixture(scope="module")
def yaw():
    return refshim.import_reference()


@pytest.fixture(scope="module")
def tmpdir():
    d = tempfile.mkdtemp(prefix="yawb_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    yield d
    shutil.rmtree(d, ignore_errors=True)


def test_boxrandoms_identical(yaw):
    from yaw.randoms import BoxRandoms as RefBox

    import yet_another_wizz_b200 as yb

    pool_z = np.random.default_rng(7).uniform(0.1, 1.0, 1000)
    pool_w = np.random.default_rng(8).uniform(0.5, 1.5, 1000)
    ref = RefBox(0, 40, -12.5, 12.5, redshifts=pool_z, weights=pool_w, seed=3)(5000)
    mine = yb.BoxRandoms(0, 40, -12.5, 12.5, redshifts=pool_z, weights=pool_w, seed=3)(5000)
    for key in ("ra", "dec", "redshifts", "weights"):
        assert_array_equal(mine[key], np.asarray(ref[key]))


def _make_ref_catalogs(yaw, tmpdir, tag, weighted, kappa=False):
    import pandas as pd
    from yaw import AngularCoordinates, Catalog

    rng = np.random.default_rng(21)
    centers = AngularCoordinates(np.deg2rad([[20.5, -0.5], [21.5, -0.5], [20.5, 0.5], [21.5, 0.5]]))
    cats = {}
    for key, n, has_z in (("ref", 1500, True), ("unk", 2500, False), ("ref_rand", 3000, True), ("unk_rand", 3000, False)):
        cols = dict(ra=rng.uniform(20, 22, n), dec=rng.uniform(-1, 1, n))
        kw = dict(ra_name="ra", dec_name="dec", patch_centers=centers)
        if has_z:
            cols["z"] = rng.uniform(0.1, 0.9, n)
            kw["redshift_name"] = "z"
        if weighted and key in ("ref", "unk"):
            cols["w"] = rng.uniform(0.5, 1.5, n)
            kw["weight_name"] = "w"
        if kappa and key == "ref":
            cols["kappa"] = rng.normal(0.0, 0.3, n)
            kw["kappa_name"] = "kappa"
        cats[key] = Catalog.from_dataframe(os.path.join(tmpdir, f"{tag}_{key}"), pd.DataFrame(cols), **kw)
    return cats


@pytest.mark.parametrize("weighted", [False, True])
def test_dropin_with_reference_objects(yaw, tmpdir, weighted):
    """genuine yaw.Configuration / yaw.Catalog objects go straight into this package's
    crosscorrelate / autocorrelate; results equal the reference's own"""
    import yet_another_wizz_b200 as yb

    cats = _make_ref_catalogs(yaw, tmpdir, f"w{int(weighted)}", weighted)
    kw = dict(rmin=[100, 400], rmax=[1000, 2000], zmin=0.1, zmax=0.9, num_bins=4)
    if weighted:
        kw.update(rweight=-1.0, resolution=15)
    config = yaw.Configuration.create(**kw)
    want = yaw.crosscorrelate(config, cats["ref"], cats["unk"], ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"])
    got = yb.crosscorrelate(config, cats["ref"], cats["unk"], ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"],
                            engine=OracleEngine())
    assert len(got) == len(want) == 2
    for g, w in zip(got, want):
        for kind in ("dd", "dr", "rd", "rr"):
            a, b = getattr(g, kind), getattr(w, kind)
            if weighted:
                assert_allclose(a.counts.counts, b.counts.counts, rtol=1e-12, atol=0)
                assert_allclose(a.sum_weights.sum_weights1, b.sum_weights.sum_weights1, rtol=1e-12)
            else:
                assert_array_equal(a.counts.counts, b.counts.counts)
                assert_array_equal(a.sum_weights.sum_weights1, b.sum_weights.sum_weights1)
            assert_allclose(a.sum_weights.sum_weights2, b.sum_weights.sum_weights2, rtol=1e-12)
        if not weighted:
            # hand-over: the reference's own estimator + jackknife run unchanged on our counts
            ref_obj = g.to_reference()
            assert ref_obj == w
            assert_array_equal(ref_obj.sample().data, w.sample().data)

    want = yaw.autocorrelate(config, cats["ref"], cats["ref_rand"])
    got = yb.autocorrelate(config, cats["ref"], cats["ref_rand"], engine=OracleEngine())
    for g, w in zip(got, want):
        for kind in ("dd", "dr", "rr"):
            a, b = getattr(g, kind), getattr(w, kind)
            assert_allclose(a.counts.counts, b.counts.counts, rtol=1e-12, atol=0)
            assert_allclose(a.sum_weights.get_array(), b.sum_weights.get_array(), rtol=1e-12)
            assert_allclose(a.sample_patch_sum().samples, b.sample_patch_sum().samples, rtol=1e-10)


def test_scalar_dropin_with_reference_objects(yaw, tmpdir):
    """yaw.Catalog objects carrying `kappa` through crosscorrelate_scalar / autocorrelate_scalar"""
    import yet_another_wizz_b200 as yb

    cats = _make_ref_catalogs(yaw, tmpdir, "kappa", True, kappa=True)
    config = yaw.Configuration.create(rmin=[100, 400], rmax=[1000, 2000], zmin=0.1, zmax=0.9, num_bins=4)
    runs = [
        (yaw.crosscorrelate_scalar(config, cats["ref"], cats["unk"], unk_rand=cats["unk_rand"]),
         yb.crosscorrelate_scalar(config, cats["ref"], cats["unk"], unk_rand=cats["unk_rand"], engine=OracleEngine())),
        (yaw.crosscorrelate_scalar(config, cats["ref"], cats["unk"]),
         yb.crosscorrelate_scalar(config, cats["ref"], cats["unk"], engine=OracleEngine())),
        (yaw.autocorrelate_scalar(config, cats["ref"]),
         yb.autocorrelate_scalar(config, cats["ref"], engine=OracleEngine())),
    ]
    for want, got in runs:
        assert len(want) == len(got) == 2
        for w, g in zip(want, got):
            for kind in ("dd", "dr"):
                b, a = getattr(w, kind), getattr(g, kind)
                assert (a is None) == (b is None)
                if b is None:
                    continue
                scale = np.abs(b.kappa_counts.counts).max()
                assert_allclose(a.kappa_counts.counts, b.kappa_counts.counts, rtol=1e-12, atol=1e-12 * scale)
                assert_allclose(a.number_counts.counts, b.number_counts.counts, rtol=1e-12, atol=0)
                assert_allclose(a.sample_patch_sum().data, b.sample_patch_sum().data, rtol=1e-9, atol=1e-12)


def test_bundled_example_reproduces_golden_nz(yaw):
    """counts of this package -> the reference's own estimator and jackknife -> the golden n(z) the
    reference ships in `examples/estimate.{dat,smp}` (7 decimals), `tests/test_setups.py:149-168`"""
    import golden_cases
    import golden_io
    from yaw import RedshiftData

    g = golden_io.load("example_2dflens")
    cross, auto = golden_cases.run_example(g, OracleEngine())
    nz = RedshiftData.from_corrfuncs(cross[0].to_reference(), auto[0].to_reference())
    assert_allclose(nz.data, g["bundled_nz_data"], rtol=0, atol=2e-7)
    assert_allclose(nz.error, g["bundled_nz_error"], rtol=0, atol=2e-7)
    assert_allclose(nz.samples, g["bundled_nz_samples"], rtol=0, atol=2e-7)
    assert_allclose(nz.data, g["nz_data"], rtol=1e-9)


def test_read_reference_cache(yaw, tmpdir):
    """`Catalog.from_cache` opens the reference's on-disk patch cache byte-compatibly"""
    import yet_another_wizz_b200 as yb

    cats = _make_ref_catalogs(yaw, tmpdir, "cache", True)
    ref = cats["ref"]
    mine = yb.Catalog.from_cache(ref.cache_directory)
    assert mine.get_num_records() == ref.get_num_records()
    assert_array_equal(mine.get_centers().data, ref.get_centers().data)
    assert_array_equal(mine.get_radii().data, ref.get_radii().data)
    assert mine.has_weights and mine.has_redshifts
    for pid in ref.keys():
        a, b = mine[pid].load_data(), ref[pid].load_data()
        for f in ("ra", "dec", "weights", "redshifts"):
            assert_array_equal(a[f], b[f])


def test_oracle_against_live_reference(yaw):
    """the restated primitive equals AngularTree.count on fresh random inputs"""
    import oracle
    from yaw import AngularCoordinates
    from yaw.catalog.trees import AngularTree

    rng = np.random.default_rng(99)
    for trial in range(3):
        a = np.column_stack([rng.uniform(0, 0.05, 3000), rng.uniform(-0.02, 0.02, 3000)])
        b = np.column_stack([rng.uniform(0, 0.05, 4000), rng.uniform(-0.02, 0.02, 4000)])
        wa = rng.uniform(0.5, 1.5, 3000) if trial else None
        ta, tb = AngularTree(AngularCoordinates(a), wa), AngularTree(AngularCoordinates(b))
        amin, amax = np.array([2e-4, 5e-4]), np.array([2e-3, 4e-3])
        kw = dict(weight_scale=-1.0, weight_res=30) if trial == 2 else {}
        want = ta.count(tb, amin, amax, **kw)
        got = oracle.tree_count(oracle.radec_to_xyz(a[:, 0], a[:, 1]), wa, oracle.radec_to_xyz(b[:, 0], b[:, 1]),
                                None, amin, amax, **kw)
        if trial == 0:
            assert_array_equal(got, want)
        else:
            assert_allclose(got, want, rtol=1e-12)
