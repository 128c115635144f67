"""
Generate the golden vectors in this directory by running the UNMODIFIED
reference (`/root/reference/src/yaw`, imported through `oracle/refshim.py`).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Each `.npz` stores the exact per-patch inputs the reference saw (read back from
its own patch cache, so patch assignment and row order are the reference's) and
the per-patch-pair outputs of `yaw.crosscorrelate` / `yaw.autocorrelate`
(`CorrFunc.{dd,dr,rd,rr}.counts.counts`, `.sum_weights.sum_weights{1,2}`), plus
the angular scales the reference's cosmology code produced for every z-bin, so
the checks do not depend on the accuracy of the astropy stand-in.
"""

from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refshim  # noqa: E402

yaw = refshim.import_reference()
import pandas as pd  # noqa: E402
from yaw import AngularCoordinates, Catalog, Configuration  # noqa: E402
from yaw.catalog.trees import AngularTree  # noqa: E402
from yaw.correlation.measurements import PatchLinkage  # noqa: E402


def box_points(rng, n, ra0, ra1, dec0, dec1):
    """uniform in RA x sin(Dec), like BoxRandoms (src/yaw/randoms.py:243-259)"""
    ra = np.deg2rad(rng.uniform(ra0, ra1, n)) % (2 * np.pi)
    y = rng.uniform(np.sin(np.deg2rad(dec0)), np.sin(np.deg2rad(dec1)), n)
    return ra, np.arcsin(y)


def grid_centers(nx, ny, ra0, ra1, dec0, dec1):
    ras = ra0 + (np.arange(nx) + 0.5) * (ra1 - ra0) / nx
    decs = dec0 + (np.arange(ny) + 0.5) * (dec1 - dec0) / ny
    rr, dd = np.meshgrid(ras, decs)
    return AngularCoordinates(
        np.deg2rad(np.column_stack([rr.ravel() % 360.0, dd.ravel()]))
    )


def make_catalog(tmp, name, ra, dec, centers, *, z=None, w=None, kappa=None):
    cols = dict(ra=ra, dec=dec)
    kw = dict(ra_name="ra", dec_name="dec", degrees=False, patch_centers=centers)
    if kappa is not None:
        cols["kappa"] = kappa
        kw["kappa_name"] = "kappa"
    if z is not None:
        cols["z"] = z
        kw["redshift_name"] = "z"
    if w is not None:
        cols["w"] = w
        kw["weight_name"] = "w"
    return Catalog.from_dataframe(os.path.join(tmp, name), pd.DataFrame(cols), **kw)


def dump_catalog(prefix, cat, out):
    """concatenate the reference's per-patch cache content, patch-id order"""
    ras, decs, zs, ws, ks, pids = [], [], [], [], [], []
    for pid in sorted(cat.keys()):
        data = cat[pid].load_data()
        if "kappa" in data.dtype.names:
            ks.append(np.asarray(data["kappa"]))
        ras.append(np.asarray(data["ra"]))
        decs.append(np.asarray(data["dec"]))
        if cat.has_redshifts:
            zs.append(np.asarray(data["redshifts"]))
        if cat.has_weights:
            ws.append(np.asarray(data["weights"]))
        pids.append(np.full(len(data), pid, dtype=np.int32))
    out[f"{prefix}_ra"] = np.concatenate(ras)
    out[f"{prefix}_dec"] = np.concatenate(decs)
    out[f"{prefix}_patch"] = np.concatenate(pids)
    if zs:
        out[f"{prefix}_z"] = np.concatenate(zs)
    if ws:
        out[f"{prefix}_w"] = np.concatenate(ws)
    if ks:
        out[f"{prefix}_kappa"] = np.concatenate(ks)
    out[f"{prefix}_centers"] = cat.get_centers().data
    out[f"{prefix}_radii"] = cat.get_radii().data


def dump_config(cfg, out):
    zmids = cfg.binning.binning.mids
    amin, amax = [], []
    for z in zmids:
        a, b = cfg.scales.scales.get_angle_radian(z, cosmology=cfg.cosmology)
        amin.append(a)
        amax.append(b)
    out["zedges"] = cfg.binning.edges
    out["closed"] = np.array(str(cfg.binning.closed))
    out["ang_min"] = np.array(amin)  # (n_bins, n_scales)
    out["ang_max"] = np.array(amax)
    out["rweight"] = np.array(np.nan if cfg.scales.rweight is None else cfg.scales.rweight)
    out["resolution"] = np.array(-1 if cfg.scales.resolution is None else cfg.scales.resolution)
    out["rmin"] = np.atleast_1d(cfg.scales.rmin).astype(float)
    out["rmax"] = np.atleast_1d(cfg.scales.rmax).astype(float)
    out["zmin"] = np.array(cfg.binning.zmin)
    out["zmax"] = np.array(cfg.binning.zmax)
    # the linkage cut, measurements.py:152-168
    from yaw.correlation.measurements import get_max_angle

    out["max_angle"] = get_max_angle(cfg).data


def dump_counts(tag, corrs, out):
    for s, corr in enumerate(corrs):
        for kind in ("dd", "dr", "rd", "rr"):
            nc = getattr(corr, kind)
            if nc is None:
                continue
            out[f"{tag}_{kind}_counts_s{s}"] = nc.counts.counts
            if s == 0:
                out[f"{tag}_{kind}_sw1"] = nc.sum_weights.sum_weights1
                out[f"{tag}_{kind}_sw2"] = nc.sum_weights.sum_weights2


def dump_links(cfg, cats, out):
    links = PatchLinkage.from_catalogs(cfg, *cats)
    pairs = sorted((i, j) for i, js in links.patch_links.items() for j in js)
    out["links"] = np.array(pairs, dtype=np.int32)


def case_cross(tmp, name, *, weighted, multiscale, box, nx, ny, n, seed, zbins, closed="right"):
    rng = np.random.default_rng(seed)
    centers = grid_centers(nx, ny, *box)
    n_ref, n_unk, n_rr, n_ur = n
    zlo, zhi = 0.1, 1.0
    cats = {}
    for key, npts, has_z in (
        ("ref", n_ref, True), ("unk", n_unk, False), ("ref_rand", n_rr, True), ("unk_rand", n_ur, False),
    ):
        ra, dec = box_points(rng, npts, *box)
        z = rng.uniform(zlo - 0.05, zhi + 0.05, npts) if has_z else None  # some rows fall outside the binning
        w = rng.uniform(0.5, 1.5, npts) if (weighted and key in ("ref", "unk")) else None
        cats[key] = make_catalog(tmp, f"{name}_{key}", ra, dec, centers, z=z, w=w)
    if multiscale:
        cfg = Configuration.create(
            rmin=[100, 300], rmax=[1000, 1500], rweight=-1.0, resolution=20,
            zmin=zlo, zmax=zhi, num_bins=zbins, closed=closed,
        )
    else:
        cfg = Configuration.create(rmin=100, rmax=1000, zmin=zlo, zmax=zhi, num_bins=zbins, closed=closed)
    corrs = yaw.crosscorrelate(
        cfg, cats["ref"], cats["unk"], ref_rand=cats["ref_rand"], unk_rand=cats["unk_rand"]
    )
    out = {}
    for key, cat in cats.items():
        dump_catalog(key, cat, out)
    dump_config(cfg, out)
    dump_counts("cross", corrs, out)
    dump_links(cfg, [cats["ref"], cats["unk"], cats["ref_rand"], cats["unk_rand"]], out)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "DD sum", corrs[0].dd.counts.counts.sum(), "RR sum", corrs[0].rr.counts.counts.sum())


def case_auto(tmp, name, *, box, nx, ny, n, seed, zbins, rweight, resolution, weighted, closed):
    rng = np.random.default_rng(seed)
    centers = grid_centers(nx, ny, *box)
    n_d, n_r = n
    zlo, zhi = 0.2, 0.8
    cats = {}
    for key, npts in (("data", n_d), ("rand", n_r)):
        ra, dec = box_points(rng, npts, *box)
        z = rng.uniform(zlo - 0.03, zhi + 0.03, npts)
        w = rng.uniform(0.5, 1.5, npts) if (weighted and key == "data") else None
        cats[key] = make_catalog(tmp, f"{name}_{key}", ra, dec, centers, z=z, w=w)
    cfg = Configuration.create(
        rmin=[200, 500], rmax=[1500, 3000], rweight=rweight, resolution=resolution,
        zmin=zlo, zmax=zhi, num_bins=zbins, closed=closed,
    )
    corrs = yaw.autocorrelate(cfg, cats["data"], cats["rand"], count_rr=True)
    out = {}
    for key, cat in cats.items():
        dump_catalog(key, cat, out)
    dump_config(cfg, out)
    dump_counts("auto", corrs, out)
    dump_links(cfg, [cats["data"], cats["rand"]], out)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "DD sum", corrs[0].dd.counts.counts.sum(), "RR sum", corrs[0].rr.counts.counts.sum())


def dump_scalar_counts(tag, corrs, out):
    for s, corr in enumerate(corrs):
        for kind in ("dd", "dr"):
            nc = getattr(corr, kind)
            if nc is None:
                continue
            out[f"{tag}_{kind}_kappa_counts_s{s}"] = nc.kappa_counts.counts
            out[f"{tag}_{kind}_number_counts_s{s}"] = nc.number_counts.counts


def case_scalar(tmp, name, *, box, nx, ny, n, seed, zbins, weighted, closed="right"):
    """`crosscorrelate_scalar` with and without unknown randoms and `autocorrelate_scalar` on the
    same kappa-carrying reference sample (measurements.py:651-794)."""
    rng = np.random.default_rng(seed)
    centers = grid_centers(nx, ny, *box)
    n_ref, n_unk, n_ur = n
    zlo, zhi = 0.1, 1.0
    cats = {}
    for key, npts, has_z in (("ref", n_ref, True), ("unk", n_unk, False), ("unk_rand", n_ur, False)):
        ra, dec = box_points(rng, npts, *box)
        z = rng.uniform(zlo - 0.05, zhi + 0.05, npts) if has_z else None
        w = rng.uniform(0.5, 1.5, npts) if (weighted and key in ("ref", "unk")) else None
        kappa = rng.normal(0.02, 0.3, npts) if key == "ref" else None
        cats[key] = make_catalog(tmp, f"{name}_{key}", ra, dec, centers, z=z, w=w, kappa=kappa)
    cfg = Configuration.create(
        rmin=[100, 300], rmax=[1000, 1500], rweight=-0.5 if weighted else None, resolution=10 if weighted else None,
        zmin=zlo, zmax=zhi, num_bins=zbins, closed=closed,
    )
    out = {}
    for key, cat in cats.items():
        dump_catalog(key, cat, out)
    dump_config(cfg, out)
    dump_scalar_counts("crossr", yaw.crosscorrelate_scalar(cfg, cats["ref"], cats["unk"], unk_rand=cats["unk_rand"]), out)
    dump_scalar_counts("cross", yaw.crosscorrelate_scalar(cfg, cats["ref"], cats["unk"]), out)
    dump_scalar_counts("auto", yaw.autocorrelate_scalar(cfg, cats["ref"]), out)
    dump_links(cfg, [cats["ref"], cats["unk"], cats["unk_rand"]], out)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "kn sum", out["crossr_dd_kappa_counts_s0"].sum(), "nn sum", out["crossr_dd_number_counts_s0"].sum(),
          "kk sum", out["auto_dd_kappa_counts_s0"].sum())


def case_example(tmp, name="example_2dflens"):
    """The reference's bundled 2dFLenS example (SURVEY.md section 8c ii): recipe of the reference's
    `create_example_data.py`, with the catalogs read from the bundled parquet files.  The generator
    itself checks that the reference, run here, reproduces its own golden `examples/estimate.{dat,smp,cov}`
    to the 7 decimals those files carry, then stores inputs, per-patch-pair counts and the n(z)."""
    from yaw import RedshiftData

    exdir = os.path.join(os.path.dirname(yaw.__file__), "examples")
    kw = dict(ra_name="RA", dec_name="Dec", redshift_name="redshift", weight_name="wei", patch_name="patch")
    data = Catalog.from_file(os.path.join(tmp, f"{name}_ref"), os.path.join(exdir, "2dflens_kidss_data.pqt"), **kw)
    unk = Catalog.from_file(os.path.join(tmp, f"{name}_unk"), os.path.join(exdir, "2dflens_kidss_data.pqt"), **kw)
    rand = Catalog.from_file(os.path.join(tmp, f"{name}_rand"), os.path.join(exdir, "2dflens_kidss_rand_5x.pqt"), **kw)
    cfg = Configuration.create(rmin=100, rmax=1000, zmin=0.15, zmax=0.7, num_bins=11)  # examples/__init__.py:271
    cross = yaw.crosscorrelate(cfg, data, unk, ref_rand=rand)
    auto = yaw.autocorrelate(cfg, data, rand)
    nz = RedshiftData.from_corrfuncs(cross[0], auto[0])

    want = np.loadtxt(os.path.join(exdir, "estimate.dat"))
    want_smp = np.loadtxt(os.path.join(exdir, "estimate.smp"))[:, 2:].T
    want_cov = np.loadtxt(os.path.join(exdir, "estimate.cov"))
    for got, ref, what in ((nz.data, want[:, 2], "dat"), (nz.error, want[:, 3], "err"),
                           (nz.samples, want_smp, "smp"), (nz.covariance, want_cov, "cov")):
        diff = np.nanmax(np.abs(np.asarray(got) - ref))
        print(name, what, "max abs difference to the bundled golden:", diff)
        assert diff < 2e-7, what

    out = {}
    dump_catalog("ref", data, out)
    dump_catalog("rand", rand, out)
    dump_config(cfg, out)
    dump_counts("cross", cross, out)
    dump_counts("auto", auto, out)
    dump_links(cfg, [data, unk, rand], out)
    out["nz_data"], out["nz_error"], out["nz_samples"] = nz.data, nz.error, nz.samples
    out["bundled_nz_data"], out["bundled_nz_error"], out["bundled_nz_samples"] = want[:, 2], want[:, 3], want_smp
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "cross DD sum", cross[0].dd.counts.counts.sum(), "auto DD sum", auto[0].dd.counts.counts.sum())


def case_edge(name, n=1500, theta=3.7e-3, seed=11):
    """adversarial: every matched pair (A_i, B_i) sits within a few ulp of the
    bin edge r = 2 sin(theta/2)."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, 3))
    a /= np.linalg.norm(a, axis=1)[:, None]
    t = np.cross(a, rng.normal(size=(n, 3)))
    t /= np.linalg.norm(t, axis=1)[:, None]
    b = np.cos(theta) * a + np.sin(theta) * t
    A = AngularCoordinates.from_3d(a)
    B = AngularCoordinates.from_3d(b)
    ta, tb = AngularTree(A), AngularTree(B)
    # store the exact doubles the reference's trees hold: numpy's sin/cos may differ by an
    # ulp between CPUs, which is enough to flip on-edge pairs
    out = dict(a_radec=A.data, b_radec=B.data, a_xyz=np.array(ta.data), b_xyz=np.array(tb.data),
               theta=np.array(theta))
    # edge exactly on theta as upper limit, as lower limit, and both (multi-bin)
    specs = {
        "upper": ([theta / 10], [theta]),
        "lower": ([theta], [theta * 3]),
        "multi": ([theta / 10, theta], [theta, theta * 3]),
    }
    for key, (amin, amax) in specs.items():
        out[f"{key}_ang_min"] = np.array(amin)
        out[f"{key}_ang_max"] = np.array(amax)
        out[f"{key}_counts"] = ta.count(tb, np.array(amin), np.array(amax))
        print(name, key, out[f"{key}_counts"])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def main():
    tmp = tempfile.mkdtemp(prefix="yaw_golden_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if "--example-only" in sys.argv:
            case_example(tmp)
            return
        case_scalar(tmp, "scalar_weighted", box=(10.0, 12.0, -1.0, 1.0), nx=3, ny=2, n=(2000, 3000, 4000),
                    seed=5, zbins=4, weighted=True)
        case_scalar(tmp, "scalar_unweighted", box=(200.0, 202.0, 39.0, 41.0), nx=2, ny=2, n=(1500, 2500, 3000),
                    seed=6, zbins=3, weighted=False, closed="left")
        if "--scalar-only" in sys.argv:
            return
        case_example(tmp)
        if "--example-only" in sys.argv:
            return
        case_cross(tmp, "cross_unweighted", weighted=False, multiscale=False,
                   box=(10.0, 12.0, -1.0, 1.0), nx=3, ny=2, n=(2000, 3000, 4000, 4000), seed=1, zbins=5)
        case_cross(tmp, "cross_weighted_multiscale", weighted=True, multiscale=True,
                   box=(10.0, 12.0, -1.0, 1.0), nx=3, ny=2, n=(1500, 2500, 3000, 3000), seed=2, zbins=4,
                   closed="left")
        case_auto(tmp, "auto_rweight_polewrap", box=(357.0, 363.0, 69.0, 71.0), nx=4, ny=2,
                  n=(3000, 6000), seed=3, zbins=4, rweight=-1.0, resolution=50, weighted=True, closed="right")
        case_auto(tmp, "auto_unweighted", box=(100.0, 102.0, -31.0, -29.0), nx=2, ny=2,
                  n=(3000, 5000), seed=4, zbins=3, rweight=None, resolution=None, weighted=False, closed="right")
        case_edge("edge_adversarial")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
