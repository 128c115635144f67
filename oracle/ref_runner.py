#!/usr/bin/env python
"""
TEST / BENCHMARK INFRASTRUCTURE ONLY -- never imported by the product.

Times the UNMODIFIED reference (`yaw`, v3.1.1) on the benchmark workload, on the host cores, in its own
process (it forks `multiprocessing.Pool`s, `src/yaw/utils/parallel.py:342`; never fork after CUDA init):

    python oracle/ref_runner.py --workload C3 --stripes 2 --steps 2 --warmup 1 [--src baseline/_ref]

The package is imported from `--src` (default: `baseline/_ref`, where `__graft_entry__.build()` installs it
with `pip install --no-deps --target baseline/_ref`; falls back to `/root/reference/src`) through the five
stubs of `oracle/refshim.py` (astropy / h5py / treecorr / strenum / `yaw._version` are absent from the image).

Workload: the first `--stripes` declination stripes of the benchmark's patch grid (each stripe = one row of
patches over the full RA range), same densities, scales, z-bins and generator recipe as `bench.py`; a full
run is `--stripes <ny>`.  Per step, exactly what `yaw.crosscorrelate` does (`measurements.py:596-626`):
`Catalog.build_trees` for the four catalogs (forced, so that every step rebuilds like the GPU arm re-indexes),
`PatchLinkage.from_catalogs`, then `count_pairs` for DD, DR, RD, RR.  Prints one JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKLOADS = {  # the rows of bench.py's table that the reference arm supports (one scale, no r-weights)
    "C1": dict(n=(100_000, 100_000, 1_000_000, 1_000_000), grid=(4, 4), zmin=0.1, zmax=1.0, bins=10),
    "C3": dict(n=(1_000_000, 10_000_000, 10_000_000, 10_000_000), grid=(8, 8), zmin=0.07, zmax=1.42, bins=30),
    "C5": dict(n=(10_000_000, 100_000_000, 100_000_000, 100_000_000), grid=(16, 16), zmin=0.07, zmax=1.42, bins=50),
}
BOX = (0.0, 40.0, -12.5, 12.5)
SEEDS = dict(ref=1, unk=2, ref_rand=3, unk_rand=4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3", choices=list(WORKLOADS))
    ap.add_argument("--stripes", type=int, default=0, help="declination stripes of the patch grid (0 = all)")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--workers", type=int, default=0, help="YAW_NUM_THREADS (0 = all the reference will take)")
    ap.add_argument("--src", default=None)
    ap.add_argument("--dump", default=None, help="write the counts of the last step to this .npz (parity checks)")
    args = ap.parse_args()

    src = args.src
    if src is None:
        cand = os.path.join(ROOT, "baseline", "_ref")
        src = cand if os.path.isdir(os.path.join(cand, "yaw")) else "/root/reference/src"
    os.environ["YAW_REFERENCE_SRC"] = src
    if args.workers > 0:
        os.environ["YAW_NUM_THREADS"] = str(args.workers)  # fixed at import (parallel.py:137-142)
    else:
        os.environ.pop("YAW_NUM_THREADS", None)
        os.environ["YAW_NUM_THREADS"] = str(os.cpu_count() or 1)  # the reference caps it at the cores per socket
    sys.path.insert(0, HERE)
    import refshim

    yaw = refshim.import_reference()
    from yaw.correlation.measurements import PatchLinkage
    from yaw.randoms import BoxRandoms
    from yaw.utils import parallel

    spec = WORKLOADS[args.workload]
    nx, ny = spec["grid"]
    stripes = ny if args.stripes <= 0 else min(args.stripes, ny)
    dec_lo = BOX[2]
    dec_hi = BOX[2] + (BOX[3] - BOX[2]) * stripes / ny
    ras = BOX[0] + (np.arange(nx) + 0.5) * (BOX[1] - BOX[0]) / nx
    decs = BOX[2] + (np.arange(ny) + 0.5) * (BOX[3] - BOX[2]) / ny
    centers = yaw.AngularCoordinates(np.deg2rad([[r, d] for d in decs[:stripes] for r in ras]))
    # BoxRandoms is uniform in sin(dec): rows in proportion to the area of the stripes
    frac = (np.sin(np.deg2rad(dec_hi)) - np.sin(np.deg2rad(dec_lo))) / (np.sin(np.deg2rad(BOX[3])) - np.sin(np.deg2rad(BOX[2])))
    pool = np.random.default_rng(7).uniform(spec["zmin"], spec["zmax"], 1_000_000)
    config = yaw.Configuration.create(rmin=100, rmax=1000, zmin=spec["zmin"], zmax=spec["zmax"], num_bins=spec["bins"])
    edges, closed = config.binning.edges, config.binning.closed

    tmp = tempfile.mkdtemp(prefix="yawb_refrun_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        t0 = time.perf_counter()
        cats = {}
        for key, n in zip(("ref", "unk", "ref_rand", "unk_rand"), spec["n"]):
            n = max(int(n * args.scale * frac), 1000)
            gen = BoxRandoms(BOX[0], BOX[1], dec_lo, dec_hi, redshifts=pool if key in ("ref", "ref_rand") else None, seed=SEEDS[key])
            cats[key] = yaw.Catalog.from_random(os.path.join(tmp, key), gen, n, patch_centers=centers, overwrite=True)
        t_cat = time.perf_counter() - t0

        steps = []
        last = None
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            cats["ref"].build_trees(edges, closed=closed, force=True)
            cats["ref_rand"].build_trees(edges, closed=closed, force=True)
            cats["unk"].build_trees(None, force=True)
            cats["unk_rand"].build_trees(None, force=True)
            t_build = time.perf_counter() - t0
            t0 = time.perf_counter()
            links = PatchLinkage.from_catalogs(config, cats["ref"], cats["unk"], cats["ref_rand"], cats["unk_rand"])
            res, t_each = {}, {}
            for tag, (a, b) in dict(DD=("ref", "unk"), DR=("ref", "unk_rand"), RD=("ref_rand", "unk"), RR=("ref_rand", "unk_rand")).items():
                t1 = time.perf_counter()
                res[tag] = links.count_pairs(cats[a], cats[b])
                t_each[tag] = time.perf_counter() - t1
            t_count = time.perf_counter() - t0
            last = res
            if step >= args.warmup:
                steps.append(dict(build_s=t_build, count_s=t_count, per_count_s=t_each))

        # naive linked pair tests of this sample: sum over linked patch pairs and z-bins of n1 * n2
        pairs = list(links.iter_patch_id_pairs(auto=False))
        pi = np.array([p[0] for p in pairs]); pj = np.array([p[1] for p in pairs])
        naive, in_scale = {}, {}
        for tag, norm in last.items():
            sw = norm[0].sum_weights
            naive[tag] = int((sw.sum_weights1[:, pi] * sw.sum_weights2[:, pj]).sum())
            in_scale[tag] = float(norm[0].counts.counts.sum())
        if args.dump:
            np.savez(args.dump, pair_i=pi, pair_j=pj, **{f"{t}_counts": n[0].counts.counts for t, n in last.items()})
        out = dict(
            impl="reference", source=src, version=getattr(yaw, "__version__", "?"), workload=args.workload, scale=args.scale,
            stripes=stripes, stripes_total=ny, patches=int(nx * stripes), patches_total=int(nx * ny), linked_pairs=len(pairs),
            workers=int(parallel.get_size(None)) if hasattr(parallel, "get_size") else None,
            host_cores=os.cpu_count(), catalogs_s=t_cat, steps=steps, naive_pair_tests=naive, pairs_in_scale=in_scale,
            rows={k: int(sum(c.get_num_records())) for k, c in cats.items()},
        )
        print(json.dumps(out), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
