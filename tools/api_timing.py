"""dev helper: wall time of the public API call (`yb.crosscorrelate`) on the C3 workload, host preparation included"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import yet_another_wizz_b200 as yb
from yet_another_wizz_b200 import measurements

wl = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "C3", 1.0, field=0)
c = wl["cats"]
for rep in range(3):
    t0 = time.perf_counter()
    corrs = yb.crosscorrelate(wl["config"], c["ref"], c["unk"], ref_rand=c["ref_rand"], unk_rand=c["unk_rand"])
    dt = time.perf_counter() - t0
    dd = corrs[0].dd.counts.counts.sum()
    rr = corrs[0].rr.counts.counts.sum()
    st = measurements.last_stats()
    print(f"crosscorrelate call {rep}: {dt * 1e3:.1f} ms wall; DD pairs {dd:.0f}, RR pairs {rr:.0f}; "
          f"kernels {sum(s['kernel_ms'] for s in st.values()):.2f} ms, index {sum(s['index_ms'] for s in st.values()):.2f} ms", flush=True)
t0 = time.perf_counter()
for key, cat in c.items():
    measurements.prepare_catalog_arrays(cat, wl["config"].binning.binning if key in ("ref", "ref_rand") else None)
print(f"host preparation alone: {(time.perf_counter() - t0) * 1e3:.1f} ms ({os.cpu_count()} cpus)")

# C2 of BASELINE.json: autocorrelate 1e6 data + 1e6 randoms, 32 patches, r-weights with resolution 50, 10 z-bins
if len(sys.argv) > 2 and sys.argv[2] == "c2":
    nx, ny = 8, 4
    box = bench.BOX
    ras = box[0] + (np.arange(nx) + 0.5) * (box[1] - box[0]) / nx
    decs = box[2] + (np.arange(ny) + 0.5) * (box[3] - box[2]) / ny
    centers = yb.AngularCoordinates(np.deg2rad([[r, d] for d in decs for r in ras]))
    pool = np.random.default_rng(7).uniform(0.1, 1.0, 1_000_000)
    data = yb.Catalog.from_random("d", yb.BoxRandoms(*box, redshifts=pool, seed=1), 1_000_000, patch_centers=centers)
    rand = yb.Catalog.from_random("r", yb.BoxRandoms(*box, redshifts=pool, seed=3), 1_000_000, patch_centers=centers)
    cfg = yb.Configuration.create(rmin=100, rmax=1000, rweight=-1.0, resolution=50, zmin=0.1, zmax=1.0, num_bins=10)
    for rep in range(3):
        t0 = time.perf_counter()
        corrs = yb.autocorrelate(cfg, data, rand)
        dt = time.perf_counter() - t0
        st = measurements.last_stats()
        print(f"C2 autocorrelate call {rep}: {dt * 1e3:.1f} ms wall; DD {corrs[0].dd.counts.counts.sum():.6g}; "
              f"kernels {sum(s['kernel_ms'] for s in st.values()):.2f} ms, index {sum(s['index_ms'] for s in st.values()):.2f} ms", flush=True)
