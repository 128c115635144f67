"""dev helper: time repeated index builds of one big synthetic catalog (rows, patches, bins from argv)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yet_another_wizz_b200 as yb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n_bins = int(sys.argv[3]) if len(sys.argv) > 3 else 50
rng = np.random.default_rng(1)
P = side * side
per = n // P
t0 = time.perf_counter()
ra0, dec0 = np.deg2rad(40.0), np.deg2rad(25.0)
xyz = np.empty((per * P, 3))
for p in range(P):
    iu, iv = p % side, p // side
    ra = (iu + rng.random(per)) * ra0 / side
    dec = -dec0 / 2 + (iv + rng.random(per)) * dec0 / side
    s = slice(p * per, (p + 1) * per)
    xyz[s, 0] = np.cos(dec) * np.cos(ra); xyz[s, 1] = np.cos(dec) * np.sin(ra); xyz[s, 2] = np.sin(dec)
off = np.arange(P + 1, dtype=np.int64) * per
zbin = rng.integers(0, n_bins, per * P).astype(np.int32)
print(f"generated {per * P} rows in {time.perf_counter() - t0:.1f}s", flush=True)
eng = yb.Engine(0)
for binned in (True, False):
    dev = eng.upload_catalog(xyz, off, zbin=zbin if binned else None, n_bins=n_bins if binned else 1)
    eng.sync()
    role = 1 if binned else 2
    for rep in range(5):
        dev.drop_index()
        eng.sync()
        t = time.perf_counter()
        ms = dev.build_index(role)
        eng.sync()
        print(f"role {role} rep {rep}: device {ms:.2f} ms, wall {1e3 * (time.perf_counter() - t):.2f} ms", flush=True)
    dev.free()

# the bench's pattern: four catalogs, all indexes dropped, rebuilt in the order the counts need them
print("four catalogs, drop all / rebuild all", flush=True)
small = per * P // 10
cats = [
    (eng.upload_catalog(xyz, off, zbin=zbin, n_bins=n_bins), 1),
    (eng.upload_catalog(xyz, off), 2),
    (eng.upload_catalog(xyz[::-1].copy() if False else xyz, off), 2),
    (eng.upload_catalog(xyz[: (per // 10) * P].reshape(-1, 3) if False else xyz[:small], np.minimum(off, small), zbin=zbin[:small], n_bins=n_bins), 1),
]
eng.sync()
for rep in range(6):
    for dev, _ in cats:
        dev.drop_index()
    eng.sync()
    t = time.perf_counter()
    parts = [dev.build_index(role) for dev, role in cats]
    eng.sync()
    print(f"rep {rep}: wall {1e3 * (time.perf_counter() - t):.2f} ms, device {[round(p, 2) for p in parts]}", flush=True)
