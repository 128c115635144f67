"""dev helper: host-side timeline of one pipelined end-to-end pass (timestamps around every C-ABI call)"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import yet_another_wizz_b200 as yb
from yet_another_wizz_b200 import pipeline

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 4
SYNC_AFTER_UPLOAD = len(sys.argv) > 2 and sys.argv[2] == "sync"
wl = bench.make_workload("C3", 1.0, field=0)
eng = yb.Engine(0)
host = {}
for key, a in wl["arrays"].items():
    h = {}
    for name in ("xyz", "weights", "zbin"):
        if a[name] is None:
            h[name] = None
            continue
        buf = eng.pinned_empty(a[name].shape, a[name].dtype)
        buf[...] = a[name]
        h[name] = buf
    h["patch_off"], h["n_bins"] = a["patch_off"], a["n_bins"]
    host[key] = h

events = []
T0 = [0.0]
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter()
        r = f(*a, **k)
        events.append((name, 1e3 * (t - T0[0]), 1e3 * (time.perf_counter() - t), r[2] if name in ("count", "count2") else None))
        return r
    setattr(obj, name, g)
wrap(eng, "upload_catalog")
wrap(eng, "count")
wrap(eng, "count2")
if SYNC_AFTER_UPLOAD:
    _orig_slices = pipeline.upload_patch_slices
    def _slices(engine, arrays, n):
        out = _orig_slices(engine, arrays, n)
        t = time.perf_counter(); engine.sync()
        events.append(("sync", 1e3 * (t - T0[0]), 1e3 * (time.perf_counter() - t), None))
        return out
    pipeline.upload_patch_slices = _slices
for rep in range(3):
    events.clear()
    eng.sync()
    T0[0] = time.perf_counter()
    counts, sums, stats, devs = pipeline.count_cross_pipelined(eng, host, wl["pair_i"], wl["pair_j"], wl["plan"].r2, groups=groups)
    total = 1e3 * (time.perf_counter() - T0[0])
    for lst in devs.values():
        for d, _, _ in lst:
            d.free()
    if rep == 0:
        print("cold pass total", total)
        for name, t, dt, st in events:
            extra = "" if st is None else f" kernel {st['kernel_ms']:.2f} index {st['index_ms']:.2f} items {st['work_items']}"
            print(f"{t:8.2f} +{dt:6.2f}  {name}{extra}")
print("total", total)
for name, t, dt, st in events:
    extra = "" if st is None else f" kernel {st['kernel_ms']:.2f} index {st['index_ms']:.2f} items {st['work_items']}"
    print(f"{t:8.2f} +{dt:6.2f}  {name}{extra}")
