#!/usr/bin/env python
"""
Benchmark of the pair-counting hot path: one "step" = one full `crosscorrelate` counting pass
(index build + DD + DR + RD + RR) over the synthetic workload named by BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3]

Prints ONE JSON line (rank 0).  Metric: effective pair tests per second =
(sum over count types, linked patch pairs and z-bins of n1*n2, a property of the workload)
/ step time.  The same numerator is used for the GPU arm and for the CPU reference arm, so the
two lines are directly comparable; `roofline` reports the tests the kernel actually EXECUTED
(after sky-cell pruning) against the FP32 CUDA-core pair-test roofline of SURVEY.md section 8d.

  value    inputs (raw unit vectors, z-bin ids) already resident in HBM; the timed region covers
           index build (keys, radix sort, gather, cell table, tiles) + the four count kernels +
           the D2H of the per-patch-pair results, timed with CUDA events on the engine's stream
  e2e      the C-ABI calls with HOST buffers: H2D upload of every catalog from pinned memory,
           index build, counts, D2H of the results -- wall clock around the calls
           (schedule: `yet_another_wizz_b200.pipeline.count_cross_pipelined`)
  --impl reference   the CPU implementation of the same path (oracle/cpu_port.py: scipy cKDTree
           dual-tree count_neighbors + multiprocessing task farm, exactly the reference's
           algorithm; /root/reference itself does not exist on the GPU box) on a bounded sample
  --workload   C3 (default, the configuration the north-star target is quoted on), C1 (small), C5 (1e8-row
           catalogs), C4 (C3 with three scales: cumulative sub-bin kernel), C3w (C3 with r-weights,
           resolution 50: general sub-bin path), C3wt (C3 with weighted data samples)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: n_ref, n_unk, n_ref_rand, n_unk_rand, grid (nx, ny), zmin, zmax, n_bins   
    "C1": dict(n=(100_000, 100_000, 1_000_000, 1_000_000), grid=(4, 4), zmin=0.1, zmax=1.0, bins=10),
    "C3": dict(n=(1_000_000, 10_000_000, 10_000_000, 10_000_000), grid=(8, 8), zmin=0.07, zmax=1.42, bins=30),
    "C5": dict(n=(10_000_000, 100_000_000, 100_000_000, 100_000_000), grid=(16, 16), zmin=0.07, zmax=1.42, bins=50),
    # C3 with three scales (sub-bin histogram path of the kernel) / with r-weights (resolution 50)
    "C4": dict(n=(1_000_000, 10_000_000, 10_000_000, 10_000_000), grid=(8, 8), zmin=0.07, zmax=1.42, bins=30,
               scales=dict(rmin=[100, 300, 500], rmax=[1000, 1500, 2000])),
    # C3 with weights on the data samples (SURVEY.md section 8d "weighted variants")
    "C3wt": dict(n=(1_000_000, 10_000_000, 10_000_000, 10_000_000), grid=(8, 8), zmin=0.07, zmax=1.42, bins=30,
                 weighted=True),
    "C3w": dict(n=(1_000_000, 10_000_000, 10_000_000, 10_000_000), grid=(8, 8), zmin=0.07, zmax=1.42, bins=30,
                scales=dict(rmin=100, rmax=1000, rweight=-1.0, resolution=50)),
}
BOX = (0.0, 40.0, -12.5, 12.5)
SEEDS = dict(ref=1, unk=2, ref_rand=3, unk_rand=4)
FP32_INSTR_PER_TEST = 6  # 3 FSUB + FMUL + 2 FFMA, SURVEY.md section 8d


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
# stdout when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the
# result line goes to a private duplicate of the original stdout.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---- workload ------------------------------------------------------------------------------------
def make_workload(name: str, scale: float = 1.0, field: int = 0):
    """BoxRandoms catalogs + configuration + host-prepared arrays of the workload.  `field` > 0 places an
    independent realisation of the same survey 45 degrees further east (weak-scaling runs: one field per GPU)."""
    import yet_another_wizz_b200 as yb
    from yet_another_wizz_b200.measurements import PatchLinkage, _angles_per_bin, prepare_catalog_arrays
    from yet_another_wizz_b200.angular import AngularBinPlan

    spec = WORKLOADS[name]
    nx, ny = spec["grid"]
    box = (BOX[0] + 45.0 * field, BOX[1] + 45.0 * field, BOX[2], BOX[3])
    ras = box[0] + (np.arange(nx) + 0.5) * (box[1] - box[0]) / nx
    decs = BOX[2] + (np.arange(ny) + 0.5) * (BOX[3] - BOX[2]) / ny
    centers = yb.AngularCoordinates(np.deg2rad([[r, d] for d in decs for r in ras]))
    pool = np.random.default_rng(7).uniform(spec["zmin"], spec["zmax"], 1_000_000)
    config = yb.Configuration.create(**spec.get("scales", dict(rmin=100, rmax=1000)), zmin=spec["zmin"], zmax=spec["zmax"],
                                     num_bins=spec["bins"])
    binning = config.binning.binning

    t0 = time.perf_counter()
    cats, arrays = {}, {}
    for key, n in zip(("ref", "unk", "ref_rand", "unk_rand"), spec["n"]):
        n = max(int(n * scale), 1000)
        has_z = key in ("ref", "ref_rand")
        wpool = np.random.default_rng(8).uniform(0.5, 1.5, 1_000_000) if (spec.get("weighted") and key in ("ref", "unk")) else None
        gen = yb.BoxRandoms(*box, redshifts=pool if has_z else None, weights=wpool, seed=SEEDS[key] + 100 * field)
        cats[key] = yb.Catalog.from_random(key, gen, n, patch_centers=centers)
    t_cat = time.perf_counter() - t0
    t0 = time.perf_counter()
    for key, cat in cats.items():
        arrays[key] = prepare_catalog_arrays(cat, binning if key in ("ref", "ref_rand") else None)
    t_prep = time.perf_counter() - t0

    links = PatchLinkage.from_catalogs(config, cats["ref"], cats["unk"], cats["ref_rand"], cats["unk_rand"])
    pair_i, pair_j = links.get_patch_id_pairs(auto=False)
    amin, amax = _angles_per_bin(config)
    plan = AngularBinPlan(amin, amax, config.scales.rweight, config.scales.resolution)

    # naive linked pair tests per count type
    naive = {}
    counts = {k: np.diff(arrays[k]["patch_off"]) for k in arrays}
    binned_counts = {}
    for k in ("ref", "ref_rand"):
        zb, off = arrays[k]["zbin"], arrays[k]["patch_off"]
        ok = (zb >= 0) & (zb < len(binning))
        binned_counts[k] = np.add.reduceat(ok.astype(np.int64), off[:-1])
    for tag, (a, b) in dict(DD=("ref", "unk"), DR=("ref", "unk_rand"), RD=("ref_rand", "unk"),
                            RR=("ref_rand", "unk_rand")).items():
        naive[tag] = int((binned_counts[a][pair_i].astype(np.float64) * counts[b][pair_j]).sum())
    return dict(name=name, spec=spec, config=config, cats=cats, arrays=arrays, pair_i=pair_i, pair_j=pair_j,
                plan=plan, naive=naive, t_catalogs=t_cat, t_host_prep=t_prep, n_patch=nx * ny)


def subset_patches(a: dict, keep: np.ndarray) -> dict:
    """host arrays of one catalog restricted to the patches in `keep` (all others become empty)"""
    off = a["patch_off"]
    sizes = np.zeros(len(off) - 1, dtype=np.int64)
    sizes[keep] = np.diff(off)[keep]
    sel = np.concatenate([np.arange(off[p], off[p + 1]) for p in keep]) if len(keep) else np.empty(0, dtype=np.int64)
    return dict(
        xyz=np.ascontiguousarray(a["xyz"][sel]), patch_off=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
        weights=None if a["weights"] is None else np.ascontiguousarray(a["weights"][sel]),
        zbin=None if a["zbin"] is None else np.ascontiguousarray(a["zbin"][sel]), n_bins=a["n_bins"],
    )


COUNT_TYPES = dict(DD=("ref", "unk"), DR=("ref", "unk_rand"), RD=("ref_rand", "unk"), RR=("ref_rand", "unk_rand"))


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples taken while the timed regions run.  In-process NVML
    (a few microseconds per sample); one `nvidia-smi` query costs ~0.3 s and was seen to stall a
    concurrently running 10 ms step, so it is only the fallback."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.rows, self._stop, self.index = [], threading.Event(), index
        self._t = threading.Thread(target=self._run, daemon=True)
        self.source = "nvml"

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
        idx = self.index
        if visible and visible[0].strip().isdigit() and idx < len(visible):
            idx = int(visible[idx])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _run(self):
        try:
            nv, h = self._nvml_handle()
            sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            while not self._stop.is_set():
                r = get_reasons(h)
                row = [nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), sm_max, nv.nvmlDeviceGetPowerUsage(h) / 1e3,
                       "", "", "", ""]
                for bit, col in bits:
                    row[col] = "Active" if (r & bit) else "Not Active"
                self.rows.append([str(c) for c in row])
                self._stop.wait(0.01)
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(self.rows[0][1]), reasons=reasons,
                    samples=len(self.rows), power_w_max=max(float(r[2]) for r in self.rows), source=self.source)


# ---- CPU arm (oracle port of the reference's dual-tree path) --------------------------------------------
def cpu_sample(wl, budget_s: float, workers: int | None = None, fixed_m: int | None = None):
    """Time the CPU implementation on a bounded sample: the first `m` patches' diagonal pairs plus all
    their links, all four count types, trees built for exactly the patches touched."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_port

    workers = workers or cpu_port.physical_cores()
    pi, pj = wl["pair_i"], wl["pair_j"]
    n_bins = len(wl["config"].binning.binning)
    plan = wl["plan"]
    ang_min = np.array([lim[:, 0] for lim in plan.limits])
    ang_max = np.array([lim[:, 1] for lim in plan.limits])

    def rows_of(key, patches, binned):
        a = wl["arrays"][key]
        out = {}
        for p in patches:
            s, e = a["patch_off"][p], a["patch_off"][p + 1]
            out[p] = (a["xyz"][s:e], None if a["weights"] is None else a["weights"][s:e],
                      a["zbin"][s:e] if binned else None)
        return out

    def run(m):
        first = set(range(m))
        sel = [(int(i), int(j)) for i, j in zip(pi, pj) if int(i) in first]
        need1 = sorted({i for i, _ in sel})
        need2 = sorted({j for _, j in sel})
        t_build = t_count = 0.0
        naive = 0
        trees = {}
        for key, binned, need in (("ref", True, need1), ("ref_rand", True, need1), ("unk", False, need2),
                                  ("unk_rand", False, need2)):
            rows = rows_of(key, need, binned)
            built, dt = cpu_port.build_catalog_trees([rows[p] for p in need], n_bins if binned else None)
            trees[key] = dict(zip(need, built))
            t_build += dt
        for tag, (a, b) in COUNT_TYPES.items():
            _, dt = cpu_port.count_pairs(trees[a], trees[b], sel, ang_min, ang_max, workers=workers)
            t_count += dt
            ca, cb = wl["arrays"][a], wl["arrays"][b]
            for i, j in sel:
                zb = ca["zbin"][ca["patch_off"][i]:ca["patch_off"][i + 1]]
                n1 = int(((zb >= 0) & (zb < n_bins)).sum())
                naive += n1 * int(cb["patch_off"][j + 1] - cb["patch_off"][j])
        return dict(patches=m, pairs=len(sel), t_build=t_build, t_count=t_count, naive=naive, workers=workers)

    if fixed_m is not None:
        return run(fixed_m)
    # grow the sample until it is worth ~budget_s of CPU work
    m = 1
    res = run(m)
    while res["t_build"] + res["t_count"] < budget_s / 3 and m < wl["n_patch"]:
        per_patch = (res["t_build"] + res["t_count"]) / m
        m = int(min(wl["n_patch"], max(m + 1, budget_s / max(per_patch, 1e-3))))
        res = run(m)
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, args.scale)
    total_naive = sum(wl["naive"].values())
    times, last = [], None
    budget = args.cpu_budget
    for step in range(args.warmup + args.steps):
        last = cpu_sample(wl, budget) if last is None else cpu_sample(wl, budget, fixed_m=last["patches"])
        if step >= args.warmup:
            times.append(last["t_build"] + last["t_count"])
    t = float(np.mean(times))
    value = last["naive"] / t / 1e9
    line = dict(
        impl="reference", metric="crosscorrelate_effective_pair_tests_per_s", value=value, unit="Gpairs/s",
        n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload=f"{wl['name']} crosscorrelate DD+DR+RD+RR, BoxRandoms, scale={args.scale}",
                    sample=f"{last['patches']} of {wl['n_patch']} first-catalog patches with all their links "
                           f"({last['pairs']} patch pairs x 4 count types), tree build included"),
        cpu_baseline=dict(value=value, unit="Gpairs/s", cores=last["workers"], kind="port",
                          sample=f"{last['patches']}/{wl['n_patch']} patches, {last['pairs']} patch pairs, "
                                 f"build {last['t_build']:.2f}s + count {last['t_count']:.2f}s",
                          extrapolated_full_job_s=t * total_naive / max(last["naive"], 1)),
        e2e=dict(value=value, unit="Gpairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    emit(line)


# ---- GPU arm ------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import yet_another_wizz_b200 as yb
    from yet_another_wizz_b200 import _lib
    from yet_another_wizz_b200.pipeline import count_cross_pipelined
    from yet_another_wizz_b200.sharding import assign_patches_contiguous, pair_costs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    weak = world > 1 and args.scaling == "weak"
    wl = make_workload(args.workload, args.scale, field=rank if weak else 0)
    if rank == 0:
        log(f"[bench] workload {wl['name']} scale {args.scale}: catalogs {wl['t_catalogs']:.1f}s, "
            f"host prep {wl['t_host_prep']:.1f}s, {len(wl['pair_i'])} linked patch pairs, naive tests {wl['naive']}")
    eng = yb.Engine(local_rank)
    pi, pj, plan = wl["pair_i"], wl["pair_j"], wl["plan"]
    total_naive = sum(wl["naive"].values())
    if weak:  # one independent field per GPU: the job is the union of the fields
        import torch

        t = torch.tensor([float(total_naive), float(len(pi)), -float(len(pi))], dtype=torch.float64, device="cuda")
        tsum, tmax = t.clone(), t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_naive = int(tsum[0].item())
        n_pairs_max = int(tmax[1].item())
    else:
        n_pairs_max = len(pi)

    # this rank's share: second-catalog patches are dealt out as spatially compact groups of equal summed
    # pair cost, a rank
    # counts every linked pair of its patches and holds only the rows it needs (its own second-catalog
    # patches, plus the first-catalog patches linked to them)
    own = np.arange(len(pi))
    arrays = wl["arrays"]
    if world > 1 and not weak:
        n1 = np.diff(arrays["ref_rand"]["patch_off"])
        n2 = np.diff(arrays["unk_rand"]["patch_off"])
        costs = pair_costs(pi, pj, n1, n2)
        patch_cost = np.bincount(pj, weights=costs, minlength=wl["n_patch"])
        centers_xyz = wl["cats"]["unk_rand"].get_centers().to_3d()
        my_patches = assign_patches_contiguous(patch_cost, centers_xyz, world)[rank]
        own = np.flatnonzero(np.isin(pj, my_patches))
        need = dict(second=np.unique(pj[own]), first=np.unique(pi[own]))
        arrays = {k: subset_patches(a, need["first" if k in ("ref", "ref_rand") else "second"])
                  for k, a in arrays.items()}
    opi, opj = pi[own], pj[own]

    # pinned host staging of the inputs (what a caller holding host buffers hands to the C ABI)
    host = {}
    h2d_bytes = 0
    for key, a in arrays.items():
        h = {}
        for name in ("xyz", "weights", "zbin"):
            if a[name] is None:
                h[name] = None
                continue
            buf = eng.pinned_empty(a[name].shape, a[name].dtype)
            buf[...] = a[name]
            h[name] = buf
            h2d_bytes += buf.nbytes
        h["patch_off"], h["n_bins"] = a["patch_off"], a["n_bins"]
        host[key] = h

    def upload_all():
        # big catalogs first: the pair counts that need them (RR, RD) then overlap with the remaining copies
        return {k: eng.upload_catalog(host[k]["xyz"], host[k]["patch_off"], weights=host[k]["weights"],
                                      zbin=host[k]["zbin"], n_bins=host[k]["n_bins"])
                for k in ("ref_rand", "unk_rand", "unk", "ref")}

    n_out = len(pi) * plan.n_bins * (plan.n_edges - 1)
    d2h_bytes = 4 * len(opi) * plan.n_bins * (plan.n_edges - 1) * 16

    def reduce_results(results):
        if world == 1:
            return results
        import torch

        # weak: every field owns one slab of the result tensor; strong: every rank owns its patch pairs
        slabs = world if weak else 1
        full = np.zeros((slabs, 4, n_pairs_max, plan.n_bins, plan.n_edges - 1), dtype=np.int64)
        for t, tag in enumerate(COUNT_TYPES):
            full[rank if weak else 0, t, own] = results[tag]
        ten = torch.from_numpy(full).cuda()
        dist.reduce(ten, dst=0, op=dist.ReduceOp.SUM)  # the single NCCL reduce of the count tensors
        torch.cuda.synchronize()
        return {tag: ten[:, t].cpu().numpy() for t, tag in enumerate(COUNT_TYPES)}

    def count_all(dev):
        results, stats = {}, {}
        # order follows the uploads (ref_rand, unk_rand, unk, ref): RR and RD run while the remaining
        # catalogs are still on their way through PCIe
        for tag in ("RR", "RD", "DR", "DD"):
            a, b = COUNT_TYPES[tag]
            ci, _, st = eng.count(dev[a], dev[b], opi, opj, plan.r2)
            results[tag], stats[tag] = ci, st
        return results, stats

    def barrier():
        if dist is not None:
            dist.barrier()
        eng.sync()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: raw inputs resident in HBM -> index build + counts ----
    dev = upload_all()
    step_ms, kernel_ms, index_ms, stats_last, launches = [], [], [], None, 0
    clocks = ClockSampler(local_rank)
    results = None
    clocks.__enter__()  # sampled over both timed regions (value and e2e)
    for step in range(args.warmup + args.steps):
        for d in dev.values():
            d.drop_index()
        barrier()
        eng.timer_start()
        results, stats = count_all(dev)  # builds the dropped indexes on first use, inside the timed region
        t_idx = sum(s["index_ms"] for s in stats.values())
        ms = eng.timer_stop()
        results = reduce_results(results)
        barrier()
        if rank == 0:
            log(f"[bench] step {step}: {ms:.2f} ms (index {t_idx:.2f}, kernels "
                f"{sum(s['kernel_ms'] for s in stats.values()):.2f})")
        if step >= args.warmup:
            step_ms.append(max_over_ranks(ms))
            kernel_ms.append(sum(s["kernel_ms"] for s in stats.values()))
            index_ms.append(t_idx)
            stats_last = stats
            launches += sum(s["launches"] for s in stats.values())
    for d in dev.values():
        d.free()

    # ---- e2e: host buffers through the C ABI ----
    e2e_s = []
    e2e_warm = max(1, args.warmup)
    for step in range(e2e_warm + args.steps):
        barrier()
        t0 = time.perf_counter()
        # the package's schedule for host-resident inputs: every copy enqueued up front, counts issued in
        # arrival order; with --e2e-groups > 1 the unbinned catalogs travel and are counted in patch slices
        results_e2e, _, _, devs = count_cross_pipelined(eng, host, opi, opj, plan.r2, groups=args.e2e_groups)
        dev = {f"{k}{n}": d for k, lst in devs.items() for n, (d, _, _) in enumerate(lst)}
        results_e2e = reduce_results(results_e2e)
        barrier()
        dt = time.perf_counter() - t0
        for d in dev.values():
            d.free()
        if rank == 0:
            log(f"[bench] e2e step {step}: {dt * 1e3:.1f} ms")
            if step == 0:
                for tag in COUNT_TYPES:
                    assert np.array_equal(results_e2e[tag], results[tag]), f"e2e result of {tag} differs"
        if step >= e2e_warm:
            e2e_s.append(max_over_ranks(dt))
    clocks.__exit__()

    # how much of an e2e step is the PCIe copy alone (uploads without any count, synchronised)
    h2d_only = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        dev = upload_all()
        eng.sync()
        h2d_only.append(time.perf_counter() - t0)
        for d in dev.values():
            d.free()
    h2d_only_ms = 1e3 * min(h2d_only)

    # statistics of the last timed step, summed over ranks
    tot_stats = np.array([sum(s[k] for s in stats_last.values()) for k in ("pair_tests", "pair_tests_naive", "rechecks")],
                         dtype=np.float64)
    if dist is not None:
        import torch

        t = torch.from_numpy(tot_stats).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot_stats = t.cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- report ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    sms = eng.num_sms
    peak_tests = sms * 128 * sm_max_mhz * 1e6 / FP32_INSTR_PER_TEST
    t_step = float(np.mean(step_ms)) / 1e3
    t_kernel = float(np.mean(kernel_ms)) / 1e3
    executed = float(tot_stats[0])
    achieved = executed / max(t_kernel, 1e-12) / world  # per GPU: rank 0's kernel time, 1/world of the tests
    in_scale = {tag: int(results[tag].sum()) for tag in COUNT_TYPES}
    clk = clocks.summary()

    line = dict(
        metric="crosscorrelate_effective_pair_tests_per_s", value=total_naive / t_step / 1e9, unit="Gpairs/s",
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=t_step * 1e3, higher_is_better=True,
        scaling="weak" if (weak or world == 1) else "strong", vs_baseline=None, dtype="f32+f64", data="synthetic",
        config=dict(
            workload=f"{wl['name']} crosscorrelate DD+DR+RD+RR: {wl['spec']['n']} rows x scale {args.scale}, "
                     f"{wl['n_patch']} patches, {wl['spec']['bins']} z-bins, {wl['spec'].get('scales', '100-1000 kpc')}, BoxRandoms {BOX}"
                     + (f"; x {world} independent fields (sky area and rows grow with the GPU count, one field per GPU)"
                        if weak else ""),
            linked_patch_pairs=int(len(pi)), naive_pair_tests=wl["naive"], pairs_in_scale=in_scale,
            l2="inputs (>= 1 GB of catalog rows) exceed the 126 MB L2; every step rebuilds the index from the raw rows",
            parallelism=(f"{world} fields of {wl['n_patch']} patches, one per GPU (no patch links between fields), one NCCL "
                         f"reduce of the count tensor" if weak else
                         f"second-catalog patches dealt to {world} GPU(s) as compact equal-cost groups, each rank holds only "
                         f"the rows of its patches and of the first-catalog patches linked to them, one NCCL reduce of the counts"),
        ),
        breakdown_ms=dict(index_build=float(np.mean(index_ms)), count_kernels=t_kernel * 1e3,
                          per_count={tag: stats_last[tag]["kernel_ms"] for tag in COUNT_TYPES},
                          work_items={tag: int(stats_last[tag]["work_items"]) for tag in COUNT_TYPES},
                          executed_tests={tag: int(stats_last[tag]["pair_tests"]) for tag in COUNT_TYPES},
                          host_prep_s=wl["t_host_prep"], catalogs_s=wl["t_catalogs"]),
        roofline=dict(
            bound="fp32", achieved=achieved / 1e9, peak=peak_tests / 1e9, unit="Gtests/s",
            frac=achieved / peak_tests,
            # dram__bytes_read.sum + dram__bytes_write.sum of the dominant (RR) launch, one `ncu --set full`
            # capture of this workload (profiles/r01_h_ncu_full_k_count_uni.txt); not an HBM-bound kernel
            traffic=1.0075e9 if (args.workload == "C3" and args.scale == 1.0 and (world == 1 or weak)) else None,
            traffic_unit="bytes per RR launch (algorithmic minimum 4.8e8: 1e7 tile rows + 1e7 candidate rows, 24 B each)",
            note=f"pair-count kernels of rank 0; executed (non-pruned) tests / kernel time vs {sms} SMs x 128 lanes x "
                 f"{sm_max_mhz:.0f} MHz / {FP32_INSTR_PER_TEST} FP32 instr per test (MEASURED_PEAKS.json sm_max_mhz)",
            executed_pair_tests=int(executed), prune_efficiency=1.0 - executed / max(float(tot_stats[1]), 1.0),
            useful_fraction=sum(in_scale.values()) / max(executed, 1),
            fp64_rechecks=int(tot_stats[2]),
            per_launch_frac={tag: (stats_last[tag]["pair_tests"] / max(stats_last[tag]["kernel_ms"], 1e-9) * 1e3) / peak_tests
                             for tag in COUNT_TYPES},
        ),
        e2e=dict(value=total_naive / float(np.mean(e2e_s)) / 1e9, unit="Gpairs/s", ms_per_step=float(np.mean(e2e_s)) * 1e3,
                 h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes),
                 h2d_only_ms=h2d_only_ms, h2d_only_gb_per_s=h2d_bytes / h2d_only_ms / 1e6),
        gpu_launches=int(launches),
        clocks=clk,
    )
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_sample(wl, args.cpu_budget)
        t_cpu = cpu["t_build"] + cpu["t_count"]
        line["cpu_baseline"] = dict(
            value=cpu["naive"] / t_cpu / 1e9, unit="Gpairs/s", cores=cpu["workers"], kind="port",
            sample=f"{cpu['patches']}/{wl['n_patch']} first-catalog patches with all links ({cpu['pairs']} patch pairs "
                   f"x 4 count types): tree build {cpu['t_build']:.2f}s + count {cpu['t_count']:.2f}s",
            extrapolated_full_job_s=t_cpu * total_naive / max(cpu["naive"], 1),
        )
    emit(line)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=list(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every catalog (development only)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=1,
                    help="patch slices per unbinned catalog in the end-to-end schedule (1 = whole catalogs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one C3-sized field per GPU (default), strong = the one field split over the GPUs")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
