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


def subset_patches(a: dict, keep: np.ndarray, fractions: dict | None = None, reach: list | None = None) -> dict:
    """host arrays of one catalog restricted to the patches in `keep` (all others become empty); `fractions`:
    patch -> (f0, f1), the share of a patch that is split between two ranks (`sharding.split_rows`); `reach`:
    list of (centre, chord radius) -- only rows inside one of these caps are kept"""
    from yet_another_wizz_b200.sharding import split_rows

    off = a["patch_off"]
    sizes = np.zeros(len(off) - 1, dtype=np.int64)
    parts = []
    for p in keep:
        rows = np.arange(off[p], off[p + 1])
        if fractions is not None and p in fractions and fractions[p] != (0.0, 1.0):
            rows = rows[split_rows(a["xyz"][off[p]:off[p + 1]], *fractions[p])]
        if reach is not None and len(rows):
            x = a["xyz"][rows]
            near = np.zeros(len(rows), dtype=bool)
            for c, r in reach:
                near |= ((x - c) ** 2).sum(axis=1) <= r * r
            rows = rows[near]
        sizes[p] = len(rows)
        parts.append(rows)
    sel = np.concatenate(parts) if parts else np.empty(0, dtype=np.int64)
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
def cpu_sample(wl, budget_s: float, workers: int | None = None, fixed_m: int | None = None, tags=None):
    """Time the CPU port on a bounded sample: the first `m` patches' diagonal pairs plus all their links, the
    count types in `tags`, trees built (on the worker pool) for exactly the patches touched.  Returns the
    per-patch-pair counts as well, so that the caller can check the GPU result against them."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_port

    workers = workers or cpu_port.physical_cores()
    tags = list(tags or COUNT_TYPES)
    pi, pj = wl["pair_i"], wl["pair_j"]
    n_bins = len(wl["config"].binning.binning)
    plan = wl["plan"]
    ang_min = np.array([lim[:, 0] for lim in plan.limits])
    ang_max = np.array([lim[:, 1] for lim in plan.limits])
    scales = wl["config"].scales

    def rows_of(key, patches, binned):
        a = wl["arrays"][key]
        out = {}
        for p in patches:
            s, e = a["patch_off"][p], a["patch_off"][p + 1]
            zb = a["zbin"][s:e].astype(np.int32) if binned else None
            out[p] = (a["xyz"][s:e], None if a["weights"] is None else a["weights"][s:e], zb)
        return out

    def run(m):
        first = set(range(m))
        sel = [(int(i), int(j)) for i, j in zip(pi, pj) if int(i) in first]
        need1 = sorted({i for i, _ in sel})
        need2 = sorted({j for _, j in sel})
        t_build = t_count = 0.0
        naive = 0
        trees = {}
        used = {k for t in tags for k in COUNT_TYPES[t]}
        for key, binned, need in (("ref", True, need1), ("ref_rand", True, need1), ("unk", False, need2),
                                  ("unk_rand", False, need2)):
            if key not in used:
                continue
            rows = rows_of(key, need, binned)
            built, dt = cpu_port.build_catalog_trees([rows[p] for p in need], n_bins if binned else None, workers=workers)
            trees[key] = dict(zip(need, built))
            t_build += dt
        counts = {}
        for tag in tags:
            a, b = COUNT_TYPES[tag]
            counts[tag], dt = cpu_port.count_pairs(trees[a], trees[b], sel, ang_min, ang_max, rweight=scales.rweight,
                                                   resolution=scales.resolution, workers=workers)
            t_count += dt
            ca, cb = wl["arrays"][a], wl["arrays"][b]
            for i, j in sel:
                zb = ca["zbin"][ca["patch_off"][i]:ca["patch_off"][i + 1]]
                n1 = int(((zb >= 0) & (zb < n_bins)).sum())
                naive += n1 * int(cb["patch_off"][j + 1] - cb["patch_off"][j])
        return dict(patches=m, pairs=len(sel), t_build=t_build, t_count=t_count, naive=naive, workers=workers, counts=counts)

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


def check_parity(wl, results: dict, cpu: dict) -> int:
    """GPU result of the full job against the CPU port's per-patch-pair counts (scipy cKDTree, the reference's
    arithmetic): bit-exact for unweighted catalogs, 1e-12 relative with weights.  Raises on any difference;
    returns the number of (count type, patch pair) results compared."""
    plan = wl["plan"]
    index = {(int(i), int(j)): k for k, (i, j) in enumerate(zip(wl["pair_i"], wl["pair_j"]))}
    checked = 0
    for tag, per_pair in cpu["counts"].items():
        got = plan.finish(results[tag])  # (n_scales, n_pairs, n_bins)
        a, b = COUNT_TYPES[tag]
        weighted = wl["arrays"][a]["weights"] is not None or wl["arrays"][b]["weights"] is not None
        exact = not weighted and wl["config"].scales.rweight is None
        for (i, j), want in per_pair.items():
            mine = got[:, index[(i, j)], :]
            ok = np.array_equal(mine, want) if exact else np.allclose(mine, want, rtol=1e-12, atol=0.0)
            if not ok:
                raise AssertionError(f"parity failure: {tag} counts of patch pair ({i}, {j}) differ from the CPU reference "
                                     f"algorithm: max |diff| = {np.abs(mine - want).max()}")
            checked += 1
    return checked


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  With the
    reference installed under `baseline/_ref` (`__graft_entry__.build()`), that is the UNMODIFIED package run
    through its own `Catalog.build_trees` / `PatchLinkage.count_pairs` (`oracle/ref_runner.py`, a separate
    process: it forks worker pools) on a bounded sample of the workload -- the first declination stripes of the
    patch grid at the full densities; otherwise the oracle's port of the same algorithm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    plain = "scales" not in spec and not spec.get("weighted")
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "yaw"))
    if have_ref and plain:
        runner = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--workload", args.workload,
                  "--scale", str(args.scale)]

        def call(stripes, steps, warmup):
            out = subprocess.run(runner + ["--stripes", str(stripes), "--steps", str(steps), "--warmup", str(warmup)],
                                 capture_output=True, text=True)
            if out.returncode != 0:
                raise RuntimeError(out.stderr[-2000:])
            return json.loads(out.stdout.strip().splitlines()[-1])

        try:
            probe = call(1, 1, 0)  # one stripe, one step: what a stripe costs on this box
            ny = probe["stripes_total"]
            t1 = probe["steps"][0]["build_s"] + probe["steps"][0]["count_s"]
            total_budget = max(args.cpu_budget, 10.0) * 12.0  # the whole --steps/--warmup run: a few minutes
            stripes = int(max(1, min(ny, total_budget / max(args.steps + args.warmup, 1) / max(t1, 1e-3))))
            res = call(stripes, args.steps, args.warmup)
            times = [s_["build_s"] + s_["count_s"] for s_ in res["steps"]]
            t = float(np.mean(times))
            naive = sum(res["naive_pair_tests"].values())
            value = naive / t / 1e9
            full = call(ny, 1, 0) if args.full_reference and stripes < ny else (res if stripes == ny else None)
            cb = dict(value=value, unit="Gpairs/s", cores=res["workers"], kind="reference",
                      sample=f"unmodified yaw {res['version']} from baseline/_ref: {res['patches']}/{res['patches_total']} patches "
                             f"({stripes} of {ny} declination stripes at full density, {res['linked_pairs']} linked patch pairs x 4 "
                             f"count types), build_trees {np.mean([s_['build_s'] for s_ in res['steps']]):.2f}s + count_pairs "
                             f"{np.mean([s_['count_s'] for s_ in res['steps']]):.2f}s per step",
                      host_cores=res["host_cores"], rows=res["rows"])
            if full is not None:
                tf = full["steps"][0]["build_s"] + full["steps"][0]["count_s"]
                cb["full_job_s"] = tf
                cb["full_job_value"] = sum(full["naive_pair_tests"].values()) / tf / 1e9
                cb["full_job_pairs_in_scale"] = full["pairs_in_scale"]
            emit(dict(
                impl="reference", metric="crosscorrelate_effective_pair_tests_per_s", value=value, unit="Gpairs/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"{args.workload} crosscorrelate DD+DR+RD+RR, BoxRandoms, scale={args.scale}", sample=cb["sample"]),
                cpu_baseline=cb, e2e=dict(value=value, unit="Gpairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0))
            return
        except Exception as exc:  # fall back to the port, and say so
            log(f"[bench] reference runner failed ({exc}); falling back to the oracle's port")
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
    from yet_another_wizz_b200.pipeline import count_cross_pipelined
    from yet_another_wizz_b200.sharding import assign_patch_fractions, pair_costs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = torch = None
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
        t = torch.tensor([float(total_naive), float(len(pi))], dtype=torch.float64, device="cuda")
        tsum, tmax = t.clone(), t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_naive = int(tsum[0].item())
        n_pairs_max = int(tmax[1].item())
    else:
        n_pairs_max = len(pi)

    # this rank's share of a strong-scaling run: second-catalog patches are dealt out as spatially compact groups
    # of equal summed pair cost (`sharding.assign_patches_contiguous`), a rank counts every linked pair of its
    # patches and holds only the rows it needs (its own second-catalog patches, plus the first-catalog patches
    # linked to them)
    own = np.arange(len(pi))
    arrays = wl["arrays"]
    part_rank, part_world = rank, world
    if args.emulate_share:  # development: one GPU runs rank r's share of a w-GPU strong-scaling run ("r/w")
        part_rank, part_world = (int(v) for v in args.emulate_share.split("/"))
    if part_world > 1 and not weak:
        n1 = np.diff(arrays["ref_rand"]["patch_off"])
        n2 = np.diff(arrays["unk_rand"]["patch_off"])
        costs = pair_costs(pi, pj, n1, n2, radii1=np.asarray(wl["cats"]["ref_rand"].get_radii().data))
        patch_cost = np.bincount(pj, weights=costs, minlength=wl["n_patch"])
        centers_xyz = wl["cats"]["unk_rand"].get_centers().to_3d()
        share = assign_patch_fractions(patch_cost, centers_xyz, part_world)[part_rank]
        fractions = {p: (f0, f1) for p, f0, f1 in share}
        own = np.flatnonzero(np.isin(pj, np.array(sorted(fractions), dtype=np.int64)))
        need = dict(second=np.unique(pj[own]), first=np.unique(pi[own]))
        second = {k: subset_patches(arrays[k], need["second"], fractions) for k in ("unk", "unk_rand")}
        # of a first-catalog patch only the rows within reach of this rank's second-catalog rows can form a pair:
        # chord to the centre of an own patch <= its radius + the largest search radius (triangle inequality)
        reach = []
        rmax_chord = float(np.sqrt(np.max(plan.r2)))
        for p in need["second"]:
            rows = np.concatenate([a["xyz"][a["patch_off"][p]:a["patch_off"][p + 1]] for a in second.values()])
            if len(rows):
                c = rows.sum(axis=0)
                c /= np.linalg.norm(c)
                reach.append((c, (np.sqrt(((rows - c) ** 2).sum(axis=1).max()) + rmax_chord) * (1.0 + 1e-9) + 1e-12))
        first = {k: subset_patches(arrays[k], need["first"], reach=reach) for k in ("ref", "ref_rand")}
        arrays = {**first, **second}
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
        return {k: eng.upload_catalog(host[k]["xyz"], host[k]["patch_off"], weights=host[k]["weights"],
                                      zbin=host[k]["zbin"], n_bins=host[k]["n_bins"])
                for k in ("ref", "ref_rand", "unk", "unk_rand")}

    n_sub = plan.n_edges - 1
    shape_own = (len(opi), plan.n_bins, n_sub)
    weighted_tag = {tag: arrays[a]["weights"] is not None or arrays[b]["weights"] is not None for tag, (a, b) in COUNT_TYPES.items()}
    any_weighted = any(weighted_tag.values())
    # the reference sample and its randoms are counted in one pass against each unbinned catalog (yawb_count2)
    # unless only one of them carries weights
    fuse = (arrays["ref"]["weights"] is None) == (arrays["ref_rand"]["weights"] is None) and not args.no_fuse
    # ... and both passes in one launch (yawb_count4) when the two unbinned catalogs are of the same kind as well
    fuse4 = fuse and not args.no_fuse4 and (arrays["unk"]["weights"] is None) == (arrays["unk_rand"]["weights"] is None)
    d2h_bytes = 4 * len(opi) * plan.n_bins * n_sub * 16

    # multi-GPU: every rank counts into device buffers, ONE NCCL sum-reduce of the (4, n_pairs, n_bins, n_sub)
    # tensor to rank 0 -- inside the timed region
    red = None
    if dist is not None:
        slabs = world if weak else 1
        red = dict(full=torch.zeros((slabs, 4, n_pairs_max, plan.n_bins, n_sub), dtype=torch.float64 if any_weighted else torch.int64,
                                    device="cuda"),
                   own=torch.from_numpy(own).cuda(),
                   stack_f=torch.zeros((4,) + shape_own, dtype=torch.float64, device="cuda"),
                   stack_i=torch.zeros((4,) + shape_own, dtype=torch.int64, device="cuda"))
        # per count type (sums, counts): views of the two stacked tensors, so that ONE index_copy_ moves all four
        red["buf"] = {tag: (red["stack_f"][t], red["stack_i"][t]) for t, tag in enumerate(COUNT_TYPES)}

    def count_all(dev, on_device: bool):
        results, stats = {}, {}

        def pick(tag, ci, cf):
            return cf if weighted_tag[tag] else ci

        if fuse and fuse4:
            # all four counts in ONE launch (yawb_count4): (ref, ref_rand) x (unk, unk_rand) -> DD, RD, DR, RR
            tags4 = ("DD", "RD", "DR", "RR")
            if on_device:
                ptrs = tuple(red["buf"][tg][k].data_ptr() if (k == 0) == weighted_tag[tg] else 0 for tg in tags4 for k in (0, 1))
                _, st = eng.count4(dev["ref"], dev["ref_rand"], dev["unk"], dev["unk_rand"], opi, opj, plan.r2, out_device=ptrs)
            else:
                out4, st = eng.count4(dev["ref"], dev["ref_rand"], dev["unk"], dev["unk_rand"], opi, opj, plan.r2)
                for tg, (ci, cf) in zip(tags4, out4):
                    results[tg] = pick(tg, ci, cf)
            stats["DD+RD+DR+RR"] = st
        elif fuse:
            for (ta, tb), k2 in ((("DD", "RD"), "unk"), (("DR", "RR"), "unk_rand")):
                if on_device:
                    # only the array the reduce takes: sums of weighted catalogs, counts otherwise
                    ptrs = tuple(red["buf"][tg][k].data_ptr() if (k == 0) == weighted_tag[tg] else 0 for tg in (ta, tb) for k in (0, 1))
                    _, _, st = eng.count2(dev["ref"], dev["ref_rand"], dev[k2], opi, opj, plan.r2, out_device=ptrs)
                else:
                    (ia, fa), (ib, fb), st = eng.count2(dev["ref"], dev["ref_rand"], dev[k2], opi, opj, plan.r2)
                    results[ta], results[tb] = pick(ta, ia, fa), pick(tb, ib, fb)
                stats[f"{ta}+{tb}"] = st
        else:
            for tag in ("DD", "RD", "DR", "RR"):
                a, b = COUNT_TYPES[tag]
                if on_device:
                    st = eng.count_into_device(dev[a], dev[b], opi, opj, plan.r2, red["buf"][tag][0].data_ptr() if weighted_tag[tag] else 0,
                                               0 if weighted_tag[tag] else red["buf"][tag][1].data_ptr())
                else:
                    ci, cf, st = eng.count(dev[a], dev[b], opi, opj, plan.r2)
                    results[tag] = pick(tag, ci, cf)
                stats[tag] = st
        return results, stats

    def reduce_device():
        """scatter this rank's rows into the job-wide tensor and sum-reduce it to rank 0 (NCCL over NVLink)"""
        full = red["full"]
        full.zero_()
        if all(weighted_tag.values()) or not any_weighted:  # one scatter for the four count types
            full[rank if weak else 0].index_copy_(1, red["own"], red["stack_f"] if any_weighted else red["stack_i"])
        else:
            for t, tag in enumerate(COUNT_TYPES):
                src = red["buf"][tag][0 if weighted_tag[tag] else 1]
                full[rank if weak else 0, t].index_copy_(0, red["own"], src.to(full.dtype))
        dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
        return full

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        eng.sync()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: raw inputs resident in HBM -> index build + counts (+ the NCCL reduce) ----
    dev = upload_all()
    step_ms, kernel_ms, count_only_ms, index_ms, stats_last, launches = [], [], [], [], None, 0
    clocks = ClockSampler(local_rank)
    results = None
    clocks.__enter__()  # sampled over both timed regions (value and e2e)
    for step in range(args.warmup + args.steps):
        for d in dev.values():
            d.drop_index()
        barrier()
        if dist is None:
            eng.timer_start()
            results, stats = count_all(dev, False)  # builds the dropped indexes on first use, inside the timed region
            ms = eng.timer_stop()
        else:
            # device timeline of this rank: the counts run on the engine's stream (each call returns once its
            # results are complete), the reduce on torch's; both lie between the two events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, stats = count_all(dev, True)
            full = reduce_device()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
        t_idx = sum(s["index_ms"] for s in stats.values())
        barrier()
        if rank == 0:
            log(f"[bench] step {step}: {ms:.2f} ms (index {t_idx:.2f}, kernels "
                f"{sum(s['kernel_ms'] for s in stats.values()):.2f})")
        if step >= args.warmup:
            step_ms.append(max_over_ranks(ms))
            kernel_ms.append(sum(s["kernel_ms"] for s in stats.values()))
            count_only_ms.append(sum(s["kernel_ms"] - s.get("plan_ms", 0.0) for s in stats.values()))
            index_ms.append(t_idx)
            stats_last = stats
            launches += sum(s["launches"] for s in stats.values())
    if dist is not None:
        full = full.cpu().numpy()
        results = {tag: (full[:, t] if weak else full[0, t]) for t, tag in enumerate(COUNT_TYPES)}
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
        ci_e2e, cf_e2e, _, devs = count_cross_pipelined(eng, host, opi, opj, plan.r2, groups=args.e2e_groups, fuse=fuse)
        results_e2e = {tag: (cf_e2e if weighted_tag[tag] else ci_e2e)[tag] for tag in COUNT_TYPES}
        if dist is not None:
            for t, tag in enumerate(COUNT_TYPES):
                red["buf"][tag][0 if weighted_tag[tag] else 1].copy_(torch.from_numpy(results_e2e[tag]))
            full = reduce_device()
            torch.cuda.synchronize()
        dev = {f"{k}{n}": d for k, lst in devs.items() for n, (d, _, _) in enumerate(lst)}
        barrier()
        dt = time.perf_counter() - t0
        for d in dev.values():
            d.free()
        if rank == 0:
            log(f"[bench] e2e step {step}: {dt * 1e3:.1f} ms")
            if step == 0:
                if dist is not None:
                    full = full.cpu().numpy()
                    results_e2e = {tag: (full[:, t] if weak else full[0, t]) for t, tag in enumerate(COUNT_TYPES)}
                for tag in COUNT_TYPES:  # integers: identical; weighted sums: the atomics' order varies
                    same = (np.allclose(results_e2e[tag], results[tag], rtol=1e-12, atol=0.0) if weighted_tag[tag]
                            else np.array_equal(results_e2e[tag], results[tag]))
                    assert same, f"e2e result of {tag} differs"
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
    h2d_only_ms = 1e3 * max_over_ranks(min(h2d_only))

    # statistics of the last timed step, summed over ranks; kernel time: the slowest rank's
    tot_stats = np.array([sum(s[k] for s in stats_last.values()) for k in ("pair_tests", "pair_tests_naive", "rechecks", "work_items")],
                         dtype=np.float64)
    t_kernel_max = max_over_ranks(float(np.mean(count_only_ms)))  # the dominant kernel alone (k_count_stream), planner excluded
    if dist is not None:
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
    t_kernel = t_kernel_max / 1e3
    executed = float(tot_stats[0])
    achieved = executed / max(t_kernel, 1e-12) / world  # per GPU: the slowest rank's kernel time, 1/world of the tests
    in_scale = {tag: float(results[tag].sum()) for tag in COUNT_TYPES}
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
            parallelism=("1 GPU" if world == 1 else
                         f"{world} fields of {wl['n_patch']} patches, one per GPU (no patch links between fields), one NCCL "
                         f"reduce of the count tensor inside the timed region" if weak else
                         f"the ONE job split over {world} GPUs: second-catalog patches dealt as compact equal-cost groups (the patch a cut falls into is shared by two ranks), each rank "
                         f"holds only the rows of its patches and the rows of the linked first-catalog patches within reach of them; ONE NCCL sum-reduce "
                         f"of the (4, n_pairs, n_bins, n_sub) count tensor to rank 0 inside the timed region"),
            fused_counts=("one launch (yawb_count4)" if fuse and fuse4 else "two launches (yawb_count2)" if fuse else False),
        ),
        breakdown_ms=dict(index_build=float(np.mean(index_ms)), count_kernels=float(np.mean(kernel_ms)),
                          per_launch={tag: s["kernel_ms"] for tag, s in stats_last.items()},
                          work_items={tag: int(s["work_items"]) for tag, s in stats_last.items()},
                          executed_tests={tag: int(s["pair_tests"]) for tag, s in stats_last.items()},
                          host_prep_s=wl["t_host_prep"], catalogs_s=wl["t_catalogs"]),
        roofline=dict(
            bound="fp32", achieved=achieved / 1e9, peak=peak_tests / 1e9, unit="Gtests/s",
            frac=achieved / peak_tests,
            traffic=None,  # dram bytes per launch come from the ncu captures under profiles/, not from a timed run
            note=f"k_count_stream launches (slowest rank, CUDA events on the launching stream; the planner k_plan, "
                 f"{float(np.mean(kernel_ms)) - float(np.mean(count_only_ms)):.3f} ms, is part of breakdown_ms.count_kernels but not of this figure); "
                 f"executed (non-pruned) tests / kernel time vs {sms} SMs x 128 lanes x "
                 f"{sm_max_mhz:.0f} MHz / {FP32_INSTR_PER_TEST} FP32 instr per test (MEASURED_PEAKS.json sm_max_mhz)",
            executed_pair_tests=int(executed), prune_efficiency=1.0 - executed / max(float(tot_stats[1]), 1.0),
            useful_fraction=sum(in_scale.values()) / max(executed, 1),
            fp64_rechecks=int(tot_stats[2]), work_items=int(tot_stats[3]),
            per_launch_frac={tag: (s["pair_tests"] / max(s["kernel_ms"] - s.get("plan_ms", 0.0), 1e-9) * 1e3) / peak_tests
                             for tag, s in stats_last.items()},
        ),
        e2e=dict(value=total_naive / float(np.mean(e2e_s)) / 1e9, unit="Gpairs/s", ms_per_step=float(np.mean(e2e_s)) * 1e3,
                 h2d_bytes_per_step=int(h2d_bytes), d2h_bytes_per_step=int(d2h_bytes),
                 h2d_only_ms=h2d_only_ms, h2d_only_gb_per_s=h2d_bytes / h2d_only_ms / 1e6),
        gpu_launches=int(launches),
        clocks=clk,
    )
    if not args.no_cpu_baseline and not weak:
        # the CPU port of the reference's algorithm on a bounded sample of THIS job: timing baseline and, at the
        # full size of the workload, the parity check of the GPU result (bit-exact / 1e-12 with weights)
        cpu = cpu_sample(wl, args.cpu_budget)
        t_cpu = cpu["t_build"] + cpu["t_count"]
        line["cpu_baseline"] = dict(
            value=cpu["naive"] / t_cpu / 1e9, unit="Gpairs/s", cores=cpu["workers"], kind="port",
            sample=f"{cpu['patches']}/{wl['n_patch']} first-catalog patches with all links ({cpu['pairs']} patch pairs "
                   f"x 4 count types): tree build {cpu['t_build']:.2f}s + count {cpu['t_count']:.2f}s",
            extrapolated_full_job_s=t_cpu * total_naive / max(cpu["naive"], 1),
        )
        if not args.emulate_share:  # one rank's share alone holds partial counts of the patches it shares
            line["parity_checked_pairs"] = check_parity(wl, results, cpu)
            line["parity"] = ("GPU counts of the full job == CPU reference algorithm (scipy cKDTree) on every sampled patch pair, "
                              + ("1e-12 relative" if any_weighted or wl["config"].scales.rweight is not None else "bit-exact"))
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
    ap.add_argument("--no-fuse", action="store_true", help="count DD, RD, DR, RR in four passes instead of two fused ones")
    ap.add_argument("--no-fuse4", action="store_true", help="two fused launches (DD+RD, DR+RR: yawb_count2) instead of one (yawb_count4)")
    ap.add_argument("--full-reference", action="store_true", help="--impl reference: also time the full job once")
    ap.add_argument("--e2e-groups", type=int, default=1,
                    help="patch slices per unbinned catalog in the end-to-end schedule (1 = whole catalogs)")
    ap.add_argument("--emulate-share", default=None, help="development: 'r/w' = count rank r's share of a w-GPU strong-scaling run")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the ONE job split over the GPUs (default, the north-star case), "
                         "weak = one independent field of the workload's size per GPU")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
