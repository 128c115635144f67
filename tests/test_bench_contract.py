"""
The benchmark's output contract, checked on CPU with the reference arm (`bench.py --impl reference`, the
CPU implementation of the path on a tiny workload): exactly one line on stdout, valid JSON, the keys the
driver reads.  The GPU arm prints the same keys plus `roofline` / `gpu_launches` / `clocks`
(exercised on the GPU box by the round-end bench run).
"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--scale", "0.05",
         "--steps", "1", "--warmup", "1", "--cpu-budget", "1"],
        capture_output=True, text=True, timeout=600, cwd=ROOT,
    )
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["higher_is_better"] is True
    # "reference": the unmodified package staged under baseline/_ref by __graft_entry__.build(); "port": the oracle's
    # restatement of its algorithm (when the reference is not installed)
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "yaw"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]
