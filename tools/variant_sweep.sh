#!/bin/bash
# dev helper (GPU box): bench every library variant under gpurun_variants/ plus the in-tree build
# usage: tools/variant_sweep.sh [bench args]
for lib in yet_another_wizz_b200/csrc/libyawb.so gpurun_variants/*.so; do
  name=$(basename $lib .so)
  YAWB_LIB=$PWD/$lib python bench.py --steps 5 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.log
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/sweep_{name}.json"))
    pc = d["breakdown_ms"]["per_launch"]
    print(name, "step %.2f" % d["ms_per_step"], "kernels %.2f" % d["breakdown_ms"]["count_kernels"], {k: round(v, 3) for k, v in pc.items()},
          "frac %.3f" % d["roofline"]["frac"], "e2e %.1f" % d["e2e"]["ms_per_step"])
except Exception as e:
    print(name, "failed", e)
PY
done
