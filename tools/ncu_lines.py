#!/usr/bin/env python
"""Dev helper: attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
    nvdisasm -g -c yawb_count.sm_100a.cubin > count.disasm      (cubin of the SAME build)
    python tools/ncu_lines.py sass.csv count.disasm <mangled-kernel-substring> [top_n]

The two listings enumerate the kernel's instructions in the same order, so they are zipped by position.
"""
import collections
import csv
import re
import sys


def disasm_lines(path, kernel):
    out, cur_line, active = [], ("?", 0), False
    for raw in open(path):
        if raw.startswith("//---") and ".text." in raw:
            active = kernel in raw
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', raw)
        if m:
            cur_line = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", raw):
            out.append((cur_line, raw.split("*/", 1)[1].strip().rstrip(";").strip()))
    return out


def main():
    sass_csv, disasm, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr):
            data.append(r)
    dis = disasm_lines(disasm, kernel)
    print(f"ncu instructions {len(data)}, disasm instructions {len(dis)}")
    n = min(len(data), len(dis))
    inst, samp = collections.Counter(), collections.Counter()
    for (line, _), r in zip(dis[:n], data[:n]):
        inst[line] += int(r[ix["Instructions Executed"]])
        samp[line] += int(r[ix["# Samples"]])
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti:.3e}, samples {ts}")
    print("file:line   inst%  samp%  source")
    cache = {}
    for (fname, line), c in samp.most_common(top):
        if fname not in cache:
            try:
                cache[fname] = open(fname).read().split("\n")
            except OSError:
                cache[fname] = []
        src = cache[fname]
        text = src[line - 1].strip()[:90] if 0 < line <= len(src) else ""
        short = fname.split("/")[-1][:18]
        print(f"{short:>18s}:{line:<5d} {inst[(fname, line)] / ti * 100:6.2f} {c / ts * 100:6.2f}  {text}")


if __name__ == "__main__":
    main()
