#!/bin/bash
# dev helper: rebuild the in-tree library, then run a command on the B200 box
# usage: ./tools_gpurun.sh <logfile> <timeout_s> '<command>'
set -e
cd "$(dirname "$0")"
python yet_another_wizz_b200/csrc/build.py > /dev/null
make -C oracle > /dev/null
log="$1"; shift
to="$1"; shift
timeout $((to + 1900)) /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1 || true
echo done >> "$log"
