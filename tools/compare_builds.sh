#!/bin/bash
# usage: compare_builds.sh [-a "extra bench args"] name...   (builds in build/variants/lib_<name>.so)
EXTRA=""
if [ "$1" = "-a" ]; then EXTRA="$2"; shift 2; fi
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $EXTRA"
for v in "$@"; do
  echo "== $v $EXTRA: $(SEADUCK_B200_LIB=$PWD/build/variants/lib_$v.so $B 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d["value"]/1e9,3), "Gc/s ms/step", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3))')"
done
