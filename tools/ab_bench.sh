#!/bin/bash
# A/B of library builds through bench.py (L2 flushed between frames): tools/ab_bench.sh cur tools/ab/x.so ...
for rnd in 1 2; do
for lib in "$@"; do
  if [ "$lib" = cur ]; then unset SWGL_B200_LIB; else export SWGL_B200_LIB=$PWD/$lib; fi
  timeout 120 python bench.py --no-cpu-baseline --no-configs --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step']*1e3,1), {k: round(v,1) for k,v in d['roofline']['stage_us'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3))"
done; done
