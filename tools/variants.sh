#!/bin/bash
# Build/layout experiments: bench.py under several settings; prints the step time and the walkRegions time of each.
# Usage (on the GPU box): bash tools/variants.sh "<lib suffix or ->:<OHMB200_TILE or ->" ...   e.g. "-:-" "320x3:-" "-:1,5"
for v in "$@"; do
  lib="${v%%:*}"; tile="${v##*:}"
  env_lib=""; [ "$lib" != "-" ] && env_lib="$PWD/ohm_b200/libohmb200_$lib.so"
  env_tile=""; [ "$tile" != "-" ] && env_tile="$tile"
  OHMB200_LIB="$env_lib" OHMB200_TILE="$env_tile" python bench.py --steps 20 --warmup 3 --cpu-reps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$v', 'step_ms', round(d['ms_per_step'],4), 'walk_ms', round(d['kernels_ms_per_step']['walkRegions'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],3))"
done
