#!/bin/bash
# Build/layout/switch experiments: bench.py under several settings; prints the step time, the walkRegions and
# prepSegments times and the end-to-end time of each.
# Usage (on the GPU box): bash tools/variants.sh "<lib suffix or ->:<OHMB200_TILE or ->[:<VAR=value>]" ...
#   e.g.  "-:-"  "320x3:-"  "-:1,5"  "-:-:OHMB200_PRODUCER=2"  "-:-:OHMB200_GRAPHS=0"
# <lib suffix> selects ohm_b200/libohmb200_<suffix>.so (a variant build, see ohm_b200/build.py for the flags).
for v in "$@"; do
  IFS=: read -r lib tile extra <<< "$v"
  env_lib=""; [ "$lib" != "-" ] && [ -n "$lib" ] && env_lib="$PWD/ohm_b200/libohmb200_$lib.so"
  env_tile=""; [ "$tile" != "-" ] && env_tile="$tile"
  env OHMB200_LIB="$env_lib" OHMB200_TILE="$env_tile" ${extra:+"$extra"} python bench.py --steps 20 --warmup 3 --cpu-reps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
k=d['kernels_ms_per_step']
print('$v', 'step_ms', round(d['ms_per_step'],4), 'walk_ms', round(k['walkRegions'],4), 'prep_ms', round(k.get('prepSegments',0),4), 'e2e_ms', round(d['e2e']['ms_per_step'],3))"
done
