#!/bin/bash
# A/B of interpolation kernel builds / tunings on the same box: bench (C2) and C4
# usage: ab_interp.sh <lib>[:tune] ...
for spec in "$@"; do
  lib=${spec%%:*}; tune=""; [[ "$spec" == *:* ]] && tune=${spec#*:}
  S3B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 3 ${tune:+--tune $tune} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 $spec', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
for spec in "$@"; do
  lib=${spec%%:*}; tune=""; [[ "$spec" == *:* ]] && tune=${spec#*:}
  S3B200_LIB=$PWD/$lib timeout 500 python scripts/run_config.py C4 ${tune:+--tune $tune} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 T=2000 $spec', round(d['interp_ms'],3), round(d['roofline_frac_of_measured'],4))"
done
