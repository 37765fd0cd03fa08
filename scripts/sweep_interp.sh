#!/bin/bash
# sweep of the warp-per-cell interpolation kernel knobs on the bench workload (C2); prints ms/step and roofline fraction
for regs in 0 1; do for warps in 4 8; do for unroll in 1 2; do
  python bench.py --steps 20 --warmup 3 --regs $regs --warps $warps --unroll $unroll 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('regs $regs warps $warps unroll $unroll', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done; done; done
