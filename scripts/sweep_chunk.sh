#!/bin/bash
# column-chunk size of the warp-per-cell kernel (s3_set_tuning key 9) on C2 (bench) and C4 (full size)
timeout 300 python -m pytest tests/test_interp_gpu.py -x -q 2>&1 | tail -2
for c in 100000 512 1024 2048; do python bench.py --steps 20 --warmup 3 --tune 9=$c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 chunk $c', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; done
for c in 100000 256 512 1024; do timeout 500 python scripts/run_config.py C4 --tune 9=$c 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 T=2000 chunk $c', round(d['interp_ms'],3), round(d['roofline_frac_of_measured'],4))"; done
