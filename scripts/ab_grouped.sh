#!/bin/bash
# A/B of the grouped interpolation kernel's variants on the bench workload (C2): usage ab_grouped.sh <tune> ...
python bench.py --steps 30 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 direct', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
for t in "$@"; do
python bench.py --steps 30 --warmup 3 --kernel grouped --tune $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 grouped $t', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
