#!/bin/bash
for staging in 0 1; do for cols in 128 256; do for kb in 40 56 100; do
  python bench.py --steps 20 --warmup 3 --kernel staged --staging $staging --chunk-cols $cols --stage-kb $kb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('staged staging $staging cols $cols kb $kb', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done; done; done
