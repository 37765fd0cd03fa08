#!/bin/bash
# ncu --set full of the interpolation kernels on short rows: usage ncu_short.sh T "variants" [--k26]
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:interp_ -o gpurun_out/short_rows -f \
  python scripts/interp_lab.py --snapshots ${1:-125} --layouts pitched --steps 1 --warmup 1 --variants "${2:-;8=4,15=0,7=4}" $3 > gpurun_out/ncu_short.log 2>&1
tail -3 gpurun_out/ncu_short.log
ncu -i gpurun_out/short_rows.ncu-rep --page raw --csv > gpurun_out/short_rows_raw.csv
ls -la gpurun_out/short_rows*
