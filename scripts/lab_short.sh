#!/bin/bash
# short-row sweep of the interpolation kernel (time windows of a sharded export). Usage: lab_short.sh "T list" "variants"
mkdir -p gpurun_out
TS=${1:-"125 250 500 1000"}
V=${2:-";8=2;8=2,12=16;8=2,12=64;8=2,7=4;8=2,7=8;8=2,13=32;8=2,12=16,7=4"}
: > gpurun_out/lab_short.jsonl
for T in $TS; do
  python scripts/interp_lab.py --snapshots $T --layouts pitched --steps 40 --variants "$V" >> gpurun_out/lab_short.jsonl 2> gpurun_out/lab_short_$T.err || tail -5 gpurun_out/lab_short_$T.err
done
cut -c1-190 gpurun_out/lab_short.jsonl
if [ -n "$3" ]; then   # 3-D style tables (k = 26), T list in $3
  for T in $3; do
    python scripts/interp_lab.py --k26 --snapshots $T --layouts pitched --steps 20 --variants "$V" 2> gpurun_out/lab_short_k26_$T.err | tee -a gpurun_out/lab_short.jsonl | cut -c1-190 || tail -5 gpurun_out/lab_short_k26_$T.err
  done
fi
