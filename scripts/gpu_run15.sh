mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:interp_ --launch-skip 6 -c 2 -o gpurun_out/r2_interp_t125 python scripts/interp_lab.py --snapshots 125 --layouts pitched --variants ";" --steps 2 --warmup 1 > gpurun_out/ncu_t125.out 2>&1
ls -la gpurun_out/r2_interp_t125.ncu-rep
