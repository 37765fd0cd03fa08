# Round-2 evidence run on one B200 (gpurun -- 'bash scripts/final_evidence.sh'); outputs under gpurun_out/, copied to profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 200 gpurun_out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_timed_region.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-svd > gpurun_out/ncu_launch.out 2>&1
for C in C3 C4 C5; do python scripts/run_config.py $C > gpurun_out/r2_cfg_$C.json 2> gpurun_out/r2_cfg_$C.err; tail -c 300 gpurun_out/r2_cfg_$C.json; done
ncu --set full --clock-control none -k regex:cells_gain_kernel -s 7 -c 1 -o gpurun_out/r2_gain_c4 python scripts/run_config.py C4 --grid-only > gpurun_out/ncu_gain.out 2>&1
ls -la gpurun_out/r2_gain_c4.ncu-rep
