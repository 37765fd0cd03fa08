mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench12_n1.json 2> gpurun_out/bench12_n1.err; tail -c 300 gpurun_out/bench12_n1.err
python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/bench12_n1_k200.json 2> gpurun_out/bench12_n1_k200.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:interp_warpcell_kernel --launch-skip 6 -c 2 -o gpurun_out/r2_interp_full python bench.py --steps 2 --warmup 3 --no-cpu --no-svd > gpurun_out/ncu_full.out 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/ref_gridgen.py --c2 > gpurun_out/ref_gridgen.json 2> gpurun_out/ref_gridgen.err; tail -c 400 gpurun_out/ref_gridgen.json
