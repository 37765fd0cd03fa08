mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -6 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench10_n1.json 2> gpurun_out/bench10_n1.err; tail -c 300 gpurun_out/bench10_n1.err
python scripts/gridgen_modes.py > gpurun_out/gridgen_modes.log 2>&1; tail -5 gpurun_out/gridgen_modes.log
