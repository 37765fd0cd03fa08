mkdir -p gpurun_out
python -m pytest tests/test_gridgen_gpu.py tests/test_edge_cases_gpu.py tests/test_user_journey_gpu.py -x -q > gpurun_out/t_grid.log 2>&1; tail -6 gpurun_out/t_grid.log
python scripts/gridgen_modes.py > gpurun_out/gridgen_modes.log 2>&1; tail -4 gpurun_out/gridgen_modes.log
python scripts/run_config.py C4 --grid-only > gpurun_out/c4_grid.json 2> gpurun_out/c4_grid.err; cat gpurun_out/c4_grid.json; tail -2 gpurun_out/c4_grid.err
