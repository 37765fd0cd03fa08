mkdir -p gpurun_out
python -m pytest tests/test_geometry_surfaces_gpu.py -x -q 2>&1 | tail -2
python scripts/stl_bench.py > gpurun_out/stl_bench.jsonl 2> gpurun_out/stl_bench.err; cat gpurun_out/stl_bench.jsonl; tail -3 gpurun_out/stl_bench.err
M=gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.sum,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,launch__registers_per_thread,launch__grid_size
ncu --metrics $M --clock-control none -k regex:"stl_inside_kernel|cells_mask_kernel" -c 12 --csv --log-file gpurun_out/ncu_stl.csv python scripts/stl_bench.py > gpurun_out/ncu_stl.out 2>&1
python scripts/run_config.py C5 --grid-only --stl-subdiv 6 > gpurun_out/c5_grid_sub6.json 2> gpurun_out/c5_grid_sub6.err; cat gpurun_out/c5_grid_sub6.json; tail -2 gpurun_out/c5_grid_sub6.err
