mkdir -p gpurun_out
python -m pytest tests/test_interp_gpu.py -x -q > gpurun_out/t_interp.log 2>&1; tail -3 gpurun_out/t_interp.log
python scripts/interp_lab.py > gpurun_out/lab2_c2.jsonl 2> gpurun_out/lab2_c2.err; tail -3 gpurun_out/lab2_c2.err
python scripts/interp_lab.py --k26 --snapshots 2000 --layouts pitched --variants ";3=0;6=1;8=2;8=2,3=0;8=2,2=1;8=4,2=1;8=2,7=2;2=2,7=2;6=1,7=2" > gpurun_out/lab2_k26.jsonl 2> gpurun_out/lab2_k26.err
python scripts/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; tail -20 gpurun_out/e2e_probe.log
M=gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__cycles_elapsed.avg,launch__registers_per_thread,launch__grid_size
ncu --metrics $M --clock-control none -k regex:interp_warp -c 40 --csv --log-file gpurun_out/ncu_lab2.csv python scripts/interp_lab.py --steps 1 --warmup 0 --layouts pitched,dense --variants ";7=4;2=2;6=1,3=1;8=2;8=4;8=4,4=512" > gpurun_out/ncu_lab2.out 2>&1
python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log
