mkdir -p gpurun_out
python -m pytest tests/test_interp_gpu.py -x -q -k "rowstage" > gpurun_out/t_interp.log 2>&1; tail -4 gpurun_out/t_interp.log
V=";rs=1;rs=2;rs=3"
timeout 300 python scripts/interp_lab.py --variants "$V" > gpurun_out/lab6_c2.jsonl 2> gpurun_out/lab6_c2.err; tail -3 gpurun_out/lab6_c2.err
timeout 300 python scripts/interp_lab.py --snapshots 2000 --variants "$V" > gpurun_out/lab6_c2_t2000.jsonl 2> gpurun_out/lab6_c2_t2000.err; tail -3 gpurun_out/lab6_c2_t2000.err
