mkdir -p gpurun_out
python -m pytest tests/test_interp_gpu.py -x -q > gpurun_out/t_interp.log 2>&1; tail -4 gpurun_out/t_interp.log
V=";rs=1;rs=2;9=0;9=0,6=0;9=0,6=0,2=1"
timeout 300 python scripts/interp_lab.py --variants "$V" > gpurun_out/lab5_c2.jsonl 2> gpurun_out/lab5_c2.err; tail -3 gpurun_out/lab5_c2.err
timeout 300 python scripts/interp_lab.py --snapshots 2000 --variants "$V" > gpurun_out/lab5_c2_t2000.jsonl 2> gpurun_out/lab5_c2_t2000.err; tail -3 gpurun_out/lab5_c2_t2000.err
V26=";rs=2;6=0"
timeout 300 python scripts/interp_lab.py --k26 --snapshots 2000 --layouts pitched,dense --variants "$V26" > gpurun_out/lab5_k26.jsonl 2> gpurun_out/lab5_k26.err; tail -3 gpurun_out/lab5_k26.err
