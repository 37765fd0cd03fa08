mkdir -p gpurun_out
python -m pytest tests/test_interp_gpu.py -x -q 2>&1 | tail -2
V=";7=1;7=4;7=8;7=8,3=1;7=8,2=2;7=8,1=4;7=8,1=16"
for T in 125 250; do
timeout 300 python scripts/interp_lab.py --snapshots $T --variants "$V" > gpurun_out/lab14_t$T.jsonl 2> gpurun_out/lab14_t$T.err; tail -2 gpurun_out/lab14_t$T.err
done
