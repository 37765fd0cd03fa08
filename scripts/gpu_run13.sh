mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale13_n$N.json 2> gpurun_out/scale13_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale13_n$N.json 2> gpurun_out/scale13_n$N.err
  fi
  tail -c 200 gpurun_out/scale13_n$N.err
done
python -m pytest tests/test_export_multi_gpu.py tests/test_svd_multi_gpu.py -q > gpurun_out/t_multi13.log 2>&1; tail -3 gpurun_out/t_multi13.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
