mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale16_n$N.json 2> gpurun_out/scale16_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale16_n$N.json 2> gpurun_out/scale16_n$N.err
  fi
done
for N in 8 4 2 1; do
  if [ $N -eq 1 ]; then
    timeout 400 python scripts/scale_c4.py --steps 5 > gpurun_out/c4_scale_n$N.json 2> gpurun_out/c4_scale_n$N.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) scripts/scale_c4.py --steps 5 > gpurun_out/c4_scale_n$N.json 2> gpurun_out/c4_scale_n$N.err
  fi
  tail -c 300 gpurun_out/c4_scale_n$N.json
done
