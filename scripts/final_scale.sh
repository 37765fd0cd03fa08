# Round-2 scaling run on one 8-GPU box (gpurun --gpus 8 -- 'bash scripts/final_scale.sh')
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  fi
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29950 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_scale_ref_n8.json 2> gpurun_out/r2_scale_ref_n8.err
python scripts/run_config.py C5 > gpurun_out/r2_cfg_C5.json 2> gpurun_out/r2_cfg_C5.err; tail -c 400 gpurun_out/r2_cfg_C5.json
