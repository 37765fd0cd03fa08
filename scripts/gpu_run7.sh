mkdir -p gpurun_out
python -m pytest tests/test_export_multi_gpu.py tests/test_svd_multi_gpu.py -x -q > gpurun_out/t_multi.log 2>&1; tail -5 gpurun_out/t_multi.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench7_n1.json 2> gpurun_out/bench7_n1.err; tail -c 300 gpurun_out/bench7_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench7_n2.json 2> gpurun_out/bench7_n2.err; tail -c 300 gpurun_out/bench7_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench7_ref_n2.json 2> gpurun_out/bench7_ref_n2.err; tail -c 300 gpurun_out/bench7_ref_n2.err
