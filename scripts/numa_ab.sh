#!/bin/bash
# A/B of the NUMA binding of bench.py's host side (e2e leg): usage numa_ab.sh [n_gpus]
N=${1:-1}
nvidia-smi topo -m 2>&1 | head -12
lscpu | grep -i "numa\|socket\|^CPU(s)"
python -c "import os; print('allowed cpus', len(os.sched_getaffinity(0)))"
for f in "" "--no-numa-bind" "" "--no-numa-bind"; do
  if [ "$N" = 1 ]; then cmd="python bench.py"; else
    cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N"; fi
  $cmd --steps 10 --warmup 3 $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N flag[$f]', 'e2e ms', round(d['e2e']['ms_per_step'],2), 'cores', d['config']['host_cores_bound_to_gpu_numa_node'], 'ms', round(d['ms_per_step'],4))"
done
