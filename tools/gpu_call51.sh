#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call51.log 2>&1
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --workload c4 --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   ', round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms/step')"
}
echo "default"; run 29521
echo "NCCL_MAX_CTAS=4"; NCCL_MAX_CTAS=4 run 29522
echo "NCCL_MAX_CTAS=2"; NCCL_MAX_CTAS=2 run 29523
echo "NCCL_MAX_CTAS=8"; NCCL_MAX_CTAS=8 run 29524
