#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run2.log 2>&1
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --workload c4 --steps 10 --warmup 3 2>gpurun_out/err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   ', round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms/step')"
  grep -i "error\|Traceback" gpurun_out/err.log | head -3
}
echo "N=2 static"; DITB200_DDP_STATIC=1 run 29583
echo "N=2 everywhere-dynamic"; DITB200_GEMM_DYNAMIC=1 run 29581
echo "N=2 backward-dynamic"; run 29582
echo "N=2 static again"; DITB200_DDP_STATIC=1 run 29584
