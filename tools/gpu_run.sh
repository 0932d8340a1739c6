#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run8.log 2>&1
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 1 --warmup 3 > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; echo "exit=$?"; tail -3 gpurun_out/bench_c3_n8.err; cat gpurun_out/bench_c3_n8.json
