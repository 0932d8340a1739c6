#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c5.log 2>&1
nvidia-smi -L
echo "== 2-rank NCCL data-parallel test"
timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s 2>&1 | tail -12
echo "== C4 training, N=2 (f32 wire)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 10 --warmup 5 > gpurun_out/r2c5_c4_n2.json 2> gpurun_out/r2c5_c4_n2.err; tail -c 1500 gpurun_out/r2c5_c4_n2.json
echo "== C4 training, N=2 (bf16 wire)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 10 --warmup 5 --grad-dtype bf16 > gpurun_out/r2c5_c4_n2_bf16.json 2> gpurun_out/r2c5_c4_n2_bf16.err; tail -c 600 gpurun_out/r2c5_c4_n2_bf16.json
echo "== C3 sampling, N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r2c5_c3_n2.json 2> gpurun_out/r2c5_c3_n2.err; tail -c 400 gpurun_out/r2c5_c3_n2.json
echo "== reference arm (N=2: rank 0 only)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -3
