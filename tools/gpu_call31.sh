#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call31.log 2>&1
nvidia-smi -L
echo "=== c3 N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench31_c3_n2.json 2> gpurun_out/bench31_c3_n2.err; echo "exit=$?"; tail -3 gpurun_out/bench31_c3_n2.err; cat gpurun_out/bench31_c3_n2.json
echo "=== c4 N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 > gpurun_out/bench31_c4_n2.json 2> gpurun_out/bench31_c4_n2.err; echo "exit=$?"; tail -3 gpurun_out/bench31_c4_n2.err; cat gpurun_out/bench31_c4_n2.json
echo "=== reference arm N=2 (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench31_ref.json 2> gpurun_out/bench31_ref.err; echo "exit=$?"; tail -3 gpurun_out/bench31_ref.err; cat gpurun_out/bench31_ref.json
echo "=== done"
