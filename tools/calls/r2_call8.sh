#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c8.log 2>&1
nvidia-smi -L | wc -l
echo "== 2-rank NCCL data-parallel test"
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s 2>&1 | tail -8
run() { # tag, nproc, extra args
  tag=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - $tag <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    print(f"{tag}: value={d['value']:.1f} ms/step={d['ms_per_step']:.2f} e2e={d['e2e']['value']:.1f} clk={d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-800:])
P
}
run r2c8_c4_n8 8 --workload c4 --steps 10 --warmup 5
run r2c8_c4_n8_overlap 8 --workload c4 --steps 10 --warmup 5 --overlap-opt
run r2c8_c4_n8_bf16 8 --workload c4 --steps 10 --warmup 5 --grad-dtype bf16
run r2c8_c4_n4 4 --workload c4 --steps 10 --warmup 5
run r2c8_c3_n8 8 --steps 2 --warmup 2
