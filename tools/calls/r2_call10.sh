#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c10.log 2>&1
run() { # tag, nproc, extra args
  tag=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - $tag <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    kb = d.get("kernel_breakdown_ms_per_step", {})
    print(f"{tag}: value={d['value']:.1f} ms/step={d['ms_per_step']:.2f} e2e={d['e2e']['value']:.1f} clk={d['clocks']['sm_mhz']} gemm={kb.get('gemm_tc',{}).get('ms',0):.2f} adamw={kb.get('adamw_ema',{}).get('ms',0):.2f}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-1200:])
P
}
run r2c10_c4_n8 8 --workload c4 --steps 10 --warmup 5
run r2c10_c4_n8_shard 8 --workload c4 --steps 10 --warmup 5 --shard-opt
NCCL_ALGO=NVLS run r2c10_c4_n8_nvls 8 --workload c4 --steps 10 --warmup 5
grep -i "nvls" gpurun_out/r2c10_c4_n8_nvls.err | head -3
