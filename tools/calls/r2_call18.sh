#!/bin/bash
# Round 2, call 18 (2 GPUs): NCCL data-parallel correctness with the fused backward, C4 at N = 2.
mkdir -p gpurun_out
exec > gpurun_out/r2c18.log 2>&1
echo "== 2-rank NCCL data-parallel tests"
timeout -k 10 600 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s > gpurun_out/r2c18_nccl.log 2>&1; echo "rc=$?"
grep -v "^rank" gpurun_out/r2c18_nccl.log | tail -5
grep "^rank 0" gpurun_out/r2c18_nccl.log | tr ',' '\n' | grep -E "grad_vs_mean|vs_replicated|torch_ddp|equal_after" | head -20
run() { # tag, nproc, extra args
  tag=$1; n=$2; shift 2
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
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
run r2c18_c4_n2 2 --workload c4 --steps 20 --warmup 5
run r2c18_c4_n2_b 2 --workload c4 --steps 20 --warmup 5
