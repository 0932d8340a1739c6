#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call20.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== C4 dgrad (trans_w), M=8192"
timeout 90 $P --m 8192 --n 4608 --k 1152 --cfgs 0x0,2x256,2x128 --trans-w
timeout 90 $P --m 8192 --n 1152 --k 4608 --cfgs 0x0,2x256,2x128 --trans-w
timeout 90 $P --m 8192 --n 1152 --k 1152 --cfgs 0x0,2x256,2x128 --trans-w
timeout 90 $P --m 8192 --n 1152 --k 3456 --cfgs 0x0,2x256,2x128 --trans-w
echo "=== C4 wgrad (trans_a, trans_w), K=8192 tokens"
for sk in 2 3 4 6 8; do
echo "--- split_k=$sk"
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 0x0,2x256 --trans-w --trans-a --split-k $sk
timeout 90 $P --m 4608 --n 1152 --k 8192 --cfgs 0x0,2x256 --trans-w --trans-a --split-k $sk
timeout 90 $P --m 1152 --n 1152 --k 8192 --cfgs 0x0,2x256 --trans-w --trans-a --split-k $sk
timeout 90 $P --m 3456 --n 1152 --k 8192 --cfgs 0x0,2x256 --trans-w --trans-a --split-k $sk
done
echo "=== cuBLAS references (NN / TN layouts via torch)"
python - <<'PY'
import torch
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
dev='cuda'
for (M,N,K) in [(8192,4608,1152),(8192,1152,4608),(8192,1152,1152),(8192,1152,3456)]:
    dy=torch.randn(M,K,device=dev).bfloat16(); w=torch.randn(K,N,device=dev).bfloat16()
    ms=t(lambda: dy@w); print(f"dgrad cuBLAS M={M} N={N} K={K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF")
for (M,N,K) in [(1152,4608,8192),(4608,1152,8192),(1152,1152,8192),(3456,1152,8192)]:
    dy=torch.randn(K,M,device=dev).bfloat16(); x=torch.randn(K,N,device=dev).bfloat16()
    ms=t(lambda: dy.t()@x); print(f"wgrad cuBLAS M={M} N={N} K={K}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF")
PY
echo "=== done"
