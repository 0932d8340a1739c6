#!/bin/bash
# Round 2, call 21: tensor-core final layer, third version (3-instruction TF32 split) (shift / scale rows cached in shared memory, contiguous
# tile ranges, interleaved MMA chains): tests, A/B against the SIMT kernel, ncu.
mkdir -p gpurun_out
exec > gpurun_out/r2c21.log 2>&1
echo "== tests"; timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_bench_config_gpu.py -m gpu -x -q 2>&1 | tail -4
L=fast_dit_b200/lib/libditb200.so
B="--steps 1 --warmup 1"
bash tools/ab_bench.sh r2c21_c3_tc $L -- $B | head -1
bash tools/ab_bench.sh r2c21_c3_simt $L DITB200_FINAL_SIMT=1 -- $B | head -1
bash tools/ab_bench.sh r2c21_c3_tc_b $L -- $B | head -1
bash tools/ab_bench.sh r2c21_c3_simt_b $L DITB200_FINAL_SIMT=1 -- $B | head -1
python - <<'P'
import json
for t in ['r2c21_c3_tc','r2c21_c3_simt','r2c21_c3_tc_b','r2c21_c3_simt_b']:
    d=json.loads(open(f'gpurun_out/{t}.json').read().strip().splitlines()[-1])
    print(t, 'final_layer us:', 1e3*d['kernel_breakdown_ms_per_denoise_step']['final_layer']['ms'])
P
echo "== ncu final layer"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export DITB200_GRAPH=0
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:final_layer_tc_kernel -s 300 -c 1 -o gpurun_out/r2c21_final $CMD > gpurun_out/r2c21_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2c21_final.ncu-rep
