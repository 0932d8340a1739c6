#!/bin/bash
# Round 2, call 19: final layer on the tensor cores (mma.sync TF32 x 3): tests, same-box A/B against the SIMT kernel
# (DITB200_FINAL_SIMT=1), ncu capture.
mkdir -p gpurun_out
exec > gpurun_out/r2c19.log 2>&1
echo "== tests"; timeout -k 10 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_bench_config_gpu.py tests/test_fork_dino_gpu.py tests/test_diffusion_api_gpu.py -m gpu -x -q 2>&1 | tail -4
L=fast_dit_b200/lib/libditb200.so
B="--steps 1 --warmup 1"
bash tools/ab_bench.sh r2c19_c3_tc $L -- $B
bash tools/ab_bench.sh r2c19_c3_simt $L DITB200_FINAL_SIMT=1 -- $B
bash tools/ab_bench.sh r2c19_c3_tc_b $L -- $B
bash tools/ab_bench.sh r2c19_c3_simt_b $L DITB200_FINAL_SIMT=1 -- $B
bash tools/ab_bench.sh r2c19_c1_tc $L -- --workload c1 --steps 3 --warmup 2
bash tools/ab_bench.sh r2c19_c1_simt $L DITB200_FINAL_SIMT=1 -- --workload c1 --steps 3 --warmup 2
echo "== ncu final layer"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export DITB200_GRAPH=0
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:final_layer_tc_kernel -s 300 -c 2 -o gpurun_out/r2c19_final $CMD > gpurun_out/r2c19_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2c19_final.ncu-rep
