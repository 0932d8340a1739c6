#!/bin/bash
# compute-sanitizer pass over the hand-written kernels: small shapes of every kernel family,
# run under memcheck (out-of-bounds / misaligned global, shared, TMEM and TMA accesses), then synccheck
# (barrier misuse) on the warp-specialised kernels.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Output: gpurun_out/sanitize_<tool>.log (+ a one-line summary each in gpurun_out/sanitize_summary.txt).
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL_MEM='test_ln_modulate or test_patch_embed or test_small_linear or test_label_embed or (test_gemm_tcgen05_bias and (256-256-64 or 300-200-72 or 64-6912)) or (test_gemm_tcgen05_epilogues and 200-384) or (test_attention_bf16 and (2-256-6-64 or 2-16-6-64 or 2-100-3-72)) or (test_attention_tcgen05_persistent and (1-128-1-64 or 3-128-5-72 or 1-512-1-64)) or test_final_layer or test_cfg_combine'
SEL_SYNC='(test_gemm_tcgen05_bias and 256-256-64) or (test_attention_tcgen05_persistent and (3-128-5-72 or 1-512-1-64)) or (test_gemm_tcgen05_epilogues and 200-384)'
: > gpurun_out/sanitize_summary.txt
run() {  # tool, selection, per-run limit (s)
  local tool=$1 sel=$2 lim=$3
  timeout "$lim" $SAN --tool "$tool" --target-processes all --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$sel" -p no:cacheprovider > gpurun_out/sanitize_$tool.log 2>&1
  local rc=$?
  echo "$tool rc=$rc $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_$tool.log | tr '\n' ' ')" >> gpurun_out/sanitize_summary.txt
}
run memcheck "$SEL_MEM" ${SAN_LIMIT:-600}
run synccheck "$SEL_SYNC" ${SAN_LIMIT:-400}
cat gpurun_out/sanitize_summary.txt
