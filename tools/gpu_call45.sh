#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call45.log 2>&1
for t in 64 128 256 512 1024; do DITB200_LN_THREADS=$t python tools/ln_probe.py; done
for t in 128 256; do M=8192 DITB200_LN_THREADS=$t python tools/ln_probe.py; done
