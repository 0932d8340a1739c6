#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/cpu.log 2>&1
python tools/cpu_bound_probe.py c4
python tools/cpu_bound_probe.py c2
