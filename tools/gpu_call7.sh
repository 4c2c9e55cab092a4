#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/m2_r2d2 $B > gpurun_out/ncu_m2.log 2>&1; echo "rc=$?"
