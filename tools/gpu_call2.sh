#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
PML_FVARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/ws_full $B > gpurun_out/ncu_ws.log 2>&1
echo "ws rc=$?"
PML_FVARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/v1_full $B > gpurun_out/ncu_v1.log 2>&1
echo "v1 rc=$?"
ls -la gpurun_out
