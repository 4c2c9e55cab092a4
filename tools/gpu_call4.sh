#!/bin/bash
mkdir -p gpurun_out
export PML_FVARIANT=2
B="python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
PML_FROWS=2 PML_FDEPTH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/m_r2 $B > gpurun_out/ncu_m_r2.log 2>&1; echo "r2 rc=$?"
PML_FROWS=1 PML_FDEPTH=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/m_r1 $B > gpurun_out/ncu_m_r1.log 2>&1; echo "r1 rc=$?"
