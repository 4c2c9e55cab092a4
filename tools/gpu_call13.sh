#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload navier_stokes_2d --steps 1 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e --jacobi-sweeps 60"
PML_JACOBI_LOOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_jacobi_loop -s 1 -c 1 -f -o gpurun_out/jac_loop $B > gpurun_out/ncu_jl.log 2>&1; echo "loop rc=$?"
PML_JACOBI_LOOP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_jacobi_sweep -s 70 -c 1 -f -o gpurun_out/jac_sweep $B > gpurun_out/ncu_js.log 2>&1; echo "sweep rc=$?"
# the reference arm the way the driver launches it at N = 8 (host processes only)
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/ref_n8.json 2> gpurun_out/ref_n8.err; echo "ref n8 rc=$?"; cat gpurun_out/ref_n8.json | head -c 1500; tail -4 gpurun_out/ref_n8.err
