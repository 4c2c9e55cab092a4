#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_slab.py -x -q -k "plane_range or tall or nccl" > gpurun_out/pytest_slab2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_slab2.log
PML_JACOBI_LOOP=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_ns_sweeps.csv python bench.py --workload navier_stokes_2d --steps 1 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e --jacobi-sweeps 200 > gpurun_out/ncu_ns2.log 2>&1
grep -E "jacobi" gpurun_out/launches_ns_sweeps.csv | awk -F'","' '{print $5, $NF}' | tail -5
# Cahn-Hilliard: DRAM traffic of the stage kernels at 256^3 and 512^3, and the fused (passthrough) alternative
for n in 256 512; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pml_stage -s 8 -c 4 --csv --log-file gpurun_out/ncu_ch_$n.csv python bench.py --workload cahn_hilliard_3d --grid $n --steps 2 --warmup 2 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/ncu_ch_$n.log 2>&1
done
for n in 256 512; do for fz in auto 1; do
PML_FUSE=$fz timeout 300 python bench.py --workload cahn_hilliard_3d --grid $n --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/bench_ch_${n}_fuse$fz.json 2>gpurun_out/bench_ch_${n}_fuse$fz.err; echo "ch $n fuse=$fz $(python -c "import json;d=json.load(open('gpurun_out/bench_ch_${n}_fuse$fz.json'));print(d['ms_per_step'], d['value'], d['roofline']['launches_per_step'])" 2>&1 | tail -1)"
done; done
