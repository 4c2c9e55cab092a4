#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_fdm.py tests/test_gpu_differentiator.py -x -q > gpurun_out/pytest_slab_ns.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_slab_ns.log
for L in 1 0; do
PML_JACOBI_LOOP=$L timeout 300 python bench.py --workload navier_stokes_2d --steps 3 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e --jacobi-sweeps 200 > gpurun_out/bench_ns_loop$L.json 2> gpurun_out/bench_ns_loop$L.err; echo "ns loop=$L rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_ns_loop$L.json'));print(d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['jacobi_sweeps_per_step'])" 2>&1 | tail -1)"
done
PML_JACOBI_LOOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ns.csv python bench.py --workload navier_stokes_2d --steps 1 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e --jacobi-sweeps 200 > gpurun_out/ncu_ns.log 2>&1
grep -E "jacobi|stage" gpurun_out/launches_ns.csv | awk -F'","' '{print $5, $NF}' | tail -12
