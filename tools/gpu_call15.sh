#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fdm.py tests/test_gpu_parareal.py tests/test_gpu_slab.py -x -q > gpurun_out/pytest_fe2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_fe2.log
python tools/fe_bench.py 512 16
PML_FVARIANT=1 python tools/fe_bench.py 512 16
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"; }
run stream1 PML_STREAM=1
run base2 PML_STREAM=0
