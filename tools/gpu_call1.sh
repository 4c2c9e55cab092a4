#!/bin/bash
# first GPU call of the session: tests, bench of both pair-kernel variants, launch list
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench v2 rc=$?"
PML_FVARIANT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; echo "bench v1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
head -c 600 gpurun_out/bench_v2.json; echo
head -c 600 gpurun_out/bench_v1.json; echo
