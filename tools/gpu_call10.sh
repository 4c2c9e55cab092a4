#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_default.json; echo
