#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"; }
run base
run r2_t18_d1 PML_FTILE=30,18 PML_FDEPTH=1
run r2_t18_d1_z64 PML_FTILE=30,18 PML_FDEPTH=1 PML_FZC=64
run r1_t14_d2 PML_FROWS=1 PML_FTILE=30,14 PML_FDEPTH=2
run r1_t10_d2 PML_FROWS=1 PML_FTILE=30,10 PML_FDEPTH=2
run r1_t6_d2 PML_FROWS=1 PML_FDEPTH=2
run r2_t14_d1 PML_FDEPTH=1
run r2_t16_d1 PML_FTILE=30,16 PML_FDEPTH=1
PML_FTILE=30,18 PML_FDEPTH=1 timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -k "burgers" 2>&1 | tail -2
