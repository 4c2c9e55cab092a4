#!/bin/bash
mkdir -p gpurun_out
export PML_FVARIANT=2
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_march2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_march2.log
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
}
run r2_t14_d1 PML_FROWS=2 PML_FDEPTH=1
run r2_t14_d2 PML_FROWS=2 PML_FDEPTH=2
run r2_t14_d2_z64 PML_FROWS=2 PML_FDEPTH=2 PML_FZC=64
run r2_t6_d2 PML_FROWS=2 PML_FDEPTH=2 PML_FTILE=30,6
run r2_t6_d2_z64 PML_FROWS=2 PML_FDEPTH=2 PML_FTILE=30,6 PML_FZC=64
run r2_t10_d2 PML_FROWS=2 PML_FDEPTH=2 PML_FTILE=30,10
run r1_t6_d2 PML_FROWS=1 PML_FDEPTH=2
run r1_t6_d2_z64 PML_FROWS=1 PML_FDEPTH=2 PML_FZC=64
run v1 PML_FVARIANT=1
