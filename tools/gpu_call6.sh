#!/bin/bash
mkdir -p gpurun_out
export PML_FVARIANT=2
PML_FSYNC=1 timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_sync1.log 2>&1; echo "pytest sync1 rc=$?"; tail -3 gpurun_out/pytest_sync1.log
PML_FSYNC=1 PML_FROWS=1 timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_sync1_r1.log 2>&1; echo "pytest sync1 r1 rc=$?"; tail -3 gpurun_out/pytest_sync1_r1.log
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
}
run s1_r2_d1 PML_FSYNC=1 PML_FROWS=2 PML_FDEPTH=1
run s1_r2_t12_d2 PML_FSYNC=1 PML_FROWS=2 PML_FDEPTH=2 PML_FTILE=30,12
run s1_r1_d1 PML_FSYNC=1 PML_FROWS=1 PML_FDEPTH=1
run s1_r1_d2 PML_FSYNC=1 PML_FROWS=1 PML_FDEPTH=2
run s0_r2_d2 PML_FSYNC=0 PML_FROWS=2 PML_FDEPTH=2
run s0_r2_t62 PML_FSYNC=0 PML_FROWS=2 PML_FDEPTH=1 PML_FTILE=62,6
for w in shallow_water_polar diffusion_2d; do
for v in 1 2; do
PML_FVARIANT=$v timeout 300 $B --workload $w > gpurun_out/bench_${w}_v$v.json 2>gpurun_out/bench_${w}_v$v.err; echo "$w v$v $(python -c "import json;d=json.load(open('gpurun_out/bench_${w}_v$v.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
done; done
