#!/bin/bash
# marching variant: parity tests, then a sweep of rows-per-thread / prefetch depth
mkdir -p gpurun_out
export PML_FVARIANT=2
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fdm.py -x -q > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_march.log
for R in 1 3; do
PML_FROWS=$R timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_march_r$R.log 2>&1; echo "pytest R=$R rc=$?"; tail -3 gpurun_out/pytest_march_r$R.log
done
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
for cfg in "2 1" "2 2" "1 1" "1 2" "3 1"; do
  set -- $cfg
  PML_FROWS=$1 PML_FDEPTH=$2 timeout 300 $B > gpurun_out/bench_m_r$1_d$2.json 2> gpurun_out/bench_m_r$1_d$2.err
  echo "R=$1 D=$2 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_m_r$1_d$2.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
done
PML_FVARIANT=1 timeout 300 $B > gpurun_out/bench_v1b.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bench_v1b.json'));print('v1', d['ms_per_step'], d['value'])"
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fused.py -x -q -k "burgers" > gpurun_out/memcheck_march.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck_march.log
