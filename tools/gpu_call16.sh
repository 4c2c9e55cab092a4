#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fdm.py tests/test_gpu_differentiator.py tests/test_gpu_batch.py tests/test_gpu_initial_conditions.py -x -q > gpurun_out/pytest_lap.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_lap.log
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'], [ (p.get('max_rel_err'), p.get('ok')) for p in (d.get('parity') or [])])" 2>&1 | tail -1)"; }
run lap1 PML_FAST_LAPLACIAN=1
run lap0 PML_FAST_LAPLACIAN=0 
run lap1_r1 PML_FAST_LAPLACIAN=1 PML_FROWS=1
for w in diffusion_2d; do
timeout 300 $B --no-parity --workload $w > gpurun_out/bench_${w}_lap.json 2>/dev/null; echo "$w $(python -c "import json;d=json.load(open('gpurun_out/bench_${w}_lap.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
done
