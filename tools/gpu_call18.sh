#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_final.json; echo
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
PML_FPATH1=0 timeout 300 $B > gpurun_out/bench_nopath1.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bench_nopath1.json'));print('nopath1', d['ms_per_step'], d['value'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/final_pair python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_b.log 2>&1
