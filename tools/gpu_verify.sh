#!/bin/bash
# One verification pass on a GPU box (run under gpurun): all GPU tests, smoke(),
# the default bench line, a launch list and one full ncu capture of the pair kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_verify.sh'
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench.json; echo
B="python bench.py --steps 2 --warmup 1 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_fused -s 2 -c 2 -f -o gpurun_out/pair_kernels $B > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
