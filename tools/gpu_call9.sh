#!/bin/bash
# 2 GPUs: NCCL parity tests, Parareal + slab bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parareal.py tests/test_gpu_slab.py -x -q -m gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
