#!/bin/bash
# 8 GPUs: NCCL parity tests (Parareal + slabs), then the driver's N = 8 bench line
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parareal.py tests/test_gpu_slab.py -x -q -m gpu -k "nccl" > gpurun_out/pytest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8.json'))
print(d['value'], d['ms_per_step'], json.dumps(d['parareal']), json.dumps(d['spatial_decomposition'])[:600])
print(json.dumps(d['parity'])[:1500])
print(json.dumps(d['e2e'])[:300])
PY
tail -3 gpurun_out/bench_n8.err
