#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
txt=[l for l in open('gpurun_out/bench_n2b.json') if l.startswith('{')][-1]
d=json.loads(txt)
print(d['value'], d['ms_per_step'], json.dumps(d['parareal']), json.dumps(d['spatial_decomposition'])[:500])
print([ (p.get('max_rel_err'), p.get('ok')) for p in d['parity']])
PY
head -c 80 gpurun_out/bench_n2b.json; echo; grep -v "Warning: \[PG ID\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2b.err | tail -5
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -k nccl 2>&1 | tail -2
