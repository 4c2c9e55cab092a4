#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
txt=[l for l in open('gpurun_out/bench_n8b.json') if l.startswith('{')][-1]
d=json.loads(txt)
print(d['value'], d['ms_per_step'], json.dumps(d['parareal']), json.dumps(d['spatial_decomposition'])[:500])
print([ (p.get('max_rel_err'), p.get('ok')) for p in d['parity']])
PY
grep -v "Warning: \[PG ID\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n8b.err | tail -5
