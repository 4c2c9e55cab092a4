#!/bin/bash
# usage: tools/sweep_fused.sh "tile:depth[:zc] ..." [bench args]  (fused stage-pair kernels)
sets="$1"; shift
for s in $sets; do
  IFS=: read tile depth zc <<< "$s"
  out=$(env PML_FUSE=1 PML_FTILE=$tile PML_FDEPTH=$depth ${zc:+PML_FZC=$zc} timeout -k 10 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1)
  echo "$s $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],3), "Gcell-steps/s", round(d["ms_per_step"],3), "ms/step frac", round(d["roofline"]["frac"],3), "finite", d["config"]["finite"])' 2>&1 | tail -1)"
done
