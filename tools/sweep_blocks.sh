#!/bin/bash
# usage: tools/sweep_blocks.sh "64,2,2 128,2,1 ..." [extra bench args]
shapes="$1"; shift
for b in $shapes; do
  out=$(PML_BLOCK=$b python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1)
  echo "$b $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],3), "Gcell-steps/s", round(d["ms_per_step"],3), "ms/step frac", round(d["roofline"]["frac"],3))')"
done
