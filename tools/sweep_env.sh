#!/bin/bash
# usage: tools/sweep_env.sh "VAR=a,VAR2=b VAR=c ..." [bench args]; each word is a comma separated env set (use ; inside values as ,)
sets="$1"; shift
for s in $sets; do
  envs=$(echo "$s" | tr '+' ' ' | tr ';' ',')
  out=$(env $envs python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1)
  echo "$s $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],3), "Gcell-steps/s", round(d["ms_per_step"],3), "ms/step frac", round(d["roofline"]["frac"],3))')"
done
