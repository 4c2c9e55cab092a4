#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
}
run base PML_FPATH1=1
run nopath1 PML_FPATH1=0
run z256 PML_FZC=256
run z512 PML_FZC=512
run r1_nopath1 PML_FPATH1=0 PML_FROWS=1
run r2_t14_tx62 PML_FTILE=62,14
