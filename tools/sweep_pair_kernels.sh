#!/bin/bash
# Geometry sweep of the stage-pair kernels on the default bench workload (run
# under gpurun).  Each line: name, ms per step, Gcell-steps/s.
#   PML_FROWS   rows per thread          PML_FTILE  tile "tx,ty"
#   PML_FDEPTH  TMA prefetch distance    PML_FZC    planes per chunk
#   PML_FSYNC   1 = per-warp mbarrier arrivals instead of __syncthreads
#   PML_FVARIANT 1 = round-1 body        PML_FPATH1 0 = no intermediate boundary variant
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-workloads --no-parity --no-cpu-baseline --no-e2e"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_$name.json'));print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"; }
run default
run r1_t6 PML_FROWS=1 PML_FTILE=30,6
run r2_t14 PML_FROWS=2 PML_FTILE=30,14
run r1_t14_d1 PML_FDEPTH=1
run r1_t14_sync1 PML_FDEPTH=1 PML_FSYNC=1
run round1_body PML_FVARIANT=1
