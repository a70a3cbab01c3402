#!/bin/bash
# sweep of the query-stationary cosine GEMM variants (C5 shard); one line per variant in gpurun_out/cos_sweep.log
out=gpurun_out/cos_sweep.log; : > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python benchmarks/bench_cosine.py --steps 3 --warmup 2 --cpu-docs 1000 2>&1 | head -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['achieved'], d['check'])" >> $out 2>&1; }
for v in "$@"; do run $v; done
cat $out
