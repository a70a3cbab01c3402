#!/bin/bash
# last pass of the round on the final tile kernel: tests, ncu capture -> limiter summary, short bench line
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_final.log
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 7 -c 1 -o gpurun_out/r2h_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_tile_r2h.log 2>&1; echo tile_rc=$?
python benchmarks/ncu_limiter.py gpurun_out/r2h_tile.ncu-rep profiles/r2_tile_limiter.json > gpurun_out/limiter_r2h.log 2>&1; echo limiter_rc=$?
cp profiles/r2_tile_limiter.json gpurun_out/r2_tile_limiter.json
ncu -i gpurun_out/r2h_tile.ncu-rep --page raw --csv > gpurun_out/r2h_ncu_k_tile_score.csv 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_r2h_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['ids_checksum'], d['roofline']['limiter']['profile_matches_shipped_source'])"
