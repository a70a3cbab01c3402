#!/bin/bash
# GPU check used during development: tests, one bench line, and an ncu source capture of the largest tile launch.
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo pytest_rc=$?
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
if [ "$2" = "ncu" ]; then
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s ${3:-7} -c 1 -o gpurun_out/prof_tile_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1; echo ncu_rc=$?
fi
