#!/bin/bash
# development helper: ncu source-level capture of one k_tile_score launch of bench.py (args: tag, skip count, bench options...)
tag=$1; skip=$2; shift 2
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary "$@" > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s $skip -c 1 -o gpurun_out/prof_tile_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary "$@" > gpurun_out/ncu_$tag.log 2>&1; echo ncu_rc=$?
