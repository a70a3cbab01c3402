#!/bin/bash
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_exp5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 7 -c 1 -o gpurun_out/exp5_tile python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_exp5_tile.log 2>&1; echo tile_rc=$?
