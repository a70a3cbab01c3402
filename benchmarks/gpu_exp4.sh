#!/bin/bash
# development helper: source-level ncu capture of the 512-tile launch + sweep of the deferral budget
bash benchmarks/gpu_sweep.sh e4 "defer_pm=700" "defer_pm=650" "defer_pm=600" "defer_pm=550" "defer_pm=700 tile_dense_min=24" "defer_pm=650 tile_dense_min=24"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_exp4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 7 -c 1 -o gpurun_out/exp4_tile python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_exp4_tile.log 2>&1; echo tile_rc=$?
