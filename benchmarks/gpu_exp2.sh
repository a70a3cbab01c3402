#!/bin/bash
# development helper: ncu source-level captures - tile kernel (1024-tile launch) and three mid-size cosine GEMM launches
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --opt sparse_mode=0 > gpurun_out/plain_exp2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 8 -c 1 -o gpurun_out/exp2_tile python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --opt sparse_mode=0 > gpurun_out/ncu_exp2_tile.log 2>&1; echo tile_rc=$?
python benchmarks/cos_once.py > gpurun_out/plain_exp2_cos.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cosine_gemm_qs -s 19 -c 3 -o gpurun_out/exp2_cos python benchmarks/cos_once.py > gpurun_out/ncu_exp2_cos.log 2>&1; echo cos_rc=$?
ls -la gpurun_out/*.ncu-rep
