#!/bin/bash
# Round evidence on one B200: default bench line, ncu launch list of the same command, ncu --set full captures of the
# tile kernel (largest launch) and of the secondary kernels.  Outputs under gpurun_out/ (copied to profiles/ by hand).
tag=${1:-r2}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo bench_rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo ref_rc=$?
# launch list (per-launch device time) of a short run of the same program
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll_${tag}.log 2>&1; echo launchlist_rc=$?
# tile kernel: the 1024-tile launch of the first step
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain2_${tag}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 8 -c 1 -o gpurun_out/${tag}_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_tile_${tag}.log 2>&1; echo tile_rc=$?
# secondary kernels: cosine GEMM (largest launch), candidate re-rank, radix select / final select / re-score of the C3 stage
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --queries 10000 > gpurun_out/plain3_${tag}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_cosine_rerank|k_rescore_heads|k_final_select|k_tighten_big|k_seed_thr|k_cold_pass' -c 12 -o gpurun_out/${tag}_secondary \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sec_${tag}.log 2>&1; echo sec_rc=$?
