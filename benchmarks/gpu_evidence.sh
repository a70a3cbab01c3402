#!/bin/bash
# Round evidence on one B200.  Order matters: the ncu capture of the tile kernel comes first and is summarised on the
# box (profiles/r2_tile_limiter.json, keyed by the SHA-256 of br_tile.cu), so that the bench line that follows carries
# `roofline.limiter` / `roofline.traffic` of the code it has just run.  Then: reference arm, ncu launch list of a short
# run of the same program, ncu --set full of the cosine GEMM.  Outputs under gpurun_out/ (copied to profiles/ by hand).
tag=${1:-r2}
light=${2:-}          # "light": skip the reference arm and the cosine GEMM capture (unchanged code)
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain2_${tag}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_tile_score -s 7 -c 1 -o gpurun_out/${tag}_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_tile_${tag}.log 2>&1; echo tile_rc=$?
python benchmarks/ncu_limiter.py gpurun_out/${tag}_tile.ncu-rep profiles/r2_tile_limiter.json > gpurun_out/limiter_${tag}.log 2>&1; echo limiter_rc=$?
cp profiles/r2_tile_limiter.json gpurun_out/r2_tile_limiter.json
ncu -i gpurun_out/${tag}_tile.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_k_tile_score.csv 2>/dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo bench_rc=$?
[ -z "$light" ] && python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo ref_rc=$?
# launch list (per-launch device time) of a short run of the same program
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ll_${tag}.log 2>&1; echo launchlist_rc=$?
[ -n "$light" ] && exit 0
# cosine GEMM with the single-site epilogue: the 64-tile ... 2048-tile launches of the second call
python benchmarks/cos_once.py > gpurun_out/plain3_${tag}.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_cosine_gemm_qs -s 19 -c 6 -o gpurun_out/${tag}_cos \
    python benchmarks/cos_once.py > gpurun_out/ncu_cos_${tag}.log 2>&1; echo cos_rc=$?
ncu -i gpurun_out/${tag}_cos.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_k_cosine_gemm_qs.csv 2>/dev/null
