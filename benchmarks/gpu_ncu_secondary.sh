#!/bin/bash
# ncu --set full captures (one launch each) of the secondary kernels, from the default bench command; summaries are
# extracted by hand into profiles/r2_ncu_secondary_kernels.csv
tag=${1:-r2}
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_sec_${tag}.log 2>&1 || exit 1
i=0
for spec in "k_cosine_rerank:0" "k_tighten_big:3" "k_select_radix:0" "k_tok_emit:0" "k_lookup:0" "k_cosine_gemm_qs:12" "k_final_select:1" "k_rescore_heads:1" "k_seed_thr:1" "k_dedupe_first_docs:0"; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none -k regex:$k -s $s -c 1 -o gpurun_out/${tag}_sec_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sec_${tag}_$k.log 2>&1
  echo "$k rc=$?"
done
