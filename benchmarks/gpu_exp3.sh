#!/bin/bash
# development helper: cosine compact epilogue A/B; tile kernel with 1024-doc sub-ranges (variant build) and groups of 2
python -m pytest tests/test_gpu_cosine.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_exp3.log 2>&1; echo pytest_rc=$?
tail -3 gpurun_out/pytest_exp3.log
python benchmarks/sweep_cosine_schedule.py "qs_epi=0" "qs_epi=1" "qs_epi=1 chunk_mult=4" "qs_epi=1 chunk0=4 chunk_mult=4" "qs_epi=1 chunk_mult=8" "qs_epi=1 chunk0=4 chunk_mult=16" > gpurun_out/cos_exp3.log 2>&1; echo cos_rc=$?
cat gpurun_out/cos_exp3.log
run() {
  tag=$1; so=$2; shift 2
  args=""
  for o in "$@"; do args="$args --opt $o"; done
  BR_B200_SO=$so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary $args > gpurun_out/bench_exp3_$tag.json 2> gpurun_out/bench_exp3_$tag.err
  echo "== $tag $so $* rc=$?"
  tail -2 gpurun_out/bench_exp3_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_exp3_$tag.json").read())
    r=d["roofline"]
    print("qps %.0f e2e %.0f ms/step %.2f kernel_ms %.2f share %.3f launches %d path %s cks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms_per_launch"]*r["launches_per_step"], r["kernel_share_of_step"], r["launches_per_step"], r["path"], d["config"]["ids_checksum"]))
except Exception as e:
    print("failed", e)
PY
}
S10=$PWD/document_retrieval_b200/libbr_b200_s10.so
run a "" sparse_mode=0 tile_g=2
run b $S10 sparse_mode=0 tile_g=2 tile_dense_min=64
run c $S10 sparse_mode=0 tile_g=2 tile_dense_min=32
run d $S10 sparse_mode=0 tile_dense_min=64
run e $S10 sparse_mode=1 tile_g=2 tile_dense_min=64
