#!/bin/bash
# development helper (round 2, session 3): A/B of the tile kernel's sparse phase and the cosine epilogue variants
python -m pytest tests/test_gpu_bm25.py tests/test_gpu_configs.py tests/test_gpu_cosine.py tests/test_gpu_properties.py -m gpu -x -q > gpurun_out/pytest_exp1.log 2>&1; echo pytest_rc=$?
tail -5 gpurun_out/pytest_exp1.log
python benchmarks/sweep_cosine_schedule.py "qs_epi=0" "qs_epi=1" "qs_epi=2" "qs_epi=2 chunk_mult=4" "qs_epi=2 chunk0=4 chunk_mult=4" "qs_epi=2 chunk_mult=8" > gpurun_out/cos_exp1.log 2>&1; echo cos_rc=$?
cat gpurun_out/cos_exp1.log
i=0
for opts in "sparse_mode=1" "sparse_mode=0" "sparse_mode=2"; do
  args=""
  for o in $opts; do args="$args --opt $o"; done
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary $args > gpurun_out/bench_exp1_$i.json 2> gpurun_out/bench_exp1_$i.err
  echo "== $opts rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_exp1_$i.json").read())
r=d["roofline"]
print("qps %.0f e2e %.0f ms/step %.2f kernel_ms %.2f share %.3f launches %d path %s cks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms_per_launch"]*r["launches_per_step"], r["kernel_share_of_step"], r["launches_per_step"], r["path"], d["config"]["ids_checksum"]))
PY
  i=$((i+1))
done
