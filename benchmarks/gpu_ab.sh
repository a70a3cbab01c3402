#!/bin/bash
# development helper: A/B of two builds of the library (libbr_b200.so vs libbr_b200_v.so) in one call, alternating
V=$PWD/document_retrieval_b200/libbr_b200_v.so
for i in 0 1; do
for so in "" $V; do
  BR_B200_SO=$so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab.json").read())
r=d["roofline"]
print("${so:-base} qps %.0f ms/step %.2f kernel_ms %.2f cks %s path %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_launch"]*r["launches_per_step"], d["config"]["ids_checksum"], r["path"]))
PY
done
done
