#!/bin/bash
# development helper: GPU tests + the default bench line (with cpu baseline and secondary figures)
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo pytest_rc=$?; tail -5 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_rc=$?
tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read())
print("qps %.0f e2e %.0f e2e_text %.0f ms %.2f share %.3f" % (d["value"], d["e2e"]["value"], d["e2e_text"]["value"], d["ms_per_step"], d["roofline"]["kernel_share_of_step"]))
print(json.dumps(d.get("secondary"), indent=0)[:1800])
print(d.get("cpu_baseline"), d["config"]["ids_checksum"], d["index_build"])
PY
