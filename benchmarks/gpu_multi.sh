#!/bin/bash
# development helper: bench.py on N GPUs (torchrun), prints the summary
n=$1; tag=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err; echo rc=$?
tail -5 gpurun_out/bench_${tag}_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}_n$n.json").read())
print("N=%d qps %.0f e2e %.0f ms %.2f share %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_share_of_step"]))
print(d.get("verify"), d["config"]["ids_checksum"])
s=d.get("secondary") or {}
if "cosine_c5" in s: print({k: s["cosine_c5"][k] for k in ("ms","tflops_per_gpu","frac_of_bf16_burst","check")})
PY
