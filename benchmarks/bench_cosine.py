#!/usr/bin/env python
"""Secondary benchmark: the dense cosine kernels (BASELINE.json configs[4] per-GPU shard and configs[2]).
  C5 shard : brute-force cosine top-10, 1.25M x 768 bf16 docs (one of 8 shards of 10M), 10k queries
             -> TFLOP/s against MEASURED_PEAKS.json bf16_tflops (tensor roofline)
  C3       : cosine re-rank of 1000 candidates per query (BM25 top-1000 shape), 207,363 docs, 10k queries
             -> GB/s of gathered rows against hbm_gbs (gather / HBM roofline)
Prints one JSON line per config.  CPU baseline: torch.matmul + topk fp32 on the host (team_run1.py:280-282)
on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200.cosine import CosineIndex  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-docs", type=int, default=100_000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    hbm, tf_burst, tf_sus, src = peaks()
    g = torch.Generator(device=dev).manual_seed(20241105 + 5)
    docs = torch.randn(args.docs, args.dim, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    qs = torch.randn(args.queries, args.dim, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ix = CosineIndex(docs)
    ms, (ids, sims) = timed(lambda: ix.topk(qs, 10), args.steps, args.warmup)
    flop = 2.0 * args.docs * args.queries * args.dim
    tfs = flop / (ms * 1e-3) / 1e12
    # spot check against torch on the GPU (fp32 of the bf16 values), 64 queries
    dn = docs.float()
    dn = dn / (dn.norm(dim=1, keepdim=True) + 1e-10)
    qn = qs[:64].float()
    qn = qn / (qn.norm(dim=1, keepdim=True) + 1e-10)
    ref = torch.topk(qn @ dn.T, 10)
    ok_ids = float((ref.indices == ids[:64]).float().mean().item())
    max_err = float((ref.values - sims[:64]).abs().max().item())
    # CPU baseline: torch fp32 on the host over a doc subsample
    cd = docs[:args.cpu_docs].float().cpu()
    cq = qs.float().cpu()
    t0 = time.time()
    cdn = cd / (cd.norm(dim=1, keepdim=True) + 1e-10)
    cqn = cq / (cq.norm(dim=1, keepdim=True) + 1e-10)
    torch.topk(cqn @ cdn.T, 10)
    cpu_s = time.time() - t0
    cpu_qps_full = args.queries / (cpu_s * args.docs / args.cpu_docs)
    print(json.dumps({
        "metric": "cosine top-10 queries/sec (brute force)", "value": args.queries / (ms * 1e-3), "unit": "queries/s",
        "ms_per_step": ms, "config": {"workload": f"C5 shard: {args.docs} x {args.dim} bf16 docs, {args.queries} queries, k=10"},
        "roofline": {"bound": "tensor", "achieved": tfs, "peak": tf_burst, "peak_sustained": tf_sus, "unit": "TFLOP/s",
                     "frac": tfs / tf_burst, "frac_of_sustained": tfs / tf_sus, "peak_source": src,
                     "note": "whole call: GEMM chunks + tighten kernels + query norms"},
        "check": {"ids_equal_torch_fp32_frac": ok_ids, "max_abs_sim_err": max_err},
        "cpu_baseline": {"value": cpu_qps_full, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"torch fp32 matmul+topk, all queries x first {args.cpu_docs} docs ({cpu_s:.1f} s), scaled to {args.docs} docs"},
    }))
    # ---- C3: re-rank of 1000 candidates per query
    n3, c = 207_363, 1000
    docs3 = docs[:n3].contiguous()
    ix3 = CosineIndex(docs3)
    cand = torch.randint(0, n3, (args.queries, c), generator=g, device=dev, dtype=torch.int32)
    ms3, (i3, s3) = timed(lambda: ix3.rerank(qs, cand, 10), args.steps, args.warmup)
    bytes3 = args.queries * (c * args.dim * 2 + args.dim * 2 + 80)
    print(json.dumps({
        "metric": "cosine re-rank queries/sec (1000 candidates -> top-10)", "value": args.queries / (ms3 * 1e-3),
        "unit": "queries/s", "ms_per_step": ms3,
        "config": {"workload": f"C3: {n3} x {args.dim} bf16 docs, {args.queries} queries x {c} candidates, k=10"},
        "roofline": {"bound": "hbm", "achieved": bytes3 / (ms3 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": bytes3 / (ms3 * 1e-3) / 1e9 / hbm, "peak_source": src,
                     "note": "algorithmic bytes = c*D*2 + D*2 + 8k per query (SURVEY 8d); the 318 MB table is L2-missing"},
    }))


if __name__ == "__main__":
    main()
